# 2-GPU session: multi-GPU golden tests, N=2 bench (ours p2p / ours nccl / reference), NVLink A/B
cd /root/repo
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_gpu_multi.py -q > gpurun_out/r2_gpu_multi_tests_n2.log 2>&1
echo "pytest multi exit $?" >> gpurun_out/r2_gpu_multi_tests_n2.log
tail -4 gpurun_out/r2_gpu_multi_tests_n2.log
timeout 600 $TR --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --keep-data > gpurun_out/r2_bench_n2_ours.json 2> gpurun_out/r2_bench_n2_ours.err
echo "ours n2 exit $?"
timeout 600 $TR --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --keep-data --exchange nccl --skip-e2e > gpurun_out/r2_bench_n2_nccl.json 2> gpurun_out/r2_bench_n2_nccl.err
echo "nccl n2 exit $?"
timeout 900 $TR --master-port 29513 bench.py --impl reference --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_n2_ref.json 2> gpurun_out/r2_bench_n2_ref.err
echo "ref n2 exit $?"
cat gpurun_out/r2_bench_n2_ours.json gpurun_out/r2_bench_n2_nccl.json gpurun_out/r2_bench_n2_ref.json
timeout 300 $TR --master-port 29514 tools/a2a_ceiling.py > gpurun_out/r2_a2a_ceiling_n2.jsonl 2> gpurun_out/r2_a2a_ceiling_n2.err
cat gpurun_out/r2_a2a_ceiling_n2.jsonl
for rb in 256 1024 4096; do
  timeout 120 python tools/kernel_bench.py --bulk-store-probe $rb --peer 1 --iters 5 >> gpurun_out/r2_bulk_store_probe_peer.jsonl 2>> gpurun_out/r2_bulk_store_probe_peer.err
done
cat gpurun_out/r2_bulk_store_probe_peer.jsonl
timeout 600 bash tools/next_experiments.sh p2p > gpurun_out/r2_next_p2p.stdout 2>&1
tail -8 gpurun_out/r2_next_p2p.stdout
