cd /root/repo
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 500 python -m pytest tests/test_gpu_multi.py -q -x > gpurun_out/r2_gpu_multi_tests_n2_final.log 2>&1
echo "pytest multi exit $?" >> gpurun_out/r2_gpu_multi_tests_n2_final.log
tail -4 gpurun_out/r2_gpu_multi_tests_n2_final.log
timeout 300 $TR --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 5 --keep-data > gpurun_out/r2_bench_n2_ours_final.json 2> gpurun_out/r2_bench_n2_ours_final.err; echo "ours n2 exit $?"
timeout 300 $TR --master-port 29542 bench.py --gpus 2 --steps 20 --warmup 5 --keep-data --skip-e2e --reducers-per-trainer 4 > gpurun_out/r2_bench_n2_ours_k7_default.json 2> gpurun_out/r2_bench_n2_ours_k7_default.err; echo "k7 default exit $?"
timeout 300 $TR --master-port 29543 bench.py --gpus 2 --steps 20 --warmup 5 --keep-data --skip-e2e --reducers-per-trainer 4 --chunk-passes 2 > gpurun_out/r2_bench_n2_ours_k7_p2.json 2> gpurun_out/r2_bench_n2_ours_k7_p2.err; echo "k7 p2 exit $?"
timeout 500 $TR --master-port 29544 bench.py --impl reference --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_n2_ref_final.json 2> gpurun_out/r2_bench_n2_ref_final.err; echo "ref n2 exit $?"
python - <<'PY'
import json
for n in ("r2_bench_n2_ours_final", "r2_bench_n2_ours_k7_default", "r2_bench_n2_ours_k7_p2", "r2_bench_n2_ref_final"):
    try:
        d = json.loads(open(f"gpurun_out/{n}.json").read().strip().splitlines()[-1])
        print(n, {k: d.get(k) for k in ("value", "ms_per_step", "ms_per_epoch", "shuffle_kernel_ms_per_epoch", "first_chunk_ms", "nvlink_egress_gbps_per_gpu")}, (d.get("e2e") or {}).get("value"), (d.get("engine") or {}).get("chunk_passes"), d.get("exactly_once", {}) and d["exactly_once"].get("ok"))
    except Exception as e:
        print(n, "unreadable", e)
PY
