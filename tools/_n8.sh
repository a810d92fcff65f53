mkdir -p gpurun_out
timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --keep-data > gpurun_out/bench_n8_v18.json 2> gpurun_out/bench_n8_v18.err
tail -c 1500 gpurun_out/bench_n8_v18.json; tail -5 gpurun_out/bench_n8_v18.err
timeout -k 10 150 python examples/horovod/ray_torch_shuffle.py --num-workers 8 --model resnet50 --image-size 64 --num-rows 65536 --num-files 8 --batch-size 256 --epochs 3 --num-reducers 8 --bf16 --log-interval 1000 --data-dir /tmp/rsdl_resnet > gpurun_out/resnet50_n8.log 2>&1
grep -E "stats over|Mean batch wait" gpurun_out/resnet50_n8.log | tail -6
