mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_dataset.py -x -q -m gpu > gpurun_out/gpu_multi_tests.log 2>&1; tail -3 gpurun_out/gpu_multi_tests.log
run() { tag=$1; shift; timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --keep-data "$@" > gpurun_out/bench_n2_$tag.json 2> gpurun_out/bench_n2_$tag.err; tail -c 1600 gpurun_out/bench_n2_$tag.json; }
run stream
RSDL_BACKPRESSURE=host run host --wait-mode host
