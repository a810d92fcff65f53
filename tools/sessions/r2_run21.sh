cd /root/repo
mkdir -p gpurun_out
timeout 60 python -m pytest tests/test_gpu_multi.py -q -x -k "two_ranks_on_one_gpu and 2" > gpurun_out/r2_gpu_two_ranks_one_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_gpu_two_ranks_one_gpu.log; tail -15 gpurun_out/r2_gpu_two_ranks_one_gpu.log
