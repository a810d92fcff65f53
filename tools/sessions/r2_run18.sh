# last call of the round: the full GPU tier on the final tree
cd /root/repo
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q > gpurun_out/r2_gpu_tests_last.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_gpu_tests_last.log; tail -4 gpurun_out/r2_gpu_tests_last.log
