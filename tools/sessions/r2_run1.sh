cd /root/repo
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2_run1_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gpu_tests_1.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_gpu_tests_1.log
timeout 600 python bench.py --steps 20 --warmup 5 --keep-data > gpurun_out/r2_bench_n1_ours.json 2> gpurun_out/r2_bench_n1_ours.err
echo "ours exit $?"
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_bench_n1_ref.json 2> gpurun_out/r2_bench_n1_ref.err
echo "ref exit $?"
tail -3 gpurun_out/r2_gpu_tests_1.log
cat gpurun_out/r2_bench_n1_ours.json gpurun_out/r2_bench_n1_ref.json
