# vectorised tail loads: tail-related golden tests + the DLRM-row A/B again
cd /root/repo
mkdir -p gpurun_out
timeout 100 python -m pytest tests -m gpu -q -x -k "split_fast or typed64_prefix or trimmed or fp8_features or window_backpressure or exactly_once or pandas_output or equals_cpu" > gpurun_out/r2_gpu_tests_tail_vec.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_gpu_tests_tail_vec.log; tail -3 gpurun_out/r2_gpu_tests_tail_vec.log
timeout 60 python tools/tail_fields_bench.py > gpurun_out/r2_tail_fields_ab_vec.jsonl 2> gpurun_out/r2_tail_fields_ab_vec.err
cat gpurun_out/r2_tail_fields_ab_vec.jsonl
