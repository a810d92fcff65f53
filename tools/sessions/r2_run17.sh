# fp8 epilogue with paired e4m3x2 conversions: golden tests + kernel timing (1 GPU)
cd /root/repo
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -k "fp8 or packed" > gpurun_out/r2_gpu_tests_fp8x2.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_gpu_tests_fp8x2.log; tail -3 gpurun_out/r2_gpu_tests_fp8x2.log
for m in 2 1 0; do
  timeout 100 python tools/kernel_bench.py --rows 12500000 --cols 64 --mode $m --iters 8 --tag "mode$m fp8x2-build" >> gpurun_out/r2_kbench_fp8x2.jsonl 2>> gpurun_out/r2_kbench_fp8x2.err
done
python - <<'PY'
import json
for line in open("gpurun_out/r2_kbench_fp8x2.jsonl"):
    d = json.loads(line); print(d["tag"], "pitch", d["row_pitch"], "ms %.3f" % d["ms_best"], "gbps %.0f" % d["gbps_best"])
PY
