# final GPU test pass (the golden's epoch production changed to the inverse-permutation gather)
# + BASELINE config 1 (CPU plumbing) on the 128-core box, both arms, two sizes
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/r2_gpu_tests_final3.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_gpu_tests_final3.log; tail -3 gpurun_out/r2_gpu_tests_final3.log
L=gpurun_out/r2_cpu_plumbing_config1.jsonl; rm -f $L
for rows in 1000000 10000000; do
  for impl in reference ours-numpy ours reference ours; do
    CUDA_VISIBLE_DEVICES="" timeout 300 python benchmarks/cpu_plumbing.py --impl $impl --num-rows $rows >> $L 2>> gpurun_out/r2_cpu_plumbing_config1.err
  done
done
python - <<'PY'
import json
for line in open("gpurun_out/r2_cpu_plumbing_config1.jsonl"):
    try:
        d = json.loads(line)
        print({k: d.get(k) for k in ("impl", "num_rows", "rows_per_sec", "seconds", "exactly_once")})
    except Exception:
        print(line[:200])
PY
