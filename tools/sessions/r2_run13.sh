# BASELINE config 1 (CPU plumbing) on the 128-core box after the per-chunk DataFrame path
cd /root/repo
mkdir -p gpurun_out
L=gpurun_out/r2_cpu_plumbing_config1_v2.jsonl; rm -f $L
for rows in 1000000 10000000; do
  for impl in ours reference ours reference ours-numpy; do
    CUDA_VISIBLE_DEVICES="" timeout 300 python benchmarks/cpu_plumbing.py --impl $impl --num-rows $rows >> $L 2>> gpurun_out/r2_cpu_plumbing_config1_v2.err
  done
done
python - <<'PY'
import json
for line in open("gpurun_out/r2_cpu_plumbing_config1_v2.jsonl"):
    try:
        d = json.loads(line)
        print({k: d.get(k) for k in ("impl", "num_rows", "rows_per_sec", "seconds", "exactly_once")})
    except Exception:
        print(line[:200])
PY
CUDA_VISIBLE_DEVICES="" timeout 300 python -m pytest tests/test_dataset_cpu.py tests/test_native_cpu.py -q -x 2>&1 | tail -2
