cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/r2_gpu_tests_5.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_gpu_tests_5.log
tail -4 gpurun_out/r2_gpu_tests_5.log
timeout 300 python tools/ingest_bench.py --threads 8 16 32 > gpurun_out/r2_ingest_bench_ring.jsonl 2> gpurun_out/r2_ingest_bench_ring.err
cat gpurun_out/r2_ingest_bench_ring.jsonl; tail -3 gpurun_out/r2_ingest_bench_ring.err
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1_ours_d.json 2> gpurun_out/r2_bench_n1_ours_d.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_n1_ours_d.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value", "ms_per_epoch", "shuffle_kernel_ms_per_epoch", "host_enqueue_ms_per_epoch", "ingest_seconds", "cold_rows_per_sec", "exactly_once")}, (d.get("e2e") or {}).get("value"))
PY
timeout 400 bash tools/sanitize.sh synccheck gpurun_out/r2_sanitize_synccheck.log
grep -E "ERROR SUMMARY|overall exit" gpurun_out/r2_sanitize_synccheck.log | sort | uniq -c
