cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/r2_gpu_tests_4.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_gpu_tests_4.log
tail -4 gpurun_out/r2_gpu_tests_4.log
timeout 400 python bench.py --steps 20 --warmup 5 --keep-data > gpurun_out/r2_bench_n1_ours_c.json 2> gpurun_out/r2_bench_n1_ours_c.err
timeout 300 python bench.py --steps 20 --warmup 5 --keep-data --skip-e2e --shuffle-priority high > gpurun_out/r2_bench_n1_ours_c_hiprio.json 2> gpurun_out/r2_bench_n1_ours_c_hiprio.err
python - <<'PY'
import json
for n in ("r2_bench_n1_ours_c", "r2_bench_n1_ours_c_hiprio"):
    d = json.loads(open(f"gpurun_out/{n}.json").read().strip().splitlines()[-1])
    print(n, {k: d.get(k) for k in ("value", "ms_per_epoch", "shuffle_kernel_ms_per_epoch", "host_enqueue_ms_per_epoch", "ingest_seconds", "exactly_once")}, (d.get("e2e") or {}).get("value"))
PY
timeout 300 python tools/ingest_bench.py > gpurun_out/r2_ingest_bench.jsonl 2> gpurun_out/r2_ingest_bench.err
cat gpurun_out/r2_ingest_bench.jsonl
for tool in synccheck racecheck; do
  timeout 400 bash tools/sanitize.sh $tool gpurun_out/r2_sanitize_$tool.log
  echo "$tool exit $?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|overall exit" gpurun_out/r2_sanitize_$tool.log | sort | uniq -c
done
