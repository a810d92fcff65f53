cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/r2_gpu_tests_final2.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_gpu_tests_final2.log
tail -3 gpurun_out/r2_gpu_tests_final2.log
timeout 400 python bench.py --steps 20 --warmup 5 --keep-data > gpurun_out/r2_bench_n1_ours_final2.json 2> gpurun_out/r2_bench_n1_ours_final2.err; echo "ours exit $?"
timeout 300 python bench.py --schema dataspec --steps 20 --warmup 5 --skip-e2e > gpurun_out/r2_bench_n1_dataspec_ours_final2.json 2> gpurun_out/r2_bench_n1_dataspec_ours_final2.err; echo "ds exit $?"
python - <<'PY'
import json
for n in ("r2_bench_n1_ours_final2", "r2_bench_n1_dataspec_ours_final2"):
    d = json.loads(open(f"gpurun_out/{n}.json").read().strip().splitlines()[-1])
    print(n, {k: d.get(k) for k in ("value", "ms_per_epoch", "shuffle_kernel_ms_per_epoch", "host_enqueue_ms_per_epoch", "ingest_seconds")}, (d.get("e2e") or {}).get("value"), d["exactly_once"]["ok"])
PY
