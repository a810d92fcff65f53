cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r2_gpu_tests_final.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_gpu_tests_final.log
tail -4 gpurun_out/r2_gpu_tests_final.log
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/r2_smoke.log
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 --keep-data > gpurun_out/r2_bench_n1_ref_final.json 2> gpurun_out/r2_bench_n1_ref_final.err; echo "ref exit $?"
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1_ours_final.json 2> gpurun_out/r2_bench_n1_ours_final.err; echo "ours exit $?"
timeout 600 python bench.py --impl reference --schema dataspec --steps 20 --warmup 5 --keep-data > gpurun_out/r2_bench_n1_dataspec_ref_final.json 2> gpurun_out/r2_bench_n1_dataspec_ref_final.err; echo "ref ds exit $?"
timeout 400 python bench.py --schema dataspec --steps 20 --warmup 5 > gpurun_out/r2_bench_n1_dataspec_ours_final.json 2> gpurun_out/r2_bench_n1_dataspec_ours_final.err; echo "ours ds exit $?"
python - <<'PY'
import json
for n in ("r2_bench_n1_ref_final", "r2_bench_n1_ours_final", "r2_bench_n1_dataspec_ref_final", "r2_bench_n1_dataspec_ours_final"):
    try:
        d = json.loads(open(f"gpurun_out/{n}.json").read().strip().splitlines()[-1])
        print(n, {k: d.get(k) for k in ("value", "ms_per_step", "steps", "epochs_timed", "ms_per_epoch", "shuffle_kernel_ms_per_epoch", "ingest_seconds", "timed_region")}, (d.get("e2e") or {}).get("value"), (d.get("engine") or {}).get("row_bytes"), d.get("exactly_once", {}) and d["exactly_once"].get("ok"))
    except Exception as e:
        print(n, "unreadable", e)
PY
