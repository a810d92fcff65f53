# tail fields folded into the fast kernel: full GPU test pass (incl. the new tail / fp8+label cases)
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r2_gpu_tests_tail.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_gpu_tests_tail.log; tail -6 gpurun_out/r2_gpu_tests_tail.log
timeout 200 python bench.py --steps 20 --warmup 5 --skip-e2e --min-timed-epochs 10 > gpurun_out/r2_bench_n1_tailbuild.json 2> gpurun_out/r2_bench_n1_tailbuild.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_n1_tailbuild.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value", "ms_per_epoch", "shuffle_kernel_ms_per_epoch")}, d["exactly_once"]["ok"])
PY
