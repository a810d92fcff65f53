# fp8: one store instruction per row (unified path): golden tests incl. 2-GPU + N=2 bench
cd /root/repo
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 400 python -m pytest tests -m gpu -q -k "fp8 or multi_gpu or packed or tail or split" > gpurun_out/r2_gpu_tests_fp8_unified.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_gpu_tests_fp8_unified.log; tail -3 gpurun_out/r2_gpu_tests_fp8_unified.log
timeout 240 $TR --master-port 29581 bench.py --gpus 2 --steps 20 --warmup 5 --skip-e2e --feature-dtype fp8 > gpurun_out/r2_bench_n2_fp8_unified.json 2> gpurun_out/r2_bench_n2_fp8_unified.err
echo "fp8 exit $?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_n2_fp8_unified.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value", "ms_per_epoch", "shuffle_kernel_ms_per_epoch", "nvlink_egress_gbps_per_gpu")}, d["engine"]["row_bytes"])
PY
