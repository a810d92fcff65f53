# N=2: fp8 features + float32 label, label folded into the fast kernel as a tail field
cd /root/repo
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 240 $TR --master-port 29571 bench.py --gpus 2 --steps 20 --warmup 5 --skip-e2e --feature-dtype fp8 --keep-data > gpurun_out/r2_bench_n2_fp8_tail.json 2> gpurun_out/r2_bench_n2_fp8_tail.err
echo "fp8 exit $?"
timeout 240 $TR --master-port 29572 bench.py --gpus 2 --steps 20 --warmup 5 --skip-e2e > gpurun_out/r2_bench_n2_f32_tailbuild.json 2> gpurun_out/r2_bench_n2_f32_tailbuild.err
echo "f32 exit $?"
python - <<'PY'
import json
for n in ("r2_bench_n2_fp8_tail", "r2_bench_n2_f32_tailbuild"):
    d = json.loads(open(f"gpurun_out/{n}.json").read().strip().splitlines()[-1])
    print(n, {k: d.get(k) for k in ("value", "ms_per_epoch", "shuffle_kernel_ms_per_epoch", "scatter_launches_per_epoch", "nvlink_egress_gbps_per_gpu")}, d["engine"]["row_bytes"])
PY
