# N=2: bytes-on-wire variants of the headline table (bf16 / block-scaled fp8 epilogues) + run-to-run spread
cd /root/repo
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
i=0
for extra in "" "" "--feature-dtype bfloat16" "--feature-dtype fp8"; do
  i=$((i+1))
  timeout 240 $TR --master-port $((29550+i)) bench.py --gpus 2 --steps 20 --warmup 5 --keep-data --skip-e2e $extra > gpurun_out/r2_bench_n2_variant_$i.json 2> gpurun_out/r2_bench_n2_variant_$i.err
  echo "variant $i ($extra) exit $?"
done
python - <<'PY'
import json
for i in range(1, 5):
    try:
        d = json.loads(open(f"gpurun_out/r2_bench_n2_variant_{i}.json").read().strip().splitlines()[-1])
        print(i, d["dtype"], {k: d.get(k) for k in ("value", "ms_per_epoch", "shuffle_kernel_ms_per_epoch", "nvlink_egress_gbps_per_gpu")}, d["engine"]["row_bytes"], (d.get("exactly_once") or {}).get("ok"))
    except Exception as e:
        print(i, "unreadable", e)
PY
