# 8-GPU session (one call): headline N=8 / N=4, K7 at scale, NCCL baseline, a2a ceiling, wide rows, ResNet-50
cd /root/repo
mkdir -p gpurun_out
TR8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
run() { name=$1; shift; timeout 420 "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; echo "$name exit $?"; }
run r2_bench_n8_ours $TR8 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 --keep-data
run r2_bench_n8_ours_k7 $TR8 --master-port 29522 bench.py --gpus 8 --steps 20 --warmup 5 --keep-data --reducers-per-trainer 4 --skip-e2e
run r2_bench_n8_nccl $TR8 --master-port 29523 bench.py --gpus 8 --steps 20 --warmup 5 --exchange nccl --skip-e2e
timeout 300 $TR8 --master-port 29524 tools/a2a_ceiling.py > gpurun_out/r2_a2a_ceiling_n8.jsonl 2> gpurun_out/r2_a2a_ceiling_n8.err
cat gpurun_out/r2_a2a_ceiling_n8.jsonl
run r2_bench_n4_ours $TR4 --master-port 29525 bench.py --gpus 4 --steps 20 --warmup 5
run r2_wide_n8_c256 $TR8 --master-port 29526 bench.py --gpus 8 --steps 20 --warmup 5 --cols 256 --rows-per-gpu 1250000 --skip-e2e
run r2_wide_n8_c1024 $TR8 --master-port 29527 bench.py --gpus 8 --steps 20 --warmup 5 --cols 1024 --rows-per-gpu 1250000 --skip-e2e
run r2_wide_n8_c4096 $TR8 --master-port 29528 bench.py --gpus 8 --steps 20 --warmup 5 --cols 4096 --rows-per-gpu 312500 --skip-e2e
run r2_wide_n8_c1024_bf16 $TR8 --master-port 29529 bench.py --gpus 8 --steps 20 --warmup 5 --cols 1024 --rows-per-gpu 1250000 --skip-e2e --feature-dtype bfloat16
run r2_resnet50_n8_ours $TR8 --master-port 29530 benchmarks/resnet50_images.py --gpus 8
for f in r2_bench_n8_ours r2_bench_n8_ours_k7 r2_bench_n8_nccl r2_bench_n4_ours r2_wide_n8_c256 r2_wide_n8_c1024 r2_wide_n8_c4096 r2_wide_n8_c1024_bf16 r2_resnet50_n8_ours; do
  python - "$f" <<'PY'
import json, sys
name = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/{name}.json").read().strip().splitlines()[-1])
    keep = {k: d.get(k) for k in ("value", "ms_per_epoch", "shuffle_kernel_ms_per_epoch", "first_chunk_ms",
                                  "nvlink_egress_gbps_per_gpu", "exactly_once", "seconds_per_epoch",
                                  "batch_wait_mean_ms", "invalid")}
    keep["e2e"] = (d.get("e2e") or {}).get("value") if isinstance(d.get("e2e"), dict) else None
    print(name, json.dumps(keep))
except Exception as e:
    print(name, "unreadable:", e)
    print(open(f"gpurun_out/{name}.err").read()[-1500:])
PY
done
