#!/bin/bash
# One gpurun call's worth of A/B measurements that were still open at the end of
# round 1 (profiles/README.md, "Open questions"). Writes gpurun_out/next_*.log.
#   gpurun --timeout 900 -- 'bash tools/next_experiments.sh'            (1 GPU part)
#   gpurun --gpus 2 --timeout 900 -- 'bash tools/next_experiments.sh p2p' (NVLink part)
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
KB="timeout 120 python tools/kernel_bench.py"
if [ "${1:-}" = "p2p" ]; then
  L=gpurun_out/next_p2p.log; rm -f $L
  # narrow rows over NVLink: does padding to whole 128-byte lines pay?
  for al in 0 128; do
    $KB --rows 12500000 --cols 21 --mode 4 --peer 1 --trainers 1 --row-align $al --tag "dataspec->f32 align$al" >> $L 2>&1
    $KB --rows 12500000 --cols 21 --mode 3 --peer 1 --trainers 1 --row-align $al --tag "dataspec native align$al" >> $L 2>&1
    $KB --rows 25000000 --cols 16 --mode 0 --peer 1 --trainers 1 --row-align $al --tag "16xf32 align$al" >> $L 2>&1
  done
  cat $L; exit 0
fi
L=gpurun_out/next_n1.log; rm -f $L
# 1. line-aligned row padding for sub-line rows (local HBM)
for al in 0 128; do
  $KB --rows 12500000 --cols 21 --mode 4 --row-align $al --tag "align$al" >> $L 2>&1
  $KB --rows 12500000 --cols 21 --mode 3 --row-align $al --tag "align$al" >> $L 2>&1
  $KB --rows 50000000 --cols 16 --mode 0 --row-align $al --tag "align$al" >> $L 2>&1
  $KB --rows 16666666 --cols 48 --mode 0 --row-align $al --tag "align$al" >> $L 2>&1
done
# 2. tensor-map loads under the new loader/index split (was 1.22 ms with the old producer)
for tm in 1 2; do
  $KB --rows 12500000 --cols 64 --mode 0 --tmap-mode $tm --tag "tmap$tm" >> $L 2>&1
done
# 3. geometry variants of the cooperative schedule for the headline shape
python - <<'PY' >> $L 2>&1
from ray_shuffling_data_loader_b200 import _build
for name, defs in [("t256s3", ["RSDL_TILE_F32=256", "RSDL_STAGES_F32=3"]),
                   ("t128s6", ["RSDL_STAGES_F32=6"])]:
    print(_build.build_variant(name, defs))
PY
V=ray_shuffling_data_loader_b200/csrc/build/variants
for v in t256s3 t128s6; do
  for s in 0 1; do
    $KB --rows 12500000 --cols 64 --mode 0 --sched $s --tag "$v sched$s" --ext $V/$v/_C.cpython-312-x86_64-linux-gnu.so >> $L 2>&1
  done
done
# 4. where do the ~0.3-0.7 ms per epoch between back-to-back shuffles go? (chrome trace)
RSDL_TRACE=gpurun_out/next_trace_n1.json timeout 300 python bench.py --gpus 1 --steps 100 --warmup 50 --skip-e2e > gpurun_out/next_bench_trace.json 2>> $L
cat $L
