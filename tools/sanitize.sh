#!/bin/bash
# Race / memory checking of the kernels (the reference has none).
# Run on a GPU box: tools/sanitize.sh [memcheck|racecheck|synccheck]
set -euo pipefail
cd "$(dirname "$0")/.."
tool="${1:-memcheck}"
for mode in 0 3 4; do
  compute-sanitizer --tool "$tool" --error-exitcode 1 \
    python tools/kernel_bench.py --rows 200000 --cols 21 --mode "$mode" --iters 1 --warmup 1 --verify
done
compute-sanitizer --tool "$tool" --error-exitcode 1 \
  python tools/kernel_bench.py --rows 200000 --cols 64 --mode 0 --generic --iters 1 --warmup 1 --verify
