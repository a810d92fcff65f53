#!/bin/bash
# CI entry point (reference: run_ci_tests.sh). CPU tier by default; `gpu` runs
# the device tier (needs a B200).
set -euxo pipefail
cd "$(dirname "$0")"
python -m ray_shuffling_data_loader_b200._build
if [ "${1:-cpu}" = "gpu" ]; then
  python -m pytest tests -v --durations=0 -x -m gpu
else
  python -m pytest tests -v --durations=0 -x -m "not gpu"
fi
