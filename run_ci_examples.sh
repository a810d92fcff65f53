#!/bin/bash
# CI smoke of the runnable modules (reference: run_ci_examples.sh runs the
# dataset.py / torch_dataset.py __main__ blocks).
set -euxo pipefail
cd "$(dirname "$0")"
python -m ray_shuffling_data_loader_b200.dataset
python -m ray_shuffling_data_loader_b200.torch_dataset
python examples/ddp/torch_shuffle.py --num-rows 200000 --num-files 4 --num-columns 16 \
  --batch-size 20000 --epochs 2 --num-reducers 4 --data-dir "${TMPDIR:-/tmp}/rsdl_ci_example"
python examples/horovod/ray_torch_shuffle.py --num-workers 2 --no-cuda --num-rows 40000 \
  --num-files 4 --num-columns 8 --batch-size 5000 --epochs 2 --num-reducers 4 \
  --mock-train-step-time 0.001 --data-dir "${TMPDIR:-/tmp}/rsdl_ci_example_hvd"
