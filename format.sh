#!/bin/bash
# Lint (reference: format.sh = yapf + flake8). Uses whatever is installed;
# byte-compiles everything as the minimum bar.
set -euo pipefail
cd "$(dirname "$0")"
python -m compileall -q ray_shuffling_data_loader_b200 benchmarks examples baseline tests bench.py __graft_entry__.py
if python -c "import flake8" 2>/dev/null; then
  python -m flake8 --max-line-length 100 ray_shuffling_data_loader_b200 benchmarks examples bench.py
fi
if command -v clang-format >/dev/null 2>&1; then
  clang-format --dry-run ray_shuffling_data_loader_b200/csrc/*.cu ray_shuffling_data_loader_b200/csrc/*.cpp \
    ray_shuffling_data_loader_b200/csrc/*.cuh ray_shuffling_data_loader_b200/csrc/*.h || true
fi
echo "format.sh: ok"
