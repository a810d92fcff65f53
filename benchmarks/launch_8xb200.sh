#!/bin/bash
# Launch helper: the analogue of the reference's benchmarks/cluster.yaml (a 4-node
# AWS Ray cluster). Here the "cluster" is one host: one process per B200, NCCL
# over NVLink 5 / NVSwitch, rendezvous on localhost.
#   benchmarks/launch_8xb200.sh [NGPUS] <script> [args...]
set -euo pipefail
ngpus="${1:-8}"; shift || true
export CUDA_DEVICE_MAX_CONNECTIONS="${CUDA_DEVICE_MAX_CONNECTIONS:-32}"
export NCCL_DEBUG="${NCCL_DEBUG:-WARN}"
exec python -m torch.distributed.run --nnodes=1 --nproc-per-node "$ngpus" \
  --master-addr 127.0.0.1 --master-port "${MASTER_PORT:-29544}" "$@"
