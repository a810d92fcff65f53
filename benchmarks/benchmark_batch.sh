#!/bin/bash
# Benchmark sweep (component C16). Same grid as the reference's
# benchmarks/benchmark_batch.sh:9-24 - files {100,50,25} x trainers x
# reducers/trainer {4,3,2}, 10 epochs, window 2, batch 250000, 2 trials - but the
# "cluster" is one 8xB200 host driven by torchrun instead of a 4-node Ray
# cluster, so trainers sweep {8,4,2} and the row count defaults to 1e8
# (override with NUM_ROWS=400000000 for the reference's size).
set -exo pipefail

data_dir="${DATA_DIR:-/tmp/benchmark_scratch}"
stats_dir="${STATS_DIR:-./results}"

num_rows="${NUM_ROWS:-100000000}"
num_columns="${NUM_COLUMNS:-64}"
num_row_groups_per_file=5
batch_size=250000
num_trials=2
num_epochs=10

max_concurrent_epochs_list=(2)
num_files_list=(100 50 25)
num_trainers_list=(8 4 2)
num_reducers_per_trainer_list=(4 3 2)

for max_concurrent_epochs in "${max_concurrent_epochs_list[@]}"; do
  for num_files in "${num_files_list[@]}"; do
    for num_trainers in "${num_trainers_list[@]}"; do
      for num_reducers_per_trainer in "${num_reducers_per_trainer_list[@]}"; do
        num_reducers=$(( num_reducers_per_trainer * num_trainers ))
        python -m torch.distributed.run --nnodes=1 --nproc-per-node "$num_trainers" \
          --master-addr 127.0.0.1 --master-port 29533 benchmarks/benchmark.py \
          --num-rows "$num_rows" --num-columns "$num_columns" \
          --num-files "$num_files" \
          --num-row-groups-per-file $num_row_groups_per_file \
          --batch-size $batch_size \
          --num-trials $num_trials \
          --cluster \
          --num-reducers "$num_reducers" \
          --num-trainers "$num_trainers" \
          --num-epochs $num_epochs \
          --max-concurrent-epochs "$max_concurrent_epochs" \
          --data-dir "$data_dir" \
          --stats-dir "$stats_dir" \
          --unique-stats --quiet
      done
    done
  done
done
