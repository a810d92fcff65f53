# the reference's own benchmark CLI on its own schema, GPU backend, all three residency modes
cd /root/repo
mkdir -p gpurun_out
L=gpurun_out/r2_benchmark_cli_dataspec_gpu.log; rm -f $L
COMMON="--num-rows 20000000 --num-files 20 --num-row-groups-per-file 5 --num-epochs 8 --max-concurrent-epochs 2 --batch-size 250000 --num-trainers 1 --num-reducers 4 --num-trials 2 --data-dir /tmp/rsdl_bm --stats-dir gpurun_out/r2_bm_stats --quiet --seed 1 --utilization-sample-period 0.2"
echo "== resident=hbm" >> $L
timeout 300 python benchmarks/benchmark.py $COMMON >> $L 2>&1; echo "hbm exit $?"
echo "== resident=host" >> $L
timeout 300 python benchmarks/benchmark.py $COMMON --use-old-data --resident host >> $L 2>&1; echo "host exit $?"
echo "== resident=disk" >> $L
timeout 400 python benchmarks/benchmark.py $COMMON --use-old-data --resident disk >> $L 2>&1; echo "disk exit $?"
echo "== chunk_passes=4 (hbm)" >> $L
timeout 300 python benchmarks/benchmark.py $COMMON --use-old-data --chunk-passes 4 >> $L 2>&1; echo "k7 exit $?"
grep -E "^==|Mean throughput|Mean over|Trial .* done" $L
ls gpurun_out/r2_bm_stats | head
