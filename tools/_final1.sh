mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests.log 2>&1; tail -2 gpurun_out/gpu_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --gpus 1 > gpurun_out/bench_n1_v19.json 2> gpurun_out/bench_n1_v19.err; cut -c1-400 gpurun_out/bench_n1_v19.json
timeout 300 python bench.py --gpus 1 --max-concurrent-epochs 1 --skip-e2e --keep-data > gpurun_out/bench_n1_v19_window1.json 2>> gpurun_out/bench_n1_v19.err; cut -c1-200 gpurun_out/bench_n1_v19_window1.json
timeout 300 python bench.py --gpus 1 --impl reference --steps 40 --warmup 10 > gpurun_out/bench_ref_n1_v19.json 2> gpurun_out/bench_ref_n1_v19.err; cut -c1-300 gpurun_out/bench_ref_n1_v19.json
timeout 200 ncu --set full --clock-control none --import-source on -k regex:scatter_tma -s 3 -c 1 -o gpurun_out/prof_v19_f32_64cols -f python tools/kernel_bench.py --rows 12500000 --cols 64 --mode 0 --iters 2 > gpurun_out/ncu_a.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:scatter_tma -s 3 -c 1 -o gpurun_out/prof_v19_mode4_21cols -f python tools/kernel_bench.py --rows 12500000 --cols 21 --mode 4 --iters 2 > gpurun_out/ncu_b.log 2>&1
tail -1 gpurun_out/ncu_a.log gpurun_out/ncu_b.log
timeout 200 python benchmarks/benchmark.py --num-rows 20000000 --num-files 8 --num-row-groups-per-file 2 --num-trainers 1 --num-reducers 8 --num-epochs 10 --max-concurrent-epochs 2 --batch-size 250000 --num-trials 1 --data-dir /tmp/rsdl_bm --stats-dir gpurun_out/bm_stats --backend cuda > gpurun_out/benchmark_cli_dataspec.log 2>&1; tail -8 gpurun_out/benchmark_cli_dataspec.log
