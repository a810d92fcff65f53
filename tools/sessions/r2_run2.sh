cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_gpu_tests_2.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_gpu_tests_2.log
tail -5 gpurun_out/r2_gpu_tests_2.log
# K7 time-to-first-batch (keeps the generated table for the runs below)
timeout 600 python tools/ttfb_bench.py > gpurun_out/r2_ttfb_n1.jsonl 2> gpurun_out/r2_ttfb_n1.err
cat gpurun_out/r2_ttfb_n1.jsonl
# TMA bulk-store probe vs row size
for rb in 256 1024 4096; do
  timeout 120 python tools/kernel_bench.py --bulk-store-probe $rb --iters 5 >> gpurun_out/r2_bulk_store_probe.jsonl 2>> gpurun_out/r2_bulk_store_probe.err
done
cat gpurun_out/r2_bulk_store_probe.jsonl
# row_align / tmap / geometry A/B (profiles open questions)
timeout 900 bash tools/next_experiments.sh > gpurun_out/r2_next_n1.stdout 2>&1
# DATA_SPEC schema through bench.py, both arms
timeout 900 python bench.py --schema dataspec --steps 20 --warmup 5 --keep-data > gpurun_out/r2_bench_n1_dataspec_ours.json 2> gpurun_out/r2_bench_n1_dataspec_ours.err
echo "dataspec ours exit $?"
timeout 900 python bench.py --schema dataspec --row-align 128 --steps 20 --warmup 5 --keep-data --skip-e2e > gpurun_out/r2_bench_n1_dataspec_ours_align128.json 2> gpurun_out/r2_bench_n1_dataspec_ours_align128.err
echo "dataspec align128 exit $?"
timeout 1200 python bench.py --impl reference --schema dataspec --steps 20 --warmup 5 > gpurun_out/r2_bench_n1_dataspec_ref.json 2> gpurun_out/r2_bench_n1_dataspec_ref.err
echo "dataspec ref exit $?"
cat gpurun_out/r2_bench_n1_dataspec_ours.json gpurun_out/r2_bench_n1_dataspec_ours_align128.json gpurun_out/r2_bench_n1_dataspec_ref.json
# ncu: top kernel, full set, once
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scatter_tma -s 3 -c 1 -f -o gpurun_out/r2_scatter_f32 python tools/kernel_bench.py --rows 12500000 --cols 64 --mode 0 --iters 2 --warmup 3 > gpurun_out/r2_ncu_f32.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scatter_tma -s 3 -c 1 -f -o gpurun_out/r2_scatter_dataspec_m4 python tools/kernel_bench.py --rows 12500000 --cols 21 --mode 4 --iters 2 --warmup 3 > gpurun_out/r2_ncu_m4.log 2>&1
ls -la gpurun_out/*.ncu-rep
