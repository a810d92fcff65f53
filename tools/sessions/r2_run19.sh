cd /root/repo
mkdir -p gpurun_out
timeout 120 python tools/tail_fields_bench.py > gpurun_out/r2_tail_fields_ab.jsonl 2> gpurun_out/r2_tail_fields_ab.err
cat gpurun_out/r2_tail_fields_ab.jsonl; tail -2 gpurun_out/r2_tail_fields_ab.err
