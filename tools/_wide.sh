L=gpurun_out/p2p_wide_v20.log; rm -f $L
for c in 64 256 1024 4096; do
  timeout 120 python tools/kernel_bench.py --rows 1250000 --cols $c --mode 0 --peer 1 --trainers 1 --iters 5 --tag "all-to-peer" >> $L 2>&1
done
timeout 120 python tools/kernel_bench.py --rows 1250000 --cols 1024 --mode 1 --peer 1 --trainers 1 --iters 5 --tag "all-to-peer bf16" >> $L 2>&1
timeout 120 python tools/kernel_bench.py --rows 12500000 --cols 21 --mode 4 --peer 1 --trainers 1 --iters 5 --tag "all-to-peer dataspec->f32" >> $L 2>&1
timeout 120 python tools/kernel_bench.py --rows 12500000 --cols 21 --mode 3 --peer 1 --trainers 1 --iters 5 --tag "all-to-peer dataspec native" >> $L 2>&1
cat $L
