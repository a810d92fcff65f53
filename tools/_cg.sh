V=ray_shuffling_data_loader_b200/csrc/build/variants
L=gpurun_out/kbench_v20.log; rm -f $L
for v in main cg64_2; do
  ext=""; [ $v != main ] && ext="--ext $V/$v/_C.cpython-312-x86_64-linux-gnu.so"
  for m in 3 4; do for c in 21 40 64; do
    timeout 100 python tools/kernel_bench.py --rows 12500000 --cols $c --mode $m --verify --tag "$v" $ext >> $L 2>&1
  done; done
done
cat $L
