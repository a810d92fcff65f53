# stage-depth / tile-height variants for the 8-byte-source modes (DATA_SPEC) - kernel only
cd /root/repo
mkdir -p gpurun_out
L=gpurun_out/r2_kbench_stage_variants.jsonl; rm -f $L
V=ray_shuffling_data_loader_b200/csrc/build/variants
SFX=$(python -c "import sysconfig;print(sysconfig.get_config_var('EXT_SUFFIX'))")
KB="timeout 120 python tools/kernel_bench.py --rows 12500000 --iters 8"
for mode in 4 3; do
  for al in 0 128; do
    $KB --cols 21 --mode $mode --row-align $al --tag "shipped m$mode align$al" >> $L 2>> gpurun_out/r2_kbench_stage_variants.err
    for v in s6_64 t256s3_64; do
      $KB --cols 21 --mode $mode --row-align $al --ext $V/$v/_C$SFX --tag "$v m$mode align$al" >> $L 2>> gpurun_out/r2_kbench_stage_variants.err
    done
  done
done
$KB --cols 16 --mode 0 --rows 50000000 --tag "shipped 16xf32" >> $L 2>> gpurun_out/r2_kbench_stage_variants.err
$KB --cols 16 --mode 0 --rows 50000000 --ext $V/s8_f32/_C$SFX --tag "s6 panel32 16xf32" >> $L 2>> gpurun_out/r2_kbench_stage_variants.err
python - <<'PY'
import json
for line in open("gpurun_out/r2_kbench_stage_variants.jsonl"):
    d = json.loads(line)
    print(d["tag"], "pitch", d["row_pitch"], "ms %.3f" % d["ms_best"], "frac %.3f" % d.get("frac_of_measured_hbm_peak", 0))
PY
