cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/r2_gpu_tests_3.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_gpu_tests_3.log
tail -6 gpurun_out/r2_gpu_tests_3.log
timeout 300 python bench.py --steps 20 --warmup 5 --skip-e2e > gpurun_out/r2_bench_n1_ours_b.json 2> gpurun_out/r2_bench_n1_ours_b.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_n1_ours_b.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value", "ingest_seconds", "cold_rows_per_sec", "exactly_once")})
PY
timeout 300 python benchmarks/resnet50_images.py --gpus 1 --images-per-gpu 4096 --epochs 3 --keep-data > gpurun_out/r2_resnet50_n1_ours.json 2> gpurun_out/r2_resnet50_n1_ours.err
echo "resnet ours exit $?"; cat gpurun_out/r2_resnet50_n1_ours.json
timeout 600 python benchmarks/resnet50_images.py --gpus 1 --images-per-gpu 4096 --epochs 3 --impl reference > gpurun_out/r2_resnet50_n1_ref.json 2> gpurun_out/r2_resnet50_n1_ref.err
echo "resnet ref exit $?"; cat gpurun_out/r2_resnet50_n1_ref.json; tail -3 gpurun_out/r2_resnet50_n1_ref.err
for tool in memcheck synccheck racecheck; do
  timeout 500 bash tools/sanitize.sh $tool gpurun_out/r2_sanitize_$tool.log
  echo "$tool exit $?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|overall exit" gpurun_out/r2_sanitize_$tool.log | sort | uniq -c
done
