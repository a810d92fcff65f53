cd /root/repo
mkdir -p gpurun_out
rm -f gpurun_out/r2_sink_ab.jsonl
for schema in f32 dataspec; do
  for cfg in 8,256 4,256 2,256 8,128 16,64; do
    RSDL_SINK_GRID=$cfg timeout 200 python bench.py --schema $schema --steps 20 --warmup 5 --keep-data --skip-e2e --min-timed-epochs 10 > /tmp/ab.json 2> /tmp/ab.err
    python - "$schema" "$cfg" <<'PY' >> gpurun_out/r2_sink_ab.jsonl
import json, sys
try:
    d = json.loads(open("/tmp/ab.json").read().strip().splitlines()[-1])
    print(json.dumps({"schema": sys.argv[1], "sink_grid": sys.argv[2], "value": d["value"], "ms_per_epoch": d["ms_per_epoch"],
                      "shuffle_kernel_ms": d["shuffle_kernel_ms_per_epoch"], "row_bytes": d["engine"]["row_bytes"]}))
except Exception as e:
    print(json.dumps({"schema": sys.argv[1], "sink_grid": sys.argv[2], "error": str(e), "err": open("/tmp/ab.err").read()[-300:]}))
PY
  done
done
cat gpurun_out/r2_sink_ab.jsonl
