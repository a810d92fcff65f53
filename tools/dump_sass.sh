#!/bin/bash
# Regenerate the committed SASS evidence (no GPU needed): one listing per kernel
# of csrc/build/shuffle_kernels.cu.o plus a table of the opcodes that prove the
# TMA / mbarrier / vector-store / conversion paths (B200_PROFILING.md section 4).
set -euo pipefail
cd "$(dirname "$0")/.."
OBJ=ray_shuffling_data_loader_b200/csrc/build/shuffle_kernels.cu.o
OUT=profiles/sass
mkdir -p "$OUT"
python -m ray_shuffling_data_loader_b200._build >/dev/null
rm -f "$OUT"/*.sass
# address + instruction only: the hex encodings double the size and prove nothing
cuobjdump -sass "$OBJ" | sed -E 's#[[:space:]]*/\* 0x[0-9a-f]+ \*/[[:space:]]*$##' | awk -v out="$OUT" '
  /Function : / { name=$3; gsub(/[^A-Za-z0-9_]/, "_", name); file=out "/" name ".sass" }
  file != "" && $0 !~ /^[[:space:]]*$/ { print > file }'
python - "$OUT" <<'PY'
import collections, glob, os, re, subprocess, sys
out = sys.argv[1]
keys = ["UBLKCP", "UTMALDG", "SYNCS.ARRIVE.TRANS64", "SYNCS.PHASECHK", "LDS.128", "LDS.64", "STS.64",
        "STG.E.128", "I2F.S64", "F2F.F32.F64", "F2FP.BF16", "F2FP.SATFINITE.E4M3", "STG.E.STRONG.SYS",
        "LDG.E.STRONG.SYS", "MEMBAR", "VOTE.ANY", "ERRBAR"]
rows = []
for f in sorted(glob.glob(os.path.join(out, "*.sass"))):
    mangled = os.path.basename(f)[:-5]
    try:
        name = subprocess.run(["c++filt", mangled], capture_output=True, text=True).stdout.strip()
    except Exception:
        name = mangled
    ops = collections.Counter()
    n = 0
    for line in open(f):
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            n += 1
            for k in keys:
                if m.group(1).startswith(k):
                    ops[k] += 1
    rows.append((name or mangled, n, ops))
with open(os.path.join(out, "OPCODES.md"), "w") as w:
    w.write("# SASS opcode evidence (sm_100a, `cuobjdump -sass`, regenerate with tools/dump_sass.sh)\n\n")
    w.write("| kernel | instructions | " + " | ".join(f"`{k}`" for k in keys) + " |\n")
    w.write("|---|---|" + "---|" * len(keys) + "\n")
    for name, n, ops in rows:
        short = re.sub(r"\(.*", "", name).replace("rsdl::", "")
        w.write(f"| `{short}` | {n} | " + " | ".join(str(ops.get(k, 0) or "") for k in keys) + " |\n")
print(open(os.path.join(out, "OPCODES.md")).read())
PY
