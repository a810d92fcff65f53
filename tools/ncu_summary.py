#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed) into profiles/<name>.md:
duration, DRAM bytes / %, key throughputs, registers, top stall sites.

    python tools/ncu_summary.py gpurun_out/prof_v3.ncu-rep profiles/scatter_f32_v3.md "title"
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum",
]


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep, out, title = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else sys.argv[1])
    raw = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units, vals = raw[0], raw[1], raw[2]
    lines = [f"# {title}", "", f"source: `{rep}` (ncu --set full --clock-control none)", "",
             "| metric | value | unit |", "|---|---|---|"]
    name_idx = hdr.index("Kernel Name") if "Kernel Name" in hdr else None
    if name_idx is not None:
        lines.insert(2, f"kernel: `{vals[name_idx]}`")
    for h, u, v in zip(hdr, units, vals):
        if h in KEYS:
            lines.append(f"| {h} | {v} | {u} |")
    src = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "source", "--csv"]))))
    if len(src) > 3:
        h = src[1]
        data = src[2:]
        i_src, i_s, i_ex = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
        stall_cols = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
        total = sum(int(r[i_s] or 0) for r in data)
        lines += ["", f"## Top stall sites (of {total} warp samples)", "",
                  "| samples | executed | SASS | dominant stalls |", "|---|---|---|---|"]
        for r in sorted(data, key=lambda r: -int(r[i_s] or 0))[:12]:
            st = sorted(((h[i], int(r[i] or 0)) for i in stall_cols if int(r[i] or 0) > 0),
                        key=lambda kv: -kv[1])[:2]
            lines.append(f"| {r[i_s]} | {r[i_ex]} | `{r[i_src][:60]}` | "
                         + ", ".join(f"{k}={v}" for k, v in st) + " |")
        ops = {}
        for r in data:
            op = r[i_src].split()[0] if r[i_src] and not r[i_src].startswith("@") else \
                (r[i_src].split()[1] if len(r[i_src].split()) > 1 else "")
            for tag in ("UTMALDG", "UBLKCP", "LDS.128", "STG.E.128", "SYNCS", "LDS.64", "STS.64"):
                if op.startswith(tag):
                    ops[tag] = ops.get(tag, 0) + int(r[i_ex] or 0)
        if ops:
            lines += ["", "## Instruction evidence (executed warp-instructions)", ""]
            lines += [f"* `{k}`: {v}" for k, v in sorted(ops.items())]
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:30]))


if __name__ == "__main__":
    main()
