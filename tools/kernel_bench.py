#!/usr/bin/env python
"""Micro-benchmark of the scatter kernels on synthetic device-resident columns
(no Parquet, no queue): CUDA-event timing, GB/s and fraction of the measured HBM
copy peak. Used under gpurun and as the ncu target.

    python tools/kernel_bench.py --rows 12500000 --cols 64 --mode 0 --iters 10
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

from ray_shuffling_data_loader_b200.ops import perm

_C = None


def load_ext(path):
    global _C
    if path:
        import importlib.util
        spec = importlib.util.spec_from_file_location("_C", path)
        _C = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(_C)
    else:
        from ray_shuffling_data_loader_b200 import _C as mod
        _C = mod


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=12_500_000)
    ap.add_argument("--cols", type=int, default=64)
    ap.add_argument("--mode", type=int, default=0,
                    help="0 f32, 1 f32->bf16, 2 f32->fp8 block-scaled, 3 int64/f64 bit copy, "
                         "4 int64/f64 -> f32")
    ap.add_argument("--generic", action="store_true",
                    help="run the same layout through the generic kernel instead of the TMA kernel")
    ap.add_argument("--trainers", type=int, default=1)
    ap.add_argument("--grid", type=int, default=0)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--verify", action="store_true")
    ap.add_argument("--no-tmap", action="store_true", help="use 1-D bulk copies")
    ap.add_argument("--tmap-mode", type=int, default=0, help="1: swizzled 32-row boxes, 2: dense box")
    ap.add_argument("--peer", type=int, default=-1,
                    help="place the destination on this other GPU (single-process P2P)")
    ap.add_argument("--peer-frac", type=float, default=1.0,
                    help="with --peer and --trainers 2: trainer 0 local, trainer 1 on the peer")
    ap.add_argument("--identity", action="store_true", help="identity permutation (sequential dst)")
    ap.add_argument("--ext", default=None, help="path to an alternative _C build")
    ap.add_argument("--row-align", type=int, default=0,
                    help="round the destination row pitch up to this power of two (0: layout default)")
    ap.add_argument("--sched", type=int, default=-1,
                    help="producer schedule: -1 auto, 0 loader warps, 1 cooperative")
    ap.add_argument("--tag", default="")
    ap.add_argument("--bulk-store-probe", type=int, default=0, metavar="ROW_BYTES",
                    help="instead of the scatter: rate of TMA bulk stores (smem -> global, "
                         "UBLKCP.G.S) of ROW_BYTES rows to random slots (local, or --peer)")
    a = ap.parse_args()
    load_ext(a.ext)
    torch.cuda.set_device(0)
    _C.set_device(0)
    sm = _C.sm_count(0)
    if a.bulk_store_probe:
        rb = a.bulk_store_probe
        total = 3_200_000_000 // rb
        where = f"cuda:{a.peer}" if a.peer >= 0 else "cuda:0"
        if a.peer >= 0:
            _C.enable_peer_access(a.peer)
        dst = torch.zeros(total * rb, dtype=torch.uint8, device=where)
        torch.cuda.synchronize()
        stream = torch.cuda.current_stream().cuda_stream
        times = []
        for i in range(a.warmup + a.iters):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _C.bulk_store_probe(dst.data_ptr(), total, rb, total, a.grid or sm * 2, stream)
            e1.record()
            torch.cuda.synchronize()
            if i >= a.warmup:
                times.append(e0.elapsed_time(e1))
        best = min(times)
        print(json.dumps({"tag": a.tag or "bulk_store_probe", "row_bytes": rb, "peer": a.peer,
                          "rows": total, "ms_best": best,
                          "store_gbps_best": total * rb / best / 1e6,
                          "grid": a.grid or sm * 2}))
        return
    n, F = a.rows, a.cols
    rows_pad = (n + 255) // 256 * 256
    ssz = 8 if a.mode >= 3 else 4
    if ssz == 8:
        # DATA_SPEC-like: int64 columns and a float64 last column
        src = torch.randint(0, 1 << 40, (F, rows_pad), dtype=torch.int64, device="cuda")
        src[F - 1] = torch.rand(rows_pad, dtype=torch.float64, device="cuda").view(torch.int64)
    else:
        src = torch.rand((F, rows_pad), dtype=torch.float32, device="cuda")
    dsz = {0: 4, 1: 2, 2: 1, 3: 8, 4: 4}[a.mode]
    payload = F * dsz
    scale_off = (payload + 15) // 16 * 16
    pitch = ((scale_off + (F + 31) // 32 if a.mode == 2 else payload) + 31) // 32 * 32
    if a.row_align:
        pitch = (pitch + a.row_align - 1) // a.row_align * a.row_align
    T = a.trainers
    per = -(-n // T)
    dst = torch.zeros((T, per, pitch), dtype=torch.uint8, device="cuda")
    ptrs = torch.tensor([src[c].data_ptr() for c in range(F)], dtype=torch.int64, device="cuda")
    dst_ptrs = [dst[t].data_ptr() for t in range(T)]
    if a.peer >= 0:
        # single-process P2P: trainers >= 1 (or all, with one trainer) live on the peer GPU
        _C.enable_peer_access(a.peer)
        dst_peer = torch.zeros((T, per, pitch), dtype=torch.uint8, device=f"cuda:{a.peer}")
        torch.cuda.synchronize(a.peer)
        for t in range(T):
            if T == 1 or t >= 1:
                dst_ptrs[t] = dst_peer[t].data_ptr()
    stream = torch.cuda.current_stream().cuda_stream
    fields = None
    src_codes = [7] * F if ssz == 4 else [4] * (F - 1) + [8]
    dst_codes = {0: [7] * F, 1: [6] * F, 2: [9] * F, 3: src_codes, 4: [7] * F}[a.mode]
    kinds = None
    if a.mode == 4:
        kinds = torch.tensor([0] * (F - 1) + [1] + [2] * ((-F) % 4), dtype=torch.uint8,
                             device="cuda")
    if a.generic:
        assert a.mode != 2, "the generic kernel has no block-scaled fp8 epilogue"
        dt = np.dtype([("src", "<u8"), ("src_code", "<u4"), ("dst_code", "<u4"),
                       ("dst_off", "<u4"), ("width", "<u4")])
        arr = np.zeros(F, dtype=dt)
        for c in range(F):
            arr[c] = (src[c].data_ptr(), src_codes[c], dst_codes[c], dsz * c, 1)
        fields = torch.from_numpy(arr.view(np.uint8).copy()).cuda()

    def launch(epoch):
        key = list(perm.make_key(n, 1234, epoch).as_words())
        if a.identity:
            key = [1, 1, 1, 0, 0, 0, 0, 0, 0]
        if a.generic:
            _C.scatter_generic(key=key, num_rows=n, num_trainers=T, fields=fields.data_ptr(),
                               num_fields=F, n_local=n, global_offset=0, row_pitch=pitch,
                               write_lo=0, write_hi=pitch, dst=dst_ptrs,
                               grid=a.grid, stream=stream)
        else:
            tiles = -(-n // _C.fast_tile_rows(a.mode))
            _C.scatter_fast(key=key, num_rows=n, num_trainers=T, cols=ptrs.data_ptr(),
                            num_cols=F, n_local=n, global_offset=0, row_pitch=pitch,
                            scale_offset=scale_off, dst=dst_ptrs, mode=a.mode,
                            grid=min(a.grid or sm * _C.fast_ctas_per_sm(a.mode), tiles),
                            stream=stream,
                            col_base=0 if (a.no_tmap or not a.tmap_mode) else src.data_ptr(),
                            col_stride=0 if (a.no_tmap or not a.tmap_mode) else rows_pad * 4,
                            rows_alloc=0 if (a.no_tmap or not a.tmap_mode) else rows_pad,
                            tmap_mode=a.tmap_mode,
                            kinds=kinds.data_ptr() if kinds is not None else 0,
                            write_end=0 if a.mode == 2 else pitch, sched=a.sched)
    for i in range(a.warmup):
        launch(i)
    torch.cuda.synchronize()
    times = []
    for i in range(a.iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        launch(100 + i)
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    if a.verify and a.mode in (0, 3, 4) and a.peer < 0 and not a.identity:
        launch(7)
        torch.cuda.synchronize()
        pos = perm.permute(np.arange(min(n, 100000), dtype=np.uint64), perm.make_key(n, 1234, 7))
        q, rem = divmod(n, T)
        assert rem == 0 or T == 1
        rows_got = dst.view(-1, pitch)[torch.from_numpy(pos.astype(np.int64)).cuda()]
        want = src[:, :len(pos)].t().contiguous()
        if a.mode == 0:
            got = rows_got.view(torch.float32)[:, :F]
        elif a.mode == 3:
            got = rows_got.view(torch.int64)[:, :F]
        else:
            got = rows_got.view(torch.float32)[:, :F]
            w = want[:, :F - 1].to(torch.float32)
            want = torch.cat([w, want[:, F - 1:].view(torch.float64).to(torch.float32)], dim=1)
        assert torch.equal(got, want), "scatter mismatch"
    bytes_moved = n * (F * ssz + pitch)
    best, med = min(times), sorted(times)[len(times) // 2]
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(
            os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
    except Exception:
        pass
    out = {"tag": a.tag, "peer": a.peer, "identity": a.identity, "tmap": (0 if a.no_tmap else a.tmap_mode), "tile_rows": 0 if a.generic else _C.fast_tile_rows(a.mode), "generic": a.generic, "sched": a.sched,
           "rows": n, "cols": F, "mode": a.mode, "trainers": T, "row_pitch": pitch,
           "ms_best": best, "ms_median": med, "gbps_best": bytes_moved / best / 1e6,
           "gbps_median": bytes_moved / med / 1e6, "bytes": bytes_moved, "grid": a.grid or sm}
    if peaks.get("hbm_gbs"):
        out["frac_of_measured_hbm_peak"] = out["gbps_best"] / peaks["hbm_gbs"]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
