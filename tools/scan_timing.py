"""Measurement aid: event stamps of the scan pipeline (HEVCB_SCAN_TIMING): per CTA and tile: 0 load issued, 1 tile seen by the
analysers, 6 analyser warp 0 done, 2 aggregate published, 3 prefix arrived, 4 writers start, 5 writers done.  Prints latencies.
Needs a library built with the stamps compiled in: make -C hevcbitstream_b200/csrc clean all EXTRA=-DHEVCB_SCAN_TIMING_BUILD."""
import ctypes as C
import os
import sys

os.environ["HEVCB_SCAN_TIMING"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from hevcbitstream_b200 import Context

def report(ctx):
    """prints the latencies of the last scan launches of `ctx` (they must have run with HEVCB_SCAN_TIMING set)"""
    L = ctx._L
    L.hevcb_scan_timing_dump.restype = C.c_int64
    L.hevcb_scan_timing_dump.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
    n = L.hevcb_scan_timing_dump(ctx._h, None, 0)
    if n == 0:
        print("no event stamps: build the library with make -C hevcbitstream_b200/csrc clean all EXTRA=-DHEVCB_SCAN_TIMING_BUILD")
        return
    buf = np.zeros(n, np.uint64)
    L.hevcb_scan_timing_dump(ctx._h, buf.ctypes.data_as(C.c_void_p), n)
    T = buf.reshape(-1, 256, 8).astype(np.int64)
    ctas = T.shape[0] - 1  # the last CTA is the scanner
    S = T[ctas]            # its rows: every 8th batch: 0 poll start, 1 batch complete, 2 running prefix taken over, 3 prefixes stored
    T = T[:ctas]
    t0 = T[:, 0, 0].min()
    S = np.where(S > 0, S - t0, -1)
    T = np.where(T > 0, T - t0, -1)
    its = slice(20, 200)
    def stat(name, a):
        a = a[a > -10**9]
        print(f"{name:44s} mean {a.mean():9.0f} ns   p10 {np.percentile(a,10):9.0f}  p50 {np.percentile(a,50):9.0f}  p90 {np.percentile(a,90):9.0f}  max {a.max():9.0f}")
    E = lambda e: T[:, its, e]
    stat("per-tile period (load issue i+1 - i)", T[:, 21:201, 0] - T[:, 20:200, 0])
    stat("load: issue -> seen by analysers", E(1) - E(0))
    stat("analysis: seen -> warp 0 done", E(6) - E(1))
    stat("   of which: seen -> zero-pair filter done", E(7) - E(1))
    stat("analysis: warp 0 done -> aggregate published", E(2) - E(6))
    stat("scan: aggregate published -> prefix arrived", E(3) - E(2))
    stat("prefix arrived -> writers start", E(4) - E(3))
    stat("write: start -> warp 0 done", E(5) - E(4))
    stat("writers done -> next load into the stage", T[:, 26:206, 0] - T[:, 20:200, 5])
    stat("whole cycle of a stage (issue i+6 - issue i)", T[:, 26:206, 0] - T[:, 20:200, 0])
    # skew between CTAs at the same iteration
    print("aggregate publication skew over CTAs at iteration 100: ", T[:, 100, 2].max() - T[:, 100, 2].min(), "ns;  prefix arrival skew:", T[:, 100, 3].max() - T[:, 100, 3].min())
    print("iteration 100: min/max publication", T[:, 100, 2].min(), T[:, 100, 2].max(), " min/max prefix arrival", T[:, 100, 3].min(), T[:, 100, 3].max())

    ok = (S[:, 0] > 0) & (S[:, 3] > 0)
    Sv = S[ok][10:]
    print(f"scanner, {ok.sum()} sampled batches (every 8th):")
    stat("  poll start -> batch complete", Sv[:, 1] - Sv[:, 0])
    stat("  batch complete -> running prefix taken over", Sv[:, 2] - Sv[:, 1])
    stat("  taken over -> prefixes stored", Sv[:, 3] - Sv[:, 2])
    stat("  time between sampled batches / 8", np.diff(Sv[:, 3]) / 8.0)


if __name__ == "__main__":
    nal = int(os.environ.get("NAL", "16384"))
    gib = float(os.environ.get("GIB", "2"))
    unit = bench.make_unit(nal, 64 << 20, 1234, False)
    reps = max(1, int(gib * (1 << 30)) // unit.size)
    d = torch.from_numpy(unit).cuda().repeat(reps)
    size = d.numel()
    d = torch.cat([d, torch.zeros(32, dtype=torch.uint8, device="cuda")])
    cap = size // max(16, nal // 2) + (1 << 16)
    ctx = Context(0)
    outs = ctx.scan_strip_device(d, size=size, cap_nals=cap, sync=False)
    for _ in range(3):
        ctx.scan_strip_device(d, size=size, cap_nals=cap, out=outs, sync=False)
    torch.cuda.synchronize()
    report(ctx)
