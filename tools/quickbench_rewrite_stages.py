"""Timing of the stages of the rewrite pipeline (scan + strip, parse, rewrite) on bench.py's config-3 stream (development aid)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from hevcbitstream_b200 import Context

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
dev = torch.device("cuda:0")
ctx = Context(0)
gib = float(os.environ.get("GIB", "2"))
pay = int(os.environ.get("PAYLOAD", "16384"))
unit_h = np.fromfile(os.path.join(ROOT, "tests", "golden", "headers_unit.bin"), dtype=np.uint8)
dh = torch.zeros(unit_h.size + 32, dtype=torch.uint8, device=dev)
dh[: unit_h.size] = torch.from_numpy(unit_h).to(dev)
sc = ctx.scan_strip_device(dh, size=unit_h.size)
pr = ctx.parse_device(dh, sc)
nh = int(sc.n_nals)
ro, re_ = sc.rbsp_off[:nh], sc.rbsp_end[:nh]
is_slice = (pr["kind"][:nh] == 4) & (pr["rc"][:nh] >= 0) & (re_ >= 0)
keep = torch.where(is_slice, pr["hdr_end"][:nh].to(torch.int64), torch.clamp(re_ - ro, min=0))
seg = keep + torch.where(is_slice, torch.full_like(keep, pay + 1), torch.zeros_like(keep))
seg_end = torch.cumsum(seg, 0)
seg_off = seg_end - seg
total = int(seg_end[-1])
g = torch.Generator(device=dev).manual_seed(4242)
nr = torch.randint(0, 256, (total + 32,), dtype=torch.uint8, device=dev, generator=g)
ar = torch.arange(int(keep.sum()), device=dev) - torch.repeat_interleave(torch.cumsum(keep, 0) - keep, keep)
nr[torch.repeat_interleave(seg_off, keep) + ar] = sc.rbsp[torch.repeat_interleave(ro, keep) + ar]
nr[(seg_end - 1)[is_slice]] = 0x80
live = seg > 0
unit_ins = ctx.insert_device(nr, seg_off[live].contiguous(), seg_end[live].contiguous(), start_code_len=4)
ub = int(unit_ins["out_bytes"])
reps = max(1, int(gib * (1 << 30)) // ub)
dr = torch.zeros(ub * reps + 32, dtype=torch.uint8, device=dev)
dr[: ub * reps].view(reps, ub).copy_(unit_ins["out"][:ub].unsqueeze(0).expand(reps, -1))
rsize = ub * reps
cap = nh * reps + 1024
edits = [(4, "slice_qp_delta", 0, 2), (2, "vui.video_full_range_flag", 2, 1)]


def timed(f, n=4):
    f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        r = f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, r


if os.environ.get("TIMING"):  # event stamps of the scan pipeline on this stream (tools/scan_timing.py)
    os.environ["HEVCB_SCAN_TIMING"] = "1"
    import scan_timing

    o2 = ctx.scan_strip_device(dr, size=rsize, cap_nals=cap, sync=False)
    for _ in range(3):
        ctx.scan_strip_device(dr, size=rsize, cap_nals=cap, out=o2, sync=False)
    torch.cuda.synchronize()
    scan_timing.report(ctx)
    sys.exit(0)
outs = ctx.scan_strip_device(dr, size=rsize, cap_nals=cap, sync=False)
ms_scan, _ = timed(lambda: ctx.scan_strip_device(dr, size=rsize, cap_nals=cap, out=outs, sync=False))
scan = ctx.scan_strip_device(dr, size=rsize, cap_nals=cap)
ms_parse, parsed = timed(lambda: ctx.parse_device(dr, scan, cap_pairs=80 * scan.n_nals))
ms_rw, out = timed(lambda: (ctx.parse_device(dr, scan, cap_pairs=80 * scan.n_nals), ctx.rewrite_device(dr, scan, parsed, edits, size=rsize))[1])
print(f"bytes {rsize} nals {scan.n_nals} epb {scan.n_epb}: scan {ms_scan:.3f} ms ({rsize / ms_scan / 1e6:.0f} GB/s), parse {ms_parse:.3f} ms, parse+rewrite {ms_rw:.3f} ms (rewrite alone ~{ms_rw - ms_parse:.3f})")
