"""Timing of the rewrite pipeline (scan+strip, parse, rewrite) on a config-1 style stream tiled on the device."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hevcbitstream_b200 import Context
from oracle import ref

ctx = Context(0)
s = ref.gen_stream(seed=0, profile=0, n_slices=10000, payload_min=6680, payload_max=6680, idr_period=100)
size1 = s.size - ref.PAD
reps = int(os.environ.get("REPS", "16"))
d = torch.from_numpy(s[:size1].copy()).cuda().repeat(reps)
size = d.numel()
d = torch.cat([d, torch.zeros(32, dtype=torch.uint8, device="cuda")])
def ev():
    return torch.cuda.Event(enable_timing=True)
for it in range(6):
    e = [ev() for _ in range(4)]
    e[0].record()
    scan = ctx.scan_strip_device(d, size=size, cap_nals=size // 1000 + 1024)
    e[1].record()
    parsed = ctx.parse_device(d, scan, cap_pairs=64 * scan.n_nals)
    e[2].record()
    out = ctx.rewrite_device(d, scan, parsed, [(4, "slice_qp_delta", 0, 2)], size=size)
    e[3].record()
    torch.cuda.synchronize()
    t = [e[i].elapsed_time(e[i + 1]) for i in range(3)]
    print(f"size {size/1e9:.2f} GB nals {scan.n_nals}: scan {t[0]:.2f} ms parse {t[1]:.2f} ms rewrite {t[2]:.2f} ms -> rewrite {size/t[2]/1e6:.1f} GB/s, pipeline {size/sum(t)/1e6:.1f} GB/s; rewritten {out['n_rewritten']} out {out['out_bytes']}", flush=True)
