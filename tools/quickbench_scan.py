"""Timing of hevcb_scan_strip_device for one workload under different HEVCB_SCAN_ANALYSERS / HEVCB_SCAN_DEBUG settings."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from hevcbitstream_b200 import Context

nal = int(os.environ.get("NAL", "16384"))
dense = bool(int(os.environ.get("DENSE", "0")))
gib = float(os.environ.get("GIB", "2"))
unit = bench.make_unit(nal, 64 << 20, 1234, dense)
reps = max(1, int(gib * (1 << 30)) // unit.size)
d = torch.from_numpy(unit).cuda().repeat(reps)
size = d.numel()
d = torch.cat([d, torch.zeros(32, dtype=torch.uint8, device="cuda")])
cap = size // max(16, nal // 2) + (1 << 16)
for setting in sys.argv[1:]:
    for kv in setting.split(","):
        if kv:
            k, v = kv.split("=")
            os.environ[k] = v
    ctx = Context(0)
    outs = ctx.scan_strip_device(d, size=size, cap_nals=cap, sync=False)
    for _ in range(3):
        ctx.scan_strip_device(d, size=size, cap_nals=cap, out=outs, sync=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ctx.scan_strip_device(d, size=size, cap_nals=cap, out=outs, sync=False)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    s = outs["summary"].cpu().numpy()
    print(f"{setting:50s} {ms:8.3f} ms  {size / ms / 1e6:8.1f} GB/s in   nals={int(s[0])} rbsp={int(s[5])}", flush=True)
    ctx.close()
    for kv in setting.split(","):
        if kv:
            os.environ.pop(kv.split("=")[0], None)
