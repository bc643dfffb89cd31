"""Timing of hevcb_insert_device on synthetic payloads (development aid, not the bench)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hevcbitstream_b200 import Context

ctx = Context(0)
g = torch.Generator(device="cuda").manual_seed(1)
n_bytes = 1 << 30
x = torch.randint(0, 256, (n_bytes,), dtype=torch.uint8, device="cuda", generator=g)
CASES = [("64B", 64, 0), ("1KiB", 1024, 0), ("4KiB", 4096, 0), ("16KiB", 16384, 0), ("1MiB", 1 << 20, 0), ("16KiB-z50", 16384, 2)]
sel = os.environ.get("CASES")
for name, seg, zero_frac in [c for c in CASES if not sel or c[0] in sel.split(",")]:
    y = x
    if zero_frac:
        m = torch.randint(0, 4, (n_bytes,), dtype=torch.uint8, device="cuda", generator=g)
        y = torch.where(m < zero_frac, torch.zeros_like(x), x)
    n = n_bytes // seg
    off = torch.arange(n, dtype=torch.int64, device="cuda") * seg
    end = off + seg
    for _ in range(2):
        ins = ctx.insert_device(y, off, end, start_code_len=4)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ins = ctx.insert_device(y, off, end, start_code_len=4, sync=False)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"{name}: {ms:.3f} ms  {n_bytes / ms / 1e6:.1f} GB/s in  inserted={int(ins['summary'][2])}", flush=True)
    del ins
