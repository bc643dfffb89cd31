"""Runs the scan+strip path a few times on one configuration (used under ncu; not a benchmark)."""
import sys
import torch
sys.path.insert(0, '.')
import hevcbitstream_b200 as hb
from oracle import ref
from tests import util

nal = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
total = int(sys.argv[2]) if len(sys.argv) > 2 else (512 << 20)
dense = len(sys.argv) > 3 and sys.argv[3] == "dense"
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 3
ctx = hb.Context(0)
unit = util.c2_stream(nal, 32 << 20, seed=nal, dense=dense)
su = unit.size - ref.PAD
d = torch.from_numpy(unit[:su].copy()).cuda().repeat(max(1, total // su))
size = d.numel()
cap = size // 60 + 1000
outs = None
for i in range(iters):
    outs = ctx.scan_strip_device(d, size=size, cap_nals=cap, want_rbsp=True, out=outs, sync=False)
torch.cuda.synchronize()
print("done", size, outs["summary"].cpu().numpy()[:7])
