import sys, ctypes as C, numpy as np, torch, os
sys.path.insert(0,'.')
import hevcbitstream_b200._lib as L
L.LIB_PATH = os.path.join(os.path.dirname(L.LIB_PATH), "libhevcb200_timing%s.so" % os.environ.get("TV","A"))
import hevcbitstream_b200 as hb
from oracle import ref
from tests import util
lib = L.load_library()
ctx = hb.Context(0)
nal = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
unit = util.c2_stream(nal, 32 << 20, seed=nal)
su = unit.size - ref.PAD
d = torch.from_numpy(unit[:su].copy()).cuda().repeat((2 << 30) // su)
size = d.numel(); cap = size // 60 + 1000
outs = ctx.scan_strip_device(d, size=size, cap_nals=cap, sync=False); torch.cuda.synchronize()
buf = (C.c_ulonglong * 64)()
lib.hevcb_debug_scan_timing(buf, 1)
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(); ctx.scan_strip_device(d, size=size, cap_nals=cap, out=outs, sync=False); e1.record(); torch.cuda.synchronize()
lib.hevcb_debug_scan_timing(buf, 0)
a = np.array(list(buf), dtype=np.float64).reshape(4, 16)
ntiles = (size + 32767) // 32768
per = ntiles / 296
print("kernel ms", e0.elapsed_time(e1), "tiles/CTA", per)
names = ["w:mbar wait", "w:fixup", "w:phase1", "w:S1 wait", "w:aggregate", "w:phase2", "w:copy-out", "w:end sync+issue", "lb:-", "lb:fixup", "lb:lookback", "lb:S1 wait", "lb:agg+publish", "lb:phase2 idle", "lb:copy idle", "lb:end sync"]
for c in range(3):
    print("CTA", c, "launches", a[c][15])
    for i, nme in enumerate(names):
        if i != 15: print(f"   {nme:18s} {a[c][i]/per:10.0f} cycles/tile")
print("LB rounds total", a[3][15], "sum pl", a[3][14], "load cycles per round", a[3][13]/max(1,a[3][15])); print("spins per j (3 CTAs total):", a[3][:10], " by lane0..3,rest:", a[3][10:15], "tiles", per*3)
