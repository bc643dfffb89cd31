import sys, os, numpy as np, torch
sys.path.insert(0,'.')
import hevcbitstream_b200._lib as L
if len(sys.argv) > 1 and sys.argv[1] != "default":
    L.LIB_PATH = os.path.join(os.path.dirname(L.LIB_PATH), sys.argv[1])
import hevcbitstream_b200 as hb
from oracle import ref
from tests import util
def run(ctx,name,d,size,want=True):
    cap=size//60+1000
    outs=ctx.scan_strip_device(d,size=size,cap_nals=cap,want_rbsp=want,sync=False)
    torch.cuda.synchronize()
    ev=[torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ts=[]
    for i in range(5):
        ev[0].record(); ctx.scan_strip_device(d,size=size,cap_nals=cap,want_rbsp=want,out=outs,sync=False); ev[1].record(); torch.cuda.synchronize(); ts.append(ev[0].elapsed_time(ev[1]))
    s=outs['summary'].cpu().numpy(); n=int(s[0]); rb=int(s[5]); t=min(ts)/1e3
    alg=size+(rb if want else 0)+24*n
    print(f"{name:22s} n={n} t={t*1e3:.3f}ms in={size/t/1e9:.1f}GB/s alg={alg/t/1e9:.1f}GB/s frac={alg/t/1e9/6544.3:.3f}", flush=True)
total=2<<30
ctx=hb.Context(0)
for k in (64,1024,16384,1<<20):
    u=util.c2_stream(k, 32<<20, seed=k); su=u.size-ref.PAD
    d=torch.from_numpy(u[:su].copy()).cuda().repeat(total//su)
    run(ctx,f"{sys.argv[1]}-nal{k}",d,d.numel())
u=util.c2_stream(4096, 32<<20, dense=True); su=u.size-ref.PAD
d=torch.from_numpy(u[:su].copy()).cuda().repeat(total//su)
run(ctx,f"{sys.argv[1]}-dense",d,d.numel())
