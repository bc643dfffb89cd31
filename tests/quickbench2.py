import sys, os, numpy as np, torch
sys.path.insert(0,'.')
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
    print(f"{name:22s} rbsp={want} t={t*1e3:.3f}ms in={size/t/1e9:.1f}GB/s alg={alg/t/1e9:.1f}GB/s frac={alg/t/1e9/6544.3:.3f}", flush=True)
total=2<<30
units={k:util.c2_stream(k, 32<<20, seed=k) for k in (1024,1<<20)}
data={}
for k,u in units.items():
    su=u.size-ref.PAD
    data[k]=torch.from_numpy(u[:su].copy()).cuda().repeat(total//su)
for stag in sys.argv[1:]:
    os.environ['HEVCB_SCAN_STAGGER']=stag
    ctx=hb.Context(0)
    for k,d in data.items():
        run(ctx,f"stag{stag}-nal{k}",d,d.numel())
    ctx.close()
