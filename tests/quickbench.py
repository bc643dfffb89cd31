import sys, time, numpy as np, torch
sys.path.insert(0,'.')
import hevcbitstream_b200 as hb
from oracle import ref
from tests import util
ctx=hb.Context(0)
def run(name, unit, total):
    size_u=unit.size-ref.PAD
    reps=max(1,total//size_u)
    d=torch.from_numpy(unit[:size_u].copy()).cuda().repeat(reps)
    size=d.numel()
    cap=size//60+1000
    out=None
    for want in (True, False):
        outs=ctx.scan_strip_device(d,size=size,cap_nals=cap,want_rbsp=want,sync=False)
        torch.cuda.synchronize()
        ev=[torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ts=[]
        for i in range(5):
            ev[0].record(); ctx.scan_strip_device(d,size=size,cap_nals=cap,want_rbsp=want,out=outs,sync=False); ev[1].record(); torch.cuda.synchronize(); ts.append(ev[0].elapsed_time(ev[1]))
        s=outs['summary'].cpu().numpy()
        n=int(s[0]); rb=int(s[5])
        t=min(ts)/1e3
        alg=size+(rb if want else 0)+24*n
        print(f"{name:14s} rbsp={want} size={size/2**30:.2f}GiB nals={n} t={t*1e3:.3f}ms in={size/t/1e9:.1f}GB/s alg={alg/t/1e9:.1f}GB/s frac={alg/t/1e9/6544.3:.3f}", flush=True)
total=int(sys.argv[1]) if len(sys.argv)>1 else (1<<30)
for nal in (64,1024,16384,1<<20):
    run(f"nal{nal}", util.c2_stream(nal, 32<<20, seed=nal), total)
run("dense4k", util.c2_stream(4096, 32<<20, dense=True), total)
