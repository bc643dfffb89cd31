"""Worker of tests/test_shard_multigpu.py (torch.distributed.run, backend nccl, one rank per GPU): byte-range sharded scan +
strip + header parse of one reference-written stream; every rank checks its share against the unsharded result computed on
its own GPU."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import hevcbitstream_b200 as hb  # noqa: E402
from hevcbitstream_b200 import shard as hs  # noqa: E402
from oracle import ref  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rank, world = dist.get_rank(), dist.get_world_size()
    ctx = hb.Context(local)
    s = ref.gen_stream(seed=12, profile=1, n_slices=20000, payload_min=1, payload_max=1500, zero_heavy_pct=20, extra_zero_pct=10, ps_period=500,
                       unsupported_pct=3)
    size = s.size - ref.PAD
    # unsharded reference result on this GPU
    d = torch.zeros(size + 32, dtype=torch.uint8, device=dev)
    d[:size] = torch.from_numpy(s[:size].copy()).to(dev)
    whole = ctx.scan_strip_device(d, size=size)
    pw = ctx.parse_device(d, whole)
    # sharded run
    bounds = hs.plan_shards(s, world, size)
    own, halo, first, last = hs.shard_flags(bounds, rank)
    lo = int(bounds[rank])
    b = torch.zeros(own + halo + 32, dtype=torch.uint8, device=dev)
    b[: own + halo] = torch.from_numpy(s[lo: lo + own + halo].copy()).to(dev)
    sc, res = hs.scan_strip_sharded(ctx, b, own, halo, first, last, extra_rbsp=hs.HEAD_BYTES)
    ps = hs.parse_sharded(ctx, b, own, halo, sc, res)
    f, n, g = int(res.first_local[rank]), int(res.n_owned[rank]), int(res.nal_base[rank])
    assert int(res.glob.n_nals) == whole.n_nals and int(res.glob.rbsp_bytes) == whole.rbsp_bytes
    bb, rb = int(res.byte_base[rank]), int(res.rbsp_base[rank])
    assert torch.equal(sc.nal_start[f:f + n] + bb, whole.nal_start[g:g + n])
    assert torch.equal(sc.nal_end[f:f + n] + bb, whole.nal_end[g:g + n])
    re_ = sc.rbsp_end[f:f + n]
    assert torch.equal(torch.where(re_ >= 0, re_ + rb, re_), whole.rbsp_end[g:g + n])
    assert torch.equal(sc.rbsp[: sc.record.rbsp_bytes], whole.rbsp[rb: rb + sc.record.rbsp_bytes])
    for key in ("rc", "nal_hdr", "kind"):
        assert torch.equal(ps[key][:n], pw[key][g:g + n]), key
    a0, a1 = int(pw["pair_off"][g]), int(pw["pair_off"][g + n])
    assert torch.equal(ps["pair_field"][: a1 - a0], pw["pair_field"][a0:a1])
    assert torch.equal(ps["pair_value"][: a1 - a0], pw["pair_value"][a0:a1])
    ok = torch.ones(1, device=dev)
    dist.all_reduce(ok)
    if rank == 0 and int(ok.item()) == world:
        print(f"SHARD_NCCL_OK world={world} nals={whole.n_nals} owned={[int(res.n_owned[r]) for r in range(world)]}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
