"""Worker of tests/test_shard_cpu.py::test_world_size_2_gloo (launched with torch.distributed.run, backend gloo).

Every rank builds the same seeded stream, scans ITS shard with the host build of the kernel logic (there is no GPU in
this test), exchanges the shard records with hevcbitstream_b200.shard.gather_records -- the same code path the NCCL
run uses -- stitches, patches its arrays, and rank 0 compares the assembled result with the oracle."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from hevcbitstream_b200 import shard as hs  # noqa: E402
from oracle import ref  # noqa: E402
from tests import shard_check, util  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    lib = C.CDLL(os.path.join(ROOT, "tests", "_hostsim", "libhostsim.so"))
    run = shard_check.hostsim_shard_runner(lib)
    ok = True
    for case in range(6):
        if case < 3:
            s = ref.gen_stream(seed=10 + case, profile=1, n_slices=300, payload_min=1, payload_max=500, zero_heavy_pct=30, extra_zero_pct=10,
                               ps_period=30)
            size = s.size - ref.PAD
            buf = util.padded(s[:size])
        else:
            rng = np.random.default_rng(case)
            size = 5000 + case
            buf = util.adversarial(rng, size, case, density=0.2)
        bounds = hs.plan_shards(buf, world, size)
        own, halo, first, last = hs.shard_flags(bounds, rank)
        lo = int(bounds[rank])
        if own > 0:
            rec, ns, ne, ro, re, img = run(buf[lo: lo + own + halo], own, halo, first, last)
        else:
            from hevcbitstream_b200._lib import ShardSummary
            rec, ns, ne, ro, re, img = ShardSummary(), *[np.zeros(1, np.int64) for _ in range(4)], np.zeros(0, np.uint8)
        records = hs.gather_records(rec, torch.device("cpu"))
        res = hs.stitch(records)
        hs.apply_patches(res, rank, ns, ne, ro, re)
        f, n = int(res.first_local[rank]), int(res.n_owned[rank])
        e = re[f:f + n].copy()
        e[e >= 0] += res.rbsp_base[rank]
        mine = dict(ns=ns[f:f + n] + res.byte_base[rank], ne=ne[f:f + n] + res.byte_base[rank], ro=ro[f:f + n] + res.rbsp_base[rank], re=e,
                    img=img[: rec.rbsp_bytes])
        parts = [None] * world
        dist.all_gather_object(parts, mine)
        if rank == 0:
            g = shard_check.Global()
            g.nal_start = np.concatenate([p["ns"] for p in parts])
            g.nal_end = np.concatenate([p["ne"] for p in parts])
            g.rbsp_off = np.concatenate([p["ro"] for p in parts])
            g.rbsp_end = np.concatenate([p["re"] for p in parts])
            image = np.concatenate([p["img"] for p in parts])
            s_ = res.glob
            g.n_nals, g.n_terminated, g.last_rc, g.last_start, g.last_end = s_.n_nals, s_.n_terminated, s_.last_rc, s_.last_start, s_.last_end
            assert len(g.nal_start) == g.n_nals
            n_checked = util.compare_scan(buf, size, g, image, tag=f"gloo{case}")
            ok = ok and n_checked > 0
    dist.barrier()
    if rank == 0 and ok:
        print("SHARD_GLOO_OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
