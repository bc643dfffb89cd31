"""Helpers shared by the sharding tests: run every shard with a given per-shard scanner, stitch, assemble the global
view and compare it with the oracle's whole-stream result."""
from __future__ import annotations

import ctypes as C

import numpy as np

from hevcbitstream_b200 import shard as hs
from hevcbitstream_b200._lib import ShardSummary
from tests import util


class Global:
    pass


def hostsim_shard_runner(lib):
    lib.hostsim_scan_strip_shard.restype = C.c_int64

    def run(piece: np.ndarray, own, halo, is_first, is_last):
        cap = own // 3 + 8
        a = [np.full(cap, -7, np.int64) for _ in range(4)]
        img = np.zeros(own + 16, np.uint8)
        rec = ShardSummary()
        padded = np.zeros(own + halo + 32, np.uint8)
        padded[: own + halo] = piece[: own + halo]
        p = lambda x: x.ctypes.data_as(C.c_void_p)
        lib.hostsim_scan_strip_shard(p(padded), C.c_int64(own), C.c_int64(halo), int(is_first), int(is_last), p(a[0]), p(a[1]), p(a[2]), p(a[3]),
                                     C.c_int64(cap), p(img), C.byref(rec))
        return rec, a[0], a[1], a[2], a[3], img
    return run


def device_shard_runner(ctx):
    import torch

    def run(piece: np.ndarray, own, halo, is_first, is_last):
        d = torch.zeros(own + halo + 32, dtype=torch.uint8, device="cuda")
        d[: own + halo] = torch.from_numpy(piece[: own + halo].copy())
        sc = hs.scan_strip_shard(ctx, d, own, halo, is_first, is_last)
        return sc.record, sc.nal_start.cpu().numpy(), sc.nal_end.cpu().numpy(), sc.rbsp_off.cpu().numpy(), sc.rbsp_end.cpu().numpy(), sc.rbsp.cpu().numpy()
    return run


def run_sharded(buf: np.ndarray, size: int, n_shards: int, runner, bounds=None):
    """Returns (Global result in whole-stream coordinates, concatenated image, stitch result, bounds)."""
    if bounds is None:
        bounds = hs.plan_shards(buf, n_shards, size)
    assert bounds[0] == 0 and bounds[-1] == size and np.all(np.diff(bounds) >= 0)
    for b in bounds[1:-1]:
        assert b == size or b == 0 or buf[b - 1] >= 2, "cut after a byte < 2"
    outs, recs = [], []
    for r in range(n_shards):
        own, halo, first, last = hs.shard_flags(bounds, r)
        if own == 0:
            recs.append(ShardSummary())
            outs.append(None)
            continue
        lo = int(bounds[r])
        rec, ns, ne, ro, re, img = runner(buf[lo: lo + own + halo], own, halo, first, last)
        recs.append(rec)
        outs.append([ns, ne, ro, re, img])
    res = hs.stitch(recs)
    g = Global()
    NS, NE, RO, RE, IM = [], [], [], [], []
    for r in range(n_shards):
        if outs[r] is None:
            continue
        ns, ne, ro, re, img = outs[r]
        hs.apply_patches(res, r, ns, ne, ro, re)
        f, n = int(res.first_local[r]), int(res.n_owned[r])
        assert int(res.nal_base[r]) == sum(len(x) for x in NS)
        NS.append(ns[f:f + n] + res.byte_base[r])
        NE.append(ne[f:f + n] + res.byte_base[r])
        RO.append(ro[f:f + n] + res.rbsp_base[r])
        e = re[f:f + n].copy()
        e[e >= 0] += res.rbsp_base[r]
        RE.append(e)
        IM.append(img[: recs[r].rbsp_bytes])
    cat = lambda xs, dt: np.concatenate(xs) if xs else np.zeros(0, dt)
    g.nal_start, g.nal_end, g.rbsp_off, g.rbsp_end = cat(NS, np.int64), cat(NE, np.int64), cat(RO, np.int64), cat(RE, np.int64)
    image = cat(IM, np.uint8)
    s = res.glob
    g.n_nals, g.n_terminated, g.last_rc, g.last_start, g.last_end = s.n_nals, s.n_terminated, s.last_rc, s.last_start, s.last_end
    g.rbsp_bytes, g.n_epb = s.rbsp_bytes, s.n_epb
    assert len(g.nal_start) == g.n_nals, f"owned NALs {len(g.nal_start)} != global n_nals {g.n_nals}"
    assert image.size == g.rbsp_bytes
    return g, image, res, bounds


def check_sharded(buf, size, n_shards, runner, tag="", bounds=None):
    g, image, res, bounds = run_sharded(buf, size, n_shards, runner, bounds)
    return util.compare_scan(buf, size, g, image, tag=f"{tag}/G{n_shards}")
