"""GPU parity tests of the byte-range sharding: every shard goes through hevcb_scan_strip_shard_device (the CUDA path),
the records are stitched by hevcb_stitch, and the assembled result must equal the oracle's whole-stream result.  All
shards run on cuda:0 one after the other; the multi-process exchange is covered by tests/test_shard_cpu.py (gloo) and
by bench.py --gpus N (NCCL)."""
import numpy as np
import pytest

from oracle import ref
from tests import shard_check, util

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built")]


@pytest.mark.parametrize("alphabet", [0, 1, 2, 3])
def test_adversarial_small(ctx, alphabet):
    run = shard_check.device_shard_runner(ctx)
    rng = np.random.default_rng(400 + alphabet)
    for it in range(60):
        size = int(rng.integers(0, 600))
        buf = util.adversarial(rng, size, alphabet, density=[1.0, 0.5, 0.1][it % 3])
        for g in (2, 3, 5):
            shard_check.check_sharded(buf, size, g, run, tag=f"a{alphabet}-{it}")


def test_adversarial_multi_tile(ctx):
    """shards of several 32 KiB tiles, cuts at arbitrary (unaligned) stream positions"""
    run = shard_check.device_shard_runner(ctx)
    rng = np.random.default_rng(13)
    for it in range(8):
        size = int(rng.integers(300_000, 2_000_000))
        buf = util.adversarial(rng, size, it, density=[1.0, 0.3, 0.01][it % 3])
        for g in (2, 8):
            shard_check.check_sharded(buf, size, g, run, tag=f"multi{it}")


@pytest.mark.parametrize("seed", [1, 2])
def test_generated_rich_streams(ctx, seed):
    run = shard_check.device_shard_runner(ctx)
    s = ref.gen_stream(seed=seed, profile=1, n_slices=20000, payload_min=1, payload_max=400, zero_heavy_pct=30, extra_zero_pct=20, ps_period=40,
                       unsupported_pct=5)
    size = s.size - ref.PAD
    for g, cut in ((2, 0), (4, 3), (8, 7)):
        buf = util.padded(s[: size - cut])
        n = shard_check.check_sharded(buf, size - cut, g, run, tag=f"gen{seed}-{cut}")
        assert n > 20000


def test_large_nals_span_shards(ctx):
    """1 MiB NALs over 8 shards of ~0.9 MiB: most shards hold no start code at all"""
    run = shard_check.device_shard_runner(ctx)
    s = util.c2_stream(1 << 20, 7 << 20, seed=3)
    size = s.size - ref.PAD
    g, image, res, bounds = shard_check.run_sharded(s, size, 8, run)
    util.compare_scan(s, size, g, image, tag="big")
    assert sum(1 for r in range(8) if res.n_owned[r] == 0) >= 1


def test_single_shard_equals_whole_stream_entry_point(ctx):
    run = shard_check.device_shard_runner(ctx)
    rng = np.random.default_rng(2)
    for it in range(20):
        size = int(rng.integers(0, 100_000))
        buf = util.adversarial(rng, size, it, density=0.2)
        g, image, res, bounds = shard_check.run_sharded(buf, size, 1, run)
        whole = ctx.scan_strip_host(buf[:size], size=size)
        assert (g.n_nals, g.n_terminated, g.last_rc, g.last_start, g.last_end, g.rbsp_bytes, g.n_epb) == (
            whole.n_nals, whole.n_terminated, whole.last_rc, whole.last_start, whole.last_end, whole.rbsp_bytes, whole.n_epb)
        n = whole.n_nals
        assert np.array_equal(g.nal_start[:n], whole.nal_start[:n]) and np.array_equal(g.nal_end[:n], whole.nal_end[:n])
        assert np.array_equal(g.rbsp_off[:n], whole.rbsp_off[:n]) and np.array_equal(g.rbsp_end[:n], whole.rbsp_end[:n])
        assert np.array_equal(image, whole.rbsp)


def test_apply_patches_device_matches_host_patching(ctx):
    import torch

    from hevcbitstream_b200 import shard as hs

    s = ref.gen_stream(seed=5, profile=1, n_slices=3000, payload_min=1, payload_max=2000, zero_heavy_pct=20, extra_zero_pct=10, ps_period=40)
    size = s.size - ref.PAD
    bounds = hs.plan_shards(s, 4, size)
    scans = []
    for r in range(4):
        own, halo, first, last = hs.shard_flags(bounds, r)
        lo = int(bounds[r])
        d = torch.zeros(own + halo + 32, dtype=torch.uint8, device="cuda")
        d[: own + halo] = torch.from_numpy(s[lo: lo + own + halo].copy())
        scans.append(hs.scan_strip_shard(ctx, d, own, halo, first, last))
    res = hs.stitch([sc.record for sc in scans])
    assert res.n_patches >= 3
    for r, sc in enumerate(scans):
        host = [t.cpu().numpy().copy() for t in (sc.nal_start, sc.nal_end, sc.rbsp_off, sc.rbsp_end)]
        hs.apply_patches(res, r, *host)
        hs.apply_patches_device(ctx, res, r, sc)
        n = int(res.first_local[r] + res.n_owned[r])
        for h, t in zip(host, (sc.nal_start, sc.nal_end, sc.rbsp_off, sc.rbsp_end)):
            assert np.array_equal(h[:n], t.cpu().numpy()[:n])


@pytest.mark.parametrize("n_shards", [2, 4, 7])
def test_device_join_equals_host_join(ctx, n_shards):
    """hevcb_stitch_apply_device (the join of the gathered records + this shard's patches as one kernel, what the NCCL step runs)
    against hevcb_stitch + host patching: the result struct byte for byte, the patched arrays entry for entry"""
    import ctypes as C

    import torch

    from hevcbitstream_b200 import shard as hs
    from hevcbitstream_b200._lib import StitchResult

    for seed, pmax in ((5, 2000), (6, 400000)):  # NALs inside shards; NALs that span several shards
        s = ref.gen_stream(seed=seed, profile=1, n_slices=3000 if pmax < 10000 else 40, payload_min=1, payload_max=pmax, zero_heavy_pct=20, extra_zero_pct=10,
                           ps_period=40)
        size = s.size - ref.PAD
        bounds = hs.plan_shards(s, n_shards, size)
        scans = []
        for r in range(n_shards):
            own, halo, first, last = hs.shard_flags(bounds, r)
            lo = int(bounds[r])
            d = torch.zeros(own + halo + 32, dtype=torch.uint8, device="cuda")
            d[: own + halo] = torch.from_numpy(s[lo: lo + own + halo].copy())
            scans.append(hs.scan_strip_shard(ctx, d, own, halo, first, last))
        res = hs.stitch([sc.record for sc in scans])
        allrec = torch.cat([sc.summary for sc in scans])  # what all_gather_into_tensor leaves on every rank
        stream = torch.cuda.current_stream().cuda_stream
        for r, sc in enumerate(scans):
            host = [t.cpu().numpy().copy() for t in (sc.nal_start, sc.nal_end, sc.rbsp_off, sc.rbsp_end)]
            hs.apply_patches(res, r, *host)
            d_res = torch.zeros(C.sizeof(StitchResult), dtype=torch.uint8, device="cuda")
            ctx._check(ctx._L.hevcb_stitch_apply_device(ctx._h, allrec.data_ptr(), n_shards, r, sc.nal_start.data_ptr(), sc.nal_end.data_ptr(),
                                                        sc.rbsp_off.data_ptr(), sc.rbsp_end.data_ptr(), sc.cap_nals, d_res.data_ptr(), stream))
            assert d_res.cpu().numpy().tobytes() == bytes(res), f"join differs on shard {r}"
            n = int(res.first_local[r] + res.n_owned[r])
            for h, t in zip(host, (sc.nal_start, sc.nal_end, sc.rbsp_off, sc.rbsp_end)):
                assert np.array_equal(h[:n], t.cpu().numpy()[:n])


def _pairs(p, k):
    a, b = int(p["pair_off"][k]), int(p["pair_off"][k + 1])
    return p["pair_field"][a:b], p["pair_value"][a:b]


@pytest.mark.parametrize("n_shards,big", [(2, False), (5, False), (40, True), (23, True)])
def test_sharded_parse_equals_whole_stream_parse(ctx, n_shards, big):
    """parameter-set hand-over: every shard parses the NALs it owns with the last SPS / PPS state of the earlier shards and the
    continuation of its last NAL; rc, NAL header, kind, header end and every syntax element must equal the unsharded parse
    (which the other tests pin against the reference)"""
    import torch

    from hevcbitstream_b200 import shard as hs

    if big:  # NALs larger than a shard: the continuation of a shard's last NAL then runs through whole shards
        s = ref.gen_stream(seed=9, profile=1, n_slices=30, payload_min=200000, payload_max=900000, zero_heavy_pct=20, extra_zero_pct=10, ps_period=9,
                           unsupported_pct=3)
    else:
        s = ref.gen_stream(seed=8, profile=1, n_slices=6000, payload_min=1, payload_max=900, zero_heavy_pct=20, extra_zero_pct=10, ps_period=300,
                           unsupported_pct=3)
    size = s.size - ref.PAD
    d = torch.zeros(size + 32, dtype=torch.uint8, device="cuda")
    d[:size] = torch.from_numpy(s[:size].copy())
    whole = ctx.scan_strip_device(d, size=size)
    pw = ctx.parse_device(d, whole)
    pw = {k: (v.cpu().numpy() if hasattr(v, "cpu") else v) for k, v in pw.items()}
    bounds = hs.plan_shards(s, n_shards, size)
    bufs, scans = [], []
    for r in range(n_shards):
        own, halo, first, last = hs.shard_flags(bounds, r)
        lo = int(bounds[r])
        b = torch.zeros(own + halo + 32, dtype=torch.uint8, device="cuda")
        b[: own + halo] = torch.from_numpy(s[lo: lo + own + halo].copy())
        bufs.append(b)
        scans.append(hs.scan_strip_shard(ctx, b, own, halo, first, last, extra_rbsp=hs.HEAD_BYTES))
    res = hs.stitch([sc.record for sc in scans])
    heads = []
    for r, sc in enumerate(scans):
        hs.apply_patches_device(ctx, res, r, sc)
        h = torch.zeros(hs.HEAD_BYTES, dtype=torch.uint8, device="cuda")
        nb = min(hs.HEAD_BYTES, int(sc.record.rbsp_bytes))
        h[:nb] = sc.rbsp[:nb]
        heads.append(h)
    states, missing, spanning = [], [], 0
    for r, sc in enumerate(scans):
        missing.append(hs.append_continuation(sc, res, r, heads)[1])
        spanning += int(res.cont_last_shard[r] > r + 1)
        states.append(hs.local_ps_contexts(ctx, bufs[r], sc, int(res.first_local[r]), int(res.n_owned[r])))
    g = 0
    crossing = 0
    for r, sc in enumerate(scans):
        own, halo, first, last = hs.shard_flags(bounds, r)
        sps_in, pps_in = hs.pick_incoming(states, r)
        ps = hs.parse_shard(ctx, bufs[r], own, halo, sc, res, r, sps_in, pps_in, missing=missing[r])
        n = int(res.n_owned[r])
        ps = {k: (v.cpu().numpy() if hasattr(v, "cpu") else v) for k, v in ps.items()}
        crossing += int(res.cont_last_shard[r] >= 0)
        assert np.array_equal(ps["rc"][:n], pw["rc"][g:g + n]), f"rc differs in shard {r}"
        assert np.array_equal(ps["nal_hdr"][:n], pw["nal_hdr"][g:g + n])
        assert np.array_equal(ps["kind"][:n], pw["kind"][g:g + n])
        sl = ps["kind"][:n] == 4
        assert np.array_equal(ps["hdr_end"][:n][sl], pw["hdr_end"][g:g + n][sl])
        cnt_s = np.diff(ps["pair_off"][: n + 1])
        cnt_w = np.diff(pw["pair_off"][g: g + n + 1])
        assert np.array_equal(cnt_s, cnt_w), f"element counts differ in shard {r}"
        a0, a1 = int(pw["pair_off"][g]), int(pw["pair_off"][g + n])
        assert np.array_equal(ps["pair_field"][: a1 - a0], pw["pair_field"][a0:a1])
        assert np.array_equal(ps["pair_value"][: a1 - a0], pw["pair_value"][a0:a1]), f"values differ in shard {r}"
        g += n
    assert g == whole.n_nals
    if big:
        assert spanning >= 1, "the stream was meant to hold NALs that span whole shards"
    else:
        assert crossing >= n_shards - 1
