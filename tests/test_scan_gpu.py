"""GPU parity tests for hevcb_scan_strip_* (the CUDA path through the C ABI) against the reference-built
oracle: every NAL offset, every nal_to_rbsp status and every RBSP byte must be identical."""
import numpy as np
import pytest

from oracle import ref
from tests import util

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built")]


def host_run(ctx, buf, size):
    res = ctx.scan_strip_host(buf[:size] if size else buf[:0], size=size)
    return res, res.rbsp


def test_tiny_and_appendix_b(ctx):
    h = lambda s: np.frombuffer(bytes.fromhex(s.replace(" ", "")), np.uint8)
    for v in ["", "00", "00 00 01", "00 00 00 01", "00 00 01 40 01 02 03 04 00 00 01 42 01",
              "00 00 00 01 40 01 02 03 04 00 00 00 01 42 01", "00 00 01 00 00 01 40 01",
              "00 00 01 40 01 02 03 04 05 00 00 01", "01 02 03 04 05 06 00 00 01 0A", "01 02 03 04 05 00 00 01 09 0A",
              "00 00 01 40 00 00 03 01 05 00 00 01 07 07", "00 00 01 40 01 80 00 00 03"]:
        a = h(v)
        buf = util.padded(a)
        res, img = host_run(ctx, buf, a.size)
        util.compare_scan(buf, a.size, res, img, tag=v)


@pytest.mark.parametrize("alphabet", [0, 1, 2, 3])
def test_adversarial_small(ctx, alphabet):
    rng = np.random.default_rng(200 + alphabet)
    for it in range(400):
        size = int(rng.integers(0, 200))
        buf = util.adversarial(rng, size, alphabet)
        res, img = host_run(ctx, buf, size)
        util.compare_scan(buf, size, res, img, tag=f"a{alphabet}-{it}")


def test_adversarial_tile_edges(ctx):
    """sizes around multiples of the 16 KiB tile and the 512 B row, dense and sparse event mixes"""
    rng = np.random.default_rng(11)
    sizes = []
    for base in (512, 16384, 32768, 49152, 16384 * 5):
        sizes += [base - 9, base - 8, base - 3, base - 1, base, base + 1, base + 7, base + 8, base + 9, base + 17]
    for it, size in enumerate(sizes):
        for density in (1.0, 0.02):
            buf = util.adversarial(rng, size, it, density=density)
            res, img = host_run(ctx, buf, size)
            util.compare_scan(buf, size, res, img, tag=f"edge{size}-{density}")


def test_adversarial_multi_tile(ctx):
    rng = np.random.default_rng(12)
    for it in range(12):
        size = int(rng.integers(200_000, 3_000_000))
        buf = util.adversarial(rng, size, it, density=[1.0, 0.3, 0.01][it % 3])
        res, img = host_run(ctx, buf, size)
        util.compare_scan(buf, size, res, img, tag=f"multi{it}")


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_generated_rich_streams(ctx, seed):
    s = ref.gen_stream(seed=seed, profile=1, n_slices=20000, payload_min=1, payload_max=400, zero_heavy_pct=30,
                       extra_zero_pct=20, ps_period=40, unsupported_pct=5)
    size = s.size - ref.PAD
    for cut in (0, 1, 4, 7):
        buf = util.padded(s[: size - cut])
        res, img = host_run(ctx, buf, size - cut)
        n = util.compare_scan(buf, size - cut, res, img, tag=f"gen{seed}-{cut}")
        assert n > 20000


def test_config1_stream_device_api(ctx):
    """BASELINE config 1: 64 MB Main 1920x1080 stream, 10k slices, through the device-pointer entry point"""
    import torch

    s = ref.gen_stream(seed=0, profile=0, n_slices=10000, payload_min=6680, payload_max=6680, idr_period=100)
    size = s.size - ref.PAD
    d = torch.from_numpy(s[:size].copy()).cuda()
    res = ctx.scan_strip_device(d, size=size, cap_nals=20000)
    assert res.n_nals == 10003
    res.nal_start = res.nal_start.cpu().numpy()
    res.nal_end = res.nal_end.cpu().numpy()
    res.rbsp_off = res.rbsp_off.cpu().numpy()
    res.rbsp_end = res.rbsp_end.cpu().numpy()
    img = res.rbsp.cpu().numpy()
    util.compare_scan(s, size, res, img, tag="config1")
    # scan-only mode gives the same offsets
    res2 = ctx.scan_strip_device(d, size=size, cap_nals=20000, want_rbsp=False)
    assert res2.n_nals == res.n_nals
    assert np.array_equal(res2.nal_start.cpu().numpy()[:10003], res.nal_start[:10003])
    assert np.array_equal(res2.rbsp_end.cpu().numpy()[:10003], res.rbsp_end[:10003])


@pytest.mark.parametrize("nal_size,dense", [(64, False), (256, False), (1000, False), (4096, False), (16384, False), (65536, False), (262144, False),
                                            (1 << 20, False), (4096, True), (67, True), (300, True)])
def test_config2_shapes(ctx, nal_size, dense):
    """BASELINE config 2 shapes at 96 MiB: fixed-size NALs with escaped random payload, and the EPB-dense worst case"""
    import torch

    total = 96 << 20
    s = util.c2_stream(nal_size, total, seed=nal_size, dense=dense)
    size = s.size - ref.PAD
    d = torch.from_numpy(s[:size].copy()).cuda()
    res = ctx.scan_strip_device(d, size=size)
    n = res.n_nals
    res.nal_start = res.nal_start[:n].cpu().numpy()
    res.nal_end = res.nal_end[:n].cpu().numpy()
    res.rbsp_off = res.rbsp_off[:n].cpu().numpy()
    res.rbsp_end = res.rbsp_end[:n].cpu().numpy()
    img = res.rbsp.cpu().numpy()
    util.compare_scan(s, size, res, img, tag=f"c2-{nal_size}-{dense}")
    if dense:
        assert res.n_epb > size // 5


def test_capacity_overflow_is_reported(ctx):
    import hevcbitstream_b200 as hb

    s = util.c2_stream(64, 1 << 16, seed=3)
    size = s.size - ref.PAD
    with pytest.raises(hb.HevcbError):
        ctx.scan_strip_host(s[:size], size=size, cap_nals=10)


def test_pipelined_host_path_matches_oracle(monkeypatch):
    """hevcb_scan_strip_host on 'large' buffers cuts the stream into shards that are copied, scanned and copied back on three
    streams; forced here with a tiny shard size so that every boundary rule is exercised"""
    import hevcbitstream_b200 as hb

    monkeypatch.setenv("HEVCB_HOST_CHUNK", "4096")
    c = hb.Context(0)
    try:
        rng = np.random.default_rng(21)
        for it in range(30):
            size = int(rng.integers(9000, 200_000))
            buf = util.adversarial(rng, size, it, density=[1.0, 0.3, 0.01][it % 3])
            res = c.scan_strip_host(buf[:size], size=size)
            util.compare_scan(buf, size, res, res.rbsp, tag=f"pipe{it}")
        s = ref.gen_stream(seed=6, profile=1, n_slices=5000, payload_min=1, payload_max=600, zero_heavy_pct=30, extra_zero_pct=20, ps_period=40,
                           unsupported_pct=5)
        size = s.size - ref.PAD
        for cut in (0, 5):
            res = c.scan_strip_host(s[: size - cut], size=size - cut)
            n = util.compare_scan(util.padded(s[: size - cut]), size - cut, res, res.rbsp, tag=f"pipegen{cut}")
            assert n > 5000
    finally:
        c.close()


@pytest.mark.parametrize("workload", ["nal16k", "epb_dense_4k"])
def test_bench_inputs_4gib_vs_oracle(ctx, workload):
    """The very buffers bench.py times (BASELINE config[1]: 4 GiB, 16 KiB NALs = the headline, and the EPB-dense worst case):
    one hevcb_scan_strip_device call over the whole 4 GiB, compared with the unmodified reference fed in 512 MiB pieces cut
    at NAL boundaries (its API takes `int` sizes): every NAL offset, every nal_to_rbsp status / size and every RBSP byte."""
    import torch

    import bench

    name, nal_size, dense = next(w for w in bench.WORKLOADS if w[0] == workload)
    unit = bench.make_unit(nal_size, bench.UNIT_BYTES, 1234, dense)  # rank 0's unit, as bench.py builds it
    assert unit[0] == 0 and unit[1] == 0 and unit[2] == 1 and unit[-1] >= 2  # a unit starts with a start code and ends a NAL
    reps = (4 << 30) // unit.size
    size = unit.size * reps
    ut = torch.from_numpy(unit).cuda()
    d = torch.zeros(size + 32, dtype=torch.uint8, device="cuda")[: size + 16]
    d[:size].view(reps, -1).copy_(ut.unsqueeze(0).expand(reps, -1))
    cap = size // max(16, (nal_size if not dense else 4096) // 2) + (1 << 16)
    res = ctx.scan_strip_device(d, size=size, cap_nals=cap)
    n = res.n_nals
    # the oracle on one piece of `upp` units (every piece holds the same bytes: the buffer is the unit tiled)
    upp = max(1, (512 << 20) // unit.size)
    piece = np.tile(unit, upp)
    pbuf = util.padded(piece)
    st, en, r = ref.scan_all_with_tail(pbuf, piece.size)
    assert r["last_rc"] == -1  # the piece ends inside its last NAL, which the next piece's start code terminates
    sr = ref.strip_all(pbuf, st, en)
    npp = len(st)
    rc_ref = torch.from_numpy(sr["rc"].astype(np.int64)).cuda()
    st_t, en_t = torch.from_numpy(st).cuda(), torch.from_numpy(en).cuda()
    want = torch.from_numpy(sr["rbsp"]).cuda()
    lens = torch.clamp(rc_ref, min=0)
    dense_off = torch.cumsum(lens, 0) - lens
    n_pieces = -(-reps // upp)
    assert res.n_terminated == n - 1 and res.last_rc == -1 and res.last_end == size
    assert res.rbsp_bytes == size - res.n_epb
    k0 = 0
    for p in range(n_pieces):
        units_here = min(upp, reps - p * upp)
        if units_here != upp:  # the shorter last piece
            piece = np.tile(unit, units_here)
            pbuf = util.padded(piece)
            st, en, r = ref.scan_all_with_tail(pbuf, piece.size)
            sr = ref.strip_all(pbuf, st, en)
            npp = len(st)
            rc_ref = torch.from_numpy(sr["rc"].astype(np.int64)).cuda()
            st_t, en_t = torch.from_numpy(st).cuda(), torch.from_numpy(en).cuda()
            want = torch.from_numpy(sr["rbsp"]).cuda()
            lens = torch.clamp(rc_ref, min=0)
            dense_off = torch.cumsum(lens, 0) - lens
        base = p * upp * unit.size
        gs, ge = res.nal_start[k0: k0 + npp], res.nal_end[k0: k0 + npp]
        go, gr = res.rbsp_off[k0: k0 + npp], res.rbsp_end[k0: k0 + npp]
        assert torch.equal(gs - base, st_t), f"{workload}: nal_start differs in piece {p}"
        assert torch.equal(ge - base, en_t), f"{workload}: nal_end differs in piece {p}"
        assert torch.equal(gr == -1, rc_ref < 0), f"{workload}: nal_to_rbsp status differs in piece {p}"
        assert torch.equal(torch.where(gr >= 0, gr - go, torch.full_like(gr, -1)), rc_ref), f"{workload}: RBSP sizes differ in piece {p}"
        idx = torch.repeat_interleave(go - dense_off, lens) + torch.arange(int(lens.sum()), device="cuda")
        assert torch.equal(res.rbsp[idx], want[: idx.numel()]), f"{workload}: RBSP bytes differ in piece {p}"
        del idx
        k0 += npp
    assert k0 == n


def test_pipelined_host_path_empty_shards(monkeypatch):
    """A run of bytes < 2 that covers whole nominal chunks leaves shards without bytes (hevcb_plan_shards moves every cut forward to
    a byte >= 2): the pipeline must skip them without losing the prefetch of the next non-empty shard."""
    import hevcbitstream_b200 as hb

    monkeypatch.setenv("HEVCB_HOST_CHUNK", "4096")
    c = hb.Context(0)
    try:
        rng = np.random.default_rng(77)
        for it, (runlen, fill) in enumerate([(4096 * 3 + 17, 0), (4096 * 5, 0), (4096 * 2 + 1, 1), (4096 * 7 + 5, 0)]):
            parts = []
            for blk in range(4):
                s = util.c2_stream([64, 300, 1000, 4096][(it + blk) % 4], 20000, seed=it * 10 + blk)
                parts.append(s[: s.size - ref.PAD])
                if blk < 3:
                    run = np.full(runlen, fill, np.uint8)
                    if fill == 1:
                        run[::2] = 0  # 00 01 00 01 ...: bytes < 2 without a start code
                    parts.append(run)
            a = np.concatenate(parts)
            buf = util.padded(a)
            res = c.scan_strip_host(buf[: a.size], size=a.size)
            n = util.compare_scan(buf, a.size, res, res.rbsp, tag=f"emptyshard{it}")
            assert n > 20
    finally:
        c.close()


@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_sparse_patterns_in_random_payload(ctx, seed):
    """Large random payloads with a zero-pair pattern every few KiB: the analysers then take their single-flagged-row path (one
    or two chunks of a 2 KiB row group need the exact masks).  The patterns include the ones whose chunk holds a start code
    FOLLOWED by another event (00 00 01 xx 00 00 00), which the ordered carry reports as 'last event: not a start code'."""
    rng = np.random.default_rng(900 + seed)
    size = 24 << 20
    x = rng.integers(4, 256, size).astype(np.uint8)  # no bytes <= 3: no accidental patterns
    pats = [[0, 0, 1, 0x26, 1], [0, 0, 0, 1, 0x40, 1], [0, 0, 1, 0x02, 0, 0, 0], [0, 0, 1, 9, 0, 0, 1, 7], [0, 0, 3, 1], [0, 0, 3, 0, 0, 3], [0, 0, 2],
            [0, 0, 3, 200], [0, 0, 0, 0, 0, 0, 0, 1, 5], [0, 0, 1, 0x42, 0, 0, 3, 0, 0, 0, 1]]
    step = [1500, 5000, 20000, 700][seed - 1]
    pos = np.cumsum(rng.integers(step // 2, step * 2, size // step))
    pos = pos[pos < size - 64]
    last = 0
    for p in pos.tolist():
        pat = pats[int(rng.integers(0, len(pats)))]
        # also exercise every alignment inside a 16-byte chunk and across 512-byte rows / 2 KiB row groups / 32 KiB tiles
        if rng.random() < 0.3:
            p = (p & ~2047) + [2047, 2046, 2045, 511, 510, 32767 & 2047, 15, 14, 13, 0, 1][int(rng.integers(0, 11))] - int(rng.integers(0, 3))
        if p < last + 32 or p > size - 64:  # patterns must not touch (two start codes back to back are a zero-length NAL: the reference loop stops there)
            continue
        x[p: p + len(pat)] = pat
        last = p + len(pat)
    buf = util.padded(x)
    import torch

    d = torch.from_numpy(buf[:size].copy()).cuda()
    res = ctx.scan_strip_device(d, size=size)
    n = res.n_nals
    res.nal_start = res.nal_start[:n].cpu().numpy()
    res.nal_end = res.nal_end[:n].cpu().numpy()
    res.rbsp_off = res.rbsp_off[:n].cpu().numpy()
    res.rbsp_end = res.rbsp_end[:n].cpu().numpy()
    util.compare_scan(buf, size, res, res.rbsp.cpu().numpy(), tag=f"sparse{seed}")
    assert n > 300


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_removed_bytes_unevenly_spread_over_a_tile(ctx, seed):
    """Tiles whose 4 KiB row groups (one writer warp each) lose very different numbers of bytes: none, one to four (the kept runs go out
    as separate copies), many (the gaps are closed in place, behind the writers' barrier), with start codes and error patterns in
    between.  The writers take the per-group counts and carries from the analysers' aggregates: every mix must reproduce the oracle."""
    rng = np.random.default_rng(1700 + seed)
    size = 12 << 20
    x = rng.integers(4, 256, size).astype(np.uint8)
    group = 4096
    last = 0

    def put(p, pat):
        nonlocal last
        if p < last + 8 or p + len(pat) > size - 64:
            return
        x[p: p + len(pat)] = pat
        last = p + len(pat)

    for g in range(size // group):
        kind = int(rng.integers(0, 8))
        base = g * group
        if kind == 0:
            continue  # a clean group
        if kind in (1, 2):  # a few removed bytes
            for _ in range(int(rng.integers(1, 5))):
                put(base + int(rng.integers(0, group - 16)), [0, 0, 3, int(rng.integers(0, 4))])
        elif kind in (3, 4):  # many removed bytes: 00 00 03 01 back to back over part of the group
            lo = base + int(rng.integers(0, group // 2))
            n = int(rng.integers(8, 300))
            for i in range(n):
                put(lo + 4 * i + (8 if i == 0 else 0), [0, 0, 3, 1])
        elif kind == 5:  # start codes (short NALs) and a removed byte
            p = base + int(rng.integers(0, 64))
            while p < base + group - 80:
                put(p, [0, 0, 1, 0x26, 1])
                p += int(rng.integers(40, 400))
            put(base + group - 40, [0, 0, 3, 2])
        elif kind == 6:  # nal_to_rbsp error patterns inside a NAL, removed bytes around them
            put(base + 100, [0, 0, 1, 0x40, 1])
            put(base + 300, [0, 0, 3, 0, 0, 3, 1])
            put(base + 900, [0, 0, 2])
            put(base + 1500, [0, 0, 3, 200])
            put(base + 2500, [0, 0, 0, 1, 0x42, 1])
        else:  # patterns straddling the group's edges (rows of neighbouring writer warps)
            put(base + group - 2, [0, 0, 3, 1, 0, 0, 3, 1])
            put(base + 2046, [0, 0, 3, 0, 0, 1, 0x26])
    buf = util.padded(x)
    import torch

    d = torch.from_numpy(buf[:size].copy()).cuda()
    res = ctx.scan_strip_device(d, size=size)
    n = res.n_nals
    res.nal_start = res.nal_start[:n].cpu().numpy()
    res.nal_end = res.nal_end[:n].cpu().numpy()
    res.rbsp_off = res.rbsp_off[:n].cpu().numpy()
    res.rbsp_end = res.rbsp_end[:n].cpu().numpy()
    util.compare_scan(buf, size, res, res.rbsp.cpu().numpy(), tag=f"uneven{seed}")
    assert n > 1000
