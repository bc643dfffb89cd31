"""GPU parity tests for hevcb_insert_* (rbsp_to_nal, h264_nal.c:92-132): every output byte and every NAL offset must
equal what the reference's rbsp_to_nal produces segment by segment; strip(insert(x)) == x as the size-independent
property."""
import numpy as np
import pytest

from oracle import port, ref
from tests import util

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built")]


@pytest.fixture(autouse=True, params=["fused", "twopass"])
def insert_path(request, monkeypatch):
    """every case runs through the single-pass assembly and through the two-pass kernels (HEVCB_INSERT_PATH, csrc/hevcb_insert.cu)"""
    monkeypatch.setenv("HEVCB_INSERT_PATH", request.param)
    return request.param



def zero_heavy(rng, n, alphabet):
    if alphabet == 0:
        vals, p = [0, 1, 2, 3, 4, 0x80], [0.55, 0.1, 0.08, 0.1, 0.07, 0.1]
    elif alphabet == 1:
        vals, p = [0, 3], [0.8, 0.2]
    elif alphabet == 2:
        vals, p = [0, 1, 0xFF], [0.34, 0.33, 0.33]
    else:
        return rng.integers(0, 256, n, dtype=np.uint8)
    return rng.choice(np.array(vals, dtype=np.uint8), size=n, p=p)


def check(ctx, rbsp, off, end, sc_len, tag, use_ref=True):
    oracle = ref.insert_all(rbsp, off, end, sc_len) if use_ref else port.insert_all(rbsp, off, end, sc_len)
    out, out_off, n_ins = ctx.insert_host(rbsp, off, end, start_code_len=sc_len)
    exp = oracle["out"]
    assert out.size == exp.size, f"{tag}: size {out.size} != {exp.size}"
    if not np.array_equal(out, exp):
        i = int(np.nonzero(out != exp)[0][0])
        raise AssertionError(f"{tag}: first diff at {i}: {out[max(0, i - 8):i + 8]} vs {exp[max(0, i - 8):i + 8]}")
    n = len(off)
    # oracle nal_off[k] = first byte after the start code
    assert np.array_equal(out_off[:n] + sc_len, oracle["nal_off"][:n]), tag
    assert out_off[n] == exp.size
    total = int((np.asarray(end) - np.asarray(off)).sum())
    assert n_ins == exp.size - total - sc_len * n, tag
    return out, out_off


@pytest.mark.parametrize("alphabet", [0, 1, 2, 3])
@pytest.mark.parametrize("sc_len", [0, 3, 4])
def test_random_segments(ctx, alphabet, sc_len):
    rng = np.random.default_rng(31 + alphabet * 3 + sc_len)
    for it in range(60):
        n_bytes = int(rng.integers(0, 6000))
        rbsp = zero_heavy(rng, n_bytes, alphabet)
        # random cut points -> adjacent or gapped segments, arbitrary alignment, including empty ones
        k = int(rng.integers(0, 40))
        cuts = np.sort(rng.integers(0, n_bytes + 1, 2 * k)).astype(np.int64)
        off, end = cuts[0::2].copy(), cuts[1::2].copy()
        check(ctx, rbsp, off, end, sc_len, f"a{alphabet}-sc{sc_len}-{it}", use_ref=(it % 2 == 0))


def test_long_zero_runs_and_row_edges(ctx):
    """zero runs that span lanes, rows and whole NALs; segment starts at every residue mod 16"""
    rng = np.random.default_rng(5)
    for run in (1, 2, 3, 15, 16, 17, 31, 32, 33, 511, 512, 513, 1023, 1025, 5000):
        for lead in range(0, 18):
            body = np.concatenate([rng.integers(1, 256, lead, dtype=np.uint8), np.zeros(run, np.uint8), np.array([rng.integers(0, 6)], np.uint8),
                                   rng.integers(0, 4, 40, dtype=np.uint8)])
            rbsp = np.concatenate([np.zeros(7, np.uint8), body, np.zeros(9, np.uint8)])
            off = np.array([7, 0, 7 + lead], np.int64)
            end = np.array([7 + body.size, 7, rbsp.size], np.int64)
            check(ctx, rbsp, off, end, 4, f"run{run}-lead{lead}")


def test_scan_strip_insert_round_trip_device(ctx):
    """strip then insert through the device entry points re-creates every NAL of a reference-written stream"""
    import torch

    s = ref.gen_stream(seed=4, profile=1, n_slices=20000, payload_min=1, payload_max=3000, zero_heavy_pct=40, extra_zero_pct=0, ps_period=50)
    size = s.size - ref.PAD
    d = torch.from_numpy(s[:size].copy()).cuda()
    res = ctx.scan_strip_device(d, size=size)
    n = res.n_nals
    ok = (res.rbsp_end[:n] >= 0)
    ins = ctx.insert_device(res.rbsp, res.rbsp_off[:n].contiguous(), res.rbsp_end[:n].contiguous(), n_nals=n, start_code_len=0)
    out = ins["out"].cpu().numpy()
    oo = ins["out_off"].cpu().numpy()
    ns, ne = res.nal_start[:n].cpu().numpy(), res.nal_end[:n].cpu().numpy()
    okh = ok.cpu().numpy()
    assert okh.sum() > 20000
    # the generator writes its NALs with rbsp_to_nal and no trailing zero bytes -> insert(strip(nal)) == nal, except
    # that nal_to_rbsp drops a trailing 00 00 03 (h264_nal.c:183-189) which rbsp_to_nal does not put back
    bad = 0
    for k in range(n):
        if not okh[k]:
            assert oo[k + 1] == oo[k]
            continue
        a = out[oo[k]:oo[k + 1]]
        b = s[ns[k]:ne[k]]
        if a.size == b.size - 1 and b.size >= 3 and tuple(b[-3:]) == (0, 0, 3):
            b = b[:-1]
        if not np.array_equal(a, b):
            bad += 1
    assert bad == 0
    # and against the oracle on the same segments
    o = ref.insert_all(res.rbsp.cpu().numpy()[: res.rbsp_bytes], res.rbsp_off[:n].cpu().numpy()[okh], res.rbsp_end[:n].cpu().numpy()[okh], 0)
    assert np.array_equal(o["out"], out[: oo[n]])


def test_large_buffer_round_trip(ctx):
    """256 MiB of 0/1/2/3-heavy payload in 64 KiB segments: strip(insert(x)) == x, checked on the device"""
    import torch

    g = torch.Generator(device="cuda").manual_seed(9)
    n_bytes = 256 << 20
    x = torch.randint(0, 256, (n_bytes,), dtype=torch.uint8, device="cuda", generator=g)
    m = torch.randint(0, 4, (n_bytes,), dtype=torch.uint8, device="cuda", generator=g)
    x = torch.where(m < 2, torch.zeros_like(x), x)  # half of the bytes are zero
    seg = 65536 - 3
    n = n_bytes // seg
    x[: n * seg].view(n, seg)[:, -1] = 0x80  # rbsp_trailing_bits: a NAL never ends in a zero byte
    off = torch.arange(n, dtype=torch.int64, device="cuda") * seg
    end = off + seg
    ins = ctx.insert_device(x, off, end, start_code_len=4)
    total = ins["out_bytes"]
    assert ins["n_inserted"] > n_bytes // 64
    res = ctx.scan_strip_device(ins["out"], size=total, cap_nals=n + 8)
    assert res.n_nals == n
    assert res.n_epb == ins["n_inserted"]
    ro, re = res.rbsp_off[:n], res.rbsp_end[:n]
    assert bool((re - ro == seg).all())
    # the stripped image = [start code + segment] * n
    img = res.rbsp[: res.rbsp_bytes].view(n, seg + 4)
    assert bool((img[:, 4:] == x[: n * seg].view(n, seg)).all())


def test_capacity_overflow_is_reported(ctx):
    from hevcbitstream_b200 import HevcbError

    rbsp = np.zeros(1000, np.uint8)
    with pytest.raises(HevcbError) as e:
        ctx.insert_host(rbsp, np.array([0], np.int64), np.array([1000], np.int64), start_code_len=3, out_cap=1100)
    assert e.value.code == -104
    out, out_off, n_ins = ctx.insert_host(rbsp, np.array([0], np.int64), np.array([1000], np.int64), start_code_len=3, out_cap=1502)
    assert out.size == 1502 and n_ins == 499


def test_empty_batch(ctx):
    out, out_off, n_ins = ctx.insert_host(np.zeros(0, np.uint8), np.zeros(0, np.int64), np.zeros(0, np.int64), start_code_len=4)
    assert out.size == 0 and out_off.tolist() == [0] and n_ins == 0


def test_full_size_strip_insert_round_trip(ctx):
    """BASELINE config-1 size (4 GiB, 16 KiB NALs, built like bench.py builds it): insert(strip(x)) re-creates x byte for byte,
    n_epb == n_inserted, every NAL found -- the size-independent properties at the full benchmark size"""
    import torch

    import bench

    unit = bench.make_unit(16384, 64 << 20, 99, False)
    reps = (4 << 30) // unit.size
    d = torch.from_numpy(unit).cuda().repeat(reps)
    size = d.numel()
    d = torch.cat([d, torch.zeros(32, dtype=torch.uint8, device="cuda")])
    res = ctx.scan_strip_device(d, size=size, cap_nals=size // 8192 + 65536)
    n = res.n_nals
    assert n % reps == 0 and n // reps >= 4000  # every unit contributes the same ~4096 NALs
    assert res.rbsp_bytes == size - res.n_epb and res.last_rc == -1
    assert bool((res.rbsp_end[:n] >= 0).all())
    ins = ctx.insert_device(res.rbsp, res.rbsp_off[:n].contiguous(), res.rbsp_end[:n].contiguous(), n_nals=n, start_code_len=3,
                            out_cap=size + size // 64 + 4096)
    assert ins["out_bytes"] == size and ins["n_inserted"] == res.n_epb
    assert torch.equal(ins["out"][:size], d[:size])


@pytest.mark.parametrize("alphabet", [0, 1, 3])
def test_long_parts_are_split(ctx, alphabet):
    """NALs of 64 KiB and more are cut into pieces walked by different warps (cuts only behind non-zero bytes): zero runs
    around every nominal cut, an all-zero NAL (no usable cut), long and short NALs mixed, every start alignment"""
    rng = np.random.default_rng(77 + alphabet)
    n_bytes = 3 << 20
    rbsp = zero_heavy(rng, n_bytes, alphabet)
    for c in range(32 << 10, n_bytes, 32 << 10):  # zero runs of various lengths across the nominal cut positions
        run = int(rng.integers(0, 6000))
        lo = max(0, c - int(rng.integers(0, run + 1)))
        rbsp[lo:lo + run] = 0
    rbsp[(1 << 20) + 1000:(1 << 20) + 200000] = 0  # a long all-zero stretch
    off, end = [], []
    p = int(rng.integers(0, 16))
    while p < n_bytes - 10:
        ln = int(rng.choice([50, 3000, 65535, 65536, 65537, 100000, 300001, 700000]))
        e = min(n_bytes, p + ln)
        off.append(p)
        end.append(e)
        p = e + int(rng.integers(0, 3))
    check(ctx, rbsp, np.array(off, np.int64), np.array(end, np.int64), 3, f"split-a{alphabet}", use_ref=True)
    # one NAL over the whole buffer, and an all-zero NAL
    check(ctx, rbsp, np.array([5], np.int64), np.array([n_bytes - 3], np.int64), 4, f"split-one-a{alphabet}")
    z = np.zeros(200000, np.uint8)
    check(ctx, z, np.array([0, 7], np.int64), np.array([200000, 150000], np.int64), 0, f"split-zero-a{alphabet}")


def test_side_list_of_pieces_can_fill_up(ctx):
    """more pieces than the side list holds (only possible when the output does not fit either): the NALs that find no room
    stay with their own warp, sizes are still exact and the overflow is reported; with room for the output the same extents
    (overlapping on purpose) come out right"""
    from hevcbitstream_b200 import HevcbError

    rng = np.random.default_rng(9)
    rbsp = zero_heavy(rng, 1 << 20, 0)
    n = 40
    off = np.arange(n, dtype=np.int64) * 3
    end = np.full(n, (1 << 20) - 5, np.int64)
    with pytest.raises(HevcbError) as e:
        ctx.insert_host(rbsp, off, end, start_code_len=3, out_cap=2 << 20)
    assert e.value.code == -104
    check(ctx, rbsp, off[:6], end[:6], 3, "overlap-6")


def test_back_to_back_launches_without_host_sync(ctx):
    """the single-pass assembly reuses its scratch (tickets, tile states, batch prefixes) from launch to launch: twenty launches queued
    without a host synchronisation in between must all produce the first launch's bytes (checked on the device)"""
    import torch

    g = torch.Generator(device="cuda").manual_seed(3)
    n_bytes = 256 << 20
    x = torch.randint(0, 256, (n_bytes,), dtype=torch.uint8, device="cuda", generator=g)
    x[::4099] = 0
    x[1::4099] = 0  # a few insertions
    for seg in (16384, 4096 + 37):
        n = n_bytes // seg
        off = torch.arange(n, dtype=torch.int64, device="cuda") * seg
        end = off + seg
        first = ctx.insert_device(x, off, end, start_code_len=3)
        total = first["out_bytes"]
        outs = [ctx.insert_device(x, off, end, start_code_len=3, out_cap=total + 64, sync=False) for _ in range(20)]
        torch.cuda.synchronize()
        for o in outs:
            assert int(o["summary"][1]) == total and int(o["summary"][2]) == first["n_inserted"]
            assert torch.equal(o["out"][:total], first["out"][:total])
            assert torch.equal(o["out_off"], first["out_off"])
