"""GPU tests of the length-prefixed framing <-> Annex-B conversions (hevcb_reframe_device, hevcb_lenpref_index_device; SURVEY 8f-4:
the container step either side of the path).  The reference has no container code, so the oracle for the framing itself is its
definition (ISO/IEC 14496-15: big-endian length + the NAL bytes), restated in numpy below; what the reference CAN check is that
nothing happened to the NAL units: the stream that comes back from the round trip is scanned, stripped and parsed with the
same results as the original (util.compare_scan against oracle/_ref)."""
import numpy as np
import pytest

from oracle import ref
from tests import util

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built")]


@pytest.fixture(autouse=True, params=["fused", "twopass"])
def insert_path(request, monkeypatch):
    """every case runs through the single-pass assembly and through the two-pass kernels (HEVCB_INSERT_PATH, csrc/hevcb_insert.cu)"""
    monkeypatch.setenv("HEVCB_INSERT_PATH", request.param)
    return request.param



def lenpref_numpy(s, st, en, len_size):
    parts = []
    for a, b in zip(st.tolist(), en.tolist()):
        parts.append(np.frombuffer(int(b - a).to_bytes(8, "big")[8 - len_size:], np.uint8))
        parts.append(s[a:b])
    return np.concatenate(parts) if parts else np.zeros(0, np.uint8)


@pytest.mark.parametrize("len_size,seed", [(4, 1), (2, 2), (4, 3), (1, 4)])
def test_round_trip_through_length_prefixed_framing(ctx, len_size, seed):
    import torch

    pmax = {4: 3000, 2: 3000, 1: 150}[len_size]
    s = ref.gen_stream(seed=seed, profile=1, n_slices=4000, payload_min=1, payload_max=pmax, zero_heavy_pct=30, extra_zero_pct=10, ps_period=40,
                       unsupported_pct=5)
    size = s.size - ref.PAD
    d = torch.zeros(size + 32, dtype=torch.uint8, device="cuda")
    d[:size] = torch.from_numpy(s[:size].copy())
    scan = ctx.scan_strip_device(d, size=size)
    n = scan.n_nals
    st, en = scan.nal_start[:n].contiguous(), scan.nal_end[:n].contiguous()
    if len_size == 1:  # keep the units that fit a one-byte length
        keep = (en - st) < 256
        st, en = st[keep].contiguous(), en[keep].contiguous()
        n = int(st.numel())
    # Annex-B -> length-prefixed
    lp = ctx.reframe_device(d, st, en, n_nals=n, len_size=len_size)
    want = lenpref_numpy(s, st.cpu().numpy(), en.cpu().numpy(), len_size)
    got = lp["out"][: lp["out_bytes"]].cpu().numpy()
    assert got.size == want.size and np.array_equal(got, want), "length-prefixed bytes differ from the definition"
    off = lp["out_off"].cpu().numpy()
    assert np.array_equal(np.diff(off), (en - st).cpu().numpy() + len_size)
    # index the length-prefixed data again: one sample per 7 NALs (as a container's sample table would give), and as ONE sample
    lpbuf = torch.zeros(lp["out_bytes"] + 32, dtype=torch.uint8, device="cuda")
    lpbuf[: lp["out_bytes"]] = lp["out"][: lp["out_bytes"]]
    sample_off = torch.from_numpy(np.append(off[:-1:7], off[-1])).cuda()
    for so in (sample_off, None):
        ns2, ne2, n2, bad = ctx.lenpref_index_device(lpbuf, size=lp["out_bytes"], len_size=len_size, sample_off=so)
        assert n2 == n and bad == 0
        assert np.array_equal(ns2[:n].cpu().numpy(), off[:-1] + len_size) and np.array_equal(ne2[:n].cpu().numpy(), off[1:])
    # length-prefixed -> Annex-B (4-byte start codes), then the reference checks the result like any other stream
    ab = ctx.reframe_device(lpbuf, ns2[:n].contiguous(), ne2[:n].contiguous(), n_nals=n, start_code_len=4)
    out = ab["out"][: ab["out_bytes"]]
    exp = np.concatenate([np.concatenate([np.array([0, 0, 0, 1], np.uint8), s[a:b]]) for a, b in zip(st.cpu().numpy().tolist(), en.cpu().numpy().tolist())])
    assert np.array_equal(out.cpu().numpy(), exp)
    buf = util.padded(exp)
    res = ctx.scan_strip_host(buf[: exp.size], size=exp.size)
    util.compare_scan(buf, exp.size, res, res.rbsp, tag=f"reframed-{len_size}")
    assert res.n_nals == n


def test_broken_length_chain_is_reported(ctx):
    import torch

    raw = np.concatenate([np.array([0, 0, 0, 3, 0x40, 1, 2], np.uint8), np.array([0, 0, 0, 9, 1, 2], np.uint8)])  # second length runs past the end
    d = torch.zeros(64, dtype=torch.uint8, device="cuda")
    d[: raw.size] = torch.from_numpy(raw)
    ns, ne, n, bad = ctx.lenpref_index_device(d, size=raw.size, len_size=4)
    assert n == 1 and bad == 1 and int(ns[0]) == 4 and int(ne[0]) == 7


def test_large_units_and_every_alignment(ctx):
    """NAL units of up to 3 MiB at every source / destination alignment (the verbatim copy works in 16-byte vectors)"""
    import torch

    rng = np.random.default_rng(5)
    sizes = [1, 2, 15, 16, 17, 31, 33, 500, 4097, 70000, 3 << 20] + rng.integers(1, 5000, 64).tolist()
    parts, st, en, pos = [], [], [], 0
    for i, z in enumerate(sizes):
        gap = int(rng.integers(0, 19))
        parts += [np.zeros(gap, np.uint8), rng.integers(4, 256, z).astype(np.uint8)]
        st.append(pos + gap)
        en.append(pos + gap + z)
        pos += gap + z
    s = np.concatenate(parts)
    d = torch.zeros(s.size + 32, dtype=torch.uint8, device="cuda")
    d[: s.size] = torch.from_numpy(s)
    st_t, en_t = torch.tensor(st, dtype=torch.int64, device="cuda"), torch.tensor(en, dtype=torch.int64, device="cuda")
    for ls, sc in ((4, 0), (2, 0), (0, 3), (0, 4), (0, 0)):
        out = ctx.reframe_device(d, st_t, en_t, start_code_len=sc, len_size=ls)
        pre = lambda z: (np.frombuffer(int(z).to_bytes(8, "big")[8 - ls:], np.uint8) if ls else np.array([0] * (sc - 1) + [1] if sc else [], np.uint8))
        want = np.concatenate([np.concatenate([pre(b - a), s[a:b]]) for a, b in zip(st, en)])
        assert np.array_equal(out["out"][: out["out_bytes"]].cpu().numpy(), want), (ls, sc)
