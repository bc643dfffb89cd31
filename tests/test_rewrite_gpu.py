"""GPU parity tests for hevcb_rewrite_device: scan + strip + parse + rewrite on the device must reproduce byte for byte
the reference's read_hevc_nal_unit -> edit -> write_hevc_nal_unit -> rbsp_to_nal composition (ref_rewrite_all)."""
import numpy as np
import pytest

from oracle import ref
from tests import rewrite_check as rc

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built")]


@pytest.fixture(autouse=True, params=["fused", "twopass"])
def insert_path(request, monkeypatch):
    """every case runs through the single-pass assembly and through the two-pass kernels (HEVCB_INSERT_PATH, csrc/hevcb_insert.cu)"""
    monkeypatch.setenv("HEVCB_INSERT_PATH", request.param)
    return request.param



def device_rewrite(ctx, s, size, qp, vui):
    import torch

    d = torch.zeros(size + 32, dtype=torch.uint8, device="cuda")
    d[:size] = torch.from_numpy(s[:size].copy())
    scan = ctx.scan_strip_device(d, size=size)
    parsed = ctx.parse_device(d, scan)
    edits = []
    if qp:
        edits.append((rc.KIND_SLICE, "slice_qp_delta", rc.EDIT_ADD, qp))
    if vui:
        edits.append((rc.KIND_SPS, "vui.video_full_range_flag", rc.EDIT_XOR, 1))
    out = ctx.rewrite_device(d, scan, parsed, edits, size=size)
    n = scan.n_nals
    return (out["out"].cpu().numpy()[: out["out_bytes"]], out["out_start"].cpu().numpy()[:n], out["out_end"].cpu().numpy()[:n], out,
            scan.nal_start.cpu().numpy()[:n], scan.nal_end.cpu().numpy()[:n])


@pytest.mark.parametrize("profile,qp,vui", [(0, 0, 0), (0, 3, 1), (1, 0, 0), (1, -2, 1), (1, 5, 0)])
def test_rewrite_matches_reference(ctx, profile, qp, vui):
    for seed in (1, 2):
        s = ref.gen_stream(seed=seed, profile=profile, n_slices=3000, payload_min=1, payload_max=3000, zero_heavy_pct=30, extra_zero_pct=10,
                           ps_period=25, unsupported_pct=5)
        size = s.size - ref.PAD
        got, os_, oe, out, st, en = device_rewrite(ctx, s, size, qp, vui)
        want = ref.rewrite_all(s, size, st, en, qp_delta_add=qp, vui_flip=vui)
        assert out["n_rewritten"] > 3000
        rc.compare_rewrite(got, os_, oe, want, tag=f"p{profile}-s{seed}-qp{qp}-vui{vui}")


def test_config1_stream_qp_edit_round_trip(ctx):
    """BASELINE config 1 shape (Main 1080p, 10k slices of ~6.7 KB): +2 on slice_qp_delta, then parse the rewritten stream
    again on the device: every slice_qp_delta moved by 2, everything else identical"""
    import torch

    s = ref.gen_stream(seed=0, profile=0, n_slices=10000, payload_min=6680, payload_max=6680, idr_period=100)
    size = s.size - ref.PAD
    got, os_, oe, out, st, en = device_rewrite(ctx, s, size, 2, 0)
    want = ref.rewrite_all(s, size, st, en, qp_delta_add=2, vui_flip=0)
    rc.compare_rewrite(got, os_, oe, want, tag="c1")
    d0 = torch.from_numpy(s[:size].copy()).cuda()
    sc0 = ctx.scan_strip_device(d0, size=size)
    p0 = ctx.parse_device(d0, sc0)
    d1 = out["out"][: out["out_bytes"] + 16]
    sc1 = ctx.scan_strip_device(d1, size=out["out_bytes"])
    p1 = ctx.parse_device(d1, sc1)
    n = sc0.n_nals
    assert sc1.n_nals == n
    k0, k1 = p0["kind"][:n].cpu().numpy(), p1["kind"][:n].cpu().numpy()
    assert np.array_equal(k0, k1)
    q0, q1 = p0["cols"][1, :n].cpu().numpy(), p1["cols"][1, :n].cpu().numpy()
    sl = k0 == rc.KIND_SLICE
    assert sl.sum() == 10000
    assert np.array_equal(q1[sl], q0[sl] + 2)
    for col in (0, 2, 3, 4):
        assert np.array_equal(p0["cols"][col, :n].cpu().numpy()[sl], p1["cols"][col, :n].cpu().numpy()[sl])


def test_corrupt_and_unsupported_nals_pass_through(ctx):
    s = ref.gen_stream(seed=9, profile=1, n_slices=1500, payload_min=20, payload_max=500, ps_period=30, unsupported_pct=20)
    size = s.size - ref.PAD
    st, en, _ = ref.scan_all_with_tail(s, size)
    a = s.copy()
    rng = np.random.default_rng(1)
    hit = 0
    for k in rng.choice(len(st), 120, replace=False):
        t = (a[st[k]] >> 1) & 0x3F
        if t <= 21 and en[k] - st[k] > 16:  # slice NALs only: nal_to_rbsp error pattern near the end
            p = int(en[k]) - 6
            a[p:p + 4] = [0x55, 0, 0, 2]
            hit += 1
    assert hit > 20
    got, os_, oe, out, st2, en2 = device_rewrite(ctx, a, size, 1, 1)
    want = ref.rewrite_all(a, size, st2, en2, qp_delta_add=1, vui_flip=1)
    rc.compare_rewrite(got, os_, oe, want, tag="corrupt")
    assert out["n_rewritten"] < len(st2)


def test_rewrite_requires_matching_parse(ctx):
    import torch

    from hevcbitstream_b200 import HevcbError

    s = ref.gen_stream(seed=3, profile=0, n_slices=50, payload_min=10, payload_max=50)
    size = s.size - ref.PAD
    d = torch.from_numpy(s[:size + 16].copy()).cuda()
    scan = ctx.scan_strip_device(d, size=size)
    parsed = ctx.parse_device(d, scan)
    s2 = ref.gen_stream(seed=3, profile=0, n_slices=70, payload_min=10, payload_max=50)
    d2 = torch.from_numpy(s2[: s2.size - ref.PAD + 16].copy()).cuda()
    scan2 = ctx.scan_strip_device(d2, size=s2.size - ref.PAD)
    with pytest.raises(HevcbError):
        ctx.rewrite_device(d2, scan2, parsed, [], size=s2.size - ref.PAD)


@pytest.mark.parametrize("slot", ["0", "16", "48"])
def test_second_header_pass_with_small_slots(ctx, monkeypatch, slot):
    """HEVCB_HDR_SLOT shrinks the per-slice slot of the first header pass, so that (nearly) every slice header is written by
    the second pass into the compact staging: both routes must give the reference's bytes"""
    monkeypatch.setenv("HEVCB_HDR_SLOT", slot)
    s = ref.gen_stream(seed=11, profile=1, n_slices=300, payload_min=1, payload_max=400, zero_heavy_pct=20, ps_period=40, unsupported_pct=3)
    size = s.size - ref.PAD
    got, os_, oe, out, st, en = device_rewrite(ctx, s, size, 2, 1)
    want = ref.rewrite_all(s, size, st, en, qp_delta_add=2, vui_flip=1)
    assert out["n_rewritten"] > 300
    rc.compare_rewrite(got, os_, oe, want, tag=f"slot{slot}")


def test_rewrite_of_a_buffer_without_nals(ctx):
    """no start code at all: nothing to parse, every byte is copied through"""
    rng = np.random.default_rng(3)
    s = np.concatenate([rng.integers(2, 256, 5000, dtype=np.uint8), np.zeros(ref.PAD, np.uint8)])
    size = s.size - ref.PAD
    got, os_, oe, out, st, en = device_rewrite(ctx, s, size, 1, 0)
    assert len(st) == 0 and np.array_equal(got, s[:size])
