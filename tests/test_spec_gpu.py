"""GPU parity tests of the SPEC-CORRECT mode (HEVCB_PARSE_SPEC) through the C ABI: hevcb_index_host / hevcb_parse_device
with the flag against the oracle of that mode (oracle/_ref/libhevcref_spec.so = the reference's own template with the spec fixes,
oracle/make_spec_ref.py): rc, h->nal, digest of every struct, slice-data extents and bytes, NAL by NAL, on streams whose slices
refer to several live SPS / PPS ids."""
import numpy as np
import pytest

from oracle import ref
from tests import parse_check

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not (ref.available() and ref.spec_available()), reason="oracle/_ref (spec build) not built")]

SPEC = 2  # HEVCB_PARSE_SPEC


@pytest.fixture()
def spec_ref():
    ref.use_spec(True)
    yield ref
    ref.use_spec(False)


@pytest.mark.parametrize("seed", [1, 2, 3, 4, 5, 6])
def test_spec_streams_with_ids(ctx, spec_ref, seed):
    s = ref.gen_stream(seed=seed, profile=1, n_slices=4000, payload_min=1, payload_max=64, zero_heavy_pct=20, extra_zero_pct=10, ps_period=23, unsupported_pct=5)
    size = s.size - ref.PAD
    idx = ctx.index_host(s[:size], size=size, flags=SPEC)
    n, ok = parse_check.compare_index(s, size, idx, tag=f"spec{seed}")
    assert n > 4000 and ok > 3500
    assert not idx.ubflag.any()
    sps = (idx.nal_hdr & 0xFF) == 33
    assert sps.sum() > 10 and (idx.rc[sps] > 0).all()  # with its trailing bits every SPS parses (the reference's loses its last byte, App. A-1)


def test_default_mode_is_unchanged_and_differs(ctx, spec_ref):
    """the same spec-written stream through the default (reference-compatible) walk: many slices resolve other parameter sets"""
    s = ref.gen_stream(seed=7, profile=1, n_slices=3000, payload_min=1, payload_max=64, ps_period=23)
    size = s.size - ref.PAD
    a = ctx.index_host(s[:size], size=size, flags=SPEC)
    b = ctx.index_host(s[:size], size=size)
    assert a.n == b.n
    cnt_a, cnt_b = np.diff(a.pair_off), np.diff(b.pair_off)
    assert (cnt_a != cnt_b).sum() > 100


def test_device_entry_point_and_mode_limits(ctx, spec_ref):
    import torch

    from hevcbitstream_b200 import HevcbError

    s = ref.gen_stream(seed=8, profile=1, n_slices=2000, payload_min=1, payload_max=200, ps_period=31)
    size = s.size - ref.PAD
    d = torch.from_numpy(s[:size].copy()).cuda()
    scan = ctx.scan_strip_device(d, size=size)
    out = ctx.parse_device(d, scan, spec=True)
    host = ctx.index_host(s[:size], size=size, flags=SPEC)
    n = scan.n_nals
    assert np.array_equal(out["rc"][:n].cpu().numpy(), host.rc) and np.array_equal(out["pair_off"].cpu().numpy(), host.pair_off)
    assert np.array_equal(out["pair_value"][: out["n_pairs"]].cpu().numpy(), host.pair_value)


@pytest.mark.parametrize("qp,vui", [(0, 0), (3, 1)])
def test_rewrite_of_spec_results_matches_the_spec_reference(ctx, spec_ref, qp, vui):
    """read -> edit -> write -> rbsp_to_nal in spec mode: byte-exact against the same composition run by the spec build of the reference
    (its writer ends the SPS with trailing bits and resolves the slices' parameter sets through the id tables)"""
    import torch

    from tests import rewrite_check as rc

    for seed in (1, 2):
        s = ref.gen_stream(seed=seed, profile=1, n_slices=3000, payload_min=1, payload_max=3000, zero_heavy_pct=30, extra_zero_pct=10, ps_period=25, unsupported_pct=5)
        size = s.size - ref.PAD
        d = torch.zeros(size + 32, dtype=torch.uint8, device="cuda")
        d[:size] = torch.from_numpy(s[:size].copy())
        scan = ctx.scan_strip_device(d, size=size)
        parsed = ctx.parse_device(d, scan, spec=True)
        edits = []
        if qp:
            edits.append((rc.KIND_SLICE, "slice_qp_delta", rc.EDIT_ADD, qp))
        if vui:
            edits.append((rc.KIND_SPS, "vui.video_full_range_flag", rc.EDIT_XOR, 1))
        out = ctx.rewrite_device(d, scan, parsed, edits, size=size)
        n = scan.n_nals
        st, en = scan.nal_start.cpu().numpy()[:n], scan.nal_end.cpu().numpy()[:n]
        want = ref.rewrite_all(s, size, st, en, qp_delta_add=qp, vui_flip=vui)
        assert out["n_rewritten"] > 3000
        rc.compare_rewrite(out["out"].cpu().numpy()[: out["out_bytes"]], out["out_start"].cpu().numpy()[:n], out["out_end"].cpu().numpy()[:n], want,
                           tag=f"spec-s{seed}-qp{qp}-vui{vui}")
