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
    # a rewrite of spec-mode results is refused (documented limit), a default parse afterwards is accepted again
    with pytest.raises(HevcbError):
        ctx.rewrite_device(d, scan, out, size=size)
    out2 = ctx.parse_device(d, scan)
    ctx.rewrite_device(d, scan, out2, size=size)
