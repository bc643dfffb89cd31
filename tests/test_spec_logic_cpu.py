"""CPU tests of the SPEC-CORRECT mode (HEVCB_PARSE_SPEC): the parser's syntax walker (hevcb_syntax.h, compiled for the
host by tests/hostsim) with the spec switch on, against the oracle of that mode -- the reference's own template with the spec fixes
applied, regenerated and compiled by oracle/make_spec_ref.py (oracle/_ref/libhevcref_spec.so).  Streams are written by that
build's writer: several SPS / PPS ids alive at once, slices that refer to any of them, inter-predicted RPS, HRD, VUI."""
import numpy as np
import pytest

from oracle import ref
from tests.test_parse_logic_cpu import sim_parse

pytestmark = pytest.mark.skipif(not (ref.available() and ref.spec_available()), reason="oracle/_ref (spec build) not built")


@pytest.fixture()
def spec_ref():
    ref.use_spec(True)
    yield ref
    ref.use_spec(False)


def check(lib, s, tag, spec=True):
    size = s.size - ref.PAD
    st, en, _ = ref.scan_all_with_tail(s, size)
    rp = ref.parse_all(s, st, en)
    ok, rec, npairs, fl = sim_parse(lib, s, st, en, fn="hostsim_parse_all_spec" if spec else "hostsim_parse_all")
    assert ok >= 0, f"[{tag}] count and emit passes disagree ({ok})"
    R = rp["rec"]
    for f in ("strip_rc", "rc", "nal_unit_type", "nal_layer_id", "nal_temporal_id_plus1", "slice_data_size", "state_hash", "slice_data_hash"):
        d = np.nonzero(R[f] != rec[f])[0]
        assert len(d) == 0, f"[{tag}] {f} differs for {len(d)} NALs, first {d[0]} (type {R['nal_unit_type'][d[0]]}): ref {R[f][d[0]]} got {rec[f][d[0]]}"
    assert ok == rp["n_ok"] and fl == 0
    return len(st), ok, R


@pytest.mark.parametrize("seed", list(range(1, 9)))
def test_rich_streams_with_ids(hostsim, spec_ref, seed):
    s = ref.gen_stream(seed=seed, profile=1, n_slices=2500, payload_min=1, payload_max=64, zero_heavy_pct=20, extra_zero_pct=10, ps_period=23, unsupported_pct=5)
    n, ok, R = check(hostsim, s, f"spec{seed}")
    assert n > 2500
    # the SPS of this mode ends on a byte boundary, so (unlike the reference's, App. A-1) every one of them parses
    sps = R["nal_unit_type"] == 33
    assert sps.sum() > 10 and (R["rc"][sps] > 0).all()


def test_config1_shape_is_the_same_in_both_modes(hostsim, spec_ref):
    """a stream without any of the constructs the fixes touch: same bytes from both writers except the SPS's trailing bits"""
    s = ref.gen_stream(seed=0, profile=0, n_slices=500, payload_min=50, payload_max=50, idr_period=100)
    n, ok, _ = check(hostsim, s, "c1-spec")
    assert n == 503 and ok == 503
