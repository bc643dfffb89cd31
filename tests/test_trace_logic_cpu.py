"""CPU tests of the read_debug (trace) variant of the syntax walker and of hevcb_trace_name: the host build of the kernels' walker
(tests/hostsim) must print, byte for byte, what the unmodified reference's read_debug_hevc_nal_unit loop prints (oracle/_ref:
ref_analyze_to_file = hevc_analyze.c's framing lines + hevc_stream.c:2343-3436).  No GPU, product library not exercised."""
import ctypes as C
import os
import tempfile

import numpy as np
import pytest

from oracle import ref

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built")


def ref_dump(s, size, verbose=1):
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "dump.txt")
        assert ref.lib().ref_analyze_to_file(s.ctypes.data_as(C.c_void_p), size, p.encode(), verbose) == 0
        return open(p, "rb").read()


def sim_dump(lib, s, size, verbose=1):
    st, en, _ = ref.scan_all_with_tail(s, size)
    st = np.ascontiguousarray(st, np.int64)
    en = np.ascontiguousarray(en, np.int64)
    cap = 400 * len(st) * 64 + (1 << 20)
    out = np.zeros(cap, np.uint8)
    lib.hostsim_trace_all.restype = C.c_int64
    n = lib.hostsim_trace_all(s.ctypes.data_as(C.c_void_p), st.ctypes.data_as(C.c_void_p), en.ctypes.data_as(C.c_void_p), C.c_int64(len(st)),
                              C.c_int(verbose), out.ctypes.data_as(C.c_void_p), C.c_int64(cap))
    assert n >= 0, f"hostsim_trace_all failed ({n})"
    return out[:n].tobytes()


def first_diff(a, b):
    la, lb = a.split(b"\n"), b.split(b"\n")
    for i, (x, y) in enumerate(zip(la, lb)):
        if x != y:
            return f"line {i}: ref {x!r} / got {y!r}"
    return f"lengths {len(la)} / {len(lb)}"


def test_config1_shape(hostsim):
    s = ref.gen_stream(seed=0, profile=0, n_slices=300, payload_min=50, payload_max=50, idr_period=100)
    size = s.size - ref.PAD
    a, b = ref_dump(s, size), sim_dump(hostsim, s, size)
    assert a == b, first_diff(a, b)
    assert a.count(b"!! Found NAL") == 303


@pytest.mark.parametrize("seed", list(range(1, 7)))
def test_rich_streams(hostsim, seed):
    s = ref.gen_stream(seed=seed, profile=1, n_slices=600, payload_min=1, payload_max=64, zero_heavy_pct=20, extra_zero_pct=10, ps_period=37,
                       unsupported_pct=5)
    size = s.size - ref.PAD
    a, b = ref_dump(s, size), sim_dump(hostsim, s, size)
    assert a == b, first_diff(a, b)
    for needle in (b"vui->", b"hrd->", b"sub_layer_hrd->", b"st_ref_pic_set->", b"pwt->", b"sld->", b"sh->rpld.", b"reserved_zero_xxbits",
                   b"general_reserved_zero_34bits", b"slice_reserved_flag", b"slice_segment_header_extension_data_byte"):
        assert needle in a, needle
    if seed in (2, 4, 5):
        assert b"sps_range_ext->" in a and b"pps_range_ext->" in a


def test_verbose_zero(hostsim):
    s = ref.gen_stream(seed=3, profile=1, n_slices=100, payload_min=1, payload_max=64, ps_period=20, unsupported_pct=5)
    size = s.size - ref.PAD
    a, b = ref_dump(s, size, 0), sim_dump(hostsim, s, size, 0)
    assert a == b and b"!!" not in a, first_diff(a, b)
