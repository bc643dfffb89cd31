"""CPU tests of the header WRITER logic (hevcb_syntax.h, write variant of the walker) compiled for the host by
tests/hostsim: parse -> edit -> write -> splice -> EPB insertion must reproduce, byte for byte, what the reference's
read_hevc_nal_unit / write_hevc_nal_unit / rbsp_to_nal composition (oracle/ref_harness.c: ref_rewrite_all) produces."""
import ctypes as C

import numpy as np
import pytest

from oracle import ref
from tests import rewrite_check as rc

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built")


def sim_rewrite(lib, buf, size, starts, ends, edits):
    n = len(starts)
    cap = size * 2 + 4096 + 64 * n
    out = np.zeros(cap, np.uint8)
    os_, oe = np.zeros(n, np.int64), np.zeros(n, np.int64)
    nrw = C.c_int64(0)
    st = np.ascontiguousarray(starts, np.int64)
    en = np.ascontiguousarray(ends, np.int64)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    lib.hostsim_rewrite_all.restype = C.c_int64
    tot = lib.hostsim_rewrite_all(p(buf), C.c_int64(size), p(st), p(en), C.c_int64(n), C.byref(edits), p(out), C.c_int64(cap), p(os_), p(oe), C.byref(nrw))
    assert tot >= 0
    return out[:tot], os_, oe, nrw.value


@pytest.mark.parametrize("profile,qp,vui", [(0, 0, 0), (0, 3, 1), (1, 0, 0), (1, -2, 1), (1, 5, 0)])
def test_rewrite_matches_reference(hostsim, profile, qp, vui):
    for seed in (1, 2, 3):
        s = ref.gen_stream(seed=seed, profile=profile, n_slices=400, payload_min=1, payload_max=300, zero_heavy_pct=30, extra_zero_pct=10,
                           ps_period=25, unsupported_pct=5)
        size = s.size - ref.PAD
        st, en, _ = ref.scan_all_with_tail(s, size)
        want = ref.rewrite_all(s, size, st, en, qp_delta_add=qp, vui_flip=vui)
        out, os_, oe, nrw = sim_rewrite(hostsim, s, size, st, en, rc.reference_edits(qp, vui))
        assert nrw > 400
        rc.compare_rewrite(out, os_, oe, want, tag=f"p{profile}-s{seed}-qp{qp}-vui{vui}")


def test_identity_rewrite_of_reference_written_stream(hostsim):
    """no edits: slices come back byte-identical (the writer re-emits what the reader parsed)"""
    s = ref.gen_stream(seed=7, profile=1, n_slices=300, payload_min=10, payload_max=200, ps_period=50)
    size = s.size - ref.PAD
    st, en, _ = ref.scan_all_with_tail(s, size)
    out, os_, oe, nrw = sim_rewrite(hostsim, s, size, st, en, rc.reference_edits(0, 0))
    want = ref.rewrite_all(s, size, st, en)
    rc.compare_rewrite(out, os_, oe, want, tag="identity")


def test_field_index_lookup():
    assert rc.field_index(rc.KIND_SLICE, "first_slice_segment_in_pic_flag") == 0
    a = rc.field_index(rc.KIND_SLICE, "pwt.luma_offset_l0[3]")
    b = rc.field_index(rc.KIND_SLICE, "pwt.luma_offset_l0[0]")
    assert a == b + 3
    assert rc.field_index(rc.KIND_SPS, "st_ref_pic_set[2].delta_poc_s0_minus1[1]") - rc.field_index(rc.KIND_SPS, "st_ref_pic_set[1].delta_poc_s0_minus1[1]") == 792 // 4
    from hevcbitstream_b200._lib import load_library
    L = load_library()
    L.hevcb_field_index.restype = C.c_int64
    L.hevcb_field_index.argtypes = [C.c_int, C.c_char_p]
    assert L.hevcb_field_index(rc.KIND_SLICE, b"no_such_field") == -1
    assert L.hevcb_field_index(rc.KIND_SLICE, b"pwt.luma_offset_l0[99]") == -1
