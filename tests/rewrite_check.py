"""Helpers for the rewrite tests: edit sets and the comparison with the reference's parse -> edit -> write composition."""
from __future__ import annotations

import ctypes as C

import numpy as np

from hevcbitstream_b200._lib import load_library

KIND_VPS, KIND_SPS, KIND_PPS, KIND_SLICE = 1, 2, 3, 4
EDIT_ADD, EDIT_SET, EDIT_XOR = 0, 1, 2


class EditRule(C.Structure):
    _fields_ = [("kind", C.c_int32), ("field", C.c_uint32), ("op", C.c_int32), ("arg", C.c_int32)]


class EditSet(C.Structure):
    _fields_ = [("n", C.c_int32), ("e", EditRule * 8)]


def field_index(kind: int, path: str) -> int:
    L = load_library()
    L.hevcb_field_index.restype = C.c_int64
    L.hevcb_field_index.argtypes = [C.c_int, C.c_char_p]
    v = L.hevcb_field_index(kind, path.encode())
    assert v >= 0, path
    return int(v)


def reference_edits(qp_delta_add: int, vui_flip: int) -> EditSet:
    """The two edits ref_rewrite_all applies: sh->slice_qp_delta += d; sps->vui.video_full_range_flag ^= 1."""
    es = EditSet()
    n = 0
    if qp_delta_add:
        es.e[n] = EditRule(KIND_SLICE, field_index(KIND_SLICE, "slice_qp_delta"), EDIT_ADD, qp_delta_add)
        n += 1
    if vui_flip:
        es.e[n] = EditRule(KIND_SPS, field_index(KIND_SPS, "vui.video_full_range_flag"), EDIT_XOR, 1)
        n += 1
    es.n = n
    return es


def compare_rewrite(out, out_starts, out_ends, ref_res, tag=""):
    exp = ref_res["out"]
    n = len(ref_res["starts"])
    assert np.array_equal(np.asarray(out_starts[:n]), ref_res["starts"]), f"{tag}: NAL starts differ at {int(np.nonzero(np.asarray(out_starts[:n]) != ref_res['starts'])[0][0])}"
    assert np.array_equal(np.asarray(out_ends[:n]), ref_res["ends"]), f"{tag}: NAL ends differ at {int(np.nonzero(np.asarray(out_ends[:n]) != ref_res['ends'])[0][0])}"
    assert out.size == exp.size, f"{tag}: size {out.size} != {exp.size}"
    if not np.array_equal(out, exp):
        i = int(np.nonzero(out != exp)[0][0])
        k = int(np.searchsorted(ref_res["ends"], i, side="right"))
        raise AssertionError(f"{tag}: first byte diff at {i} (NAL {k}): {out[max(0, i - 6):i + 6]} vs {exp[max(0, i - 6):i + 6]}")
