"""ctypes loader for libhevcb200.so (built in-tree by hevcbitstream_b200/csrc/Makefile)."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HEVCB_LIB", os.path.join(_HERE, "libhevcb200.so"))  # HEVCB_LIB: alternative build (experiments)

HEVCB_OK = 0
ERRORS = {
    -100: "HEVCB_E_NODEVICE",
    -101: "HEVCB_E_CUDA",
    -102: "HEVCB_E_ARG",
    -103: "HEVCB_E_ALIGN",
    -104: "HEVCB_E_CAPACITY",
    -105: "HEVCB_E_NOMEM",
}


class HevcbError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"{ERRORS.get(code, code)}: {msg}")
        self.code = code


class ScanSummary(C.Structure):
    _fields_ = [
        ("n_nals", C.c_int64),
        ("n_terminated", C.c_int64),
        ("last_rc", C.c_int32),
        ("overflow", C.c_int32),
        ("last_start", C.c_int64),
        ("last_end", C.c_int64),
        ("rbsp_bytes", C.c_int64),
        ("n_epb", C.c_int64),
    ]


class ParseBuffers(C.Structure):
    _fields_ = [
        ("rc", C.c_void_p), ("nal_hdr", C.c_void_p), ("kind", C.c_void_p), ("ubflag", C.c_void_p), ("hdr_end", C.c_void_p),
        ("cols", C.c_void_p), ("pair_off", C.c_void_p), ("pair_field", C.c_void_p), ("pair_value", C.c_void_p), ("cap_pairs", C.c_int64),
        ("pair_pos", C.c_void_p),  # None: plain parse; an array: the trace (read_debug) variant
        ("flags", C.c_uint32), ("pad", C.c_uint32),  # HEVCB_PARSE_AUX = 1: extension mode (AUD / EOS / EOB / filler / SEI are parsed)
    ]


class BsOp(C.Structure):
    _fields_ = [("kind", C.c_int32), ("n", C.c_int32), ("value", C.c_int32), ("pad", C.c_int32)]


class ParseSummary(C.Structure):
    _fields_ = [
        ("n_nals", C.c_int64), ("n_ok", C.c_int64), ("n_pairs", C.c_int64), ("n_vps", C.c_int64), ("n_sps", C.c_int64),
        ("n_pps", C.c_int64), ("n_slices", C.c_int64), ("overflow", C.c_int32), ("pad", C.c_int32),
    ]


class InsertSummary(C.Structure):
    _fields_ = [("n_nals", C.c_int64), ("out_bytes", C.c_int64), ("n_inserted", C.c_int64), ("overflow", C.c_int32), ("pad", C.c_int32)]


class EditRule(C.Structure):
    _fields_ = [("kind", C.c_int32), ("field", C.c_uint32), ("op", C.c_int32), ("arg", C.c_int32)]


class EditSet(C.Structure):
    _fields_ = [("n", C.c_int32), ("e", EditRule * 8)]


class RewriteSummary(C.Structure):
    _fields_ = [("n_nals", C.c_int64), ("n_rewritten", C.c_int64), ("out_bytes", C.c_int64), ("n_inserted", C.c_int64), ("overflow", C.c_int32),
                ("pad", C.c_int32)]


class ShardSummary(C.Structure):
    _fields_ = [
        ("own", C.c_int64), ("n_nals", C.c_int64), ("first_empty", C.c_int64), ("first_empty_start", C.c_int64), ("rbsp_bytes", C.c_int64),
        ("n_epb", C.c_int64), ("head_end", C.c_int64), ("head_rbsp_end", C.c_int64), ("last_nal_start", C.c_int64),
        ("last_rbsp_off", C.c_int64), ("last_nal_end", C.c_int64), ("is_first", C.c_int32), ("is_last", C.c_int32),
        ("open_at_end", C.c_int32), ("open_err", C.c_int32), ("overflow", C.c_int32), ("tail_len", C.c_int32), ("tail", C.c_uint8 * 32),
        ("head_last3", C.c_uint8 * 3), ("pad", C.c_uint8),
    ]


MAX_SHARDS = 64


class StitchPatch(C.Structure):
    _fields_ = [("shard", C.c_int32), ("set_start", C.c_int32), ("index", C.c_int64), ("nal_start", C.c_int64), ("rbsp_off", C.c_int64),
                ("nal_end", C.c_int64), ("rbsp_end", C.c_int64), ("ends_003", C.c_int32), ("pad", C.c_int32)]


class ParseChain(C.Structure):
    _fields_ = [("sps_in", C.c_void_p), ("pps_in", C.c_void_p), ("sps_out", C.c_void_p), ("pps_out", C.c_void_p), ("buf_size", C.c_int64)]


class StitchResult(C.Structure):
    _fields_ = [
        ("glob", ScanSummary), ("n_shards", C.c_int32), ("n_patches", C.c_int32),
        ("byte_base", C.c_int64 * MAX_SHARDS), ("rbsp_base", C.c_int64 * MAX_SHARDS), ("first_local", C.c_int64 * MAX_SHARDS),
        ("n_owned", C.c_int64 * MAX_SHARDS), ("nal_base", C.c_int64 * MAX_SHARDS), ("cont_last_shard", C.c_int32 * MAX_SHARDS),
        ("cont_last_bytes", C.c_int64 * MAX_SHARDS), ("cont_bytes", C.c_int64 * MAX_SHARDS), ("patches", StitchPatch * (MAX_SHARDS + 8)),
    ]


class StreamIndex(C.Structure):
    _fields_ = [
        ("cap_nals", C.c_int64), ("nal_start", C.c_void_p), ("nal_end", C.c_void_p), ("rbsp_off", C.c_void_p), ("rbsp_end", C.c_void_p),
        ("rbsp", C.c_void_p), ("p", ParseBuffers), ("scan", ScanSummary), ("parse", ParseSummary),
    ]


_lib = None


def load_library() -> C.CDLL:
    """Loads the CUDA library; fails loudly when it has not been built (there is no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `make -C hevcbitstream_b200/csrc` or "
            "`python -c 'import __graft_entry__ as g; g.build()'`. There is no CPU fallback."
        )
    L = C.CDLL(LIB_PATH)
    vp, i64, i32 = C.c_void_p, C.c_int64, C.c_int32
    L.hevcb_create.restype = C.c_int
    L.hevcb_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.hevcb_destroy.restype = None
    L.hevcb_destroy.argtypes = [vp]
    L.hevcb_last_error.restype = C.c_char_p
    L.hevcb_last_error.argtypes = [vp]
    L.hevcb_version.restype = C.c_int
    L.hevcb_launch_count.restype = i64
    L.hevcb_launch_count.argtypes = [vp]
    L.hevcb_sm_count.restype = C.c_int
    L.hevcb_sm_count.argtypes = [vp]
    L.hevcb_scan_strip_device.restype = C.c_int
    L.hevcb_scan_strip_device.argtypes = [vp, vp, i64, vp, vp, i64, vp, vp, vp, vp, vp]
    L.hevcb_scan_strip_host.restype = C.c_int
    L.hevcb_scan_strip_host.argtypes = [vp, vp, i64, vp, vp, i64, vp, vp, vp, C.POINTER(ScanSummary)]
    L.hevcb_scan_strip_shard_device.restype = C.c_int
    L.hevcb_scan_strip_shard_device.argtypes = [vp, vp, i64, i64, C.c_int, C.c_int, vp, vp, i64, vp, vp, vp, vp, vp]
    L.hevcb_apply_patches_device.restype = C.c_int
    L.hevcb_apply_patches_device.argtypes = [vp, C.POINTER(StitchResult), C.c_int, vp, vp, vp, vp, i64, vp]
    L.hevcb_plan_shards.restype = C.c_int
    L.hevcb_plan_shards.argtypes = [vp, i64, C.c_int, vp]
    L.hevcb_stitch_apply_device.restype = C.c_int
    L.hevcb_stitch_apply_device.argtypes = [vp, vp, C.c_int, C.c_int, vp, vp, vp, vp, i64, vp, vp]
    L.hevcb_stitch.restype = C.c_int
    L.hevcb_stitch.argtypes = [C.POINTER(ShardSummary), C.c_int, C.POINTER(StitchResult)]
    L.hevcb_rewrite_device.restype = C.c_int
    L.hevcb_rewrite_device.argtypes = [vp, vp, i64, vp, vp, vp, vp, vp, i64, C.POINTER(ParseBuffers), C.POINTER(EditSet), vp, i64, vp, vp, vp, vp]
    L.hevcb_field_index.restype = i64
    L.hevcb_field_index.argtypes = [C.c_int, C.c_char_p]
    L.hevcb_reframe_device.restype = C.c_int
    L.hevcb_reframe_device.argtypes = [vp, vp, vp, vp, i64, C.c_int, C.c_int, vp, i64, vp, vp, vp]
    L.hevcb_lenpref_index_device.restype = C.c_int
    L.hevcb_lenpref_index_device.argtypes = [vp, vp, i64, C.c_int, vp, i64, vp, vp, i64, vp, vp]
    L.hevcb_bs_read_host.restype = C.c_int
    L.hevcb_bs_read_host.argtypes = [vp, vp, i64, C.POINTER(BsOp), C.c_int, vp, vp, vp]
    L.hevcb_bs_write_host.restype = C.c_int
    L.hevcb_bs_write_host.argtypes = [vp, C.POINTER(BsOp), C.c_int, vp, i64, C.POINTER(i64), C.POINTER(C.c_int32)]
    L.hevcb_parse_rbsp_host.restype = C.c_int
    L.hevcb_parse_rbsp_host.argtypes = [vp, vp, i64, vp, vp, i64, C.POINTER(ParseBuffers), C.POINTER(ParseSummary), vp]
    L.hevcb_trace_name.restype = C.c_int
    L.hevcb_trace_name.argtypes = [C.c_int, C.c_uint32, C.c_char_p, C.c_int]
    L.hevcb_insert_device.restype = C.c_int
    L.hevcb_insert_device.argtypes = [vp, vp, vp, vp, i64, C.c_int, vp, i64, vp, vp, vp]
    L.hevcb_insert_host.restype = C.c_int
    L.hevcb_insert_host.argtypes = [vp, vp, i64, vp, vp, i64, C.c_int, vp, i64, vp, C.POINTER(InsertSummary)]
    L.hevcb_parse_device.restype = C.c_int
    L.hevcb_parse_device.argtypes = [vp, vp, vp, vp, vp, vp, vp, i64, C.POINTER(ParseBuffers), vp, vp]
    L.hevcb_parse_shard_device.restype = C.c_int
    L.hevcb_parse_shard_device.argtypes = [vp, vp, vp, vp, vp, vp, vp, i64, C.POINTER(ParseBuffers), vp, C.POINTER(ParseChain), vp]
    L.hevcb_ps_context_bytes.restype = C.c_int
    L.hevcb_ps_context_bytes.argtypes = [C.POINTER(i64), C.POINTER(i64)]
    L.hevcb_index_host.restype = C.c_int
    L.hevcb_index_host.argtypes = [vp, vp, i64, C.POINTER(StreamIndex)]
    L.hevcb_materialize.restype = C.c_int
    L.hevcb_materialize.argtypes = [C.POINTER(StreamIndex), i64, vp, vp, vp, vp, vp]
    _lib = L
    return L
