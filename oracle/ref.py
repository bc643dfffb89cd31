"""oracle/ref.py -- TEST INFRASTRUCTURE: ctypes binding of oracle/_ref/libhevcref.so.

libhevcref.so is the UNMODIFIED reference (leslie-wang/hevcbitstream) compiled from /root/reference
by oracle/Makefile together with oracle/ref_harness.c.  It is the strongest parity oracle this repo
has and also the CPU baseline (`cpu_baseline.kind == "reference"`).  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_ref", "libhevcref.so")
ANALYZE_BIN = os.path.join(_HERE, "_ref", "hevc_analyze")

PAD = 16  # zero bytes kept after every buffer handed to the reference (it reads up to buf[size+2])


def available() -> bool:
    return os.path.exists(_LIB_PATH)


class GenParams(C.Structure):
    _fields_ = [
        ("seed", C.c_uint64),
        ("profile", C.c_int32),
        ("idr_period", C.c_int32),
        ("n_slices", C.c_int64),
        ("payload_min", C.c_int32),
        ("payload_max", C.c_int32),
        ("zero_heavy_pct", C.c_int32),
        ("extra_zero_pct", C.c_int32),
        ("ps_period", C.c_int32),
        ("unsupported_pct", C.c_int32),
    ]


NAL_RECORD_DTYPE = np.dtype(
    [
        ("rc", "<i4"),
        ("strip_rc", "<i4"),
        ("nal_unit_type", "<i4"),
        ("nal_layer_id", "<i4"),
        ("nal_temporal_id_plus1", "<i4"),
        ("slice_data_size", "<i4"),
        ("state_hash", "<u8"),
        ("slice_data_hash", "<u8"),
    ]
)

_lib = None
_SPEC_LIB_PATH = os.path.join(_HERE, "_ref", "libhevcref_spec.so")
_libs = {}
_spec = False


def spec_available() -> bool:
    return os.path.exists(_SPEC_LIB_PATH)


def use_spec(flag: bool) -> None:
    """Selects which build every function of this module talks to: the unmodified reference (default) or the reference with
    the spec fixes of oracle/make_spec_ref.py (the oracle of the spec-correct mode)."""
    global _spec, _lib
    _spec = bool(flag)
    _lib = _libs.get(_spec)


def lib():
    global _lib
    if _lib is None:
        path = _SPEC_LIB_PATH if _spec else _LIB_PATH
        if not os.path.exists(path):
            raise RuntimeError(f"{path} missing: run `make -C oracle ref` / `python oracle/make_spec_ref.py` where /root/reference exists")
        L = C.CDLL(path)
        p8 = C.POINTER(C.c_uint8)
        p64 = C.POINTER(C.c_int64)
        p32 = C.POINTER(C.c_int32)
        L.ref_scan_all.restype = C.c_int64
        L.ref_scan_all.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, p32, p64, p64]
        L.ref_strip_all.restype = C.c_int64
        L.ref_strip_all.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_insert_all.restype = C.c_int64
        L.ref_insert_all.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int64, C.c_void_p]
        L.ref_parse_all.restype = C.c_int64
        L.ref_parse_all.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_int32, p32, C.c_void_p, C.c_int32, p32, C.c_void_p, C.c_int32, p32]
        L.ref_gen_stream.restype = C.c_int64
        L.ref_gen_stream.argtypes = [C.POINTER(GenParams), C.c_void_p, C.c_int64]
        L.ref_rewrite_all.restype = C.c_int64
        L.ref_rewrite_all.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int,
                                      C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        L.ref_time_loop.restype = C.c_double
        L.ref_time_loop.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int, p64]
        L.ref_find_nal_unit.restype = C.c_int
        L.ref_find_nal_unit.argtypes = [C.c_void_p, C.c_int, p32, p32]
        L.ref_nal_to_rbsp.restype = C.c_int
        L.ref_nal_to_rbsp.argtypes = [C.c_void_p, p32, C.c_void_p, p32]
        L.ref_rbsp_to_nal.restype = C.c_int
        L.ref_rbsp_to_nal.argtypes = [C.c_void_p, p32, C.c_void_p, p32]
        L.ref_hash_ints.restype = C.c_uint64
        L.ref_hash_ints.argtypes = [C.c_uint64, C.c_void_p, C.c_int64]
        L.ref_hash_bytes.restype = C.c_uint64
        L.ref_hash_bytes.argtypes = [C.c_uint64, C.c_void_p, C.c_int64]
        L.ref_sizeof.restype = C.c_int
        L.ref_sizeof.argtypes = [C.c_int]
        L.ref_analyze_to_file.restype = C.c_int
        L.ref_analyze_to_file.argtypes = [C.c_void_p, C.c_int64, C.c_char_p, C.c_int]
        L.ref_writer_reset.restype = None
        L.ref_writer_struct.restype = C.c_void_p
        L.ref_writer_struct.argtypes = [C.c_int]
        L.ref_writer_write.restype = C.c_int
        L.ref_writer_write.argtypes = [C.c_void_p, C.c_int]
        L.ref_writer_read.restype = C.c_int
        L.ref_writer_read.argtypes = [C.c_void_p, C.c_int]
        _lib = L
        _libs[_spec] = L
    return _lib


def padded(data) -> np.ndarray:
    """uint8 array = data followed by PAD zero bytes (returns the padded array; logical size is len(data))."""
    a = np.frombuffer(bytes(data), dtype=np.uint8) if not isinstance(data, np.ndarray) else data
    out = np.zeros(a.size + PAD, dtype=np.uint8)
    out[: a.size] = a
    return out


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


SIZEOF = {"vps": 0, "sps": 1, "pps": 2, "sh": 3, "nal": 4, "stream": 5, "bs": 6}


def sizeof(name: str) -> int:
    return lib().ref_sizeof(SIZEOF[name])


def find_nal_unit(data: bytes):
    """Single reference call on `data` (zero padded). Returns (rc, nal_start, nal_end)."""
    buf = padded(data)
    s = C.c_int32(0)
    e = C.c_int32(0)
    rc = lib().ref_find_nal_unit(_ptr(buf), len(data), C.byref(s), C.byref(e))
    return rc, s.value, e.value


def nal_to_rbsp(nal: bytes):
    """Returns (rc, nal_size, rbsp bytes)."""
    src = padded(nal)
    n = len(nal)
    dst = np.zeros(n + PAD, dtype=np.uint8)
    ns = C.c_int32(n)
    rs = C.c_int32(n)
    rc = lib().ref_nal_to_rbsp(_ptr(src), C.byref(ns), _ptr(dst), C.byref(rs))
    return rc, ns.value, bytes(dst[: max(rc, 0)])


def rbsp_to_nal(rbsp: bytes) -> bytes:
    src = padded(rbsp)
    n = len(rbsp)
    dst = np.zeros(n * 3 // 2 + PAD, dtype=np.uint8)
    rs = C.c_int32(n)
    ns = C.c_int32(dst.size)
    rc = lib().ref_rbsp_to_nal(_ptr(src), C.byref(rs), _ptr(dst), C.byref(ns))
    assert rc >= 0
    return bytes(dst[:rc])


def scan_all(buf: np.ndarray, size: int, cap: int | None = None):
    """Canonical find_nal_unit loop. `buf` must hold >= size+PAD bytes (zero padded).
    Returns dict(starts, ends, n, last_rc, last_start, last_end)."""
    assert buf.dtype == np.uint8 and buf.size >= size + 8
    if cap is None:
        cap = size // 3 + 2
    starts = np.zeros(cap, dtype=np.int64)
    ends = np.zeros(cap, dtype=np.int64)
    rc = C.c_int32(0)
    ls = C.c_int64(0)
    le = C.c_int64(0)
    n = lib().ref_scan_all(_ptr(buf), size, _ptr(starts), _ptr(ends), cap, C.byref(rc), C.byref(ls), C.byref(le))
    assert n <= cap
    return dict(starts=starts[:n], ends=ends[:n], n=int(n), last_rc=rc.value, last_start=ls.value, last_end=le.value)


def scan_all_with_tail(buf: np.ndarray, size: int):
    """starts/ends of every NAL a reader visits: the rc>0 NALs plus the unterminated last NAL
    (rc == -1, end == size), as hevc_analyze.c:190-205 does."""
    r = scan_all(buf, size)
    starts, ends = r["starts"], r["ends"]
    if r["last_rc"] == -1:
        starts = np.append(starts, r["last_start"])
        ends = np.append(ends, r["last_end"])
    return starts.astype(np.int64), ends.astype(np.int64), r


def strip_all(buf: np.ndarray, starts: np.ndarray, ends: np.ndarray):
    n = len(starts)
    total = int((ends - starts).sum()) if n else 0
    out = np.zeros(total + PAD, dtype=np.uint8)
    off = np.zeros(n + 1, dtype=np.int64)
    rc = np.zeros(n, dtype=np.int32)
    ns = np.zeros(n, dtype=np.int32)
    starts = np.ascontiguousarray(starts, dtype=np.int64)
    ends = np.ascontiguousarray(ends, dtype=np.int64)
    tot = lib().ref_strip_all(_ptr(buf), _ptr(starts), _ptr(ends), n, _ptr(out), out.size, _ptr(off), _ptr(rc), _ptr(ns))
    assert tot >= 0
    return dict(rbsp=out[:tot], rbsp_off=off, rc=rc, nal_size=ns)


def insert_all(rbsp: np.ndarray, rbsp_off: np.ndarray, rbsp_end: np.ndarray, sc_len: int = 0):
    n = len(rbsp_off)
    total = int((rbsp_end - rbsp_off).sum()) if n else 0
    out = np.zeros(total * 3 // 2 + (sc_len + 16) * (n + 1) + PAD, dtype=np.uint8)
    nal_off = np.zeros(n + 1, dtype=np.int64)
    src = padded(rbsp) if rbsp.size == 0 or True else rbsp
    rbsp_off = np.ascontiguousarray(rbsp_off, dtype=np.int64)
    rbsp_end = np.ascontiguousarray(rbsp_end, dtype=np.int64)
    tot = lib().ref_insert_all(_ptr(src), _ptr(rbsp_off), _ptr(rbsp_end), n, sc_len, _ptr(out), out.size, _ptr(nal_off))
    assert tot >= 0
    return dict(out=out[:tot], nal_off=nal_off)


def parse_all(buf: np.ndarray, starts: np.ndarray, ends: np.ndarray, dump_sh: bool = False, ps_cap: int = 0):
    """read_hevc_nal_unit over all NALs with one hevc_stream_t. Returns dict(rec, sh, vps, sps, pps)."""
    n = len(starts)
    rec = np.zeros(n, dtype=NAL_RECORD_DTYPE)
    starts = np.ascontiguousarray(starts, dtype=np.int64)
    ends = np.ascontiguousarray(ends, dtype=np.int64)
    shw = sizeof("sh") // 4
    sh = np.zeros((n, shw), dtype=np.int32) if dump_sh else None
    dumps = {}
    counts = {}
    for name in ("vps", "sps", "pps"):
        w = sizeof(name) // 4
        dumps[name] = np.zeros((ps_cap, w), dtype=np.int32) if ps_cap > 0 else None
        counts[name] = C.c_int32(0)
    ok = lib().ref_parse_all(
        _ptr(buf), _ptr(starts), _ptr(ends), n, _ptr(rec), _ptr(sh) if dump_sh else None,
        _ptr(dumps["vps"]) if ps_cap else None, ps_cap, C.byref(counts["vps"]),
        _ptr(dumps["sps"]) if ps_cap else None, ps_cap, C.byref(counts["sps"]),
        _ptr(dumps["pps"]) if ps_cap else None, ps_cap, C.byref(counts["pps"]),
    )
    out = dict(rec=rec, sh=sh, n_ok=int(ok))
    for name in ("vps", "sps", "pps"):
        out["n_" + name] = counts[name].value
        out[name] = dumps[name][: min(ps_cap, counts[name].value)] if ps_cap else None
    return out


def gen_stream(seed=0, profile=0, n_slices=100, payload_min=100, payload_max=100, idr_period=100,
               zero_heavy_pct=0, extra_zero_pct=0, ps_period=0, unsupported_pct=0) -> np.ndarray:
    """Synthetic Annex-B stream written by the reference's own writer. Returns a uint8 array that is
    zero padded by PAD bytes; the logical stream is out[:-PAD]."""
    gp = GenParams(seed, profile, idr_period, n_slices, payload_min, payload_max, zero_heavy_pct,
                   extra_zero_pct, ps_period, unsupported_pct)
    cap = int(n_slices) * (int(payload_max) * 3 // 2 + 1200) + (1 << 20)
    if profile == 1 and ps_period > 0:
        cap += (int(n_slices) // ps_period + 1) * 20000
    out = np.zeros(cap + PAD, dtype=np.uint8)
    n = lib().ref_gen_stream(C.byref(gp), _ptr(out), cap)
    if n < 0:
        raise RuntimeError("ref_gen_stream failed (capacity or writer error)")
    res = out[: n + PAD].copy()
    res[n:] = 0
    return res


def rewrite_all(buf: np.ndarray, size: int, starts, ends, qp_delta_add=0, vui_flip=0):
    n = len(starts)
    cap = size * 2 + 4096 + 64 * n
    out = np.zeros(cap + PAD, dtype=np.uint8)
    os_ = np.zeros(n, dtype=np.int64)
    oe = np.zeros(n, dtype=np.int64)
    starts = np.ascontiguousarray(starts, dtype=np.int64)
    ends = np.ascontiguousarray(ends, dtype=np.int64)
    tot = lib().ref_rewrite_all(_ptr(buf), size, _ptr(starts), _ptr(ends), n, qp_delta_add, vui_flip,
                                _ptr(out), cap, _ptr(os_), _ptr(oe))
    assert tot >= 0
    return dict(out=out[:tot], starts=os_, ends=oe)


def time_loop(buf: np.ndarray, size: int, mode: int, reps: int = 3):
    """Best-of-reps seconds for the reference CPU loop. mode 0 scan, 1 scan+strip, 2 scan+read."""
    n = C.c_int64(0)
    t = lib().ref_time_loop(_ptr(buf), size, mode, reps, C.byref(n))
    return t, n.value


def hash_ints(a: np.ndarray, seed: int = 0) -> int:
    a = np.ascontiguousarray(a, dtype=np.int32)
    return int(lib().ref_hash_ints(seed, _ptr(a), a.size))


def hash_bytes(a: np.ndarray, seed: int = 0) -> int:
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return int(lib().ref_hash_bytes(seed, _ptr(a), a.size))
