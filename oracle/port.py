"""oracle/port.py -- TEST INFRASTRUCTURE: ctypes binding of oracle/liboracle.so, the plain-C restatement of the
reference's byte/bit layer (oracle/oracle_port.c).  Pinned against the reference by tests/test_oracle_cpu.py."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
PAD = 16
_lib = None


class BS(C.Structure):
    _fields_ = [("start", C.c_void_p), ("p", C.c_void_p), ("end", C.c_void_p), ("bits_left", C.c_int)]


def build():
    src = os.path.join(_HERE, "oracle_port.c")
    if not os.path.exists(_LIB_PATH) or os.path.getmtime(src) > os.path.getmtime(_LIB_PATH):
        subprocess.check_call(["gcc", "-O2", "-std=gnu99", "-fPIC", "-shared", "-o", _LIB_PATH, src])


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        vp, i64 = C.c_void_p, C.c_int64
        p64 = C.POINTER(C.c_int64)
        L.oracle_find_nal_unit.restype = i64
        L.oracle_find_nal_unit.argtypes = [vp, i64, p64, p64]
        L.oracle_scan_all.restype = i64
        L.oracle_scan_all.argtypes = [vp, i64, vp, vp, i64, C.POINTER(C.c_int32), p64, p64]
        L.oracle_nal_to_rbsp.restype = i64
        L.oracle_nal_to_rbsp.argtypes = [vp, p64, vp, p64]
        L.oracle_rbsp_to_nal.restype = i64
        L.oracle_rbsp_to_nal.argtypes = [vp, i64, vp, p64]
        L.oracle_strip_all.restype = i64
        L.oracle_strip_all.argtypes = [vp, vp, vp, i64, vp, vp, vp, vp]
        L.oracle_insert_all.restype = i64
        L.oracle_insert_all.argtypes = [vp, vp, vp, i64, C.c_int, vp, vp]
        L.oracle_bs_init.argtypes = [C.POINTER(BS), vp, i64]
        L.oracle_bs_read_u1.restype = C.c_uint32
        L.oracle_bs_read_u1.argtypes = [C.POINTER(BS)]
        L.oracle_bs_read_u.restype = C.c_uint32
        L.oracle_bs_read_u.argtypes = [C.POINTER(BS), C.c_int]
        L.oracle_bs_read_ue.restype = C.c_uint32
        L.oracle_bs_read_ue.argtypes = [C.POINTER(BS)]
        L.oracle_bs_read_se.restype = C.c_int32
        L.oracle_bs_read_se.argtypes = [C.POINTER(BS)]
        L.oracle_bs_eof.argtypes = [C.POINTER(BS)]
        L.oracle_bs_overrun.argtypes = [C.POINTER(BS)]
        L.oracle_bs_pos_byte.restype = i64
        L.oracle_bs_pos_byte.argtypes = [C.POINTER(BS)]
        L.oracle_bs_bits_left.argtypes = [C.POINTER(BS)]
        L.oracle_hash_ints.restype = C.c_uint64
        L.oracle_hash_ints.argtypes = [C.c_uint64, vp, i64]
        L.oracle_hash_bytes.restype = C.c_uint64
        L.oracle_hash_bytes.argtypes = [C.c_uint64, vp, i64]
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def padded(data) -> np.ndarray:
    a = np.frombuffer(bytes(data), dtype=np.uint8) if not isinstance(data, np.ndarray) else data
    out = np.zeros(a.size + PAD, dtype=np.uint8)
    out[: a.size] = a
    return out


def find_nal_unit(data: bytes):
    buf = padded(data)
    s, e = C.c_int64(0), C.c_int64(0)
    rc = lib().oracle_find_nal_unit(_ptr(buf), len(data), C.byref(s), C.byref(e))
    return int(rc), s.value, e.value


def nal_to_rbsp(nal: bytes):
    src = padded(nal)
    dst = np.zeros(len(nal) + PAD, np.uint8)
    ns, rs = C.c_int64(len(nal)), C.c_int64(len(nal))
    rc = lib().oracle_nal_to_rbsp(_ptr(src), C.byref(ns), _ptr(dst), C.byref(rs))
    return int(rc), ns.value, bytes(dst[: max(rc, 0)])


def rbsp_to_nal(rbsp: bytes) -> bytes:
    src = padded(rbsp)
    dst = np.zeros(len(rbsp) * 3 // 2 + PAD, np.uint8)
    ns = C.c_int64(0)
    rc = lib().oracle_rbsp_to_nal(_ptr(src), len(rbsp), _ptr(dst), C.byref(ns))
    return bytes(dst[:rc])


def scan_all(buf: np.ndarray, size: int):
    cap = size // 3 + 2
    starts = np.zeros(cap, np.int64)
    ends = np.zeros(cap, np.int64)
    rc, ls, le = C.c_int32(0), C.c_int64(0), C.c_int64(0)
    n = lib().oracle_scan_all(_ptr(buf), size, _ptr(starts), _ptr(ends), cap, C.byref(rc), C.byref(ls), C.byref(le))
    return dict(starts=starts[:n], ends=ends[:n], n=int(n), last_rc=rc.value, last_start=ls.value, last_end=le.value)


def scan_all_with_tail(buf: np.ndarray, size: int):
    r = scan_all(buf, size)
    starts, ends = r["starts"], r["ends"]
    if r["last_rc"] == -1:
        starts = np.append(starts, r["last_start"])
        ends = np.append(ends, r["last_end"])
    return starts.astype(np.int64), ends.astype(np.int64), r


def strip_all(buf: np.ndarray, starts, ends):
    n = len(starts)
    total = int((np.asarray(ends) - np.asarray(starts)).sum()) if n else 0
    out = np.zeros(total + PAD, np.uint8)
    off = np.zeros(n + 1, np.int64)
    rc = np.zeros(n, np.int32)
    ns = np.zeros(n, np.int32)
    starts = np.ascontiguousarray(starts, np.int64)
    ends = np.ascontiguousarray(ends, np.int64)
    tot = lib().oracle_strip_all(_ptr(buf), _ptr(starts), _ptr(ends), n, _ptr(out), _ptr(off), _ptr(rc), _ptr(ns))
    return dict(rbsp=out[:tot], rbsp_off=off, rc=rc, nal_size=ns)


def insert_all(rbsp: np.ndarray, off, end, sc_len=0):
    n = len(off)
    total = int((np.asarray(end) - np.asarray(off)).sum()) if n else 0
    out = np.zeros(total * 3 // 2 + (sc_len + 16) * (n + 1) + PAD, np.uint8)
    nal_off = np.zeros(n + 1, np.int64)
    src = padded(rbsp)
    off = np.ascontiguousarray(off, np.int64)
    end = np.ascontiguousarray(end, np.int64)
    tot = lib().oracle_insert_all(_ptr(src), _ptr(off), _ptr(end), n, sc_len, _ptr(out), _ptr(nal_off))
    return dict(out=out[:tot], nal_off=nal_off)


def read_syntax(data: bytes, ops):
    """Run a list of ('u', n) / ('ue',) / ('se',) / ('u1',) reads; returns (values, byte_pos, bits_left, eof, overrun)."""
    buf = padded(data)
    b = BS()
    L = lib()
    L.oracle_bs_init(C.byref(b), _ptr(buf), len(data))
    vals = []
    for op in ops:
        if op[0] == "u":
            vals.append(int(L.oracle_bs_read_u(C.byref(b), op[1])))
        elif op[0] == "u1":
            vals.append(int(L.oracle_bs_read_u1(C.byref(b))))
        elif op[0] == "ue":
            vals.append(int(L.oracle_bs_read_ue(C.byref(b))))
        elif op[0] == "se":
            vals.append(int(L.oracle_bs_read_se(C.byref(b))))
    return vals, int(L.oracle_bs_pos_byte(C.byref(b))), int(L.oracle_bs_bits_left(C.byref(b))), int(L.oracle_bs_eof(C.byref(b))), int(L.oracle_bs_overrun(C.byref(b)))


def hash_ints(a: np.ndarray, seed: int = 0) -> int:
    a = np.ascontiguousarray(a, dtype=np.int32)
    return int(lib().oracle_hash_ints(seed, _ptr(a), a.size))


def hash_bytes(a: np.ndarray, seed: int = 0) -> int:
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return int(lib().oracle_hash_bytes(seed, _ptr(a), a.size))
