"""CPU tests: the parser's host/device syntax walker (hevcb_syntax.h) compiled for the host by tests/hostsim, run NAL by
NAL with the kernels' dependency rule, against the reference's read_hevc_nal_unit loop (oracle/_ref): return codes,
h->nal, digest of every materialised struct, slice-data extents and bytes.  No GPU, product library not exercised."""
import ctypes as C

import numpy as np
import pytest

from oracle import ref

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built")


def sim_parse(lib, buf, st, en, fn="hostsim_parse_all"):
    n = len(st)
    rec = np.zeros(n, dtype=ref.NAL_RECORD_DTYPE)
    st = np.ascontiguousarray(st, np.int64)
    en = np.ascontiguousarray(en, np.int64)
    npairs = C.c_int64(0)
    fl = C.c_uint32(0)
    f = getattr(lib, fn)
    f.restype = C.c_int64
    ok = f(buf.ctypes.data_as(C.c_void_p), st.ctypes.data_as(C.c_void_p), en.ctypes.data_as(C.c_void_p), C.c_int64(n),
                               rec.ctypes.data_as(C.c_void_p), None, C.byref(npairs), C.byref(fl))
    return ok, rec, npairs.value, fl.value


def check(lib, s, tag):
    size = s.size - ref.PAD
    st, en, _ = ref.scan_all_with_tail(s, size)
    rp = ref.parse_all(s, st, en)
    ok, rec, npairs, fl = sim_parse(lib, s, st, en)
    assert ok >= 0, f"[{tag}] count and emit passes disagree ({ok})"
    R = rp["rec"]
    for f in ("strip_rc", "rc", "nal_unit_type", "nal_layer_id", "nal_temporal_id_plus1", "slice_data_size", "state_hash", "slice_data_hash"):
        d = np.nonzero(R[f] != rec[f])[0]
        assert len(d) == 0, f"[{tag}] {f} differs for {len(d)} NALs, first {d[0]} (type {R['nal_unit_type'][d[0]]}): ref {R[f][d[0]]} got {rec[f][d[0]]}"
    assert ok == rp["n_ok"] and fl == 0
    return len(st), ok


def test_config1_shape(hostsim):
    s = ref.gen_stream(seed=0, profile=0, n_slices=2000, payload_min=50, payload_max=50, idr_period=100)
    n, ok = check(hostsim, s, "c1")
    assert n == 2003 and ok == 2003


@pytest.mark.parametrize("seed", list(range(1, 9)))
def test_rich_streams(hostsim, seed):
    s = ref.gen_stream(seed=seed, profile=1, n_slices=2500, payload_min=1, payload_max=64, zero_heavy_pct=20, extra_zero_pct=10,
                       ps_period=37, unsupported_pct=5)
    n, ok = check(hostsim, s, f"rich{seed}")
    assert n > 2500 and ok < n


def test_appendix_b_writer_vectors(hostsim):
    # reference-written Main-profile VPS and an IDR slice header
    vps = bytes.fromhex("40010C01FFFF016000000300900000030000030078" + "15C090")
    stream = np.frombuffer(b"\x00\x00\x00\x01" + vps + b"\x00\x00\x01" + bytes.fromhex("2601AF3E80"), np.uint8)
    s = ref.padded(stream)
    check(hostsim, s, "appB")
