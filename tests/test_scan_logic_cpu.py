"""CPU tests: the kernels' host/device logic (hevcb_scan_core.h), compiled for the host by tests/hostsim,
against the reference-built oracle.  No GPU involved; the product library is not exercised here."""
import ctypes as C

import numpy as np
import pytest

from oracle import ref
from tests import util

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built")


class Summ(C.Structure):
    _fields_ = [("n_nals", C.c_int64), ("n_terminated", C.c_int64), ("last_rc", C.c_int32), ("overflow", C.c_int32),
                ("last_start", C.c_int64), ("last_end", C.c_int64), ("rbsp_bytes", C.c_int64), ("n_epb", C.c_int64)]


class SimResult:
    pass


def run_sim(lib, buf, size, fast=1):
    cap = size // 3 + 8
    r = SimResult()
    r.nal_start = np.full(cap, -7, np.int64)
    r.nal_end = np.full(cap, -7, np.int64)
    r.rbsp_off = np.full(cap, -7, np.int64)
    r.rbsp_end = np.full(cap, -7, np.int64)
    img = np.zeros(size + 16, np.uint8)
    s = Summ()
    lib.hostsim_scan_strip.restype = C.c_int64
    lib.hostsim_scan_strip(buf.ctypes.data_as(C.c_void_p), C.c_int64(size), r.nal_start.ctypes.data_as(C.c_void_p),
                           r.nal_end.ctypes.data_as(C.c_void_p), r.rbsp_off.ctypes.data_as(C.c_void_p),
                           r.rbsp_end.ctypes.data_as(C.c_void_p), C.c_int64(cap), img.ctypes.data_as(C.c_void_p), C.byref(s), fast)
    for f, _ in Summ._fields_:
        setattr(r, f, getattr(s, f))
    return r, img


@pytest.mark.parametrize("alphabet", [0, 1, 2, 3])
def test_adversarial_small(hostsim, alphabet):
    rng = np.random.default_rng(100 + alphabet)
    for it in range(4000):
        size = int(rng.integers(0, 90))
        buf = util.adversarial(rng, size, alphabet)
        r, img = run_sim(hostsim, buf, size)
        util.compare_scan(buf, size, r, img, tag=f"a{alphabet}-{it}")


def test_adversarial_mid(hostsim):
    rng = np.random.default_rng(7)
    for it in range(200):
        size = int(rng.integers(100, 6000))
        buf = util.adversarial(rng, size, it, density=0.3)
        r, img = run_sim(hostsim, buf, size, fast=it & 1)
        util.compare_scan(buf, size, r, img, tag=f"mid{it}")


@pytest.mark.parametrize("seed", [1, 2])
def test_generated_streams(hostsim, seed):
    s = ref.gen_stream(seed=seed, profile=1, n_slices=1500, payload_min=1, payload_max=300, zero_heavy_pct=30,
                       extra_zero_pct=20, ps_period=40, unsupported_pct=5)
    size = s.size - ref.PAD
    for cut in (0, 1, 2, 3, 5, 7):
        buf = util.padded(s[: size - cut])
        r, img = run_sim(hostsim, buf, size - cut)
        n = util.compare_scan(buf, size - cut, r, img, tag=f"gen{seed}-{cut}")
        assert n > 1500


def test_appendix_b_vectors(hostsim):
    h = lambda s: np.frombuffer(bytes.fromhex(s.replace(" ", "")), np.uint8)
    for v in ["00 00 01 40 01 02 03 04 00 00 01 42 01", "00 00 00 01 40 01 02 03 04 00 00 00 01 42 01",
              "09 09 00 00 01 40 01 02 03 04 05 00 00 01 09 09", "00 00 01 40 01 02 00 00 00 00 00 01 09 09 09",
              "00 00 01 00 00 01 40 01", "00 00 01 40 01 02 03 04 05 06 07 08", "00 00 01 40 01 02 03 04 05 00 00 01",
              "01 02 03 04 05 06 00 00 01 0A", "01 02 03 04 05 00 00 01 09 0A", "00 00 01 40 00 00 03 01 05 00 00 01 07 07",
              "", "00", "00 00 01", "00 00 00 01"]:
        a = h(v)
        buf = util.padded(a)
        r, img = run_sim(hostsim, buf, a.size)
        util.compare_scan(buf, a.size, r, img, tag=v)
