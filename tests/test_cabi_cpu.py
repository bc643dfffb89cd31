"""CPU tests of the drop-in boundary: the C-ABI library loads here (no GPU), exports every symbol include/hevcb.h
declares, and fails loudly -- not silently on a CPU path -- when no CUDA device exists."""
import ctypes as C
import os
import re
import subprocess

import pytest

import hevcbitstream_b200 as hb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    lib = hb.load_library()
    hdr = open(os.path.join(ROOT, "include", "hevcb.h")).read()
    names = re.findall(r"HEVCB_API\s+[\w\s\*]+?\b(hevcb_\w+)\s*\(", hdr)
    assert len(names) >= 11, names
    for name in names:
        assert getattr(lib, name) is not None, name


def test_no_cpu_fallback_without_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(hb.HevcbError) as e:
        hb.Context(0)
    assert e.value.code == -100 and "no CPU fallback" in str(e.value)


def test_product_library_does_not_link_the_oracle():
    out = subprocess.run(["ldd", hb.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out and "hevcref" not in out and "hostsim" not in out
    syms = subprocess.run(["nm", "-D", "--defined-only", hb.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle_" not in syms and "ref_" not in syms and "hostsim_" not in syms


def test_layout_header_compiles_as_c():
    src = '#include "include/hevcb.h"\n#include "include/hevcb_layout.h"\nint main(void){return (int)sizeof(hevc_sps_t) == 76256 ? 0 : 1;}\n'
    p = os.path.join(ROOT, "tests", "_hostsim", "layout_check.c")
    os.makedirs(os.path.dirname(p), exist_ok=True)
    open(p, "w").write(src)
    exe = p[:-2]
    subprocess.check_call(["gcc", "-std=gnu99", "-Wall", "-I" + ROOT, "-o", exe, p])
    assert subprocess.run([exe]).returncode == 0


def test_compat_library_exports_the_reference_api_and_fails_loudly_without_a_device():
    """include/hevcb_compat.h: the reference's own function names; without a GPU hevc_new() returns NULL (no CPU fallback)"""
    import ctypes as C

    import torch

    path = os.path.join(ROOT, "hevcbitstream_b200", "libhevcb200_compat.so")
    hb.load_library()
    L = C.CDLL(path)
    hdr = open(os.path.join(ROOT, "include", "hevcb_compat.h")).read()
    names = re.findall(r"HEVCB_COMPAT_API\s+[\w\s\*]+?\b(\w+)\s*\(", hdr)
    assert set(names) == {"hevc_new", "hevc_free", "find_nal_unit", "nal_to_rbsp", "rbsp_to_nal", "read_hevc_nal_unit", "write_hevc_nal_unit",
                          "peek_hevc_nal_unit"}
    for name in names:
        assert getattr(L, name) is not None, name
    out = subprocess.run(["ldd", path], capture_output=True, text=True).stdout
    assert "libhevcb200.so" in out and "hevcref" not in out
    if not torch.cuda.is_available():
        L.hevc_new.restype = C.c_void_p
        assert L.hevc_new() is None
        s, e = C.c_int(0), C.c_int(0)
        buf = (C.c_uint8 * 16)(0, 0, 1, 0x40, 1, 2, 3)
        assert L.find_nal_unit(buf, 7, C.byref(s), C.byref(e)) == -1
