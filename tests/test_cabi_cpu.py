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
    names = [ln.split()[-1] for ln in syms.splitlines() if ln.strip()]
    assert not [nm for nm in names if nm.startswith(("oracle_", "ref_", "hostsim_"))]


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
                          "peek_hevc_nal_unit", "read_debug_hevc_nal_unit", "debug_bytes"}
    for name in names:
        assert getattr(L, name) is not None, name
    # everything the reference's headers declare next to them (include/compat/h264_stream.h, h264_sei.h) is exported too
    for name in ("h264_dbgfile", "more_rbsp_data", "more_rbsp_trailing_data", "_read_ff_coded_number", "_write_ff_coded_number", "read_rbsp_trailing_bits",
                 "intlog2", "is_slice_type", "sei_new", "sei_free", "read_sei_end_bits", "read_sei_payload", "write_sei_payload", "read_debug_sei_payload"):
        assert getattr(L, name) is not None, name
    out = subprocess.run(["ldd", path], capture_output=True, text=True).stdout
    assert "libhevcb200.so" in out and "hevcref" not in out
    if not torch.cuda.is_available():
        L.hevc_new.restype = C.c_void_p
        assert L.hevc_new() is None
        s, e = C.c_int(0), C.c_int(0)
        buf = (C.c_uint8 * 16)(0, 0, 1, 0x40, 1, 2, 3)
        assert L.find_nal_unit(buf, 7, C.byref(s), C.byref(e)) == -1


def test_compat_headers_bs_h_matches_the_oracle_port():
    """include/compat/bs.h (fresh implementation of the reference's bit reader / writer API) against SURVEY Appendix B's known
    answers and against oracle/liboracle.so on random reads and writes; also the host helpers the compat library exports
    (more_rbsp_data, ff-coded numbers, intlog2)."""
    src = r"""
#include <stdio.h>
#include <string.h>
#include "h264_stream.h"
static unsigned long long rng = 88172645463325252ull;
static unsigned rnd(void) { rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17; return (unsigned)(rng >> 11); }
int main(void) {
    /* Appendix B: A6 42 98 E2 04 8A decodes as ue 0..7 */
    uint8_t a[] = {0xA6, 0x42, 0x98, 0xE2, 0x04, 0x8A};
    bs_t b; bs_init(&b, a, sizeof(a));
    for (int i = 0; i < 8; i++) printf("%u ", bs_read_ue(&b));
    bs_init(&b, a, sizeof(a));
    for (int i = 0; i < 8; i++) printf("%d ", bs_read_se(&b));
    uint8_t z[] = {0, 0}; bs_init(&b, z, 2); printf("| %u ", bs_read_ue(&b)); printf("%d ", bs_overrun(&b));
    uint8_t o1[] = {1}; bs_init(&b, o1, 1); printf("%u ", bs_read_ue(&b));
    uint8_t o2[] = {2}; bs_init(&b, o2, 1); printf("%u ", bs_read_ue(&b)); printf("%d ", bs_overrun(&b));
    uint8_t w[16]; memset(w, 0, sizeof(w)); bs_init(&b, w, sizeof(w));
    bs_write_ue(&b, 0); bs_write_ue(&b, 1); bs_write_ue(&b, 2); bs_write_ue(&b, 255); bs_write_ue(&b, 65535); bs_write_se(&b, -3); bs_write_se(&b, 3);
    bs_write_u(&b, 32, 0xDEADBEEF);
    printf("| "); for (int i = 0; i < 12; i++) printf("%02X ", w[i]); printf("%d.%d\n", bs_pos(&b), b.bits_left);
    /* random script: write then read back, positions, eof / overrun past the end */
    uint8_t buf[64]; memset(buf, 0xAA, sizeof(buf)); bs_init(&b, buf, 40);
    unsigned vals[200], kinds[200]; int nops = 0;
    while (nops < 200 && !bs_eof(&b)) {
        unsigned k = rnd() % 4, v = rnd();
        kinds[nops] = k;
        if (k == 0) { int n = 1 + rnd() % 32; vals[nops] = n == 32 ? v : (v & ((1u << n) - 1)); kinds[nops] |= n << 8; bs_write_u(&b, n, vals[nops]); }
        else if (k == 1) { vals[nops] = v % 70000; bs_write_ue(&b, vals[nops]); }
        else if (k == 2) { vals[nops] = (unsigned)((int)(v % 60000) - 30000); bs_write_se(&b, (int)vals[nops]); }
        else { vals[nops] = v & 0xFF; bs_write_u8(&b, vals[nops]); }
        nops++;
    }
    printf("w %d %d %d | ", bs_pos(&b), bs_pos_out(&b), bs_overrun(&b));
    bs_init(&b, buf, 40);
    unsigned long long h = 0; int bad = 0;
    for (int i = 0; i < nops - 3; i++) {
        unsigned k = kinds[i] & 0xFF, r;
        if (k == 0) r = bs_read_u(&b, kinds[i] >> 8); else if (k == 1) r = bs_read_ue(&b); else if (k == 2) r = (unsigned)bs_read_se(&b); else r = bs_read_u8(&b);
        if (r != vals[i]) bad++;
        h = h * 1000003ull + r;
    }
    printf("r bad=%d %d.%d ", bad, bs_pos(&b), b.bits_left); printf("more=%d | ", more_rbsp_data(&b));
    uint8_t ff[] = {0xFF, 0xFF, 0x07, 0x80}; bs_init(&b, ff, 4); printf("ff %d ", _read_ff_coded_number(&b)); printf("more=%d ", more_rbsp_data(&b));
    uint8_t fo[4] = {0}; bs_init(&b, fo, 4); _write_ff_coded_number(&b, 600); printf("%02X%02X%02X ", fo[0], fo[1], fo[2]);
    printf("log %d %d %d %d %d\n", intlog2(0), intlog2(1), intlog2(2), intlog2(5), intlog2(1024));
    return 0;
}
"""
    d = os.path.join(ROOT, "tests", "_hostsim")
    os.makedirs(d, exist_ok=True)
    p = os.path.join(d, "bs_check.c")
    open(p, "w").write(src)
    exe = p[:-2]
    libdir = os.path.join(ROOT, "hevcbitstream_b200")
    subprocess.check_call(["gcc", "-std=gnu99", "-O1", "-I" + os.path.join(ROOT, "include", "compat"), "-o", exe, p, "-L" + libdir, "-lhevcb200_compat",
                           "-lhevcb200", "-Wl,-rpath," + libdir])
    out = subprocess.run([exe], capture_output=True, text=True).stdout
    lines = out.splitlines()
    assert lines[0] == "0 1 2 3 4 5 6 7 0 1 -1 2 -2 3 -3 4 | 32767 1 127 63 0 | A6 01 00 00 00 80 00 1C DB D5 B7 DD 12.5", lines[0]
    assert "r bad=0" in lines[1] and "ff 517 more=0" in lines[1] and "FFFF5A" in lines[1] and lines[1].endswith("log 0 0 1 3 10"), lines[1]
    # the same program against the REFERENCE's own headers and library must print the same (where the reference is available)
    from oracle import ref

    refdir = "/root/reference"
    if os.path.isdir(refdir) and ref.available():
        exe2 = exe + "_ref"
        subprocess.check_call(["gcc", "-std=gnu99", "-O1", "-w", "-I" + refdir, "-o", exe2, p, os.path.join(refdir, "h264_stream.c"), os.path.join(refdir, "h264_nal.c")])
        out2 = subprocess.run([exe2], capture_output=True, text=True).stdout
        assert out2 == out, (out2, out)
