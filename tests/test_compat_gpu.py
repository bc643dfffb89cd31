"""GPU tests of the compatibility layer (include/hevcb_compat.h, libhevcb200_compat.so): the reference's own per-NAL API --
find_nal_unit, nal_to_rbsp, rbsp_to_nal, hevc_new / read_hevc_nal_unit / hevc_free, peek_hevc_nal_unit -- served by the CUDA
library must behave like the reference's functions called the same way (same return values, outputs and struct contents)."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import ref
from tests import util

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built")]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class Stream(C.Structure):  # hevc_stream_t (hevc_stream.h:556-569)
    _fields_ = [("nal", C.POINTER(C.c_int32)), ("vps", C.c_void_p), ("sps", C.c_void_p), ("pps", C.c_void_p), ("aud", C.c_void_p), ("sh", C.c_void_p),
                ("slice_data", C.c_void_p), ("sps_table", C.c_void_p * 32), ("pps_table", C.c_void_p * 256)]


@pytest.fixture(scope="module")
def compat(ctx):  # ctx: makes sure a device is there and the main library is loaded first
    L = C.CDLL(os.path.join(ROOT, "hevcbitstream_b200", "libhevcb200_compat.so"))
    L.hevc_new.restype = C.POINTER(Stream)
    L.hevc_free.argtypes = [C.POINTER(Stream)]
    L.find_nal_unit.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.nal_to_rbsp.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.c_void_p, C.POINTER(C.c_int)]
    L.rbsp_to_nal.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.c_void_p, C.POINTER(C.c_int)]
    L.read_hevc_nal_unit.argtypes = [C.POINTER(Stream), C.c_void_p, C.c_int]
    L.peek_hevc_nal_unit.argtypes = [C.POINTER(Stream), C.c_void_p, C.c_int]
    return L


def loop(find, buf, size):
    """the canonical reader loop (hevc_analyze.c:135-176) with a given find_nal_unit; returns every call's (rc, start, end)"""
    calls = []
    off = 0
    base = buf.ctypes.data
    s, e = C.c_int(0), C.c_int(0)
    while True:
        rc = find(base + off, size - off, C.byref(s), C.byref(e))
        calls.append((rc, s.value + off, e.value + off))
        if rc <= 0:
            break
        off += e.value
    return calls


def test_find_nal_unit_loop_matches_reference(compat):
    Lr = ref.lib()
    Lr.ref_find_nal_unit.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    rng = np.random.default_rng(3)
    for it in range(40):
        size = int(rng.integers(0, 20000))
        buf = util.adversarial(rng, size, it, density=[1.0, 0.2, 0.02][it % 3])
        want = loop(Lr.ref_find_nal_unit, buf, size)
        got = loop(compat.find_nal_unit, buf, size)
        assert got == want, f"case {it}: first difference at call {[a == b for a, b in zip(got, want)].index(False) if len(got) == len(want) else 'count'}"
    s = ref.gen_stream(seed=2, profile=1, n_slices=3000, payload_min=1, payload_max=400, zero_heavy_pct=20, extra_zero_pct=20, ps_period=50)
    size = s.size - ref.PAD
    assert loop(compat.find_nal_unit, s, size) == loop(Lr.ref_find_nal_unit, s, size)


def test_nal_to_rbsp_and_back(compat):
    rng = np.random.default_rng(4)
    alph = [np.array([0, 0, 0, 1, 2, 3, 3, 4, 200], np.uint8), np.array([0, 3, 0, 0, 5], np.uint8)]
    n_err = n_ok = 0
    for it in range(300):
        n = int(rng.integers(0, 120))
        a = alph[it % 2]
        nal = a[rng.integers(0, len(a), n)] if it % 3 else rng.integers(0, 256, n, dtype=np.uint8)
        rc, nsz, rb = ref.nal_to_rbsp(bytes(nal))
        src = util.padded(nal)
        dst = np.zeros(n + 16, np.uint8)
        ns, rs = C.c_int(n), C.c_int(n)
        got = compat.nal_to_rbsp(src.ctypes.data, C.byref(ns), dst.ctypes.data, C.byref(rs))
        assert got == rc, f"case {it}: rc {got} != {rc} for {bytes(nal).hex()}"
        if rc >= 0:
            n_ok += 1
            assert ns.value == nsz and rs.value == rc and bytes(dst[:rc]) == rb
            back_ref = ref.rbsp_to_nal(rb)
            out = np.zeros(len(rb) * 2 + 16, np.uint8)
            rsz, osz = C.c_int(len(rb)), C.c_int(out.size)
            got2 = compat.rbsp_to_nal(util.padded(np.frombuffer(rb, np.uint8)).ctypes.data, C.byref(rsz), out.ctypes.data, C.byref(osz))
            assert got2 == len(back_ref) and osz.value == got2 and bytes(out[:got2]) == back_ref
        else:
            n_err += 1
    assert n_ok > 50 and n_err > 50


def test_read_hevc_nal_unit_matches_reference(compat):
    """NAL by NAL through hevc_new / read_hevc_nal_unit: return value, h->nal and the struct the NAL wrote (hashed) equal the
    reference's; the parameter-set state is carried from call to call"""
    s = ref.gen_stream(seed=5, profile=1, n_slices=400, payload_min=1, payload_max=300, zero_heavy_pct=20, extra_zero_pct=10, ps_period=40,
                       unsupported_pct=5)
    size = s.size - ref.PAD
    st, en, _ = ref.scan_all_with_tail(s, size)
    want = ref.parse_all(s, st, en)["rec"]
    words = {1: ref.sizeof("vps") // 4, 2: ref.sizeof("sps") // 4, 3: ref.sizeof("pps") // 4, 4: ref.sizeof("sh") // 4}
    h = compat.hevc_new()
    assert h
    n_checked = 0
    for k in range(len(st)):
        nal = np.ascontiguousarray(s[st[k]: en[k]])
        t = compat.peek_hevc_nal_unit(h, nal.ctypes.data, nal.size)
        rc = compat.read_hevc_nal_unit(h, nal.ctypes.data, nal.size)
        assert rc == want["rc"][k], f"NAL {k}: rc {rc} != {want['rc'][k]}"
        if want["strip_rc"][k] < 0:
            continue
        nalv = np.ctypeslib.as_array(h.contents.nal, shape=(4,))
        assert (nalv[1], nalv[2], nalv[3]) == (want["nal_unit_type"][k], want["nal_layer_id"][k], want["nal_temporal_id_plus1"][k])
        assert t == (nalv[1] if 0 < nalv[1] <= 40 else -1)
        typ = int(nalv[1])
        kind = 4 if (typ <= 9 or 16 <= typ <= 21) else {32: 1, 33: 2, 34: 3}.get(typ, 0)
        if kind and want["state_hash"][k]:
            ptr = {1: h.contents.vps, 2: h.contents.sps, 3: h.contents.pps, 4: h.contents.sh}[kind]
            arr = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_int32)), shape=(words[kind],))
            assert ref.hash_ints(arr) == int(want["state_hash"][k]), f"NAL {k} (type {typ}): struct differs from the reference"
            n_checked += 1
    compat.hevc_free(h)
    assert n_checked > 400


def test_write_hevc_nal_unit_matches_reference(compat):
    """read every NAL with both libraries, edit slice_qp_delta / a VUI flag in both hevc_stream_t, write with both:
    return values and bytes must be equal (write_hevc_nal_unit from caller-owned, edited structs)"""
    Lr = ref.lib()
    Lr.ref_writer_struct.restype = C.c_void_p
    Lr.ref_writer_read.argtypes = [C.c_void_p, C.c_int]
    Lr.ref_writer_write.argtypes = [C.c_void_p, C.c_int]
    compat.write_hevc_nal_unit.argtypes = [C.POINTER(Stream), C.c_void_p, C.c_int]
    from tests import rewrite_check as rcx

    f_qp = rcx.field_index(rcx.KIND_SLICE, "slice_qp_delta")
    f_fr = rcx.field_index(rcx.KIND_SPS, "vui.video_full_range_flag")
    for profile in (0, 1):
        s = ref.gen_stream(seed=11 + profile, profile=profile, n_slices=250, payload_min=1, payload_max=200, zero_heavy_pct=20, ps_period=30,
                           unsupported_pct=5)
        size = s.size - ref.PAD
        st, en, _ = ref.scan_all_with_tail(s, size)
        Lr.ref_writer_reset()
        h = compat.hevc_new()
        n_written = 0
        for k in range(len(st)):
            nal = np.ascontiguousarray(s[st[k]: en[k]])
            r_ref = Lr.ref_writer_read(nal.ctypes.data, nal.size)
            r_c = compat.read_hevc_nal_unit(h, nal.ctypes.data, nal.size)
            assert r_ref == r_c
            if r_ref < 0:
                continue
            typ = int(np.ctypeslib.as_array(h.contents.nal, shape=(4,))[1])
            if typ <= 21:  # edit the slice header in both objects
                for ptr in (Lr.ref_writer_struct(3), h.contents.sh):
                    C.cast(ptr, C.POINTER(C.c_int32))[f_qp] += 3
            elif typ == 33:
                for ptr in (Lr.ref_writer_struct(1), h.contents.sps):
                    C.cast(ptr, C.POINTER(C.c_int32))[f_fr] ^= 1
            cap = nal.size * 2 + 64
            o_ref, o_c = np.zeros(cap + 16, np.uint8), np.zeros(cap + 16, np.uint8)
            w_ref = Lr.ref_writer_write(o_ref.ctypes.data, cap)
            w_c = compat.write_hevc_nal_unit(h, o_c.ctypes.data, cap)
            assert w_ref == w_c, f"profile {profile} NAL {k} type {typ}: write rc {w_c} != {w_ref}"
            if w_ref > 0:
                assert np.array_equal(o_ref[:w_ref], o_c[:w_ref]), f"profile {profile} NAL {k} type {typ}: written bytes differ"
                n_written += 1
            if typ == 33:  # keep both objects in the state a real tool would have: re-read what was written (App. A-1)
                Lr.ref_writer_read(o_ref.ctypes.data, w_ref)
                compat.read_hevc_nal_unit(h, o_c.ctypes.data, w_c)
        compat.hevc_free(h)
        assert n_written > 250


def test_analyze_tool_on_the_compat_api(tmp_path):
    """tools/hevcb_analyze.c (a reader in the shape of hevc_analyze.c, plain C against hevcb_compat.h): its '!! Found NAL' lines
    must be the reference CLI loop's, and its per-NAL summaries must agree with the reference parse"""
    import re
    import subprocess

    exe = str(tmp_path / "hevcb_analyze")
    libdir = os.path.join(ROOT, "hevcbitstream_b200")
    subprocess.check_call(["gcc", "-O2", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tools", "hevcb_analyze.c"), "-L" + libdir,
                           "-lhevcb200_compat", "-lhevcb200", "-Wl,-rpath," + libdir, "-o", exe])
    s = ref.gen_stream(seed=21, profile=1, n_slices=300, payload_min=1, payload_max=400, zero_heavy_pct=20, extra_zero_pct=20, ps_period=40,
                       unsupported_pct=5)
    size = s.size - ref.PAD
    path = str(tmp_path / "s.h265")
    s[:size].tofile(path)
    out = subprocess.run([exe, "-v", path], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    got = [l for l in out.stdout.splitlines() if l.startswith("!! Found NAL")]
    refpath = str(tmp_path / "ref.txt")
    assert ref.lib().ref_analyze_to_file(s.ctypes.data_as(C.c_void_p), C.c_int64(size), refpath.encode(), 1) == 0
    want = [l.rstrip("\n") for l in open(refpath, errors="replace") if l.startswith("!! Found NAL")]
    assert len(want) > 300 and got == want
    st, en, _ = ref.scan_all_with_tail(s, size)
    rec = ref.parse_all(s, st, en)["rec"]
    lines = [l for l in out.stdout.splitlines() if l.startswith("nal_unit_type")]
    assert len(lines) == len(st)
    for k, l in enumerate(lines):
        if rec["strip_rc"][k] >= 0:
            assert int(re.match(r"nal_unit_type (\d+)", l).group(1)) == rec["nal_unit_type"][k]
        assert ("not parsed" in l) == (rec["rc"][k] < 0)
        m = re.search(r"slice_data (-?\d+) bytes", l)
        if m:
            assert int(m.group(1)) == rec["slice_data_size"][k]
