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


def test_nal_to_rbsp_exhaustive_short_nals(compat):
    """every NAL of up to 5 bytes over the alphabet {0, 1, 2, 3, 4}: return value, consumed size and bytes as the reference
    (the two all-zero NALs {00} and {00 00} are not a NAL behind a start code and are answered without the scanner)"""
    import itertools

    n_cases = 0
    for n in range(1, 6):
        for t in itertools.product(range(5), repeat=n):
            nal = np.array(t, np.uint8)
            rc, nsz, rb = ref.nal_to_rbsp(bytes(nal))
            src = util.padded(nal)
            dst = np.zeros(n + 16, np.uint8)
            ns, rs = C.c_int(n), C.c_int(n)
            got = compat.nal_to_rbsp(src.ctypes.data, C.byref(ns), dst.ctypes.data, C.byref(rs))
            assert got == rc, f"rc {got} != {rc} for {bytes(nal).hex()}"
            if rc >= 0:
                assert ns.value == nsz and rs.value == rc and bytes(dst[:rc]) == rb, bytes(nal).hex()
            n_cases += 1
    assert n_cases == 5 + 25 + 125 + 625 + 3125


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


def _analyze_streams():
    """(name, stream bytes) pairs for the CLI comparisons: the BASELINE config-1 shape cut to the reference's 32 MiB window, rich
    generator streams (VUI / HRD / scaling lists / RPS / pred-weight / unsupported types / nal_to_rbsp failures), and the two ways a
    stream makes the reference's loop end on return code 0 (its "last NAL" of size 0 is then dumped, hevc_analyze.c:190-205)."""
    out = []
    s = ref.gen_stream(seed=0, profile=0, n_slices=1200, payload_min=6680, payload_max=6680, idr_period=100)
    out.append(("c1", s[: s.size - ref.PAD]))
    for seed in (21, 22):
        s = ref.gen_stream(seed=seed, profile=1, n_slices=400, payload_min=1, payload_max=400, zero_heavy_pct=20, extra_zero_pct=20, ps_period=40,
                           unsupported_pct=5)
        out.append((f"rich{seed}", s[: s.size - ref.PAD]))
    s = ref.gen_stream(seed=23, profile=1, n_slices=60, payload_min=1, payload_max=200, ps_period=20)
    body = s[: s.size - ref.PAD]
    # (a stream that ends in 00 00 00 takes the same path but makes the reference itself crash: malloc(-1) + memcpy, App. A-11)
    out.append(("tail_startcode", np.concatenate([body, np.array([0, 0, 1], np.uint8)])))
    return out


def _run_cli(exe, path, args, tmp_path, tag):
    import subprocess

    ofile = str(tmp_path / f"{tag}.dbg")
    argv = [exe] + [a if a != "@O" else ofile for a in args] + [path]
    r = subprocess.run(argv, capture_output=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    dbg = open(ofile, "rb").read() if "@O" in args else b""
    return r.stdout, dbg


@pytest.mark.parametrize("args", [[], ["-v", "0"], ["-o", "@O"]])
def test_reference_cli_unmodified_on_the_compat_library(tmp_path, args):
    """oracle/_ref/hevc_analyze_compat is the reference's hevc_analyze.c, UNMODIFIED, compiled against include/compat/ (this
    repo's bs.h / h264_stream.h / hevc_stream.h) and linked with libhevcb200_compat instead of the reference's library
    (oracle/Makefile): find_nal_unit, read_debug_hevc_nal_unit and debug_bytes are then served by the CUDA library.  Its stdout
    and its -o file must equal the reference binary's byte for byte."""
    exe_ref, exe_b200 = ref.ANALYZE_BIN, os.path.join(os.path.dirname(ref.ANALYZE_BIN), "hevc_analyze_compat")
    if not os.path.exists(exe_b200):
        pytest.skip("oracle/_ref/hevc_analyze_compat not built")
    for name, data in _analyze_streams():
        path = str(tmp_path / f"{name}.h265")
        data.tofile(path)
        want = _run_cli(exe_ref, path, args, tmp_path, "ref_" + name)
        got = _run_cli(exe_b200, path, args, tmp_path, "b200_" + name)
        assert len(want[0]) > 1000
        assert got[0] == want[0], f"{name} {args}: stdout differs at byte {next(i for i, (x, y) in enumerate(zip(got[0] + b'~', want[0] + b'~')) if x != y)}"
        assert got[1] == want[1], f"{name} {args}: -o file differs"


@pytest.mark.parametrize("args", [[], ["-v", "0"], ["-o", "@O"]])
def test_batched_analyze_tool_prints_the_reference_dump(tmp_path, args):
    """tools/hevcb_analyze.c: hevc_analyze's output from ONE batched call (hevcb_index_host in its trace variant + host formatting)"""
    import subprocess

    exe = str(tmp_path / "hevcb_analyze")
    libdir = os.path.join(ROOT, "hevcbitstream_b200")
    subprocess.check_call(["gcc", "-O2", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tools", "hevcb_analyze.c"), "-L" + libdir,
                           "-lhevcb200", "-Wl,-rpath," + libdir, "-o", exe])
    for name, data in _analyze_streams():
        path = str(tmp_path / f"{name}.h265")
        data.tofile(path)
        want = _run_cli(ref.ANALYZE_BIN, path, args, tmp_path, "ref_" + name)
        got = _run_cli(exe, path, args, tmp_path, "tool_" + name)
        assert got[0] == want[0], f"{name} {args}: stdout differs at byte {next(i for i, (x, y) in enumerate(zip(got[0] + b'~', want[0] + b'~')) if x != y)}"
        assert got[1] == want[1], f"{name} {args}: -o file differs"
