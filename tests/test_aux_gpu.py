"""GPU tests of the extension mode of the batched parser (HEVCB_PARSE_AUX): access unit delimiter, end of sequence /
bitstream, filler data and SEI NAL units.

The reference defines readers for them but never dispatches them (hevc_stream.in.c:499-573; HAVE_SEI is never defined), so
read_hevc_nal_unit returns -1: that is the default here too and is checked against the reference below.  In extension mode
the oracle is the composition of the reference's OWN exported functions over its own bs_t (oracle/_ref: nal_to_rbsp,
read_hevc_access_unit_delimiter_rbsp, read_filler_data_rbsp, _read_ff_coded_number, read_sei_payload, more_rbsp_data), driven
from here the way the dead code would drive them."""
import ctypes as C

import numpy as np
import pytest

from oracle import ref

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built")]


class BS(C.Structure):  # bs_t, bs.h:34-40
    _fields_ = [("start", C.c_void_p), ("p", C.c_void_p), ("end", C.c_void_p), ("bits_left", C.c_int)]


class SEI(C.Structure):  # sei_t, h264_sei.h:37-46
    _fields_ = [("payloadType", C.c_int), ("payloadSize", C.c_int), ("data", C.c_void_p)]


def ref_aux(nal: bytes):
    """What the reference's readers produce for one AUD / EOS / EOB / FD / SEI NAL: (rc, list of values)."""
    L = ref.lib()
    rc, nal_size, rbsp = ref.nal_to_rbsp(nal)
    if rc < 0:
        return -1, []
    buf = np.frombuffer(rbsp + b"\0" * 16, np.uint8).copy()
    b = BS(buf.ctypes.data, buf.ctypes.data + 2, buf.ctypes.data + len(rbsp), 8)  # behind the two header bytes
    t = (nal[0] >> 1) & 0x3F
    vals = []
    if t == 35:
        L.hevc_new.restype = C.c_void_p
        h = L.hevc_new()
        L.read_hevc_access_unit_delimiter_rbsp(C.c_void_p(h), C.byref(b))
        aud = C.cast(h + 4 * 8, C.POINTER(C.c_void_p))[0]  # h->aud: fifth pointer of hevc_stream_t
        vals = [("aud", C.cast(aud, C.POINTER(C.c_int))[0])]
    elif t == 38:
        before = b.p
        L.read_filler_data_rbsp(C.byref(b))
        n_ff = sum(1 for x in rbsp[2:] if x == 0xFF) if False else None
        # ff bytes = bytes consumed in front of the trailing bits byte
        vals = [("ff", (b.p - before) - 1 if b.bits_left == 8 else (b.p - before))]
    elif t in (39, 40):
        while True:
            L._read_ff_coded_number.restype = C.c_int
            pt = L._read_ff_coded_number(C.byref(b))
            ps = L._read_ff_coded_number(C.byref(b))
            off = b.p - b.start
            s = SEI(pt, ps, None)
            L.read_sei_payload(C.byref(s), C.byref(b))
            data = bytes(C.cast(s.data, C.POINTER(C.c_uint8))[i] for i in range(ps)) if ps > 0 else b""
            vals.append(("sei", pt, ps, off, data))
            if not L.more_rbsp_data(C.byref(b)) or len(vals) > 1000:
                break
        L.read_rbsp_trailing_bits(C.byref(b))
    overrun = b.p > b.end
    return (-1 if overrun else nal_size), vals


def build_stream(rng, n):
    """Annex-B stream of aux NALs between reference-written parameter sets and slices."""
    base = ref.gen_stream(seed=5, profile=1, n_slices=40, payload_min=1, payload_max=64, ps_period=10)
    bsize = base.size - ref.PAD
    st, en, _ = ref.scan_all_with_tail(base, bsize)
    parts, kinds = [], []
    for i in range(n):
        if i % 3 == 0 and i // 3 < len(st):
            k = i // 3
            parts.append(np.concatenate([np.array([0, 0, 0, 1], np.uint8), base[st[k]:en[k]]]))
            kinds.append(None)
            continue
        t = int(rng.choice([35, 36, 37, 38, 39, 40]))
        hdr = [t << 1, 1]
        if t == 35:
            body = [int(rng.integers(0, 8)) << 5 | 0x10]
        elif t in (36, 37):
            body = []
        elif t == 38:
            body = [0xFF] * int(rng.integers(0, 40)) + [0x80]
        else:
            body = []
            for _ in range(int(rng.integers(1, 4))):
                pt, ps = int(rng.integers(0, 700)), int(rng.integers(0, 600))
                body += [0xFF] * (pt // 255) + [pt % 255] + [0xFF] * (ps // 255) + [ps % 255]
                pay = rng.integers(0, 256, ps)
                pay[rng.random(ps) < 0.3] = 0
                body += pay.tolist()
            body += [0x80]
        nal = ref.rbsp_to_nal(bytes(hdr + body))
        parts.append(np.concatenate([np.array([0, 0, 1], np.uint8), np.frombuffer(nal, np.uint8)]))
        kinds.append(t)
    return np.concatenate(parts), kinds


@pytest.mark.parametrize("seed", [1, 2])
def test_extension_mode_matches_the_reference_functions(ctx, seed):
    import torch

    rng = np.random.default_rng(seed)
    s, kinds = build_stream(rng, 300)
    size = s.size
    d = torch.zeros(size + 32, dtype=torch.uint8, device="cuda")
    d[:size] = torch.from_numpy(s)
    scan = ctx.scan_strip_device(d, size=size)
    n = scan.n_nals
    assert n == len(kinds)
    st, en = scan.nal_start[:n].cpu().numpy(), scan.nal_end[:n].cpu().numpy()
    # default (compat) mode: exactly the reference's read_hevc_nal_unit results, i.e. -1 for every aux NAL
    pc = ctx.parse_device(d, scan)
    R = ref.parse_all(ref.padded(s), st, en)["rec"]
    assert np.array_equal(pc["rc"][:n].cpu().numpy(), R["rc"])
    aux = np.array([k is not None for k in kinds])
    assert (pc["rc"][:n].cpu().numpy()[aux] == -1).all() and (pc["kind"][:n].cpu().numpy()[aux] == 0).all()
    # extension mode
    px = ctx.parse_device(d, scan, aux=True)
    rc, kind = px["rc"][:n].cpu().numpy(), px["kind"][:n].cpu().numpy()
    po, pf, pv = px["pair_off"].cpu().numpy(), px["pair_field"].cpu().numpy(), px["pair_value"].cpu().numpy()
    img, ro = scan.rbsp.cpu().numpy(), scan.rbsp_off[:n].cpu().numpy()
    assert np.array_equal(rc[~aux], R["rc"][~aux]) and np.array_equal(kind[~aux], pc["kind"][:n].cpu().numpy()[~aux])
    seen = set()
    for k in np.nonzero(aux)[0].tolist():
        want_rc, vals = ref_aux(s[st[k]:en[k]].tobytes())
        assert rc[k] == want_rc, (k, kinds[k], rc[k], want_rc)
        assert kind[k] == 5
        f, v = pf[po[k]:po[k + 1]].tolist(), pv[po[k]:po[k + 1]].tolist()
        t = kinds[k]
        seen.add(t)
        if t == 35:
            assert f == [0] and v == [vals[0][1]]
        elif t in (36, 37):
            assert f == []
        elif t == 38:
            assert f == [8] and v == [vals[0][1]], (v, vals)
        else:
            assert f == [16, 17, 18] * len(vals)
            for i, (_, pt, ps, off, data) in enumerate(vals):
                assert v[3 * i: 3 * i + 3] == [pt, ps, off]
                assert img[ro[k] + off: ro[k] + off + ps].tobytes() == data
    assert seen == {35, 36, 37, 38, 39, 40}
