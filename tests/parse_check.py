"""Shared parity helpers for the batched parser: compare an index produced by the CUDA path with the reference's
read_hevc_nal_unit loop (oracle/_ref) NAL by NAL: return code, h->nal, digest of the struct the NAL wrote, slice data."""
from __future__ import annotations

import numpy as np

from oracle import ref

STRUCT_WORDS = {1: ref.sizeof("vps") // 4 if ref.available() else 107034, 2: 19064, 3: 492, 4: 1006}
_M = np.uint64(0x9E3779B97F4A7C15)
_pow_cache = {}


def _powers(words: int) -> np.ndarray:
    """M^(words-1-i) mod 2^64 for i in 0..words-1"""
    if words not in _pow_cache:
        p = np.ones(words, np.uint64)
        with np.errstate(over="ignore"):
            for i in range(words - 2, -1, -1):
                p[i] = p[i + 1] * _M
        _pow_cache[words] = p
    return _pow_cache[words]


def digests_from_pairs(kind, pair_off, pair_field, pair_value):
    """Vectorised digest of every parsed NAL's struct (zero-fill + scatter, last write wins), equal to
    ref_hash_ints(0, struct) = sum((x_i + 1) * M^(W-1-i)) mod 2^64."""
    n = len(kind)
    out = np.zeros(n, np.uint64)
    cnt = np.diff(pair_off).astype(np.int64)
    nal_of_pair = np.repeat(np.arange(n, dtype=np.int64), cnt)
    with np.errstate(over="ignore"):
        for kd, words in STRUCT_WORDS.items():
            sel = kind == kd
            if not sel.any():
                continue
            pw = _powers(words)
            base = pw.sum(dtype=np.uint64)  # all-zero struct: every element contributes (0 + 1) * M^..
            pm = sel[nal_of_pair]
            nal = nal_of_pair[pm]
            fld = pair_field[: len(nal_of_pair)][pm].astype(np.int64)
            val = pair_value[: len(nal_of_pair)][pm]
            assert (fld < words).all()
            # last occurrence of every (nal, field)
            key = nal * np.int64(words) + fld
            order = np.argsort(key, kind="stable")
            ks = key[order]
            last = np.ones(len(ks), bool)
            last[:-1] = ks[1:] != ks[:-1]
            o = order[last]
            contrib = val[o].astype(np.int64).astype(np.uint64) & np.uint64(0xFFFFFFFF)
            contrib = contrib * pw[fld[o]]
            acc = np.zeros(n, np.uint64)
            np.add.at(acc, nal[o], contrib)
            out[sel] = acc[sel] + base
    return out


def compare_index(stream: np.ndarray, size: int, idx, tag: str = "", check_materialize: int = 300):
    """idx: object with numpy attributes nal_start, nal_end, rbsp_off, rbsp_end, rbsp, rc, nal_hdr, kind, hdr_end,
    pair_off, pair_field, pair_value (+ optional .materialize)."""
    st, en, r = ref.scan_all_with_tail(stream, size)
    n = len(st)
    assert len(idx.nal_start) == n, f"[{tag}] n_nals {len(idx.nal_start)} != {n}"
    assert np.array_equal(idx.nal_start, st) and np.array_equal(idx.nal_end, en), f"[{tag}] NAL offsets differ"
    rp = ref.parse_all(stream, st, en)
    R = rp["rec"]
    # return codes
    bad = np.nonzero(R["rc"] != idx.rc)[0]
    assert len(bad) == 0, f"[{tag}] rc differs for {len(bad)} NALs, first {bad[0]}: ref {R['rc'][bad[0]]} got {idx.rc[bad[0]]} type {R['nal_unit_type'][bad[0]]}"
    # h->nal after each call: the reference keeps the previous header when nal_to_rbsp fails
    hdr = idx.nal_hdr.astype(np.int64)
    ok = hdr != -1
    assert np.array_equal(ok, R["strip_rc"] >= 0), f"[{tag}] strip status differs"
    for name, shift in (("nal_unit_type", 0), ("nal_layer_id", 8), ("nal_temporal_id_plus1", 16)):
        got = (hdr >> shift) & 0xFF
        assert np.array_equal(got[ok], R[name][ok]), f"[{tag}] {name} differs"
    # struct digests
    dg = digests_from_pairs(idx.kind, idx.pair_off, idx.pair_field, idx.pair_value)
    want = R["state_hash"]
    has = idx.kind != 0
    assert np.array_equal(has, want != 0), f"[{tag}] set of NALs that wrote a struct differs"
    bad = np.nonzero(dg[has] != want[has])[0]
    if len(bad):
        k = int(np.nonzero(has)[0][bad[0]])
        raise AssertionError(f"[{tag}] struct digest differs for {len(bad)} NALs, first k={k} type={R['nal_unit_type'][k]} kind={idx.kind[k]}")
    # slice data extents and bytes
    sl = idx.kind == 4
    sd_off = idx.rbsp_off + idx.hdr_end.astype(np.int64) + 1
    sd_size = (idx.rbsp_end - sd_off).astype(np.int64)
    assert np.array_equal(sd_size[sl], R["slice_data_size"][sl].astype(np.int64)), f"[{tag}] slice data size differs"
    if idx.rbsp is not None:
        ks = np.nonzero(sl)[0]
        step = max(1, len(ks) // 2000)
        for k in ks[::step].tolist():
            if sd_size[k] > 0:
                h = ref.hash_bytes(idx.rbsp[sd_off[k]: sd_off[k] + sd_size[k]])
                assert h == int(R["slice_data_hash"][k]), f"[{tag}] slice data bytes differ at NAL {k}"
    # hevcb_materialize on a prefix of the stream (sequential semantics incl. stale h->nal)
    if hasattr(idx, "materialize") and check_materialize:
        m = min(n, check_materialize)
        nal = np.zeros(4, np.int32)
        bufs = {1: np.zeros(STRUCT_WORDS[1], np.int32), 2: np.zeros(STRUCT_WORDS[2], np.int32), 3: np.zeros(STRUCT_WORDS[3], np.int32),
                4: np.zeros(STRUCT_WORDS[4], np.int32)}
        for k in range(m):
            rc = idx.materialize(k, nal, bufs[1], bufs[2], bufs[3], bufs[4])
            assert rc == int(R["rc"][k])
            assert (nal[1], nal[2], nal[3]) == (R["nal_unit_type"][k], R["nal_layer_id"][k], R["nal_temporal_id_plus1"][k]), f"[{tag}] h->nal at {k}"
            kd = int(idx.kind[k])
            if kd:
                assert ref.hash_ints(bufs[kd]) == int(R["state_hash"][k]), f"[{tag}] materialised struct differs at NAL {k} kind {kd}"
    return n, int(rp["n_ok"])


def smoke_parse(ctx):
    """Tiny parse parity check used by __graft_entry__.smoke(): needs oracle/_ref for the stream and the answers."""
    if not ref.available():
        return
    s = ref.gen_stream(seed=3, profile=1, n_slices=400, payload_min=1, payload_max=200, zero_heavy_pct=20, extra_zero_pct=10, ps_period=23, unsupported_pct=5)
    size = s.size - ref.PAD
    idx = ctx.index_host(s[:size], size=size)
    n, ok = compare_index(s, size, idx, tag="smoke", check_materialize=100)
    print(f"parse smoke ok: {n} NALs, {ok} parsed, {int(idx.parse.n_pairs)} syntax elements")
