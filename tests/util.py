"""Shared helpers for the parity tests: stream builders and result comparison against the oracle."""
from __future__ import annotations

import numpy as np

from oracle import ref

PAD = ref.PAD

ALPHABETS = [
    np.array([0, 0, 0, 1, 1, 2, 3, 3, 4, 255], np.uint8),
    np.array([0, 0, 1, 3], np.uint8),
    np.array([0, 1], np.uint8),
    np.array([0, 0, 0, 0, 1, 2, 3, 9], np.uint8),
]


def padded(arr: np.ndarray) -> np.ndarray:
    out = np.zeros(arr.size + PAD, np.uint8)
    out[: arr.size] = arr
    return out


def adversarial(rng, size: int, alphabet: int, density: float = 1.0) -> np.ndarray:
    """`size` bytes rich in 00/01/02/03 (padded)."""
    a = ALPHABETS[alphabet % len(ALPHABETS)]
    if density >= 1.0:
        x = a[rng.integers(0, len(a), size)]
    else:
        x = rng.integers(0, 256, size).astype(np.uint8)
        m = rng.random(size) < density
        x[m] = a[rng.integers(0, len(a), int(m.sum()))]
    return padded(x.astype(np.uint8))


def c2_stream(nal_size: int, total: int, seed: int = 1, dense: bool = False, sc4_every: int = 0) -> np.ndarray:
    """BASELINE config-2 style stream: NALs of ~nal_size bytes (start code + 2-byte TRAIL_R header + escaped
    random payload + 0x80), or the EPB-dense worst case (payload 00 00 03 01 repeated).  Padded."""
    rng = np.random.default_rng(seed)
    n = max(1, total // nal_size)
    body = max(4, nal_size - 3)  # bytes after the 3-byte start code
    if dense:
        reps = max(1, (body - 3) // 4)
        nal = np.concatenate([np.array([0, 0, 1, 0x02, 0x01], np.uint8), np.tile(np.array([0, 0, 3, 1], np.uint8), reps), np.array([0x80], np.uint8)])
        out = np.tile(nal, n)
        return padded(out)
    # random RBSP payloads, escaped with the reference's rbsp_to_nal in one batched call
    pay = body - 3
    rbsp = rng.integers(0, 256, n * (pay + 3), dtype=np.uint8).reshape(n, pay + 3)
    rbsp[:, 0] = 0x02
    rbsp[:, 1] = 0x01
    rbsp[:, -1] = 0x80
    flat = rbsp.reshape(-1)
    off = np.arange(n, dtype=np.int64) * (pay + 3)
    r = ref.insert_all(flat, off, off + pay + 3, sc_len=3)
    return padded(r["out"])


def compare_scan(buf: np.ndarray, size: int, res, image: np.ndarray | None, tag: str = "", check_bytes: bool = True):
    """res: object with n_nals, n_terminated, last_rc, last_start, last_end and arrays nal_start, nal_end,
    rbsp_off, rbsp_end (numpy, >= n_nals entries).  image: EPB-free image (numpy) or None."""
    st, en, r = ref.scan_all_with_tail(buf, size)
    ctxmsg = f"[{tag}] size={size}"
    assert res.n_terminated == r["n"], f"{ctxmsg}: terminated NALs {res.n_terminated} != ref {r['n']}"
    assert res.last_rc == r["last_rc"], f"{ctxmsg}: last_rc {res.last_rc} != ref {r['last_rc']}"
    assert res.n_nals == len(st), f"{ctxmsg}: n_nals {res.n_nals} != ref {len(st)}"
    assert res.last_start == r["last_start"] and res.last_end == r["last_end"], (
        f"{ctxmsg}: last ({res.last_start},{res.last_end}) != ref ({r['last_start']},{r['last_end']})")
    n = len(st)
    ns = np.asarray(res.nal_start[:n])
    ne = np.asarray(res.nal_end[:n])
    if not np.array_equal(ns, st):
        k = int(np.nonzero(ns != st)[0][0])
        raise AssertionError(f"{ctxmsg}: nal_start[{k}] {ns[k]} != ref {st[k]}")
    if not np.array_equal(ne, en):
        k = int(np.nonzero(ne != en)[0][0])
        raise AssertionError(f"{ctxmsg}: nal_end[{k}] {ne[k]} != ref {en[k]}")
    if n == 0:
        return 0
    sr = ref.strip_all(buf, st, en)
    ro = np.asarray(res.rbsp_off[:n])
    re = np.asarray(res.rbsp_end[:n])
    ref_rc = sr["rc"].astype(np.int64)
    bad = ref_rc < 0
    if not np.array_equal(re == -1, bad):
        k = int(np.nonzero((re == -1) != bad)[0][0])
        raise AssertionError(f"{ctxmsg}: strip status NAL {k}: rbsp_end {re[k]} vs ref rc {ref_rc[k]}")
    good = ~bad
    sizes = re - ro
    if not np.array_equal(sizes[good], ref_rc[good]):
        k = int(np.nonzero(good & (sizes != ref_rc))[0][0])
        raise AssertionError(f"{ctxmsg}: rbsp size NAL {k}: {sizes[k]} vs ref {ref_rc[k]}")
    if image is not None and check_bytes:
        # gather the per-NAL RBSPs out of the EPB-free image and compare with the reference's dense output
        lens = ref_rc[good]
        total = int(lens.sum())
        if total:
            starts_img = ro[good]
            idx = np.repeat(starts_img - np.concatenate([[0], np.cumsum(lens)[:-1]]), lens) + np.arange(total)
            got = np.asarray(image)[idx]
            want = sr["rbsp"][:total]
            if not np.array_equal(got, want):
                j = int(np.nonzero(got != want)[0][0])
                k = int(np.searchsorted(np.cumsum(lens), j, side="right"))
                raise AssertionError(f"{ctxmsg}: RBSP byte mismatch in good-NAL #{k} at stream byte {j}")
    return n
