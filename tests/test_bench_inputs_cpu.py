"""bench.py builds its inputs with numpy only; check that generator against the oracle (CPU)."""
import numpy as np

import bench
from oracle import port


def test_escape_matches_rbsp_to_nal():
    rng = np.random.default_rng(3)
    alph = np.array([0, 0, 0, 1, 2, 3, 4, 200], np.uint8)
    for it in range(400):
        n = int(rng.integers(0, 60))
        r = alph[rng.integers(0, len(alph), n)]
        assert bench.escape_rbsp(r).tobytes() == port.rbsp_to_nal(r.tobytes()), r.tobytes().hex()


def test_units_are_valid_streams():
    for name, nal, dense in bench.WORKLOADS:
        u = bench.make_unit(nal, 2 << 20, 11, dense)
        buf = port.padded(u)
        st, en, r = port.scan_all_with_tail(buf, u.size)
        assert r["last_rc"] == -1
        expect = max(1, (2 << 20) // (nal if not dense else (5 + 4 * max(1, (nal - 6) // 4) + 1)))
        assert len(st) == expect, (name, len(st), expect)
        sr = port.strip_all(buf, st, en)
        assert (sr["rc"] >= 0).all(), name
        if dense:
            assert sr["rc"].sum() < 0.8 * u.size
