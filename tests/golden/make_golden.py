"""Generates tests/golden/byte_layer.json by RUNNING THE UNMODIFIED REFERENCE (oracle/_ref/libhevcref.so, built from
/root/reference by `make -C oracle ref`).  The reference ships no tests or vectors of its own, so
these are the pinned known answers for the byte layer.  Run from the repo root:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import ref  # noqa: E402

HEX = lambda b: bytes(b).hex()


def main():
    rng = np.random.default_rng(20261017)
    out = {"source": "oracle/_ref (unmodified leslie-wang/hevcbitstream) via tests/golden/make_golden.py", "find_nal_unit": [],
           "nal_to_rbsp": [], "rbsp_to_nal": [], "scan_loop": []}
    fixed = ["00 00 01 40 01 02 03 04 00 00 01 42 01", "00 00 00 01 40 01 02 03 04 00 00 00 01 42 01",
             "09 09 00 00 01 40 01 02 03 04 05 00 00 01 09 09", "00 00 01 40 01 02 00 00 00 00 00 01 09 09 09",
             "00 00 01 00 00 01 40 01", "00 00 01 40 01 02 03 04 05 06 07 08", "00 00 01 40 01 02 03 04 05 00 00 01",
             "00 00 01 40 01 02 03 04 00 00 01 07", "01 02 03 04 05 06 07 08", "01 02 03 04 05 06 00 00 01 0A",
             "01 02 03 04 05 00 00 01 09 0A", "00 00 01 40 00 00 03 01 05 00 00 01 07 07", "", "00", "00 00 01", "00 00 00 01"]
    alph = [np.array([0, 0, 0, 1, 1, 2, 3, 3, 4, 255], np.uint8), np.array([0, 0, 1, 3], np.uint8), np.array([0, 1], np.uint8)]
    cases = [bytes.fromhex(v.replace(" ", "")) for v in fixed]
    for i in range(120):
        a = alph[i % 3]
        cases.append(bytes(a[rng.integers(0, len(a), int(rng.integers(0, 40)))]))
    for c in cases:
        rc, s, e = ref.find_nal_unit(c)
        out["find_nal_unit"].append({"buf": HEX(c), "rc": rc, "start": s, "end": e})
        buf = ref.padded(c)
        r = ref.scan_all(buf, len(c))
        out["scan_loop"].append({"buf": HEX(c), "starts": r["starts"].tolist(), "ends": r["ends"].tolist(), "last_rc": r["last_rc"],
                                 "last_start": r["last_start"], "last_end": r["last_end"]})
    nals = ["40 01 00 00 03 01 05", "40 01 00 00 03 04 05", "40 01 00 00 03 03 05", "40 00 00 03 00 00 03 01", "40 00 00 00 05",
            "40 00 00 01 05", "40 00 00 02 05", "40 01 80 00 00 03", "00 00 03 01 00 00 03 01 00 00 03 01", "00 00 03 00 00 03 00 00 03", "40"]
    ncases = [bytes.fromhex(v.replace(" ", "")) for v in nals]
    for i in range(150):
        a = alph[i % 2]
        ncases.append(bytes([0x40]) + bytes(a[rng.integers(0, len(a), int(rng.integers(0, 30)))]))
    for c in ncases:
        rc, ns, rb = ref.nal_to_rbsp(c)
        out["nal_to_rbsp"].append({"nal": HEX(c), "rc": rc, "nal_size": ns if rc >= 0 else None, "rbsp": HEX(rb) if rc >= 0 else None})
    rbs = ["40 00 00 01 05", "40 00 00 03 05", "40 00 00 04 05", "40 00 00 00 00 00 07", "40 00 00 00 00 00 00 07",
           "40 00 00 00 00 00 00 02", "40 01 00 00", "40 01 00 00 00", ""]
    rcases = [bytes.fromhex(v.replace(" ", "")) for v in rbs]
    for i in range(150):
        a = alph[i % 3]
        rcases.append(bytes(a[rng.integers(0, len(a), int(rng.integers(0, 40)))]))
    for c in rcases:
        out["rbsp_to_nal"].append({"rbsp": HEX(c), "nal": HEX(ref.rbsp_to_nal(c))})
    # header-bearing stream written by the reference's own writer (BASELINE config 3 shape), used by bench.py
    hs = ref.gen_stream(seed=2026, profile=1, n_slices=4000, payload_min=1, payload_max=48, zero_heavy_pct=10, extra_zero_pct=5,
                        ps_period=400, unsupported_pct=2)
    hpath = os.path.join(os.path.dirname(os.path.abspath(__file__)), "headers_unit.bin")
    hs[: hs.size - ref.PAD].tofile(hpath)
    print("wrote", hpath, hs.size - ref.PAD, "bytes")
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "byte_layer.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=0)
    print("wrote", path, {k: len(v) for k, v in out.items() if isinstance(v, list)})


if __name__ == "__main__":
    main()
