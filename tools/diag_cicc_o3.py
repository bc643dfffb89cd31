"""Diagnosis of the cicc -O3 divergence of hevcb_parse.cu (csrc/Makefile): parse the same generated streams with the shipped
library (cicc -O1) and with `make -C hevcbitstream_b200/csrc o3` (libhevcb200_cicc_o3.so), print every NAL whose results differ.

    python tools/diag_cicc_o3.py            # parent: runs itself twice (HEVCB_LIB), then diffs
"""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SEEDS = [1, 2, 3, 4, 5, 6, 21]


def child(out):
    import hevcbitstream_b200 as hb
    from oracle import ref

    ctx = hb.Context(0)
    res = {}
    for seed in SEEDS:
        s = ref.gen_stream(seed=seed, profile=1, n_slices=6000, payload_min=1, payload_max=64, zero_heavy_pct=20, extra_zero_pct=10, ps_period=37,
                           unsupported_pct=5)
        size = s.size - ref.PAD
        idx = ctx.index_host(s[:size], size=size)
        for nm in ("rc", "nal_hdr", "kind", "ubflag", "hdr_end", "pair_off", "pair_field", "pair_value", "nal_start"):
            res[f"{seed}_{nm}"] = np.asarray(getattr(idx, nm))
    np.savez(out, **res)


def main():
    if len(sys.argv) > 1:
        return child(sys.argv[1])
    libs = {"o1": os.path.join(ROOT, "hevcbitstream_b200", "libhevcb200.so"), "o3": os.path.join(ROOT, "hevcbitstream_b200", "libhevcb200_cicc_o3.so")}
    outs = {}
    for k, lib in libs.items():
        out = f"/tmp/diag_{k}.npz"
        subprocess.check_call([sys.executable, __file__, out], env=dict(os.environ, HEVCB_LIB=lib))
        outs[k] = np.load(out)
    import hevcbitstream_b200 as hb

    ctx = hb.Context(0)
    total = 0
    for seed in SEEDS:
        a = {nm: outs["o1"][f"{seed}_{nm}"] for nm in ("rc", "nal_hdr", "kind", "ubflag", "hdr_end", "pair_off", "pair_field", "pair_value", "nal_start")}
        b = {nm: outs["o3"][f"{seed}_{nm}"] for nm in a}
        n = len(a["rc"])
        for k in range(n):
            pa = (a["pair_field"][a["pair_off"][k]:a["pair_off"][k + 1]], a["pair_value"][a["pair_off"][k]:a["pair_off"][k + 1]])
            pb = (b["pair_field"][b["pair_off"][k]:b["pair_off"][k + 1]], b["pair_value"][b["pair_off"][k]:b["pair_off"][k + 1]])
            same = a["rc"][k] == b["rc"][k] and a["hdr_end"][k] == b["hdr_end"][k] and a["ubflag"][k] == b["ubflag"][k] and \
                len(pa[0]) == len(pb[0]) and np.array_equal(pa[0], pb[0]) and np.array_equal(pa[1], pb[1])
            if same:
                continue
            total += 1
            if total > 12:
                continue
            print(f"seed {seed} NAL {k} @{a['nal_start'][k]} type {a['nal_hdr'][k] & 0xFF} kind {a['kind'][k]}: rc {a['rc'][k]}/{b['rc'][k]} hdr_end {a['hdr_end'][k]}/{b['hdr_end'][k]} "
                  f"ubflag {a['ubflag'][k]}/{b['ubflag'][k]} pairs {len(pa[0])}/{len(pb[0])}")
            m = min(len(pa[0]), len(pb[0]))
            d = [i for i in range(m) if pa[0][i] != pb[0][i] or pa[1][i] != pb[1][i]]
            first = d[0] if d else m
            for i in range(max(0, first - 3), min(max(len(pa[0]), len(pb[0])), first + 6)):
                fa = (ctx.trace_name(int(a["kind"][k]), int(pa[0][i])), int(pa[1][i])) if i < len(pa[0]) else None
                fb = (ctx.trace_name(int(b["kind"][k]), int(pb[0][i])), int(pb[1][i])) if i < len(pb[0]) else None
                print(f"    [{i}] O1 {fa}   O3 {fb}" + ("   <--" if fa != fb else ""))
    print(f"NALs that differ between cicc -O1 and -O3: {total}")


if __name__ == "__main__":
    main()
