"""GPU parity tests of the batched header parser (hevcb_index_host / hevcb_parse_device through the C ABI) against the
reference's read_hevc_nal_unit loop: return codes, h->nal, every parsed struct (digest of the materialised
hevc_vps_t / hevc_sps_t / hevc_pps_t / hevc_slice_header_t), slice data extents and bytes."""
import numpy as np
import pytest

from oracle import ref
from tests import parse_check, util

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built")]


def test_config1_shape(ctx):
    s = ref.gen_stream(seed=0, profile=0, n_slices=3000, payload_min=900, payload_max=900, idr_period=100)
    size = s.size - ref.PAD
    idx = ctx.index_host(s[:size], size=size)
    n, ok = parse_check.compare_index(s, size, idx, tag="c1")
    assert n == 3003 and ok == 3003


@pytest.mark.parametrize("seed", [1, 2, 3, 4, 5, 6])
def test_rich_streams(ctx, seed):
    """multi-slice, tiles / WPP entry points, long-term refs, slice-local + inter RPS, pred weight tables, list
    modification, VUI + HRD, scaling lists, range extensions, re-sent parameter sets, unsupported NAL types,
    NALs truncated by trailing zero bytes (rc -1)"""
    s = ref.gen_stream(seed=seed, profile=1, n_slices=6000, payload_min=1, payload_max=64, zero_heavy_pct=20, extra_zero_pct=10,
                       ps_period=37, unsupported_pct=5)
    size = s.size - ref.PAD
    idx = ctx.index_host(s[:size], size=size)
    n, ok = parse_check.compare_index(s, size, idx, tag=f"rich{seed}")
    assert n > 6000 and 0 < ok < n


def test_strip_errors_do_not_touch_state(ctx):
    """corrupt some NALs so that nal_to_rbsp fails: rc -1, h->nal keeps the previous header, later NALs unaffected"""
    s = ref.gen_stream(seed=11, profile=1, n_slices=1500, payload_min=8, payload_max=64, ps_period=50)
    size = s.size - ref.PAD
    st, en, _ = ref.scan_all_with_tail(s, size)
    rng = np.random.default_rng(1)
    s = s.copy()
    for k in rng.choice(len(st), 60, replace=False):
        # only slice NALs: a parameter set that fails to strip would make later slices parse against older state,
        # which can drive the REFERENCE into its malloc(-1) crash (App. A-11)
        if en[k] - st[k] >= 12 and ((int(s[st[k]]) >> 1) & 0x3F) < 32:
            p = int(en[k]) - 6
            s[p - 1: p + 3] = [0x55, 0, 0, 2]  # 00 00 02 near the end of the NAL: nal_to_rbsp fails, nothing is parsed
    idx = ctx.index_host(s[:size], size=size)
    n, ok = parse_check.compare_index(s, size, idx, tag="striperr")
    assert (idx.nal_hdr == -1).sum() >= 30


def test_device_api_one_million_headers(ctx):
    """BASELINE config 3 shape: 1M header-bearing NALs (payload <= 64 B), device-resident entry points"""
    import torch

    unit = ref.gen_stream(seed=21, profile=1, n_slices=50000, payload_min=1, payload_max=64, zero_heavy_pct=10, extra_zero_pct=5,
                          ps_period=500, unsupported_pct=2)
    usz = unit.size - ref.PAD
    reps = 20
    stream = np.concatenate([np.tile(unit[:usz], reps), np.zeros(ref.PAD, np.uint8)])
    size = usz * reps
    d = torch.from_numpy(stream[:size].copy()).cuda()
    scan = ctx.scan_strip_device(d, size=size, cap_nals=size // 8)
    out = ctx.parse_device(d, scan)
    n = scan.n_nals
    assert n > 1_000_000

    class Idx:
        pass

    idx = Idx()
    for nm in ("nal_start", "nal_end", "rbsp_off", "rbsp_end"):
        setattr(idx, nm, getattr(scan, nm)[:n].cpu().numpy())
    idx.rbsp = scan.rbsp[: scan.rbsp_bytes].cpu().numpy()
    for nm in ("rc", "nal_hdr", "kind", "hdr_end"):
        setattr(idx, nm, out[nm][:n].cpu().numpy())
    idx.pair_off = out["pair_off"].cpu().numpy()
    npairs = out["n_pairs"]
    idx.pair_field = out["pair_field"][:npairs].cpu().numpy().view(np.uint32)
    idx.pair_value = out["pair_value"][:npairs].cpu().numpy()
    n2, ok = parse_check.compare_index(stream, size, idx, tag="1M")
    assert n2 == n and ok == out["n_ok"]
    # SoA columns agree with the pairs for slices
    cols = out["cols"][:, :n].cpu().numpy()
    sl = idx.kind == 4
    assert sl.sum() == out["n_slices"]
    assert set(np.unique(cols[0][sl]).tolist()) <= {0, 1, 2}


def test_cicc_o3_build_of_the_parser_matches_the_reference():
    """hevcb_parse.cu is shipped with the NVVM optimiser at -O1 (csrc/Makefile): in round 1 the default -O3 made a few NALs parse
    differently and the cause was never found.  With the current source the divergence no longer reproduces (tools/diag_cicc_o3.py:
    0 of ~44 000 NALs differ); this test keeps the hazard tracked: when the diagnosis build exists (`make -C hevcbitstream_b200/csrc
    o3`, ~9 minutes of cicc), the parity suite runs against it in a separate process."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib = os.path.join(root, "hevcbitstream_b200", "libhevcb200_cicc_o3.so")
    if not os.path.exists(lib):
        pytest.skip("diagnosis build libhevcb200_cicc_o3.so not present")
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "import hevcbitstream_b200 as hb\n"
        "from oracle import ref\n"
        "from tests import parse_check\n"
        "ctx = hb.Context(0)\n"
        "tot = 0\n"
        "for seed in (1, 2, 3, 4, 5, 6, 21):\n"
        "    s = ref.gen_stream(seed=seed, profile=1, n_slices=6000, payload_min=1, payload_max=64, zero_heavy_pct=20, extra_zero_pct=10, ps_period=37, unsupported_pct=5)\n"
        "    size = s.size - ref.PAD\n"
        "    idx = ctx.index_host(s[:size], size=size)\n"
        "    n, ok = parse_check.compare_index(s, size, idx, tag='o3-%%d' %% seed)\n"
        "    tot += n\n"
        "print('O3_PARITY_OK', tot)\n" % root
    )
    out = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, HEVCB_LIB=lib), capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and "O3_PARITY_OK" in out.stdout, out.stdout[-1500:] + out.stderr[-3000:]
