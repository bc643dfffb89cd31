#!/usr/bin/env python
"""bench.py -- headline benchmark of the bitstream hot path (BASELINE.json `metric`, config[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--size-gib G] [--nal-size B] [--no-sweep]

One "step" = one pass of the fused start-code scan + EPB strip (hevcb_scan_strip_device) over a synthetic Annex-B
buffer that is already resident in HBM.  N = 1 runs BASELINE config[1]: a 4 GiB buffer; the headline `value` is quoted
on 16 KiB NALs (escaped uniform-random payload) and `sweep` carries the other NAL sizes (64 B .. 1 MiB) plus the
EPB-dense worst case.  N > 1 (torchrun, one rank per GPU): the ranks' 4 GiB buffers are consecutive BYTE RANGES of one
N x 4 GiB stream (weak scaling, BASELINE config[4]: 16-byte halo, NCCL all_gather of the ~190-byte shard records, stitch,
patch of the boundary entries inside every step).  Before timing, a reference-written stream goes through the sharded
scan + strip + header parse over the real process group and every rank checks its share against the reference
(`sharded_parity`; the run fails otherwise).

Printed (rank 0, one JSON line): metric/value/unit + `roofline` (algorithmic bytes / CUDA-event time of the step vs the
measured HBM peak), `cpu_baseline` (the UNMODIFIED reference's find_nal_unit + nal_to_rbsp loop, oracle/_ref, one host
thread, bounded sample), `e2e` (the same pass through the host-pointer C ABI with pinned buffers, H2D + D2H inside the
timed region), `clocks`, `gpu_launches`.

`--impl reference` times the reference's own CPU implementation on the same workload shape (bounded sample per step).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "annexb_scan_epb_strip_input_GBps"
UNIT = "GB/s"
UNIT_BYTES = 64 << 20  # synthetic stream unit that is tiled to the full buffer on the device


# ------------------------------------------------------------------------------------------------
# synthetic Annex-B generator (numpy only -- the oracle is not used to build inputs)
# ------------------------------------------------------------------------------------------------

def escape_rbsp(rbsp: np.ndarray) -> np.ndarray:
    """Emulation prevention as rbsp_to_nal does it (h264_nal.c:92-132): insert 03 before a byte <= 3 that follows two
    zeros, restarting the zero count after each insertion.  Candidates are found vectorised, then walked in order."""
    z = rbsp == 0
    cand = np.nonzero(z[:-2] & z[1:-1] & (rbsp[2:] <= 3))[0] + 2  # positions whose two predecessors are zero
    if cand.size == 0:
        return rbsp
    ins = []
    last_ins = -10
    for p in cand.tolist():
        # zero count at p: zeros immediately before p, counted since the last insertion point
        c = 0
        q = p - 1
        while q >= 0 and rbsp[q] == 0 and q >= last_ins and c < 2:
            c += 1
            q -= 1
        if c == 2:
            ins.append(p)
            last_ins = p
    return np.insert(rbsp, ins, 3) if ins else rbsp


def make_unit(nal_size: int, total: int, seed: int, dense: bool) -> np.ndarray:
    """`total` bytes (approximately) of NALs: 00 00 01 + 2-byte TRAIL_R header + escaped payload + 0x80."""
    rng = np.random.default_rng(seed)
    body = max(4, nal_size - 3)
    if dense:
        reps = max(1, (body - 3) // 4)
        nal = np.concatenate([np.array([0, 0, 1, 0x02, 0x01], np.uint8), np.tile(np.array([0, 0, 3, 1], np.uint8), reps), np.array([0x80], np.uint8)])
        return np.tile(nal, max(1, total // nal.size))
    n = max(1, total // nal_size)
    pay = body - 3
    rb = rng.integers(0, 256, (n, pay + 6), dtype=np.uint8)
    rb[:, 0] = 0
    rb[:, 1] = 0
    rb[:, 2] = 1
    rb[:, 3] = 0x02
    rb[:, 4] = 0x01
    rb[:, -1] = 0x80
    # escape the payload region of every NAL; start codes must stay intact, so escape NAL by NAL only where needed
    flat = rb.reshape(-1)
    z = flat == 0
    hit = np.nonzero(z[:-2] & z[1:-1] & (flat[2:] <= 3))[0] + 2
    hit = hit[(hit % (pay + 6)) >= 5]  # start codes (columns 0..2) are not payload
    rows = np.unique(hit // (pay + 6)).tolist()
    if not rows:
        return flat
    out = []
    prev = 0
    for r in rows:
        out.append(flat[prev * (pay + 6): r * (pay + 6)])
        row = rb[r]
        out.append(np.concatenate([row[:3], escape_rbsp(row[3:])]))
        prev = r + 1
    out.append(flat[prev * (pay + 6):])
    return np.concatenate(out)


WORKLOADS = [("nal64", 64, False), ("nal256", 256, False), ("nal1k", 1024, False), ("nal4k", 4096, False), ("nal16k", 16384, False),
             ("nal64k", 65536, False), ("nal256k", 262144, False), ("nal1m", 1 << 20, False), ("epb_dense_4k", 4096, True)]


# ------------------------------------------------------------------------------------------------
# helpers
# ------------------------------------------------------------------------------------------------

def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    def __init__(self, device_index: int):
        self.dev = device_index
        self.samples = []
        self.reasons = set()
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.dev)],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append((float(out[0]), float(out[1])))
                for nme, v in zip(names, out[2:]):
                    if "Active" in v and "Not" not in v:
                        self.reasons.add(nme)
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": sorted(self.reasons)}
        sm = sorted(s[0] for s in self.samples)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.samples[0][1], "reasons": sorted(self.reasons), "samples": len(sm)}


def cpu_reference_sample(unit: np.ndarray, reps: int = 3):
    """find_nal_unit + nal_to_rbsp loop of the UNMODIFIED reference (oracle/_ref) on one host thread."""
    from oracle import ref

    size = unit.size
    buf = ref.padded(unit)
    t, n = ref.time_loop(buf, size, 1, reps)
    return size / t / 1e9, n, t


def sharded_parity_check(ctx, dev, world, rank):
    """N > 1: a reference-written stream through the byte-range sharded scan + strip + header parse over the REAL process group
    (NCCL): every rank checks the NALs it owns against the unmodified reference (oracle/_ref, the checker only): offsets,
    nal_to_rbsp status / size, RBSP bytes, read_hevc_nal_unit return code, h->nal and the digest of every parsed struct.
    Returns a short description; raises when any rank disagrees."""
    import torch
    import torch.distributed as dist

    from hevcbitstream_b200 import shard as hs
    from oracle import ref

    if not ref.available():
        return "skipped: oracle/_ref not built on this box"
    from tests import parse_check

    s = ref.gen_stream(seed=12, profile=1, n_slices=20000, payload_min=1, payload_max=1500, zero_heavy_pct=20, extra_zero_pct=10, ps_period=500,
                       unsupported_pct=3)
    size = s.size - ref.PAD
    st, en, _ = ref.scan_all_with_tail(s, size)
    sr = ref.strip_all(s, st, en)
    R = ref.parse_all(s, st, en)["rec"]
    bounds = hs.plan_shards(s, world, size)
    own, halo, first, last = hs.shard_flags(bounds, rank)
    lo = int(bounds[rank])
    b = torch.zeros(own + halo + 32, dtype=torch.uint8, device=dev)
    b[: own + halo] = torch.from_numpy(s[lo: lo + own + halo].copy()).to(dev)
    err = ""
    try:
        sc, res = hs.scan_strip_sharded(ctx, b, own, halo, first, last, extra_rbsp=hs.HEAD_BYTES)
        ps = hs.parse_sharded(ctx, b, own, halo, sc, res)
        f, n, g = int(res.first_local[rank]), int(res.n_owned[rank]), int(res.nal_base[rank])
        bb = int(res.byte_base[rank])
        assert int(res.glob.n_nals) == len(st), "global NAL count"
        assert np.array_equal(sc.nal_start[f:f + n].cpu().numpy() + bb, st[g:g + n]), "nal_start"
        assert np.array_equal(sc.nal_end[f:f + n].cpu().numpy() + bb, en[g:g + n]), "nal_end"
        ro, re_ = sc.rbsp_off[f:f + n].cpu().numpy(), sc.rbsp_end[f:f + n].cpu().numpy()
        rc = sr["rc"][g:g + n].astype(np.int64)
        assert np.array_equal(re_ < 0, rc < 0), "nal_to_rbsp status"
        assert np.array_equal((re_ - ro)[rc >= 0], rc[rc >= 0]), "RBSP sizes"
        img = sc.rbsp.cpu().numpy()
        local_bytes = int(sc.record.rbsp_bytes)
        for k in np.nonzero((rc >= 0) & (re_ <= local_bytes))[0][:: max(1, n // 4000)].tolist():  # NALs whose RBSP lies in the local image
            a0 = int(sr["rbsp_off"][g + k])
            assert np.array_equal(img[ro[k]: re_[k]], sr["rbsp"][a0: a0 + rc[k]]), f"RBSP bytes of NAL {g + k}"
        prc, hdr, kind = ps["rc"][:n].cpu().numpy(), ps["nal_hdr"][:n].cpu().numpy().astype(np.int64), ps["kind"][:n].cpu().numpy()
        assert np.array_equal(prc, R["rc"][g:g + n]), "read_hevc_nal_unit return codes"
        okh = hdr != -1
        assert np.array_equal((hdr & 0xFF)[okh], R["nal_unit_type"][g:g + n][okh]), "nal_unit_type"
        npairs = int(ps["n_pairs"])
        dg = parse_check.digests_from_pairs(kind, ps["pair_off"][: n + 1].cpu().numpy(), ps["pair_field"][:npairs].cpu().numpy().view(np.uint32),
                                            ps["pair_value"][:npairs].cpu().numpy())
        has = kind != 0
        assert np.array_equal(dg[has], R["state_hash"][g:g + n][has]), "parsed struct digests"
    except Exception as ex:  # every rank must reach the collective below
        err = f"rank {rank}: {type(ex).__name__}: {ex}"
    flag = torch.tensor([0.0 if err else 1.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if float(flag.item()) != 1.0:
        raise SystemExit(f"sharded parity check FAILED ({err or 'on another rank'})")
    return f"ok: {len(st)} NALs of a reference-written stream over {world} ranks (NCCL), offsets + RBSP + parsed structs vs oracle/_ref on every rank"


def cpu_reference_extras():
    """the reference's CLI and its edit loop on the host, on BASELINE config[0]'s stream (64 MB, Main
    1920x1080, VPS/SPS/PPS + 10k slices written by the reference's writer).  One thread; bounded (a few seconds)."""
    import tempfile

    from oracle import ref

    out = {}
    s = ref.gen_stream(seed=0, profile=0, n_slices=10000, payload_min=6680, payload_max=6680, idr_period=100)
    size = s.size - ref.PAD
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "c1.h265")
        s[:size].tofile(path)
        best = None
        for _ in range(2):
            t0 = time.perf_counter()
            with open(os.devnull, "wb") as dn:
                subprocess.run([ref.ANALYZE_BIN, path], stdout=dn, stderr=dn, check=False)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        out["hevc_analyze_c1"] = {"seconds": best, "MBps": size / best / 1e6, "bytes": int(size),
                                  "what": "wall time of `hevc_analyze c1.h265 > /dev/null` (oracle/_ref/hevc_analyze, -O2), best of 2"}
    t, n = ref.time_loop(s, size, 2, 3)
    out["read_loop_c1"] = {"seconds": t, "GBps": size / t / 1e9, "nal_per_s": n / t, "what": "in-memory find_nal_unit + read_hevc_nal_unit loop, best of 3"}
    st, en, _ = ref.scan_all_with_tail(s, size)
    t0 = time.perf_counter()
    rw = ref.rewrite_all(s, size, st, en, qp_delta_add=2, vui_flip=1)
    dt = time.perf_counter() - t0
    out["rewrite_c4_composition"] = {"seconds": dt, "GBps": size / dt / 1e9, "bytes_out": int(rw["out"].size),
                                     "what": "read_hevc_nal_unit -> edit slice_qp_delta / VUI flag -> write_hevc_nal_unit -> rbsp_to_nal over the same 64 MB "
                                             "(SURVEY 3.4 composition, oracle/ref_harness.c), one pass"}
    return out


# ------------------------------------------------------------------------------------------------
# arms
# ------------------------------------------------------------------------------------------------

def workload_config(args, name, nal_size, size_bytes, world):
    """`config` of the JSON line: identical in both arms (the reference arm walks the same bytes on the host)."""
    return {"workload": f"BASELINE config[1]: fused start-code scan + EPB strip, {args.size_gib:g} GiB Annex-B buffer per GPU, NAL size {nal_size} B "
                        f"({name}); 64 MiB synthetic unit (escaped uniform-random payload) tiled to the full size; input > L2 so no flush needed",
            "nal_size": nal_size, "bytes_per_gpu": int(size_bytes), "l2": "inputs larger than L2 (4 GiB vs 126 MB)",
            "sharding": ("one stream cut by byte range, one shard per rank: 16-byte halo, NCCL all_gather of the shard records, stitch "
                         "(BASELINE config[4]); every step includes the exchange") if world > 1 else "single GPU"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import ref

    if not ref.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libhevcref.so missing (reference sources not compiled)"}))
        return
    name, nal_size, dense = next(w for w in WORKLOADS if w[0] == args.workload)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    unit = make_unit(nal_size, UNIT_BYTES, 1234, dense)  # rank 0's unit of the GPU arm
    reps = max(1, int(args.size_gib * (1 << 30)) // unit.size)
    size_total = unit.size * reps
    # One step = the reference's find_nal_unit + nal_to_rbsp loop over the same bytes the GPU arm holds: the buffer is the unit
    # tiled `reps` times and a unit begins with a start code, so the loop over one unit, run `reps` times, visits exactly the NALs
    # of the whole buffer (the reference's API takes `int` sizes: pieces < 2 GiB cut at NAL boundaries).  --ref-passes
    # bounds the passes per step (default: all of them); the throughput is per byte either way.
    passes = reps if args.ref_passes <= 0 else min(reps, args.ref_passes)
    buf = ref.padded(unit)
    size = unit.size

    def step():
        t = 0.0
        n = 0
        for _ in range(passes):
            dt, k = ref.time_loop(buf, size, 1, 1)
            t += dt
            n += k
        return t, n

    for _ in range(args.warmup):
        ref.time_loop(buf, size, 1, 1)
    ts = []
    nn = 0
    for _ in range(args.steps):
        t, nn = step()
        ts.append(t)
    mean = float(np.mean(ts))
    val = size * passes / mean / 1e9
    cfg = workload_config(args, name, nal_size, size_total, world)
    sample = (f"{passes} of {reps} passes over the 64 MiB unit per step (= {size * passes / 2**30:.2f} GiB of the {size_total / 2**30:.2f} GiB buffer), "
              f"find_nal_unit + nal_to_rbsp loop of the unmodified reference (oracle/_ref), 1 thread (the reference is single-threaded), {nn} NALs per step")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": mean * 1e3 * (reps / passes), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": 1, "kind": "reference", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def run_ours(args):
    import torch
    import torch.distributed as dist

    import hevcbitstream_b200 as hb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # stdout carries the JSON line only: libraries that print to fd 1 (NCCL's version banner) are sent to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ctx = hb.Context(local)
    size_target = int(args.size_gib * (1 << 30))
    peak, peak_src = measured_peak()
    sharded_parity = sharded_parity_check(ctx, dev, world, rank) if world > 1 else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def time_workload(name, nal_size, dense, steps, warmup, sampler=None):
        unit = make_unit(nal_size, UNIT_BYTES, 1234 + rank * 7919, dense)
        reps = max(1, size_target // unit.size)
        ut = torch.from_numpy(unit).to(dev)
        size = ut.numel() * reps
        d = torch.zeros(size + 32, dtype=torch.uint8, device=dev)[: size + 16]  # + room for the halo of the next shard
        d[:size].view(reps, -1).copy_(ut.unsqueeze(0).expand(reps, -1))
        cap = size // max(16, (nal_size if not dense else 4096) // 2) + (1 << 16)
        stitched = None
        if world == 1:
            outs = ctx.scan_strip_device(d, size=size, cap_nals=cap, want_rbsp=True, sync=False)

            def step():
                ctx.scan_strip_device(d, size=size, cap_nals=cap, out=outs, sync=False)
        else:
            # BASELINE config[4]: the ranks' buffers are consecutive byte ranges of ONE stream.  Halo = the first 16 bytes of
            # the next rank's range (all_gather of 16 B, once); per step: shard scan, all_gather of the ~190-byte shard
            # records over NCCL, host stitch, patch of the boundary entries.
            from hevcbitstream_b200 import shard as hs

            assert unit[-1] >= 2, "rank boundary must follow a byte >= 2 (hevcb_plan_shards rule)"
            heads = [torch.empty(16, dtype=torch.uint8, device=dev) for _ in range(world)]
            dist.all_gather(heads, d[:16].clone())
            is_first, is_last = rank == 0, rank == world - 1
            halo = 0 if is_last else 16
            if not is_last:
                d[size: size + 16] = heads[rank + 1]
            outs = hs.alloc_shard_outputs(d, size, cap)
            box = {}

            def step():  # the join of the records and the patches run on the device: a step is queued without waiting for the host
                box["sc"], _ = hs.scan_strip_sharded(ctx, d, size, halo, is_first, is_last, out=outs, sync=False)
        step()
        torch.cuda.synchronize()
        for _ in range(warmup):
            step()
        barrier()
        l0 = ctx.launch_count
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1) / steps
        launches = ctx.launch_count - l0
        if world == 1:
            s = outs["summary"].cpu().numpy()
            n_nals, rbsp_bytes, n_epb = int(s[0]), int(s[5]), int(s[6])
            assert int(s[2] >> 32) == 0, "NAL capacity overflow in bench"
        else:
            sc = box["sc"]
            res = hs.fetch_stitch(sc)
            n_nals, rbsp_bytes, n_epb = int(res.n_owned[rank]), int(sc.record.rbsp_bytes), int(sc.record.n_epb)
            stitched = {"global_n_nals": int(res.glob.n_nals), "global_rbsp_bytes": int(res.glob.rbsp_bytes), "last_rc": int(res.glob.last_rc),
                        "patches": int(res.n_patches), "owned_per_rank": [int(res.n_owned[r]) for r in range(world)]}
        # NCCL exchange of the per-shard counts (the only collective of the sharded path)
        tot = torch.tensor([ms, float(size), float(n_nals), float(rbsp_bytes)], dtype=torch.float64, device=dev)
        if world > 1:
            allv = [torch.zeros_like(tot) for _ in range(world)]
            dist.all_gather(allv, tot)
            allv = torch.stack(allv).cpu().numpy()
        else:
            allv = tot.cpu().numpy()[None, :]
        ms_max = float(allv[:, 0].max())
        tot_size = float(allv[:, 1].sum())
        tot_nals = float(allv[:, 2].sum())
        tot_rbsp = float(allv[:, 3].sum())
        alg = tot_size + tot_rbsp + 24.0 * tot_nals
        res = dict(name=name, nal_size=nal_size, ms=ms_max, size=tot_size, n_nals=tot_nals, rbsp_bytes=tot_rbsp, n_epb=n_epb,
                   in_gbs=tot_size / (ms_max * 1e-3) / 1e9, alg_gbs=alg / (ms_max * 1e-3) / 1e9, launches=launches,
                   alg_bytes_per_gpu=(size + rbsp_bytes + 24.0 * n_nals), unit=unit, d=d, outs=outs, cap=cap, size_local=size, stitched=stitched)
        if stitched is not None:
            assert stitched["global_n_nals"] == int(tot_nals) and stitched["global_rbsp_bytes"] == int(tot_rbsp), (stitched, tot_nals, tot_rbsp)
        return res

    name, nal_size, dense = next(w for w in WORKLOADS if w[0] == args.workload)
    with ClockSampler(local) as cs:
        head = time_workload(name, nal_size, dense, args.steps, args.warmup)
    clocks = cs.summary()

    # ---- e2e: host-pointer C ABI with pinned buffers, H2D + D2H inside the timed region (per rank, summed)
    e2e = None
    try:
        import ctypes as C

        from hevcbitstream_b200._lib import ScanSummary

        size = min(head["size_local"], int(args.e2e_gib * (1 << 30)))
        size -= size % head["unit"].size if size >= head["unit"].size else 0
        cap = head["cap"]
        h_in = torch.empty(size, dtype=torch.uint8).pin_memory()
        h_in.copy_(head["d"][:size].cpu())
        h_rbsp = torch.empty(size + 16, dtype=torch.uint8).pin_memory()
        h_arr = [torch.empty(cap, dtype=torch.int64).pin_memory() for _ in range(4)]
        sm = ScanSummary()
        L = ctx._L

        def one():
            rc = L.hevcb_scan_strip_host(ctx._h, h_in.data_ptr(), size, h_arr[0].data_ptr(), h_arr[1].data_ptr(), cap, h_rbsp.data_ptr(),
                                         h_arr[2].data_ptr(), h_arr[3].data_ptr(), C.byref(sm))
            assert rc == 0, rc

        one()
        barrier()
        t0 = time.perf_counter()
        k = max(1, min(args.steps, 5))
        for _ in range(k):
            one()
        barrier()
        dt = (time.perf_counter() - t0) / k
        v = torch.tensor([dt, float(size)], dtype=torch.float64, device=dev)
        if world > 1:
            allv = [torch.zeros_like(v) for _ in range(world)]
            dist.all_gather(allv, v)
            allv = torch.stack(allv).cpu().numpy()
        else:
            allv = v.cpu().numpy()[None, :]
        # the ceiling of this path on this box: the same bytes copied in and out at the same time by every rank, no kernel at all
        # (pinned host memory <-> device over each GPU's own link; what limits it is the host side the ranks share)
        s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
        d_in = torch.empty(size, dtype=torch.uint8, device=dev)
        d_out = torch.empty(size + 16, dtype=torch.uint8, device=dev)

        def copies():
            with torch.cuda.stream(s_in):
                d_in.copy_(h_in, non_blocking=True)
            with torch.cuda.stream(s_out):
                h_rbsp.copy_(d_out, non_blocking=True)
            s_in.synchronize()
            s_out.synchronize()

        copies()
        barrier()
        t0 = time.perf_counter()
        for _ in range(k):
            copies()
        barrier()
        dtc = (time.perf_counter() - t0) / k
        vc = torch.tensor([dtc], dtype=torch.float64, device=dev)
        if world > 1:
            allc = [torch.zeros_like(vc) for _ in range(world)]
            dist.all_gather(allc, vc)
            dtc = float(torch.stack(allc).max().item())
        ceiling = float(allv[:, 1].sum() / dtc / 1e9)
        e2e = {"value": float(allv[:, 1].sum() / allv[:, 0].max() / 1e9), "unit": UNIT, "h2d_bytes_per_step": int(size),
               "d2h_bytes_per_step": int(sm.rbsp_bytes + 4 * 8 * sm.n_nals + C.sizeof(ScanSummary)), "bytes_per_rank": int(size),
               "memcpy_ceiling": {"value": ceiling, "unit": UNIT, "frac_of_ceiling": float(allv[:, 1].sum() / allv[:, 0].max() / 1e9) / ceiling,
                                  "what": "the step's bytes copied host->device and device->host at the same time by all ranks, pinned memory, no kernel (max over ranks)"},
               "note": "hevcb_scan_strip_host: pinned host buffers, copies inside the timed region (64 MiB shards pipelined over copy-in / scan / copy-out streams, stitched on the host); bound by the host<->device copies, see memcpy_ceiling"}
        del h_in, h_rbsp, h_arr, d_in, d_out
    except Exception as ex:  # report, never fake
        e2e = {"value": None, "unit": UNIT, "error": repr(ex)}

    # ---- sweep over the other BASELINE config[1] shapes (fewer steps each)
    sweep = {}
    if not args.no_sweep and world == 1:
        unit_keep = head["unit"]
        for wname, wnal, wdense in WORKLOADS:
            if wname == name:
                r = head
            else:
                head_d = None
                r = time_workload(wname, wnal, wdense, max(2, args.steps // 4), 3)
            sweep[wname] = {"nal_size": wnal, "ms": round(r["ms"], 4), "input_GBps": round(r["in_gbs"], 1), "algorithmic_GBps": round(r["alg_gbs"], 1),
                            "frac_of_peak": round(r["alg_gbs"] / peak, 4), "n_nals": int(r["n_nals"]), "n_epb": int(r["n_epb"])}
            if r is not head:
                del r
                torch.cuda.empty_cache()

    # ---- BASELINE config[2]: batched header parse of >= 1M header-bearing NALs.  Primary number: 1 M DISTINCT headers written by the
    # reference's own writer (oracle/_ref's generator: multi-slice, tiles / WPP entry points, long-term refs, RPS, pred-weight,
    # VUI + HRD, re-sent parameter sets, unsupported types); second number: the small reference-written fixture tiled (every
    # warp then holds identical headers after the shape sort: the best case for divergence).
    parse = None
    if not args.no_parse:
        def time_parse(unit_h, reps, label):
            dh = torch.from_numpy(unit_h).to(dev)
            if reps > 1:
                dh = dh.repeat(reps)
            dh = torch.cat([dh, torch.zeros(32, dtype=torch.uint8, device=dev)])
            hsize = dh.numel() - 32
            scan = ctx.scan_strip_device(dh, size=hsize, cap_nals=hsize // 8 + 1024)
            n_h = int(scan.n_nals)
            for _ in range(3):
                pout = ctx.parse_device(dh, scan)
            barrier()
            l0 = ctx.launch_count
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            k = max(3, args.steps // 2)
            e0.record()
            for _ in range(k):
                pout = ctx.parse_device(dh, scan, sync=False)
            e1.record()
            barrier()
            pms = e0.elapsed_time(e1) / k
            pl = (ctx.launch_count - l0) // k
            pout = ctx.parse_device(dh, scan)
            e0.record()
            for _ in range(k):
                ctx.scan_strip_device(dh, size=hsize, cap_nals=hsize // 8 + 1024, sync=False)
            e1.record()
            barrier()
            sms = e0.elapsed_time(e1) / k
            return {"n_nals": n_h, "ms_parse": pms, "nal_headers_per_s": n_h / (pms * 1e-3), "syntax_elements": int(pout["n_pairs"]),
                    "elements_per_s": pout["n_pairs"] / (pms * 1e-3), "n_ok": int(pout["n_ok"]), "launches_per_parse": int(pl),
                    "ms_scan_strip": sms, "nal_headers_per_s_incl_scan_strip": n_h / ((pms + sms) * 1e-3), "workload": label}

        try:
            from oracle import ref as _ref

            fixture = np.fromfile(os.path.join(ROOT, "tests", "golden", "headers_unit.bin"), dtype=np.uint8)
            if _ref.available():
                t0 = time.perf_counter()
                gs = _ref.gen_stream(seed=21 + rank, profile=1, n_slices=args.parse_nals, payload_min=1, payload_max=64, zero_heavy_pct=10, extra_zero_pct=5,
                                     ps_period=500, unsupported_pct=2)
                gen_s = time.perf_counter() - t0
                distinct = gs[: gs.size - _ref.PAD]
                parse = time_parse(distinct, 1, f"{args.parse_nals} DISTINCT slice headers (+ parameter sets every 500) written by the reference's writer "
                                                f"(oracle/_ref generator, seed {21 + rank}, payload <= 64 B; generated on the host in {gen_s:.1f} s); per rank")
                if rank == 0 and world == 1:  # CPU baseline of this leg: the reference's own loop over a bounded sample of the same stream
                    # (a complete stream of its own: the reference crashes on a NAL that is cut short, App. A-11)
                    n_cs = min(args.parse_nals, 300000)
                    cs = _ref.gen_stream(seed=21 + rank, profile=1, n_slices=n_cs, payload_min=1, payload_max=64, zero_heavy_pct=10, extra_zero_pct=5,
                                         ps_period=500, unsupported_pct=2)
                    t_cpu, n_cpu = _ref.time_loop(cs, cs.size - _ref.PAD, 2, 3)
                    parse["cpu_reference_nal_headers_per_s"] = n_cpu / t_cpu
                    parse["cpu_reference_note"] = f"find_nal_unit + read_hevc_nal_unit loop of the unmodified reference, 1 thread, {n_cs} slices of the same generator, best of 3"
            else:
                parse = {"note": "oracle/_ref not built: no generator for distinct headers on this box"}
            reps = max(1, -(-args.parse_nals // 4106))
            parse["tiled_fixture"] = time_parse(fixture, reps, f"tests/golden/headers_unit.bin (reference-written, 4106 NALs) tiled x{reps}: identical headers side by side after the shape sort")
        except Exception as ex:
            parse = {"error": repr(ex)}

    # ---- rbsp_to_nal on the headline stream: re-insert the emulation prevention bytes of every NAL of the image just produced
    insert = None
    if not args.no_insert and world == 1:
        try:
            sres = ctx.scan_strip_device(head["d"], size=head["size_local"], cap_nals=head["cap"], out=head["outs"])
            nn = int(sres.n_nals)
            roff, rend = sres.rbsp_off[:nn].contiguous(), sres.rbsp_end[:nn].contiguous()
            ocap = int(head["size_local"]) + int(head["size_local"]) // 32 + 4096
            for _ in range(2):
                ins = ctx.insert_device(sres.rbsp, roff, rend, n_nals=nn, start_code_len=3, out_cap=ocap)
            barrier()
            k = max(3, args.steps // 4)
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(k):
                ins = ctx.insert_device(sres.rbsp, roff, rend, n_nals=nn, start_code_len=3, out_cap=ocap, sync=False)
            e1.record()
            barrier()
            ims = e0.elapsed_time(e1) / k
            s_ = ins["summary"].cpu().numpy()
            alg_i = float(int(sres.rbsp_bytes)) + float(s_[1])  # N_rbsp + N_nal (a second read of the payload earns nothing)
            insert = {"ms": ims, "n_nals": nn, "out_bytes": int(s_[1]), "epb_inserted": int(s_[2]), "input_GBps": int(sres.rbsp_bytes) / (ims * 1e-3) / 1e9,
                      "algorithmic_GBps": alg_i / (ims * 1e-3) / 1e9, "frac_of_peak": alg_i / (ims * 1e-3) / 1e9 / peak,
                      "round_trip_identical": bool(int(s_[1]) == int(head["size_local"]) and torch.equal(ins["out"][: int(s_[1])], head["d"][: int(s_[1])])),
                      "workload": "hevcb_insert_device (start codes + rbsp_to_nal) over the image and extents of the headline stream; output must equal the input stream"}
            del ins
        except Exception as ex:
            insert = {"error": repr(ex)}

    # ---- BASELINE config[3]: round-trip rewrite (parse, edit slice_qp_delta + a VUI flag, write_hevc_nal_unit, rbsp_to_nal) on a
    # stream of reference-written headers carrying 16 KiB payloads.  The unit is built on the device with the product's own
    # kernels (headers of tests/golden/headers_unit.bin + random payload -> hevcb_insert_device), then tiled.
    rewrite = None
    if not args.no_rewrite:
        try:
            unit_h = np.fromfile(os.path.join(ROOT, "tests", "golden", "headers_unit.bin"), dtype=np.uint8)
            dh = torch.zeros(unit_h.size + 32, dtype=torch.uint8, device=dev)
            dh[: unit_h.size] = torch.from_numpy(unit_h).to(dev)
            sc = ctx.scan_strip_device(dh, size=unit_h.size)
            pr = ctx.parse_device(dh, sc)
            nh = int(sc.n_nals)
            ro, re_ = sc.rbsp_off[:nh], sc.rbsp_end[:nh]
            is_slice = (pr["kind"][:nh] == 4) & (pr["rc"][:nh] >= 0) & (re_ >= 0)
            keep = torch.where(is_slice, pr["hdr_end"][:nh].to(torch.int64), torch.clamp(re_ - ro, min=0))  # header bytes / whole RBSP
            pay = args.rewrite_payload
            seg = keep + torch.where(is_slice, torch.full_like(keep, pay + 1), torch.zeros_like(keep))
            seg_end = torch.cumsum(seg, 0)
            seg_off = seg_end - seg
            total = int(seg_end[-1])
            g = torch.Generator(device=dev).manual_seed(4242 + rank)
            nr = torch.randint(0, 256, (total + 32,), dtype=torch.uint8, device=dev, generator=g)
            dst = torch.repeat_interleave(seg_off, keep) + (torch.arange(int(keep.sum()), device=dev) - torch.repeat_interleave(torch.cumsum(keep, 0) - keep, keep))
            src = torch.repeat_interleave(ro, keep) + (torch.arange(int(keep.sum()), device=dev) - torch.repeat_interleave(torch.cumsum(keep, 0) - keep, keep))
            nr[dst] = sc.rbsp[src]
            nr[(seg_end - 1)[is_slice]] = 0x80  # rbsp_trailing_bits
            live = seg > 0  # a zero-length NAL would end the reference loop
            unit_ins = ctx.insert_device(nr, seg_off[live].contiguous(), seg_end[live].contiguous(), start_code_len=4)
            ub = int(unit_ins["out_bytes"])
            reps = max(1, int(args.rewrite_gib * (1 << 30)) // ub)
            dr = torch.zeros(ub * reps + 32, dtype=torch.uint8, device=dev)
            dr[: ub * reps].view(reps, ub).copy_(unit_ins["out"][:ub].unsqueeze(0).expand(reps, -1))
            rsize = ub * reps
            del nr, unit_ins, src, dst
            cap = nh * reps + 1024
            edits = [(4, "slice_qp_delta", 0, 2), (2, "vui.video_full_range_flag", 2, 1)]

            def pipeline():
                scan = ctx.scan_strip_device(dr, size=rsize, cap_nals=cap)
                parsed = ctx.parse_device(dr, scan, cap_pairs=80 * scan.n_nals)
                return scan, parsed, ctx.rewrite_device(dr, scan, parsed, edits, size=rsize)

            for _ in range(2):
                scan, parsed, out = pipeline()
            del scan, parsed, out
            barrier()
            l0 = ctx.launch_count
            k = max(3, args.steps // 4)
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(k):
                scan, parsed, out = pipeline()
            e1.record()
            barrier()
            rms = e0.elapsed_time(e1) / k
            rewrite = {"bytes_in": rsize, "n_nals": int(scan.n_nals), "n_rewritten": int(out["n_rewritten"]), "bytes_out": int(out["out_bytes"]),
                       "epb_inserted": int(out["n_inserted"]), "ms_pipeline": rms, "input_GBps": rsize / (rms * 1e-3) / 1e9,
                       "nals_per_s": int(scan.n_nals) / (rms * 1e-3), "launches_per_pipeline": int((ctx.launch_count - l0) // k),
                       "workload": f"scan + strip + parse + rewrite (slice_qp_delta += 2, vui.video_full_range_flag ^= 1) of reference-written headers "
                                   f"with {pay} B random payloads, {rsize / 2**30:.2f} GiB per rank; byte-exactness vs the reference writer is asserted in tests/test_rewrite_gpu.py"}
            del dr, scan, parsed, out
            torch.cuda.empty_cache()
        except Exception as ex:
            rewrite = {"error": repr(ex)}

    # ---- CPU baseline: the unmodified reference on rank 0's host cores, bounded sample of the same workload
    cpu = None
    if rank == 0 and world == 1 or (rank == 0 and args.cpu_baseline_multi):
        try:
            sample = head["unit"][: args.ref_sample_mib << 20]
            v, n, t = cpu_reference_sample(sample, reps=3)
            cpu = {"value": v, "unit": UNIT, "cores": 1, "kind": "reference",
                   "sample": f"first {sample.size >> 20} MiB of the {name} stream, find_nal_unit + nal_to_rbsp loop (oracle/_ref), best of 3, {n} NALs"}
            if not args.no_cpu_extras:
                cpu["also"] = cpu_reference_extras()
        except Exception as ex:
            cpu = {"value": None, "unit": UNIT, "error": repr(ex)}

    # DRAM traffic of the dominant kernel from the committed ncu --set full capture of this same workload (profiles/)
    traffic = None
    try:
        pj = json.load(open(os.path.join(ROOT, "profiles", "r2_scan_strip_ncu.json")))
        if pj.get("workload") == name and abs(pj.get("size_gib", 0) - args.size_gib) < 1e-6:
            traffic = pj["traffic_bytes_per_launch"]
    except Exception:
        pass
    if rank == 0:
        per_gpu_alg_gbs = head["alg_bytes_per_gpu"] / (head["ms"] * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": head["in_gbs"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": head["ms"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": workload_config(args, name, nal_size, head["size_local"], world),
            "detail": {"nals_per_step": int(head["n_nals"]), "nal_headers_located_per_s": head["n_nals"] / (head["ms"] * 1e-3)},
            "roofline": {"bound": "hbm", "achieved": per_gpu_alg_gbs, "peak": peak, "unit": "GB/s", "frac": per_gpu_alg_gbs / peak,
                         "traffic": traffic, "algorithmic_bytes_per_launch": head["alg_bytes_per_gpu"], "peak_source": peak_src,
                         "algorithmic_bytes": "N_in + N_rbsp + 24*NALs per launch (SURVEY 8d), per GPU; time = CUDA-event mean over the timed steps (memset + init + scan + finalize launches)"},
            "cpu_baseline": cpu,
            "e2e": e2e,
            "gpu_launches": int(head["launches"]),
            "clocks": clocks,
        }
        if head.get("stitched"):
            line["stitched"] = head["stitched"]
        if sharded_parity is not None:
            line["sharded_parity"] = sharded_parity
        if sweep:
            line["sweep"] = sweep
        if parse:
            line["parse"] = parse
        if insert:
            line["insert"] = insert
        if rewrite:
            line["rewrite"] = rewrite
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size-gib", type=float, default=4.0)
    ap.add_argument("--e2e-gib", type=float, default=1.0)
    ap.add_argument("--workload", default="nal16k", choices=[w[0] for w in WORKLOADS])
    ap.add_argument("--no-sweep", action="store_true")
    ap.add_argument("--no-parse", action="store_true")
    ap.add_argument("--parse-nals", type=int, default=1_000_000)
    ap.add_argument("--no-rewrite", action="store_true")
    ap.add_argument("--no-insert", action="store_true")
    ap.add_argument("--rewrite-gib", type=float, default=4.0)
    ap.add_argument("--rewrite-payload", type=int, default=16384)
    ap.add_argument("--ref-sample-mib", type=int, default=64)
    ap.add_argument("--ref-passes", type=int, default=16, help="reference arm: passes over the 64 MiB unit per step (0 = the whole buffer)")
    ap.add_argument("--cpu-baseline-multi", action="store_true")
    ap.add_argument("--no-cpu-extras", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
