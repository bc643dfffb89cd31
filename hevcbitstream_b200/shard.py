"""Byte-range sharding of one Annex-B stream over several GPUs (include/hevcb.h: hevcb_plan_shards,
hevcb_scan_strip_shard_device, hevcb_stitch).

One process per GPU.  Every rank scans + strips its own bytes (plus a halo of the 16 bytes that follow them); the only
exchange is an all_gather of one ~170-byte record per shard, after which every rank runs the same host-side stitch and
patches the (at most two) entries of its own arrays that depend on a neighbour.  Payload bytes only move when a caller
asks for the continuation of a NAL that crosses into the next shard (`fetch_continuation`).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import HevcbError, ShardSummary, StitchResult, load_library

HALO = 16


def plan_shards(buf: np.ndarray, n_shards: int, size=None) -> np.ndarray:
    """Cut points (n_shards + 1 offsets) such that no start-code / EPB pattern reaches back across a cut."""
    assert buf.dtype == np.uint8
    size = int(buf.size if size is None else size)
    bounds = np.zeros(n_shards + 1, dtype=np.int64)
    rc = load_library().hevcb_plan_shards(buf.ctypes.data_as(C.c_void_p), size, n_shards, bounds.ctypes.data_as(C.c_void_p))
    if rc != 0:
        raise HevcbError(rc, "hevcb_plan_shards")
    return bounds


def shard_flags(bounds: np.ndarray, r: int):
    """(own, halo, is_first, is_last) of shard r for the cut points of plan_shards."""
    size = int(bounds[-1])
    lo, hi = int(bounds[r]), int(bounds[r + 1])
    own = hi - lo
    is_first = own > 0 and lo == 0
    is_last = own > 0 and hi == size
    halo = 0 if is_last else min(HALO, size - hi)
    return own, halo, is_first, is_last


def stitch(records) -> StitchResult:
    """records: ShardSummary per shard, in stream order."""
    n = len(records)
    arr = (ShardSummary * n)(*records)
    out = StitchResult()
    rc = load_library().hevcb_stitch(arr, n, C.byref(out))
    if rc != 0:
        raise HevcbError(rc, "hevcb_stitch: inconsistent shard records")
    return out


def record_to_tensor(rec: ShardSummary, device):
    import torch

    raw = np.frombuffer(bytes(rec), dtype=np.uint8).copy()
    return torch.from_numpy(raw).to(device)


def gather_records(rec: ShardSummary, device, group=None):
    """all_gather of the shard records (NCCL on the GPU box, gloo in the CPU tests): the only collective of the scan."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    mine = record_to_tensor(rec, device)
    out = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(out, mine, group=group)
    return [ShardSummary.from_buffer_copy(t.cpu().numpy().tobytes()) for t in out]


def patches_for(res: StitchResult, shard: int):
    return [res.patches[i] for i in range(res.n_patches) if res.patches[i].shard == shard]


def apply_patches(res: StitchResult, shard: int, nal_start, nal_end, rbsp_off, rbsp_end):
    """Writes the stitched entries into shard `shard`'s arrays (numpy arrays or torch tensors, local coordinates)."""
    for p in patches_for(res, shard):
        k = int(p.index)
        if p.set_start:
            nal_start[k] = int(p.nal_start)
            rbsp_off[k] = int(p.rbsp_off)
        nal_end[k] = int(p.nal_end)
        rbsp_end[k] = int(p.rbsp_end)


class ShardScan:
    """Result of scan_strip_shard on one rank: device arrays in local coordinates + the shard record."""

    def __init__(self, record, nal_start, nal_end, rbsp_off, rbsp_end, rbsp):
        self.record, self.nal_start, self.nal_end, self.rbsp_off, self.rbsp_end, self.rbsp = record, nal_start, nal_end, rbsp_off, rbsp_end, rbsp


def alloc_shard_outputs(buf, own: int, cap_nals: int, want_rbsp=True, extra_rbsp=0):
    import torch

    dev = buf.device
    return dict(arrays=[torch.empty(cap_nals, dtype=torch.int64, device=dev) for _ in range(4)],
                rbsp=torch.empty(own + 16 + extra_rbsp, dtype=torch.uint8, device=dev) if want_rbsp else None,
                summary=torch.zeros(C.sizeof(ShardSummary), dtype=torch.uint8, device=dev), cap_nals=cap_nals)


def scan_strip_shard(ctx, buf, own: int, halo: int, is_first: bool, is_last: bool, cap_nals=None, want_rbsp=True, extra_rbsp=0, out=None,
                     sync=True) -> ShardScan:
    """buf: torch.uint8 CUDA tensor with own + halo bytes (16-byte aligned).  extra_rbsp: spare bytes behind the image for
    the continuation of the last NAL.  out: buffers of alloc_shard_outputs to reuse.  sync=False leaves the record on the
    device (ShardScan.record is None, ShardScan.summary holds the raw bytes)."""
    import torch

    assert buf.is_cuda and buf.dtype == torch.uint8 and buf.numel() >= own + halo
    auto_cap = out is None and cap_nals is None and sync
    if out is None:
        # default: one NAL per 64 bytes (0.5 bytes of arrays per input byte); a denser shard reports its true count in the
        # record and the synchronous call repeats with it
        out = alloc_shard_outputs(buf, own, own // 64 + 1024 if cap_nals is None else cap_nals, want_rbsp, extra_rbsp)
    a, rbsp, d_sum, cap_nals = out["arrays"], out["rbsp"], out["summary"], out["cap_nals"]
    stream = torch.cuda.current_stream(buf.device).cuda_stream
    rc = ctx._L.hevcb_scan_strip_shard_device(ctx._h, buf.data_ptr(), own, halo, int(is_first), int(is_last), a[0].data_ptr(), a[1].data_ptr(),
                                              cap_nals, rbsp.data_ptr() if rbsp is not None else None, a[2].data_ptr(), a[3].data_ptr(),
                                              d_sum.data_ptr(), stream)
    ctx._check(rc)
    sc = ShardScan(None, a[0], a[1], a[2], a[3], rbsp)
    sc.summary, sc.cap_nals = d_sum, cap_nals
    if sync:
        sc.record = ShardSummary.from_buffer_copy(d_sum.cpu().numpy().tobytes())
        if sc.record.overflow:
            if auto_cap:  # the record holds the true count of the shard
                return scan_strip_shard(ctx, buf, own, halo, is_first, is_last, cap_nals=max(int(sc.record.n_nals) + 8, own // 3 + 8), want_rbsp=want_rbsp,
                                        extra_rbsp=extra_rbsp)
            raise HevcbError(-104, f"{sc.record.n_nals} NALs exceed cap_nals {cap_nals}")
    return sc


def apply_patches_device(ctx, res: StitchResult, shard: int, sc: ShardScan):
    """One tiny kernel writes the stitched entries of this shard into its device arrays (hevcb_apply_patches_device)."""
    import torch

    stream = torch.cuda.current_stream(sc.nal_start.device).cuda_stream
    ctx._check(ctx._L.hevcb_apply_patches_device(ctx._h, C.byref(res), shard, sc.nal_start.data_ptr(), sc.nal_end.data_ptr(), sc.rbsp_off.data_ptr(),
                                                 sc.rbsp_end.data_ptr(), sc.cap_nals, stream))


def scan_strip_sharded(ctx, buf, own, halo, is_first, is_last, group=None, sync=True, **kw):
    """The distributed pass: local shard scan, all_gather of the device-resident records (the only collective), then ONE small kernel
    that joins the records on the device and writes this rank's patches (hevcb_stitch_apply_device): nothing in the step waits for
    the host.  Returns (ShardScan, StitchResult); the global index of local NAL j is res.nal_base[rank] + j - res.first_local[rank].
    sync=False: the join's result stays on the device (the second value is None; `fetch_stitch(sc)` reads it when it is wanted),
    so that steps can be queued back to back."""
    import torch
    import torch.distributed as dist

    sc = scan_strip_shard(ctx, buf, own, halo, is_first, is_last, sync=False, **kw)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    allrec = torch.empty(world * sc.summary.numel(), dtype=torch.uint8, device=buf.device)
    dist.all_gather_into_tensor(allrec, sc.summary, group=group)
    d_res = torch.empty(C.sizeof(StitchResult), dtype=torch.uint8, device=buf.device)
    stream = torch.cuda.current_stream(buf.device).cuda_stream
    ctx._check(ctx._L.hevcb_stitch_apply_device(ctx._h, allrec.data_ptr(), world, rank, sc.nal_start.data_ptr(), sc.nal_end.data_ptr(), sc.rbsp_off.data_ptr(),
                                                sc.rbsp_end.data_ptr(), sc.cap_nals, d_res.data_ptr(), stream))
    sc.allrec, sc.d_stitch, sc.world, sc.rank = allrec, d_res, world, rank
    if not sync:
        return sc, None
    return sc, fetch_stitch(sc)


def fetch_stitch(sc: ShardScan) -> StitchResult:
    """Device -> host read of what the last scan_strip_sharded(..., sync=False) left on the device: the gathered records (sc.record =
    this rank's) and the join's result.  Raises together on every rank (the inputs are identical everywhere)."""
    raw = sc.allrec.cpu().numpy().tobytes()
    n = C.sizeof(ShardSummary)
    records = [ShardSummary.from_buffer_copy(raw[i * n:(i + 1) * n]) for i in range(sc.world)]
    sc.record = records[sc.rank]
    over = [r for r in range(sc.world) if records[r].overflow]
    if over:  # decided from the gathered records: every rank raises together, none is left waiting in a later collective
        raise HevcbError(-104, f"shard(s) {over}: more NALs than cap_nals ({records[over[0]].n_nals} on shard {over[0]}, cap {sc.cap_nals})")
    res = StitchResult.from_buffer_copy(sc.d_stitch.cpu().numpy().tobytes())
    if res.n_patches < 0:
        raise HevcbError(-101, "hevcb_stitch: the shard records do not describe one stream (is_first / is_last flags)")
    return res


# ---- sharded header parse: parameter-set hand-over between ranks -------------------------------------------------------
HEAD_BYTES = 65536  # bytes of every shard's image that are all-gathered so that a NAL crossing into the next shard can be parsed


def ps_context_bytes():
    from ._lib import load_library
    a, b = C.c_int64(0), C.c_int64(0)
    load_library().hevcb_ps_context_bytes(C.byref(a), C.byref(b))
    return int(a.value), int(b.value)


def _parse_call(ctx, buf, ns, ne, rbsp, ro, re, n, cap_pairs, chain):
    """hevcb_parse_shard_device on already sliced device arrays; returns the dict api.Context.parse_device returns."""
    import torch

    from ._lib import ParseBuffers

    dev = buf.device
    m = max(n, 1)
    out = dict(rc=torch.empty(m, dtype=torch.int32, device=dev), nal_hdr=torch.empty(m, dtype=torch.int32, device=dev),
               kind=torch.empty(m, dtype=torch.uint8, device=dev), ubflag=torch.empty(m, dtype=torch.uint8, device=dev),
               hdr_end=torch.empty(m, dtype=torch.int32, device=dev), cols=torch.empty((8, m), dtype=torch.int32, device=dev),
               pair_off=torch.empty(n + 1, dtype=torch.int64, device=dev), pair_field=torch.empty(cap_pairs, dtype=torch.int32, device=dev),
               pair_value=torch.empty(cap_pairs, dtype=torch.int32, device=dev), summary=torch.zeros(8, dtype=torch.int64, device=dev))
    pb = ParseBuffers(out["rc"].data_ptr(), out["nal_hdr"].data_ptr(), out["kind"].data_ptr(), out["ubflag"].data_ptr(), out["hdr_end"].data_ptr(),
                      out["cols"].data_ptr(), out["pair_off"].data_ptr(), out["pair_field"].data_ptr(), out["pair_value"].data_ptr(), cap_pairs)
    stream = torch.cuda.current_stream(dev).cuda_stream
    ctx._check(ctx._L.hevcb_parse_shard_device(ctx._h, buf.data_ptr(), ns.data_ptr(), ne.data_ptr(), rbsp.data_ptr(), ro.data_ptr(), re.data_ptr(), n,
                                               C.byref(pb), out["summary"].data_ptr(), C.byref(chain), stream))
    s = out["summary"].cpu().numpy()
    out["n"], out["cap_pairs"] = n, cap_pairs
    out["n_ok"], out["n_pairs"], out["n_vps"], out["n_sps"], out["n_pps"], out["n_slices"] = (int(x) for x in s[1:7])
    if int(s[7]) & 0xFFFFFFFF:
        raise HevcbError(-104, f"{out['n_pairs']} syntax elements exceed cap_pairs {cap_pairs}")
    return out


def append_continuation(sc: ShardScan, res: StitchResult, rank: int, heads):
    """Copies the continuation of this shard's last NAL (it ends in a later shard) behind the shard's image, at most HEAD_BYTES of
    it: header bytes are all the parser needs.  heads[q]: the first HEAD_BYTES image bytes of shard q.  A NAL that spans whole
    shards takes the (complete) images of the shards in between from their heads as long as they fit.  Decided from the
    StitchResult, which is identical on every rank, so no rank can raise alone between two collectives.
    Returns (bytes appended, bytes of the continuation that were NOT appended)."""
    q = int(res.cont_last_shard[rank])
    if q < 0:
        return 0, 0
    total = int(res.cont_bytes[rank])
    last = int(res.cont_last_bytes[rank])
    base = int(sc.record.rbsp_bytes)
    room = min(HEAD_BYTES, int(sc.rbsp.numel()) - base - 16)
    assert room >= min(total, HEAD_BYTES), "scan_strip_shard(extra_rbsp=HEAD_BYTES) is required for the sharded parse"
    take = 0
    for r in range(rank + 1, q + 1):
        avail = last if r == q else int(res.rbsp_base[r + 1] - res.rbsp_base[r])  # image bytes of shard r that belong to the NAL
        use = min(avail, room - take, int(heads[r].numel()))
        if use > 0:
            sc.rbsp[base + take: base + take + use] = heads[r][:use]
            take += use
        if use < avail:
            break
    return take, total - take


def local_ps_contexts(ctx, buf, sc: ShardScan, first: int, n: int):
    """State after this shard's last SPS / PPS NAL: (has_sps, sps_blob, has_pps, pps_blob) as numpy uint8 arrays."""
    import torch

    from ._lib import ParseChain

    sps_b, pps_b = ps_context_bytes()
    sps = np.zeros(sps_b, np.uint8)
    pps = np.zeros(pps_b, np.uint8)
    has = [0, 0]
    if n > 0:
        ns = sc.nal_start[first: first + n]
        ok = sc.rbsp_end[first: first + n] >= 0
        types = (buf[ns].to(torch.int32) >> 1) & 0x3F
        for slot, (t, blob) in enumerate(((33, sps), (34, pps))):
            idx = torch.nonzero((types == t) & ok).flatten()
            if idx.numel() == 0:
                continue
            k = first + int(idx[-1])
            chain = ParseChain(None, None, blob.ctypes.data if slot == 0 else None, blob.ctypes.data if slot == 1 else None, 0)
            _parse_call(ctx, buf, sc.nal_start[k:k + 1], sc.nal_end[k:k + 1], sc.rbsp, sc.rbsp_off[k:k + 1], sc.rbsp_end[k:k + 1], 1, 1 << 16, chain)
            has[slot] = 1
    return has[0], sps, has[1], pps


def pick_incoming(states, rank: int):
    """states[r] = (has_sps, sps_blob, has_pps, pps_blob) of every rank; returns the (sps, pps) blobs entering `rank` (None = zero state)."""
    sps = pps = None
    for r in range(rank - 1, -1, -1):
        if sps is None and states[r][0]:
            sps = states[r][1]
        if pps is None and states[r][2]:
            pps = states[r][3]
    return sps, pps


def parse_shard(ctx, buf, own: int, halo: int, sc: ShardScan, res: StitchResult, rank: int, sps_in, pps_in, cap_pairs=None, missing=0):
    """Header parse of the NALs this shard owns, with the parameter-set state that enters it.  missing: bytes at the end of the
    last NAL's RBSP that are not present behind the local image (append_continuation): the bit reader is kept off them."""
    import torch

    from ._lib import ParseChain

    f, n = int(res.first_local[rank]), int(res.n_owned[rank])
    if cap_pairs is None:
        cap_pairs = 80 * n + 4096
    chain = ParseChain(sps_in.ctypes.data if sps_in is not None else None, pps_in.ctypes.data if pps_in is not None else None, None, None, own + halo)
    rend = sc.rbsp_end[f: f + n]
    if missing > 0 and n > 0 and int(rend[n - 1]) >= 0:
        # the tail of the last NAL lives on another rank: a header never reaches that far (HEAD_BYTES of it are here), but a
        # malformed one must not walk off the local image
        rend = rend.clone()
        rend[n - 1] -= missing
    out = _parse_call(ctx, buf, sc.nal_start[f: f + n], sc.nal_end[f: f + n], sc.rbsp, sc.rbsp_off[f: f + n], rend, n, cap_pairs, chain)
    # read_hevc_nal_unit reports one byte less for a NAL that ends in 00 00 03; for a NAL that ends in a later shard only the
    # stitch has seen those bytes
    for p in patches_for(res, rank):
        k = int(p.index) - f
        if p.ends_003 and 0 <= k < n and int(p.nal_end) > own + halo and int(out["rc"][k]) >= 0:
            out["rc"][k] -= 1
    return out


def parse_sharded(ctx, buf, own: int, halo: int, sc: ShardScan, res: StitchResult, group=None, cap_pairs=None):
    """The distributed header parse (after scan_strip_sharded with extra_rbsp=HEAD_BYTES): all_gather of the image heads (a NAL
    that crosses into the next shard) and of each rank's last SPS / PPS state (~9 KB), then a local parse."""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = buf.device
    head = torch.zeros(HEAD_BYTES, dtype=torch.uint8, device=dev)
    nb = min(HEAD_BYTES, int(sc.record.rbsp_bytes))
    head[:nb] = sc.rbsp[:nb]
    heads = [torch.empty_like(head) for _ in range(world)]
    dist.all_gather(heads, head, group=group)
    _, missing = append_continuation(sc, res, rank, heads)
    f, n = int(res.first_local[rank]), int(res.n_owned[rank])
    hs_, sps, hp_, pps = local_ps_contexts(ctx, buf, sc, f, n)
    blob = torch.from_numpy(np.concatenate([np.array([hs_, hp_], np.uint8), sps, pps])).to(dev)
    blobs = [torch.empty_like(blob) for _ in range(world)]
    dist.all_gather(blobs, blob, group=group)
    sps_b, _ = ps_context_bytes()
    states = []
    for b in blobs:
        h = b.cpu().numpy()
        states.append((int(h[0]), h[2: 2 + sps_b].copy(), int(h[1]), h[2 + sps_b:].copy()))
    sps_in, pps_in = pick_incoming(states, rank)
    return parse_shard(ctx, buf, own, halo, sc, res, rank, sps_in, pps_in, cap_pairs, missing=missing)
