// hevcb_stitch.cu -- the join of the byte-range sharding (include/hevcb.h: hevcb_plan_shards, hevcb_stitch, hevcb_stitch_apply_device).
//
// The per-shard passes run on the GPUs (hevcb_scan_strip_shard_device); what is left is O(shards) integer bookkeeping over the
// all-gathered shard records, plus the reference's end-of-buffer rules (h264_nal.c:46-72, restated in hevcb_scan_tail) applied once to
// the last 8 bytes of the stream, which travel inside the last record.  The same function runs on the host (hevcb_stitch) and, so that
// a distributed step needs no device->host round trip, as a one-thread kernel that also writes the shard's patches
// (hevcb_stitch_apply_device).
#include <stdint.h>
#include <string.h>

#include <cuda_runtime.h>

#include "hevcb_internal.h"
#include "hevcb_scan_core.h"

extern "C" HEVCB_API int hevcb_plan_shards(const uint8_t* buf, int64_t size, int n_shards, int64_t* bounds)
{
    if (size < 0 || n_shards < 1 || n_shards > HEVCB_MAX_SHARDS || !bounds || (size > 0 && !buf)) { return HEVCB_E_ARG; }
    bounds[0] = 0;
    for (int g = 1; g < n_shards; g++) {
        // nominal cut, moved forward until the byte in front of it cannot take part in a pattern that crosses the cut
        int64_t p = (int64_t)(((__int128)size * g) / n_shards);
        if (p < bounds[g - 1]) { p = bounds[g - 1]; }
        while (p < size && !(p > 0 && buf[p - 1] >= 2)) { p++; }
        bounds[g] = p;
    }
    bounds[n_shards] = size;
    // the end-of-stream rules look at the last ~10 bytes: keep them inside one shard
    for (int g = n_shards - 1; g >= 1; g--) {
        if (bounds[g] < size && size - bounds[g] < 64) { bounds[g] = size; }
    }
    for (int g = 1; g <= n_shards; g++) {
        if (bounds[g] < bounds[g - 1]) { bounds[g] = bounds[g - 1]; }
    }
    return HEVCB_OK;
}

namespace {
struct TailFetch {
    const uint8_t* bytes;
    int64_t first; // local position of bytes[0]
    __host__ __device__ uint32_t operator()(int64_t pos) const { return pos >= first ? (uint32_t)bytes[pos - first] : 0xFFu; }
};
} // namespace

static __host__ __device__ int stitch_core(const hevcb_shard_summary* sh, int n_shards, hevcb_stitch_result* out)
{
    if (!sh || !out || n_shards < 1 || n_shards > HEVCB_MAX_SHARDS) { return HEVCB_E_ARG; }
    memset(out, 0, sizeof(*out));
    out->n_shards = n_shards;
    int last = -1, first = -1;
    int64_t bytes = 0, kept = 0, epb = 0;
    for (int r = 0; r < n_shards; r++) {
        out->byte_base[r] = bytes;
        out->rbsp_base[r] = kept;
        out->cont_last_shard[r] = -1;
        if (sh[r].own > 0) {
            if (first < 0) { first = r; }
            last = r;
            bytes += sh[r].own;
            kept += sh[r].rbsp_bytes;
            epb += sh[r].n_epb;
            if (sh[r].overflow) { out->global.overflow = 1; }
        }
    }
    out->global.rbsp_bytes = kept;
    out->global.n_epb = epb;
    for (int r = 0; r < n_shards; r++) { // flags must describe the position of the shard in the stream
        if (sh[r].own > 0 && ((sh[r].is_first != 0) != (r == first) || (sh[r].is_last != 0) != (r == last))) { return HEVCB_E_ARG; }
    }

    bool open = false, stopped = false, have_closed = false;
    int owner = -1, stop_shard = -1;
    int64_t owner_idx = 0, open_start = 0, cont_mid = 0, last_closed_end = 0, count = 0;
    uint32_t open_err = 0;
    auto add_patch = [&](int shard, int64_t index, bool set_start, int64_t ns, int64_t ro, int64_t ne, int64_t re) {
        hevcb_stitch_patch& p = out->patches[out->n_patches++];
        p.shard = shard; p.set_start = set_start ? 1 : 0; p.index = index; p.nal_start = ns; p.rbsp_off = ro; p.nal_end = ne; p.rbsp_end = re;
        p.ends_003 = 0; p.pad = 0;
    };
    // raw byte at a global position close to a shard boundary: the end of a NAL closed in shard q (its last three bytes) or
    // in the last bytes of the stream; served from the records (head_last3 of q, tail[] of the shards)
    auto raw_near = [&](int q, int64_t gpos, int64_t gend_in_q) -> int {
        for (int r = q; r >= 0; r--) {
            if (sh[r].own <= 0) { continue; }
            const int64_t b = out->byte_base[r];
            if (gpos < b) { continue; }
            const int64_t local = gpos - b;
            if (r == q && gend_in_q >= 0 && local >= gend_in_q - 3 && local < gend_in_q) { return sh[r].head_last3[local - (gend_in_q - 3)]; }
            const int64_t tfirst = sh[r].own - sh[r].tail_len;
            if (local >= tfirst && local < sh[r].own) { return sh[r].tail[local - tfirst]; }
            return -1;
        }
        return -1;
    };
    auto ends_003 = [&](int q, int64_t gstart, int64_t gend, int64_t end_local_in_q) -> int {
        if (gend - gstart < 3) { return 0; }
        const int b1 = raw_near(q, gend - 1, end_local_in_q), b2 = raw_near(q, gend - 2, end_local_in_q), b3 = raw_near(q, gend - 3, end_local_in_q);
        return (b1 == 3 && b2 == 0 && b3 == 0) ? 1 : 0;
    };

    for (int r = 0; r < n_shards && !stopped; r++) {
        const hevcb_shard_summary& S = sh[r];
        if (S.own <= 0) { out->nal_base[r] = count; continue; }
        const int64_t base = out->byte_base[r];
        const int64_t v = S.is_first ? 0 : 1;
        const int64_t n = S.n_nals;
        out->first_local[r] = v;
        out->nal_base[r] = count;
        if (v) {
            const bool piece_closed = (n > 1) || !S.open_at_end;
            if (open) {
                if (piece_closed) { // the NAL that entered the shard ends at the shard's first event
                    const bool err = open_err || S.head_rbsp_end < 0;
                    const int64_t cont = err ? 0 : cont_mid + S.head_rbsp_end;
                    add_patch(owner, owner_idx, false, 0, 0, base + S.head_end - out->byte_base[owner], err ? -1 : sh[owner].rbsp_bytes + cont);
                    out->patches[out->n_patches - 1].ends_003 = ends_003(r, open_start, base + S.head_end, S.head_end);
                    out->cont_last_shard[owner] = err ? -1 : r;
                    out->cont_last_bytes[owner] = err ? 0 : S.head_rbsp_end;
                    out->cont_bytes[owner] = cont;
                    last_closed_end = base + S.head_end;
                    have_closed = true;
                    open = false;
                } else { // the whole shard lies inside that NAL
                    open_err |= (uint32_t)S.open_err;
                    if (r != last) { cont_mid += S.rbsp_bytes; } // the last shard's share is added with the end-of-stream rules
                }
            }
        }
        const bool still_entering = v && open; // no event in this shard
        if (S.first_empty >= v && S.first_empty < n) { // the reference loop stops at a zero-length NAL
            out->n_owned[r] = S.first_empty - v;
            count += S.first_empty - v;
            out->global.n_nals = count;
            out->global.n_terminated = count;
            out->global.last_rc = 0;
            out->global.last_start = base + S.first_empty_start;
            out->global.last_end = out->global.last_start;
            stopped = true;
            stop_shard = r;
            break;
        }
        out->n_owned[r] = n - v;
        count += n - v;
        if (n - v >= 1) {
            if (S.open_at_end) {
                open = true; owner = r; owner_idx = n - 1; open_err = (uint32_t)S.open_err; open_start = base + S.last_nal_start; cont_mid = 0;
            } else {
                open = false; last_closed_end = base + S.last_nal_end; have_closed = true;
            }
        } else if (!still_entering) {
            open = false; // a piece that belongs to no NAL
        }
    }
    if (stopped) { // shards behind the stop own nothing
        for (int r = stop_shard + 1; r < n_shards; r++) { out->nal_base[r] = count; out->first_local[r] = sh[r].is_first ? 0 : 1; }
        return HEVCB_OK;
    }

    // ---- end-of-stream rules on the last shard (or on an empty stream)
    const int64_t base = last >= 0 ? out->byte_base[last] : 0;
    const int64_t own = last >= 0 ? sh[last].own : 0;
    hevcb_tail_in in;
    in.size = own;
    in.n = 1;
    in.kind = open ? HEVCB_KIND_SC3 : HEVCB_KIND_Z3;
    in.err = open ? open_err : 0u;
    in.open_start = open ? open_start - base : 0;
    in.prev_end = (have_closed ? last_closed_end : 0) - base;
    in.kept_total = last >= 0 ? sh[last].rbsp_bytes : 0;
    TailFetch fetch;
    fetch.bytes = last >= 0 ? sh[last].tail : nullptr;
    fetch.first = last >= 0 ? own - sh[last].tail_len : 0;
    hevcb_tail_out t;
    hevcb_scan_tail(in, fetch, t);
    const int64_t local_next = last >= 0 ? sh[last].n_nals : 0; // local index of the first NAL opened by these rules
    for (int i = 0; i < t.n_new && i < 4; i++) {
        if (t.closes_open && i == 0) {
            int64_t re = t.nal[0].rbsp_end;
            if (re >= 0 && owner != last) { // continuation: image of the last shard up to the end position
                out->cont_last_shard[owner] = last;
                out->cont_last_bytes[owner] = re;
                out->cont_bytes[owner] = cont_mid + re;
                re = sh[owner].rbsp_bytes + cont_mid + re;
            }
            add_patch(owner, owner_idx, false, 0, 0, base + t.nal[0].end - out->byte_base[owner], re);
            out->patches[out->n_patches - 1].ends_003 = ends_003(last, open_start, base + t.nal[0].end, -1);
        } else {
            add_patch(last, local_next + (i - (t.closes_open ? 1 : 0)), true, t.nal[i].start, t.nal[i].rbsp_off, t.nal[i].end, t.nal[i].rbsp_end);
        }
    }
    const int64_t gbase = count - (t.closes_open ? 1 : 0);
    int64_t total = gbase + t.n_new;
    if (t.first_empty >= 0) { total = gbase + t.first_empty; }
    if (last >= 0) {
        int64_t owned = total - out->nal_base[last];
        out->n_owned[last] = owned > 0 ? owned : 0;
    }
    out->global.n_nals = total;
    out->global.n_terminated = total - (t.last_is_nal ? 1 : 0);
    out->global.last_rc = t.last_rc;
    out->global.last_start = base + t.last_start;
    out->global.last_end = base + t.last_end;
    return HEVCB_OK;
}

extern "C" HEVCB_API int hevcb_stitch(const hevcb_shard_summary* sh, int n_shards, hevcb_stitch_result* out) { return stitch_core(sh, n_shards, out); }

namespace {
// the join on the device: one thread runs the same arithmetic over the gathered records (they are in device memory after the
// all_gather) and writes the entries of this rank's arrays that the join decides (<= 2 per rank)
__global__ void stitch_apply_kernel(const hevcb_shard_summary* __restrict__ recs, int n_shards, int shard, int64_t* ns, int64_t* ne, int64_t* ro,
                                    int64_t* re, int64_t cap, hevcb_stitch_result* __restrict__ res)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) { return; }
    const int rc = stitch_core(recs, n_shards, res);
    if (rc != HEVCB_OK) { res->n_patches = -1; return; } // (reported by the host wrapper when the result is fetched)
    for (int i = 0; i < res->n_patches; i++) {
        const hevcb_stitch_patch& p = res->patches[i];
        if (p.shard != shard || p.index < 0 || p.index >= cap) { continue; }
        if (p.set_start) { ns[p.index] = p.nal_start; ro[p.index] = p.rbsp_off; }
        ne[p.index] = p.nal_end;
        re[p.index] = p.rbsp_end;
    }
}
} // namespace

extern "C" HEVCB_API int hevcb_stitch_apply_device(hevcb_ctx* ctx, const hevcb_shard_summary* d_records, int n_shards, int shard, int64_t* d_nal_start,
                                                   int64_t* d_nal_end, int64_t* d_rbsp_off, int64_t* d_rbsp_end, int64_t cap_nals,
                                                   hevcb_stitch_result* d_result, void* stream)
{
    if (!ctx || !d_records || !d_result || n_shards < 1 || n_shards > HEVCB_MAX_SHARDS || shard < 0 || shard >= n_shards || !d_nal_start || !d_nal_end ||
        !d_rbsp_off || !d_rbsp_end) {
        return HEVCB_E_ARG;
    }
    HEVCB_CUDA(ctx, cudaSetDevice(ctx->device));
    stitch_apply_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(d_records, n_shards, shard, d_nal_start, d_nal_end, d_rbsp_off, d_rbsp_end, cap_nals, d_result);
    ctx->launches++;
    HEVCB_CUDA(ctx, cudaGetLastError());
    return HEVCB_OK;
}
