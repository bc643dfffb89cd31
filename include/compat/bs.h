/*
 * bs.h -- drop-in for the reference's header-only bit reader / writer (leslie-wang/hevcbitstream bs.h:34-382): the same `bs_t`
 * layout, the same function names, signatures and observable behaviour, written from scratch for hevcbitstream-b200.
 *
 * Behaviour kept on purpose:
 *   - the cursor is (p, bits_left) with bits_left in 8..1; reads past `end` return 0 bits but still advance the cursor,
 *     writes past `end` are dropped but still advance it; bs_overrun() <=> p strictly beyond end; bs_pos() is in whole bytes;
 *   - bs_read_ue counts at most 32 leading zeros and also stops when the bit it just consumed made bs_eof() true; a 32-bit
 *     prefix adds nothing to the suffix (what `1 << 32` evaluates to on the x86 reference);
 *   - bs_read_u8 / bs_write_u8 take the byte-aligned fast path of FAST_U8 (same bits either way).
 * Unlike the reference, multi-bit accesses work on whole bytes at a time instead of one call per bit.  This header is host
 * code for callers that include it (more_rbsp_data and friends take a bs_t*); the batched CUDA parser has its own device
 * reader (hevcb_syntax.h).
 */
#ifndef _H264_BS_H
#define _H264_BS_H 1

#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct
{
    uint8_t* start;
    uint8_t* p;
    uint8_t* end;
    int bits_left;
} bs_t;

#define _OPTIMIZE_BS_ 1
#ifndef FAST_U8
#define FAST_U8
#endif

static inline bs_t* bs_init(bs_t* b, uint8_t* buf, size_t size)
{
    b->start = buf;
    b->p = buf;
    b->end = buf + size;
    b->bits_left = 8;
    return b;
}
static inline bs_t* bs_new(uint8_t* buf, size_t size) { return bs_init((bs_t*)malloc(sizeof(bs_t)), buf, size); }
static inline void bs_free(bs_t* b) { free(b); }
/* the clone starts where the source's cursor is (its `start` is the source's p) */
static inline bs_t* bs_clone(bs_t* dest, const bs_t* src)
{
    dest->start = src->p;
    dest->p = src->p;
    dest->end = src->end;
    dest->bits_left = src->bits_left;
    return dest;
}
static inline uint32_t bs_byte_aligned(bs_t* b) { return b->bits_left == 8; }
static inline int bs_eof(bs_t* b) { return b->p >= b->end ? 1 : 0; }
static inline int bs_overrun(bs_t* b) { return b->p > b->end ? 1 : 0; }
static inline int bs_pos(bs_t* b) { return (int)((b->p > b->end ? b->end : b->p) - b->start); }
static inline int bs_pos_out(bs_t* b) { return (int)(b->p - b->start); }
static inline int bs_bytes_left(bs_t* b) { return (int)(b->end - b->p); }

/* advance the cursor by n bits inside the current byte (n <= bits_left) */
static inline void bs__advance(bs_t* b, int n)
{
    b->bits_left -= n;
    if (b->bits_left == 0) { b->p++; b->bits_left = 8; }
}

static inline uint32_t bs_read_u1(bs_t* b)
{
    const uint32_t r = (b->p < b->end) ? (uint32_t)((*b->p >> (b->bits_left - 1)) & 1u) : 0u;
    bs__advance(b, 1);
    return r;
}
static inline void bs_skip_u1(bs_t* b) { bs__advance(b, 1); }
static inline uint32_t bs_peek_u1(bs_t* b) { return (b->p < b->end) ? (uint32_t)((*b->p >> (b->bits_left - 1)) & 1u) : 0u; }

/* n bits, most significant first, taken byte-wise; widths above 32 keep the low 32 bits */
static inline uint32_t bs_read_u(bs_t* b, int n)
{
    uint32_t r = 0;
    while (n > 0) {
        const int take = n < b->bits_left ? n : b->bits_left;
        const uint32_t cur = (b->p < b->end) ? *b->p : 0u;
        const uint32_t bits = (cur >> (b->bits_left - take)) & ((1u << take) - 1u);
        r = (r << take) | bits;
        bs__advance(b, take);
        n -= take;
    }
    return r;
}
static inline void bs_skip_u(bs_t* b, int n)
{
    while (n > 0) {
        const int take = n < b->bits_left ? n : b->bits_left;
        bs__advance(b, take);
        n -= take;
    }
}
static inline uint32_t bs_read_f(bs_t* b, int n) { return bs_read_u(b, n); }
static inline uint32_t bs_read_u8(bs_t* b)
{
    if (b->bits_left == 8 && b->p < b->end) { return *b->p++; }
    return bs_read_u(b, 8);
}
static inline uint32_t bs_read_ue(bs_t* b)
{
    int zeros = 0;
    /* the terminating bit is consumed; the count stops at 32 and when that bit was the last one of the buffer */
    while (bs_read_u1(b) == 0 && zeros < 32 && !bs_eof(b)) { zeros++; }
    const uint32_t suffix = bs_read_u(b, zeros);
    return suffix + (zeros < 32 ? ((1u << zeros) - 1u) : 0u);
}
static inline int32_t bs_read_se(bs_t* b)
{
    const int32_t k = (int32_t)bs_read_ue(b);
    return (k & 1) ? (int32_t)((uint32_t)k + 1u) / 2 : -(k / 2); /* wraps for k = INT_MAX like the x86 reference */
}

static inline void bs_write_u1(bs_t* b, uint32_t v)
{
    if (b->p < b->end) {
        const uint8_t mask = (uint8_t)(1u << (b->bits_left - 1));
        *b->p = (uint8_t)((*b->p & ~mask) | ((v & 1u) ? mask : 0u));
    }
    bs__advance(b, 1);
}
static inline void bs_write_u(bs_t* b, int n, uint32_t v)
{
    int i;
    if (n > 32) { /* shift counts wrap on the x86 reference: bit i comes from (v >> ((n - i - 1) & 31)) */
        for (i = 0; i < n; i++) { bs_write_u1(b, (v >> ((n - i - 1) & 31)) & 1u); }
        return;
    }
    while (n > 0) {
        const int take = n < b->bits_left ? n : b->bits_left;
        if (b->p < b->end) {
            const uint32_t field = (v >> (n - take)) & ((1u << take) - 1u);
            const int sh = b->bits_left - take;
            const uint8_t mask = (uint8_t)(((1u << take) - 1u) << sh);
            *b->p = (uint8_t)((*b->p & ~mask) | (uint8_t)(field << sh));
        }
        bs__advance(b, take);
        n -= take;
    }
}
static inline void bs_write_f(bs_t* b, int n, uint32_t v) { bs_write_u(b, n, v); }
static inline void bs_write_u8(bs_t* b, uint32_t v)
{
    if (b->bits_left == 8 && b->p < b->end) { *b->p++ = (uint8_t)v; return; }
    bs_write_u(b, 8, v);
}
static inline void bs_write_ue(bs_t* b, uint32_t v)
{
    if (v == 0) { bs_write_u1(b, 1); return; }
    v++;
    { /* len = bit length of v + 1 (1 when v + 1 wrapped to 0, as the reference's table lookup yields) */
        int len = 1;
        uint32_t t = v;
        if (t) { len = 0; while (t) { len++; t >>= 1; } }
        bs_write_u(b, 2 * len - 1, v);
    }
}
static inline void bs_write_se(bs_t* b, int32_t v)
{
    if (v <= 0) { bs_write_ue(b, (uint32_t)(-v * 2)); } else { bs_write_ue(b, (uint32_t)(v * 2 - 1)); }
}

/* byte operations: the count actually transferred is clamped to the buffer, the cursor moves by the full request */
static inline int bs__clamp(bs_t* b, int len)
{
    int n = len;
    if (b->end - b->p < n) { n = (int)(b->end - b->p); }
    return n < 0 ? 0 : n;
}
static inline int bs_read_bytes(bs_t* b, uint8_t* buf, int len)
{
    const int n = bs__clamp(b, len);
    memcpy(buf, b->p, (size_t)n);
    b->p += len < 0 ? 0 : len;
    return n;
}
static inline int bs_write_bytes(bs_t* b, uint8_t* buf, int len)
{
    const int n = bs__clamp(b, len);
    memcpy(b->p, buf, (size_t)n);
    b->p += len < 0 ? 0 : len;
    return n;
}
static inline int bs_skip_bytes(bs_t* b, int len)
{
    const int n = bs__clamp(b, len);
    b->p += len < 0 ? 0 : len;
    return n;
}
static inline uint32_t bs_next_bits(bs_t* bs, int nbits)
{
    bs_t t;
    bs_clone(&t, bs);
    return bs_read_u(&t, nbits);
}
static inline uint64_t bs_next_bytes(bs_t* bs, int nbytes)
{
    uint64_t v = 0;
    int i;
    if (nbytes > 8 || nbytes < 1 || bs->p + nbytes > bs->end) { return 0; }
    for (i = 0; i < nbytes; i++) { v = (v << 8) | bs->p[i]; }
    return v;
}

#define bs_print_state(b) fprintf(stderr, "%s:%d@%s: b->p=0x%02hhX, b->left = %d\n", __FILE__, __LINE__, __FUNCTION__, *b->p, b->bits_left)

#ifdef __cplusplus
}
#endif

#endif
