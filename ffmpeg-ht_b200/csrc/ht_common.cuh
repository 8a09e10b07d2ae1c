/*
 * ht_common.cuh -- device helpers shared by the HT block-decode kernels (ht_block.cu: the general
 * warp-per-block kernel; ht_cleanup.cu: the split un-stuff / VLC / MagSgn pipeline).
 * References: libavcodec/jpeg2000htdec.c (MEL :429-495, VLC :145-201,:283-331, U-VLC :338-388) and
 * libavcodec/jpeg2000dec.c:2098-2181 (dequantisation).
 */
#ifndef B200J2K_HT_COMMON_CUH
#define B200J2K_HT_COMMON_CUH

#include <cuda_runtime.h>
#include <stdint.h>

#define CXTVLC_TABLE_ATTR __device__
#include "cxtvlc_tables.h"
#include "j2k_kernels.h"
#include "j2k_plan.h"

namespace b200j2k {
namespace {

#ifdef B200J2K_WARPS_PER_CTA                     /* tuning builds (tools/build_variant.sh) */
constexpr int WARPS_PER_CTA = B200J2K_WARPS_PER_CTA;
#else
constexpr int WARPS_PER_CTA = 4;
#endif
/* Bounds assertions of a debug build (-DB200J2K_DEBUG_BOUNDS; tools/build_variant.sh dbg ht_cleanup.cu ...): the
 * fuzz tests run against that build with guard zones around every device buffer (B200J2K_REDZONE=1) -- the stand-in
 * for compute-sanitizer, which is not available on the GPU pool.  A failed assertion is counted, the decode of the
 * frame then fails on the host (decoder.cu finish_slot). */
#ifdef B200J2K_DEBUG_BOUNDS
__device__ int g_bounds_failures;
#define J2K_BOUNDS(cond) do { if (!(cond)) atomicAdd(&g_bounds_failures, 1); } while (0)
#else
#define J2K_BOUNDS(cond) do { } while (0)
#endif
constexpr int TABLE_WORDS = 1024 + 128;         /* 2 x 1024 x u16 CxtVLC + 256 x u16 U-VLC pair table */
constexpr unsigned FULL = 0xFFFFFFFFu;

__device__ __constant__ uint8_t c_mel_exp[13] = { 0, 0, 0, 1, 1, 1, 2, 2, 2, 3, 3, 4, 5 };

/* ---- H3: MEL (jpeg2000htdec.c:429-440, :462-495): MSB-first bits over Dcup[Pcup..Lcup), 7 bits after
 * 0xFF, 0xFF past the end.  The few MEL bytes are read straight from global memory (L1-resident after
 * the first touch); modDcup (:1277-1278) is applied on the fly to the last two bytes. */
struct Mel {
    const uint8_t *p;       /* Dcup + Pcup */
    int len;                /* Scup */
    int pos, left;
    unsigned cur;
    int k, run, one;
};

__device__ __forceinline__ int mel_bit(Mel &m)
{
    if (m.left == 0) {
        m.left = m.cur == 0xFF ? 7 : 8;
        unsigned b = 0xFF;
        if (m.pos < m.len) {
            b = m.p[m.pos];
            if (m.pos == m.len - 2)
                b |= 0x0F;
            else if (m.pos == m.len - 1)
                b = 0xFF;
            m.pos++;
        }
        m.cur = b;
    }
    m.left--;
    return (m.cur >> m.left) & 1;
}

__device__ __forceinline__ int mel_symbol(Mel &m)
{
    if (m.run == 0 && !m.one) {
        int e = c_mel_exp[m.k];
        if (mel_bit(m)) {
            m.run = 1 << e;
            m.k = min(12, m.k + 1);
        } else {
            int r = 0;
            while (e-- > 0)
                r = 2 * r + mel_bit(m);
            m.run = r;
            m.k = max(0, m.k - 1);
            m.one = 1;
        }
    }
    if (m.run > 0) {
        m.run--;
        return 0;
    }
    m.one = 0;
    return 1;
}

/* ---- H5: U-VLC (jpeg2000htdec.c:338-388) on a 32-bit window ------------------------------------------ */
__device__ __forceinline__ int uvlc_prefix(uint32_t &win, int &used)
{
    const int f = __ffs((win & 7) | 8);         /* 1..4: position of the first 1 among three bits */
    const int len = min(f, 3);
    win >>= len;
    used += len;
    return f <= 3 ? f : 5;
}

__device__ __forceinline__ int uvlc_suffix(uint32_t &win, int &used, int pfx)
{
    const int len = pfx < 3 ? 0 : (pfx == 3 ? 1 : 5);
    const int v = win & ((1u << len) - 1);
    win >>= len;
    used += len;
    return v;
}

__device__ __forceinline__ int uvlc_ext(uint32_t &win, int &used, int sfx)
{
    const int len = sfx >= 28 ? 4 : 0;
    const int v = win & ((1u << len) - 1);
    win >>= len;
    used += len;
    return v;
}

__device__ __forceinline__ int uvlc_whole(uint32_t &win, int &used)
{
    const int p = uvlc_prefix(win, used);
    const int s = uvlc_suffix(win, used, p);
    return p + s + 4 * uvlc_ext(win, used, s);
}

/* U-VLC pair table for the non-initial quad rows (:833-854): index = mode << 6 | six stream bits, mode =
 * u_off0 | u_off1 << 1; entry = total prefix length | prefix0 << 3 | prefix1 << 6 | suffix length0 << 9 |
 * suffix length1 << 12.  Built once per CTA in shared memory. */
__device__ __forceinline__ uint16_t uvlc_pair_entry(int idx)
{
    const int mode = idx >> 6;
    uint32_t w = idx & 63;
    int used = 0, p0 = 0, p1 = 0;
    if (mode & 1)
        p0 = uvlc_prefix(w, used);
    if (mode & 2)
        p1 = uvlc_prefix(w, used);
    const int s0 = p0 < 3 ? 0 : (p0 == 3 ? 1 : 5), s1 = p1 < 3 ? 0 : (p1 == 3 ? 1 : 5);
    return (uint16_t)(used | (p0 << 3) | (p1 << 6) | (s0 << 9) | (s1 << 12));
}

/* read nb (1..31) bits at bit position pos of the un-stuffed MagSgn array */
__device__ __forceinline__ uint32_t ms_read(const uint32_t *ms, uint32_t pos, int nb)
{
    uint32_t w = pos >> 5, s = pos & 31;
    uint32_t lo = ms[w], hi = ms[w + 1];
    uint32_t v = __funnelshift_r(lo, hi, s);
    return v & ((1u << nb) - 1);
}

/* D1-D3: sign-magnitude sample -> plane value (as raw 32 bits) */
__device__ __forceinline__ uint32_t dequant(uint32_t sm, int mode, int M_b, int roi_shift, const BandDesc &bd)
{
    /* ROI up-shift, then "val | sign" (jpeg2000htdec.c:1318-1331) */
    uint32_t t = sm & 0x7FFFFFFFu;
    if (roi_shift) {
        const uint32_t mask = 0xFFFFFFFFu >> (M_b + 1);
        if ((t & ~mask) == 0)
            t <<= roi_shift;
    }
    const uint32_t c = t | (sm & 0x80000000u);
    const bool neg = (c >> 31) != 0;
    const uint32_t mag = c & 0x7FFFFFFFu;
    if (mode == 3)                                   /* stage-level entry point: t1->data as is */
        return c;
    if (mode == J2K_DWT53) {
        int32_t v = (int32_t)(mag >> (31 - M_b));
        if (neg)
            v = -v;
        if (bd.i_stepsize != 32768)
            v = (int32_t)(((int64_t)v * bd.i_stepsize) / 65536);
        return (uint32_t)v;
    } else if (mode == J2K_DWT97) {
        int32_t v = neg ? -(int32_t)mag : (int32_t)mag;
        return __float_as_uint(__fmul_rn((float)v, bd.fscale));
    } else {
        int32_t v = neg ? -(int32_t)mag : (int32_t)mag;
        v = (v + 32) >> 6;
        int64_t m = (int64_t)v * bd.iscale;
        m = m > 0 ? (m + 32768) >> 16 : (m + 32767) >> 16;   /* RSHIFT(m, 16) */
        return (uint32_t)(int32_t)m;
    }
}

/* sample state bits in the refinement path (jpeg2000htdec.c:62-65) */
constexpr unsigned ST_SIGMA = 1, ST_NEWSIG = 8, ST_SCANNED = 16;

/* SigProp bit reader (jpeg2000htdec.c:442-460), global memory, lane 0 only */
struct SpBits {
    const uint8_t *d;
    int len, pos, left;
    unsigned cur, last;
};
__device__ __forceinline__ int sp_bit(SpBits &s)
{
    if (s.left == 0) {
        s.left = s.last == 0xFF ? 7 : 8;
        s.cur = s.pos < s.len ? s.d[s.pos++] : 0;
        s.last = s.cur;
    }
    int b = s.cur & 1;
    s.cur >>= 1;
    s.left--;
    return b;
}

/* four bytes p[0..3] (p[0] in the low byte) fetched as two aligned words; the caller guarantees that the
 * aligned words around p are readable (the byte buffer has 64 bytes of slack and a 256-byte aligned base) */
__device__ __forceinline__ uint32_t load4(const uint8_t *p)
{
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint32_t *w = reinterpret_cast<const uint32_t *>(a & ~(uintptr_t)3);
    return __funnelshift_r(w[0], w[1], (int)(a & 3) * 8);
}

/*
 * Parallel removal of byte stuffing.  The five bit streams of an HT code-block differ only in direction,
 * in the stuffing rule and in how a byte's bits are consumed; all become contiguous LSB-first bit arrays
 * (first bit consumed = bit 0 of word 0):
 *   US_MAGSGN  Dcup[0..Pcup) forward; 7 bits after 0xFF, bit 7 of such a byte is OR-ed into the next
 *              byte's bit 0 as the reference's accumulator does (jpeg2000htdec.c:207-221)
 *   US_MEL     Dcup[Pcup..Lcup) forward, MSB first, 7 bits after 0xFF, modDcup on the last two bytes
 *              (:429-440, :1277-1278)
 *   US_VLC     Dcup[Lcup-2] down to Dcup[Pcup]; 7 bits when (b & 0x7F) == 0x7F and the byte read before is
 *              > 0x8F; first byte |= 0x0F, the byte "read before" it is 0xFF (:145-201, :283-295)
 *   US_SIGPROP Dref[0..Lref) forward, LSB first, 7 bits after 0xFF (:442-460)
 *   US_MAGREF  Dref[Lref-1] down to Dref[0], VLC stuffing rule, first "read before" is 0xFF (:123-131)
 * first: address of the first byte in read order; n: bytes.  Four bytes per lane and chunk, fetched as two
 * aligned words one chunk ahead; a warp prefix sum over the 7/8-bit lengths places them.  out[] must be
 * zeroed.  Returns the number of bits written (warp-uniform).
 */
enum { US_MAGSGN, US_MEL, US_VLC, US_SIGPROP, US_MAGREF };

/* inclusive warp prefix sum; shfl.up's validity predicate guards the add (no lane compare per step) */
__device__ __forceinline__ int warp_scan_add(int v)
{
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
        asm volatile("{ .reg .pred p; .reg .s32 t;\n"
                     "  shfl.sync.up.b32 t|p, %0, %1, 0, 0xffffffff;\n"
                     "  @p add.s32 %0, %0, t; }"
                     : "+r"(v)
                     : "r"(o));
    return v;
}

/* 0x01 in every byte of x that equals the corresponding byte of y */
__device__ __forceinline__ uint64_t bytes_eq(uint64_t x, uint64_t y)
{
    uint64_t d = x ^ y;                                  /* zero byte <=> equal */
    d |= d >> 4;
    d |= d >> 2;
    d |= d >> 1;
    return ~d & 0x0101010101010101ull;
}

/* 0x01 in every byte of x that is > 0x8F */
__device__ __forceinline__ uint64_t bytes_gt8f(uint64_t x)
{
    /* b > 0x8F  <=>  bit 7 set and (b & 0x70) != 0 */
    const uint64_t hi = x >> 7, any = (x >> 6) | (x >> 5) | (x >> 4);
    return hi & any & 0x0101010101010101ull;
}

template <int KIND>
__device__ __forceinline__ uint32_t unstuff(const uint8_t *first, int n, uint32_t *out, int lane, unsigned before0 = 0xFF)
{
    constexpr bool BACKWARD = KIND == US_VLC || KIND == US_MAGREF;
    constexpr bool RULE_A = BACKWARD;           /* 7 bits when (b & 0x7F) == 0x7F && before > 0x8F; else: before == 0xFF */
    constexpr int BPL = 8;                      /* bytes per lane and chunk */
    uint32_t bits = 0;
    unsigned carry = RULE_A ? before0 : 0;       /* the byte "read before" the first one */
    /* eight bytes in read order, first byte read = low byte */
    auto fetch = [&](int idx) -> uint64_t {
        if (idx >= n)
            return 0;
        const uint8_t *p = BACKWARD ? first - 7 - idx : first + idx;
        const uintptr_t a = reinterpret_cast<uintptr_t>(p);
        const uint32_t *w = reinterpret_cast<const uint32_t *>(a & ~(uintptr_t)3);
        const int sh = (int)(a & 3) * 8;
        const uint32_t w0 = w[0], w1 = w[1], w2 = w[2];
        const uint32_t lo = __funnelshift_r(w0, w1, sh), hi = __funnelshift_r(w1, w2, sh);
        return BACKWARD ? ((uint64_t)__byte_perm(lo, 0, 0x0123) << 32) | __byte_perm(hi, 0, 0x0123)
                        : ((uint64_t)hi << 32) | lo;
    };
    uint64_t nxt = fetch(lane * BPL);
    for (int chunk = 0; chunk < n; chunk += 32 * BPL) {
        const int idx = chunk + lane * BPL;
        uint64_t x = nxt;
        nxt = fetch(idx + 32 * BPL);
        const int nvalid = min(max(n - idx, 0), BPL);
        const uint64_t vmask = nvalid >= 8 ? ~0ull : ((1ull << (8 * nvalid)) - 1);
        x &= vmask;
        if (KIND == US_VLC && idx == 0)
            x |= 0x0F;                                   /* modDcup: low nibble of the first byte read */
        if (KIND == US_MEL) {                            /* modDcup: last two bytes of the segment */
            const int k2 = n - 2 - idx, k1 = n - 1 - idx;
            if (k2 >= 0 && k2 < 8)
                x |= 0x0Full << (8 * k2);
            if (k1 >= 0 && k1 < 8)
                x |= 0xFFull << (8 * k1);
        }
        /* the byte read before each byte */
        unsigned before = __shfl_up_sync(FULL, (unsigned)(x >> 56), 1);
        if (lane == 0)
            before = carry;
        const uint64_t prev = (x << 8) | before;
        /* 0x01 per byte that carries only 7 bits */
        uint64_t st;
        if (RULE_A)
            st = bytes_eq(x | 0x8080808080808080ull, ~0ull) & bytes_gt8f(prev);
        else
            st = bytes_eq(prev, ~0ull);
        st &= vmask;
        const int tot = 8 * nvalid - __popcll(st);
        uint64_t v = x;
        if (KIND == US_MEL) {                            /* consumed MSB first: reverse the bits of every byte */
            const uint32_t lo = __byte_perm(__brev((uint32_t)x), 0, 0x0123), hi = __byte_perm(__brev((uint32_t)(x >> 32)), 0, 0x0123);
            v = ((uint64_t)hi << 32) | lo;
        }
        /* squeeze out the unused bit of every 7-bit byte, highest byte first so that lower positions stay valid */
        while (st) {
            const int k8 = 63 - __clzll((long long)st);  /* bit 8k */
            st ^= 1ull << k8;
            const int pos = KIND == US_MEL ? k8 : k8 + 7;
            const uint64_t low = (1ull << pos) - 1;
            const uint64_t keep = KIND == US_MAGSGN ? v & (1ull << pos) : 0;    /* OR-ed into the next byte's bit 0 */
            v = (v & low) | ((v >> 1) & ~low) | keep;
        }
        const int incl = warp_scan_add(tot);
        const uint32_t off = bits + (uint32_t)(incl - tot);
        if (tot) {
            const uint32_t wd = off >> 5, sft = off & 31;
            const uint64_t lo = v << sft;
            const uint32_t hi = sft ? (uint32_t)(v >> (64 - sft)) : 0;
            /* n bytes give at most 8 n bits: no word at or behind bit 8 n is ever touched */
            J2K_BOUNDS(32u * (wd + (hi ? 2u : (uint32_t)(lo >> 32) ? 1u : 0u)) < 8u * (uint32_t)n);
            if ((uint32_t)lo)
                atomicOr(&out[wd], (uint32_t)lo);
            if ((uint32_t)(lo >> 32))
                atomicOr(&out[wd + 1], (uint32_t)(lo >> 32));
            if (hi)
                atomicOr(&out[wd + 2], hi);
        }
        bits += (uint32_t)__shfl_sync(FULL, incl, 31);
        carry = __shfl_sync(FULL, (unsigned)(x >> 56), 31);
    }
    __syncwarp();
    return bits;
}

/* fill everything behind `bits` with ones (streams that read 0xFF past their end) */
__device__ __forceinline__ void ones_behind(uint32_t *out, uint32_t bits, int words, int lane)
{
    const int first = (int)(bits >> 5);
    if (lane == 0)
        out[first] |= 0xFFFFFFFFu << (bits & 31);
    for (int i = first + 1 + lane; i < words; i += 32)
        out[i] = 0xFFFFFFFFu;
    __syncwarp();
}

/*
 * What the reference's backward reader (jpeg2000_bitbuf_refill_backwards, jpeg2000htdec.c:145-201: VLC and
 * MagRef) yields once its R bytes base[R-1] .. base[0] are used up.  It pulls groups of four bytes; the last,
 * partial group is zero-filled and the position is parked at 0, so every later group is base[0] followed by
 * three zero bytes, base[0] giving 7 bits when (base[0] & 0x7F) == 0x7F and base[1] > 0x8F.  Only damaged
 * segments ever read that far, but then the decoded garbage has to match the reference's garbage (quirk Q1
 * keeps such blocks when no exponent bound trips).
 *   Bv   bits of the R un-stuffed bytes plus the zero bytes that complete the partial group (a whole group
 *        of zeros when R is a multiple of four)
 *   pat  the 7 or 8 bits of base[0], P = their number + 24: the period of everything behind Bv
 */
struct BwdTail {
    uint32_t Bv, pat;
    int P;
};

__device__ __forceinline__ BwdTail bwd_tail_make(uint32_t data_bits, int R, unsigned b0, unsigned b1)
{
    BwdTail t;
    t.Bv = data_bits + 8u * (uint32_t)(4 - (R & 3));
    const int nb0 = ((b0 & 0x7F) == 0x7F && b1 > 0x8F) ? 7 : 8;
    t.pat = b0 & ((1u << nb0) - 1);
    t.P = nb0 + 24;
    return t;
}

/* 32 bits of the periodic tail starting `t` bits behind Bv.  Out of line on purpose: only damaged segments
 * get here, and inlined it would be if-converted into the hot refill path of the serial decoders. */
__device__ __noinline__ uint32_t bwd_tail_word(uint32_t pat, int P, uint32_t t)
{
    const uint32_t phase = t % (uint32_t)P;
    const uint64_t pp = (uint64_t)pat | ((uint64_t)pat << P);               /* two periods: phase + 32 <= 2 P */
    return (uint32_t)(pp >> phase);
}
__device__ __forceinline__ uint32_t bwd_tail_word(const BwdTail &tl, uint32_t t) { return bwd_tail_word(tl.pat, tl.P, t); }

/* one word describing the tail, kept in the last word of an un-stuffed backward array */
__device__ __forceinline__ uint32_t bwd_tail_pack(const BwdTail &t) { return t.Bv | (t.pat << 20) | ((uint32_t)(t.P == 31) << 28); }
__device__ __forceinline__ BwdTail bwd_tail_unpack(uint32_t w)
{
    BwdTail t;
    t.Bv = w & 0xFFFFF;                          /* < 2^20: segments are shorter than 65536 bytes */
    t.pat = (w >> 20) & 0xFF;
    t.P = (w >> 28) & 1 ? 31 : 32;
    return t;
}

/* out[0 .. words-1): the zeroed array already holding the un-stuffed bytes; completes it with the tail
 * pattern and stores the descriptor in out[words-1] */
__device__ __forceinline__ void bwd_tail_fill(uint32_t *out, int words, const BwdTail &tl, int lane)
{
    const int first = (int)(tl.Bv >> 5);
    for (int i = first + lane; i < words - 1; i += 32) {
        const int t0 = 32 * i - (int)tl.Bv;
        if (t0 < 0)
            out[i] |= bwd_tail_word(tl, 0) << (tl.Bv & 31);
        else
            out[i] = bwd_tail_word(tl, (uint32_t)t0);
    }
    if (lane == 0)
        out[words - 1] = bwd_tail_pack(tl);
    __syncwarp();
}

/* word i of a backward array of `words` words (the last one is the descriptor) */
__device__ __forceinline__ uint32_t bwd_word(const uint32_t *arr, int words, const BwdTail &tl, int i)
{
    return i < words - 1 ? arr[i] : bwd_tail_word(tl, (uint32_t)(32 * i) - tl.Bv);
}

/* ---- H2: VLC window of the warp-per-block kernel.  The backward byte stream is un-stuffed once, in
 * parallel, into a forward LSB-first bit array in shared memory (unstuff<US_VLC>), completed with the
 * reference's tail (BwdTail).  The serial decoder fetches 64 bits at the current bit position once per quad
 * pair (a pair consumes at most 7 + 7 + 24 bits); everything after that is 32-bit shifts. */
struct VlcWin {
    uint32_t lo, hi;
};
struct VlcArr {
    const uint32_t *vl;
    int words;
    BwdTail tail;
};

__device__ __forceinline__ VlcWin vlc_fetch(const VlcArr &va, uint32_t pos)
{
    const int i = (int)(pos >> 5);
    const uint32_t s = pos & 31;
    const uint32_t a = bwd_word(va.vl, va.words, va.tail, i), b = bwd_word(va.vl, va.words, va.tail, i + 1),
                   c = bwd_word(va.vl, va.words, va.tail, i + 2);
    VlcWin w;
    w.lo = __funnelshift_r(a, b, s);
    w.hi = __funnelshift_r(b, c, s);
    return w;
}

/*
 * The serial, warp-uniform part for one segment of up to 32 quads of one quad row: MEL + CxtVLC + U-VLC
 * (jpeg2000htdec.c:632-768 first row, :784-973 other rows).  Every lane executes it on identical data and
 * stores quad j's result in qinfo[j] = u_off | rho<<1 | emb_k<<5 | emb_1<<9 | u<<16 (shared memory,
 * same-address broadcast stores).
 * carry: first row -> context carried from the previous quad; other rows -> (sigma_W[3] | sigma_W[2]).
 * mA / mC: other rows -> per-quad bits (sigma_N[1] | sigma_NW[3]) and (sigma_N[3] | sigma_NE[1]).
 */
template <bool FIRST, bool PAIR>
__device__ __forceinline__ void decode_pair(Mel &mel, const VlcArr &vl, uint32_t &vpos,
                                            const uint16_t *table, const uint16_t *uvtab, unsigned mA, unsigned mC,
                                            int &carry, uint32_t *qout)
{
    const VlcWin vw = vlc_fetch(vl, vpos);
    uint32_t win = vw.lo;
    int used;
    unsigned q0, q1 = 0;
    {
        const int ctx = FIRST ? carry : (int)((mA & 1) | (carry << 1) | ((mC & 1) << 2));
        bool coded = ctx != 0;
        if (!coded)
            coded = mel_symbol(mel) != 0;
        unsigned e = table[(win & 0x7F) | ((unsigned)ctx << 7)];
        e = coded ? e : 0;
        used = e & 7;
        win >>= used;
        q0 = e >> 3;
        carry = FIRST ? (int)(((q0 & 6) ? 1 : 0) | ((q0 >> 2) & 6)) : (int)((q0 & 0x18) != 0);
    }
    if (PAIR) {
        const int ctx = FIRST ? carry : (int)(((mA >> 1) & 1) | (carry << 1) | (((mC >> 1) & 1) << 2));
        bool coded = ctx != 0;
        if (!coded)
            coded = mel_symbol(mel) != 0;
        unsigned e = table[(win & 0x7F) | ((unsigned)ctx << 7)];
        e = coded ? e : 0;
        const int l = e & 7;
        used += l;
        q1 = e >> 3;
        carry = FIRST ? (int)(((q1 & 6) ? 1 : 0) | ((q1 >> 2) & 6)) : (int)((q1 & 0x18) != 0);
    }
    int u0 = 0, u1 = 0;
    const int mode = (int)((q0 & 1) | ((q1 & 1) << 1));
    if (mode) {
        uint32_t uw = __funnelshift_r(vw.lo, vw.hi, used);      /* used <= 14 */
        if (FIRST) {
            const int uo0 = mode & 1, uo1 = mode >> 1;
            if (uo0 & uo1) {                             /* :671-704 */
                if (mel_symbol(mel)) {
                    const int p0 = uvlc_prefix(uw, used), p1 = uvlc_prefix(uw, used);
                    const int s0 = uvlc_suffix(uw, used, p0), s1 = uvlc_suffix(uw, used, p1);
                    const int e0 = uvlc_ext(uw, used, s0), e1 = uvlc_ext(uw, used, s1);
                    u0 = 2 + p0 + s0 + 4 * e0;
                    u1 = 2 + p1 + s1 + 4 * e1;
                } else {
                    const int p0 = uvlc_prefix(uw, used);
                    if (p0 > 2) {
                        u1 = (int)(uw & 1) + 1;
                        uw >>= 1;
                        used += 1;
                        const int s0 = uvlc_suffix(uw, used, p0);
                        u0 = p0 + s0 + 4 * uvlc_ext(uw, used, s0);
                    } else {
                        const int p1 = uvlc_prefix(uw, used);
                        const int s0 = uvlc_suffix(uw, used, p0), s1 = uvlc_suffix(uw, used, p1);
                        const int e0 = uvlc_ext(uw, used, s0), e1 = uvlc_ext(uw, used, s1);
                        u0 = p0 + s0 + 4 * e0;
                        u1 = p1 + s1 + 4 * e1;
                    }
                }
            } else if (uo0) {
                u0 = uvlc_whole(uw, used);
            } else {
                u1 = uvlc_whole(uw, used);
            }
        } else {                                         /* :833-854, table driven */
            const unsigned t = uvtab[(mode << 6) | (uw & 63)];
            const int plen = t & 7, sl0 = (t >> 9) & 7, sl1 = (t >> 12) & 7;
            uint32_t sw = uw >> plen;
            const int s0 = sw & ((1u << sl0) - 1);
            sw >>= sl0;
            const int s1 = sw & ((1u << sl1) - 1);
            used += plen + sl0 + sl1;
            u0 = ((t >> 3) & 7) + s0;
            u1 = ((t >> 6) & 7) + s1;
            if ((s0 | s1) >= 28) {                       /* rare: 4-bit extensions, quad 0 first */
                sw >>= sl1;
                if (s0 >= 28) {
                    u0 += 4 * (int)(sw & 15);
                    sw >>= 4;
                    used += 4;
                }
                if (s1 >= 28) {
                    u1 += 4 * (int)(sw & 15);
                    used += 4;
                }
            }
        }
    }
    vpos += used;
    qout[0] = q0 | ((unsigned)u0 << 16);
    if (PAIR)
        qout[1] = q1 | ((unsigned)u1 << 16);
}

template <bool FIRST>
__device__ __forceinline__ void decode_segment(Mel &mel, const VlcArr &vl, uint32_t &vpos,
                                               const uint16_t *table, const uint16_t *uvtab, int nq, unsigned mA,
                                               unsigned mC, int &carry, uint32_t *qinfo)
{
    const int npairs = nq >> 1;
    for (int j = 0; j < npairs; j++) {
        decode_pair<FIRST, true>(mel, vl, vpos, table, uvtab, mA, mC, carry, qinfo + 2 * j);
        mA >>= 2;
        mC >>= 2;
    }
    if (nq & 1)
        decode_pair<FIRST, false>(mel, vl, vpos, table, uvtab, mA, mC, carry, qinfo + 2 * npairs);
}

}  // namespace
}  // namespace b200j2k
#endif
