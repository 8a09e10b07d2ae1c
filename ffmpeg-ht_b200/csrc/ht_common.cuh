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

constexpr int WARPS_PER_CTA = 4;
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

/* ---- H2: VLC.  The backward byte stream (jpeg2000htdec.c:145-201, :283-295: bytes from Dcup[Lcup-2]
 * down to Dcup[Pcup]; a byte whose low 7 bits are ones gives 7 bits when the byte read before it is
 * > 0x8F) is un-stuffed once, in parallel, into a forward LSB-first bit array in shared memory.  The
 * serial decoder fetches 64 bits at the current bit position once per quad pair (a pair consumes at most
 * 7 + 7 + 24 bits); everything after that is 32-bit shifts.  Past the end the stream reads as zeros. */
struct VlcWin {
    uint32_t lo, hi;
};

__device__ __forceinline__ VlcWin vlc_fetch(const uint32_t *vl, uint32_t pos, uint32_t limit)
{
    pos = min(pos, limit);
    const uint32_t *p = vl + (pos >> 5);
    const uint32_t s = pos & 31;
    const uint32_t a = p[0], b = p[1], c = p[2];
    VlcWin w;
    w.lo = __funnelshift_r(a, b, s);
    w.hi = __funnelshift_r(b, c, s);
    return w;
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

/*
 * The serial, warp-uniform part for one segment of up to 32 quads of one quad row: MEL + CxtVLC + U-VLC
 * (jpeg2000htdec.c:632-768 first row, :784-973 other rows).  Every lane executes it on identical data and
 * stores quad j's result in qinfo[j] = u_off | rho<<1 | emb_k<<5 | emb_1<<9 | u<<16 (shared memory,
 * same-address broadcast stores).
 * carry: first row -> context carried from the previous quad; other rows -> (sigma_W[3] | sigma_W[2]).
 * mA / mC: other rows -> per-quad bits (sigma_N[1] | sigma_NW[3]) and (sigma_N[3] | sigma_NE[1]).
 */
template <bool FIRST, bool PAIR>
__device__ __forceinline__ void decode_pair(Mel &mel, const uint32_t *vl, uint32_t &vpos, uint32_t vlimit,
                                            const uint16_t *table, const uint16_t *uvtab, unsigned mA, unsigned mC,
                                            int &carry, uint32_t *qout)
{
    const VlcWin vw = vlc_fetch(vl, vpos, vlimit);
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
__device__ __forceinline__ void decode_segment(Mel &mel, const uint32_t *vl, uint32_t &vpos, uint32_t vlimit,
                                               const uint16_t *table, const uint16_t *uvtab, int nq, unsigned mA,
                                               unsigned mC, int &carry, uint32_t *qinfo)
{
    const int npairs = nq >> 1;
    for (int j = 0; j < npairs; j++) {
        decode_pair<FIRST, true>(mel, vl, vpos, vlimit, table, uvtab, mA, mC, carry, qinfo + 2 * j);
        mA >>= 2;
        mC >>= 2;
    }
    if (nq & 1)
        decode_pair<FIRST, false>(mel, vl, vpos, vlimit, table, uvtab, mA, mC, carry, qinfo + 2 * npairs);
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

}  // namespace
}  // namespace b200j2k
#endif
