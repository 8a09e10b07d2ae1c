/*
 * ht_cleanup.cu -- the HT cleanup pass (Rec. ITU-T T.814) as a three-kernel pipeline for sm_100a.
 *
 * Replaces, for code-blocks up to 64 samples wide, jpeg2000_decode_ht_cleanup_segment
 * (libavcodec/jpeg2000htdec.c:547-1014) with its bit readers (:145-221, :283-295, :429-440), the MEL
 * decoder (:462-495), CxtVLC (:300-331, :510-531) and U-VLC (:338-388), plus the dequantisation of
 * jpeg2000dec.c:2098-2181 fused into the coefficient write-back.  SURVEY.md section 8(a) rows H1-H6, D1-D3.
 *
 * Why three kernels.  ncu on the warp-per-block kernel (ht_block.cu, profiles/r1_ncu_full_v1_summary.md)
 * shows it ISSUE-bound (70 % of issue slots) with 32 lanes executing the context-adaptive MEL / CxtVLC /
 * U-VLC chain redundantly: that chain is inherently serial inside a block but independent between blocks.
 * So the work is split by its shape:
 *   K0 ht_unstuff_kernel  warp per block   byte-stuffing removed from the VLC (backward) and MEL (forward)
 *                                          streams in parallel: warp prefix sum over 7/8-bit byte lengths;
 *                                          both become contiguous LSB-first bit arrays in HBM scratch
 *   K1 ht_vlc_kernel      LANE per block   the serial chain, 32 blocks per warp, 64-bit register windows
 *                                          with one-word prefetch, CxtVLC / U-VLC tables in shared memory;
 *                                          writes one word per quad: u_off | rho<<1 | emb_k<<5 | emb_1<<9 | u<<16
 *   K2 ht_magsgn_kernel   warp per block   lane = quad: exponent predictor from the row above (shuffles),
 *                                          warp prefix sum over the quads' MagSgn bit counts, four samples per
 *                                          lane extracted from the un-stuffed MagSgn array (shared memory),
 *                                          dequantised and stored as row segments
 * Blocks wider than 64 samples and blocks with SigProp / MagRef passes stay on the general kernel.
 */
#include "ht_common.cuh"

namespace b200j2k {

namespace {

/* H1 (jpeg2000htdec.c:1237-1273): Scup (>= 2) when the cleanup segment is decodable, 0 when there is
 * nothing to decode, -1 when the segment header is damaged (block = zeros, quirk Q1) */
__device__ __forceinline__ int ht_block_scup(const CblkDesc &d, const uint8_t *Dcup)
{
    if (d.npasses == 0)
        return 0;
    const int rem = d.npasses % 3;
    const int plhd = rem ? d.npasses - rem : d.npasses - 3;
    if (d.npasses - plhd <= 0)
        return 0;
    const int Lcup = d.Lcup;
    if (Lcup < 2)
        return -1;
    const int Scup = (Dcup[Lcup - 1] << 4) + (Dcup[Lcup - 2] & 0x0F);
    if (Scup < 2 || Scup > Lcup || Scup > 4079)
        return -1;
    const int S_blk = (plhd / 3 + d.zbp) & 0xFF;
    if (S_blk + 1 >= 32 || 30 - S_blk < 1)
        return -1;
    return Scup;
}

/* ---- K0 ------------------------------------------------------------------------------------------------ */
__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
ht_unstuff_kernel(const CblkDesc *__restrict__ descs, const uint32_t *__restrict__ order,
                  const uint32_t *__restrict__ scr_off, int nblocks, const uint8_t *__restrict__ bytes,
                  uint32_t *__restrict__ scratch, int warp_words)
{
    extern __shared__ __align__(16) uint32_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gw = blockIdx.x * WARPS_PER_CTA + warp;
    if (gw >= nblocks)
        return;
    const uint32_t bi = order[gw];
    const CblkDesc d = descs[bi];
    const int Lcup = d.Lcup;
    const uint8_t *Dcup = bytes + d.data_off;
    const int Scup = ht_block_scup(d, Dcup);
    if (Scup <= 0)
        return;
    const int Pcup = Lcup - Scup;
    const int vlw = ht_stream_words(Scup);
    uint32_t *vl = smem + (size_t)warp * warp_words;
    uint32_t *ml = vl + vlw;
    for (int i = lane; i < 2 * vlw; i += 32)
        vl[i] = 0;
    __syncwarp();
    /* VLC: read order k = 0.. is Dcup[Lcup-2-k]; a byte whose low 7 bits are ones carries 7 bits when the
     * byte read before it is > 0x8F (jpeg2000htdec.c:179-194); modDcup sets the low nibble of the first
     * byte and makes the byte "read before it" 0xFF (:1277-1278, :283-295) */
    {
        const int nv = Scup - 1;
        uint32_t bits = 0;
        unsigned carry = 0xFF;
        for (int chunk = 0; chunk < nv; chunk += 128) {
            const int idx = chunk + lane * 4;
            unsigned b[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                b[k] = idx + k < nv ? Dcup[Lcup - 2 - (idx + k)] : 0;
                if (idx + k == 0)
                    b[k] |= 0x0F;
            }
            unsigned before = __shfl_up_sync(FULL, b[3], 1);
            if (lane == 0)
                before = carry;
            uint64_t v = 0;
            int tot = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int nb = idx + k < nv ? (((b[k] & 0x7F) == 0x7F && before > 0x8F) ? 7 : 8) : 0;
                v |= (uint64_t)(b[k] & ((1u << nb) - 1)) << tot;
                tot += nb;
                before = b[k];
            }
            int incl = tot;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(FULL, incl, o);
                if (lane >= o)
                    incl += t;
            }
            const uint32_t off = bits + (uint32_t)(incl - tot);
            if (tot) {
                const uint32_t wd = off >> 5, sft = off & 31;
                const uint64_t sh = v << sft;
                atomicOr(&vl[wd], (uint32_t)sh);
                if ((uint32_t)(sh >> 32))
                    atomicOr(&vl[wd + 1], (uint32_t)(sh >> 32));
            }
            bits += (uint32_t)__shfl_sync(FULL, incl, 31);
            carry = __shfl_sync(FULL, b[3], 31);
        }
    }
    /* MEL: bytes Dcup[Pcup..Lcup) forward, MSB first, 7 bits from the byte after a 0xFF (:429-440), with
     * modDcup applied to the last two; stored in consumption order LSB first; ones past the end */
    {
        uint32_t bits = 0;
        unsigned carry = 0;
        for (int chunk = 0; chunk < Scup; chunk += 128) {
            const int idx = chunk + lane * 4;
            unsigned b[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                b[k] = idx + k < Scup ? Dcup[Pcup + idx + k] : 0;
                if (idx + k == Scup - 2)
                    b[k] |= 0x0F;
                else if (idx + k == Scup - 1)
                    b[k] = 0xFF;
            }
            unsigned before = __shfl_up_sync(FULL, b[3], 1);
            if (lane == 0)
                before = carry;
            uint64_t v = 0;
            int tot = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                if (idx + k < Scup) {
                    const int nb = before == 0xFF ? 7 : 8;
                    v |= (uint64_t)__brev(b[k] << (32 - nb)) << tot;
                    tot += nb;
                }
                before = b[k];
            }
            int incl = tot;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(FULL, incl, o);
                if (lane >= o)
                    incl += t;
            }
            const uint32_t off = bits + (uint32_t)(incl - tot);
            if (tot) {
                const uint32_t wd = off >> 5, sft = off & 31;
                const uint64_t sh = v << sft;
                atomicOr(&ml[wd], (uint32_t)sh);
                if ((uint32_t)(sh >> 32))
                    atomicOr(&ml[wd + 1], (uint32_t)(sh >> 32));
            }
            bits += (uint32_t)__shfl_sync(FULL, incl, 31);
            carry = __shfl_sync(FULL, b[3], 31);
        }
        __syncwarp();
        if (lane == 0) {
            ml[bits >> 5] |= 0xFFFFFFFFu << (bits & 31);
            for (int i = (int)(bits >> 5) + 1; i < vlw; i++)
                ml[i] = 0xFFFFFFFFu;
        }
    }
    __syncwarp();
    uint32_t *out = scratch + scr_off[bi];
    for (int i = lane; i < 2 * vlw; i += 32)
        out[i] = vl[i];
}

/* ---- K1 ------------------------------------------------------------------------------------------------ */
/* LSB-first bit reader over an un-stuffed array: a 64-bit window plus one prefetched word, so the load of
 * the next word is never on the dependent chain.  After refill() at least 33 bits are valid. */
struct LsbReader {
    uint64_t win;
    int avail;
    const uint32_t *p, *end;
    uint32_t nxt, fill;
    __device__ __forceinline__ void init(const uint32_t *ptr, int words, uint32_t f)
    {
        win = (uint64_t)ptr[0] | ((uint64_t)ptr[1] << 32);      /* words >= 3 by construction */
        avail = 64;
        nxt = ptr[2];
        p = ptr + 3;
        end = ptr + words;
        fill = f;
    }
    __device__ __forceinline__ void refill()
    {
        if (avail <= 32) {
            win |= (uint64_t)nxt << avail;
            avail += 32;
            nxt = p < end ? *p : fill;
            p++;
        }
    }
    __device__ __forceinline__ void drop(int n)
    {
        win >>= n;
        avail -= n;
    }
};

/* MEL_E[k] of jpeg2000htdec.c:68, one nibble per state */
constexpr uint64_t MEL_EXP_NIBBLES = 0x5433222111000ull;

struct MelLane {
    LsbReader r;
    int k, run, one;
};

/* H3: jpeg2000_decode_mel_sym (jpeg2000htdec.c:462-495) */
__device__ __forceinline__ int mel_symbol_lane(MelLane &m)
{
    if (m.run == 0 && !m.one) {
        m.r.refill();
        const int e = (int)((MEL_EXP_NIBBLES >> (4 * m.k)) & 15);
        const uint32_t w = (uint32_t)m.r.win;
        if (w & 1) {
            m.run = 1 << e;
            m.k = min(12, m.k + 1);
            m.r.drop(1);
        } else {
            m.run = (int)((__brev(w >> 1) >> 1) >> (31 - e));   /* e bits, first one most significant */
            m.k = max(0, m.k - 1);
            m.one = 1;
            m.r.drop(1 + e);
        }
    }
    if (m.run > 0) {
        m.run--;
        return 0;
    }
    m.one = 0;
    return 1;
}

/* One quad pair (jpeg2000htdec.c:632-768 first row, :784-973 other rows) decoded by ONE lane.
 * carry: first row -> context carried from the previous quad; other rows -> (sigma_W[3] | sigma_W[2]).
 * a0/c0, a1/c1: other rows -> (sigma_N[1] | sigma_NW[3]) and (sigma_N[3] | sigma_NE[1]) of the two quads. */
template <bool FIRST>
__device__ __forceinline__ void vlc_pair_lane(LsbReader &v, MelLane &mel, const uint16_t *table, const uint16_t *uvtab,
                                              bool two, unsigned a0, unsigned c0, unsigned a1, unsigned c1,
                                              int &carry, uint32_t &out0, uint32_t &out1)
{
    v.refill();
    uint32_t win = (uint32_t)v.win;
    unsigned q0, q1 = 0;
    int used;
    {
        const int ctx = FIRST ? carry : (int)(a0 | (carry << 1) | (c0 << 2));
        bool coded = ctx != 0;
        if (!coded)
            coded = mel_symbol_lane(mel) != 0;
        unsigned e = table[(win & 0x7F) | ((unsigned)ctx << 7)];
        e = coded ? e : 0;
        used = e & 7;
        win >>= used;
        q0 = e >> 3;
        carry = FIRST ? (int)(((q0 & 6) ? 1 : 0) | ((q0 >> 2) & 6)) : (int)((q0 & 0x18) != 0);
    }
    if (two) {
        const int ctx = FIRST ? carry : (int)(a1 | (carry << 1) | (c1 << 2));
        bool coded = ctx != 0;
        if (!coded)
            coded = mel_symbol_lane(mel) != 0;
        unsigned e = table[(win & 0x7F) | ((unsigned)ctx << 7)];
        e = coded ? e : 0;
        used += e & 7;
        q1 = e >> 3;
        carry = FIRST ? (int)(((q1 & 6) ? 1 : 0) | ((q1 >> 2) & 6)) : (int)((q1 & 0x18) != 0);
    }
    v.drop(used);
    int u0 = 0, u1 = 0;
    const int mode = (int)((q0 & 1) | ((q1 & 1) << 1));
    if (mode) {
        v.refill();
        uint32_t uw = (uint32_t)v.win;
        int ulen = 0;
        if (FIRST) {
            if (mode == 3) {                             /* :671-704 */
                if (mel_symbol_lane(mel)) {
                    const int p0 = uvlc_prefix(uw, ulen), p1 = uvlc_prefix(uw, ulen);
                    const int s0 = uvlc_suffix(uw, ulen, p0), s1 = uvlc_suffix(uw, ulen, p1);
                    const int e0 = uvlc_ext(uw, ulen, s0), e1 = uvlc_ext(uw, ulen, s1);
                    u0 = 2 + p0 + s0 + 4 * e0;
                    u1 = 2 + p1 + s1 + 4 * e1;
                } else {
                    const int p0 = uvlc_prefix(uw, ulen);
                    if (p0 > 2) {
                        u1 = (int)(uw & 1) + 1;
                        uw >>= 1;
                        ulen += 1;
                        const int s0 = uvlc_suffix(uw, ulen, p0);
                        u0 = p0 + s0 + 4 * uvlc_ext(uw, ulen, s0);
                    } else {
                        const int p1 = uvlc_prefix(uw, ulen);
                        const int s0 = uvlc_suffix(uw, ulen, p0), s1 = uvlc_suffix(uw, ulen, p1);
                        const int e0 = uvlc_ext(uw, ulen, s0), e1 = uvlc_ext(uw, ulen, s1);
                        u0 = p0 + s0 + 4 * e0;
                        u1 = p1 + s1 + 4 * e1;
                    }
                }
            } else if (mode == 1) {
                u0 = uvlc_whole(uw, ulen);
            } else {
                u1 = uvlc_whole(uw, ulen);
            }
        } else {                                         /* :833-854, table driven */
            const unsigned t = uvtab[(mode << 6) | (uw & 63)];
            const int plen = t & 7, sl0 = (t >> 9) & 7, sl1 = (t >> 12) & 7;
            uint32_t sw = uw >> plen;
            const int s0 = sw & ((1u << sl0) - 1);
            sw >>= sl0;
            const int s1 = sw & ((1u << sl1) - 1);
            ulen = plen + sl0 + sl1;
            u0 = ((t >> 3) & 7) + s0;
            u1 = ((t >> 6) & 7) + s1;
            if ((s0 | s1) >= 28) {                       /* rare: 4-bit extensions, quad 0 first */
                sw >>= sl1;
                if (s0 >= 28) {
                    u0 += 4 * (int)(sw & 15);
                    sw >>= 4;
                    ulen += 4;
                }
                if (s1 >= 28) {
                    u1 += 4 * (int)(sw & 15);
                    ulen += 4;
                }
            }
        }
        v.drop(ulen);
    }
    out0 = q0 | ((unsigned)u0 << 16);
    out1 = q1 | ((unsigned)u1 << 16);
}

constexpr int VLC_LANES_PER_CTA = 32;

__global__ void __launch_bounds__(VLC_LANES_PER_CTA)
ht_vlc_kernel(const CblkDesc *__restrict__ descs, const uint32_t *__restrict__ order,
              const uint32_t *__restrict__ scr_off, int nblocks, const uint8_t *__restrict__ bytes,
              uint32_t *__restrict__ scratch)
{
    __shared__ uint16_t tab[2048 + 256];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) {
        tab[i] = CXTVLC_FIRST_ROW[i];
        tab[1024 + i] = CXTVLC_OTHER_ROWS[i];
    }
    uint16_t *uvtab = tab + 2048;
    for (int i = threadIdx.x; i < 256; i += blockDim.x)
        uvtab[i] = uvlc_pair_entry(i);
    __syncthreads();
    const int gt = blockIdx.x * blockDim.x + threadIdx.x;
    if (gt >= nblocks)
        return;
    const uint32_t bi = order[gt];
    const CblkDesc d = descs[bi];
    const int Scup = ht_block_scup(d, bytes + d.data_off);
    if (Scup <= 0)
        return;
    const int QW = (d.w + 1) >> 1, QH = (d.h + 1) >> 1; /* QW <= 32: the row's significance fits two words */
    const int vlw = ht_stream_words(Scup);
    uint32_t *base = scratch + scr_off[bi];
    LsbReader v;
    v.init(base, vlw, 0);
    v.drop(4);                                           /* the Scup nibble (:283-295) */
    MelLane mel;
    mel.r.init(base + vlw, vlw, 0xFFFFFFFFu);
    mel.k = mel.run = mel.one = 0;
    uint32_t *q = base + 2 * vlw;

    uint32_t p1 = 0, p3 = 0;                             /* sigma bits 1 and 3 of every quad of the row above */
    {
        int carry = 0;
        for (int qx = 0; qx < QW; qx += 2) {
            const bool two = qx + 1 < QW;
            uint32_t o0, o1;
            vlc_pair_lane<true>(v, mel, tab, uvtab, two, 0, 0, 0, 0, carry, o0, o1);
            q[qx] = o0;
            if (two)
                q[qx + 1] = o1;
            p1 |= (((o0 >> 2) & 1) | (((o1 >> 2) & 1) << 1)) << qx;
            p3 |= (((o0 >> 4) & 1) | (((o1 >> 4) & 1) << 1)) << qx;
        }
    }
    for (int qy = 1; qy < QH; qy++) {
        q += QW;
        const uint32_t mA = p1 | (p3 << 1);              /* sigma_N[1] | sigma_NW[3] */
        const uint32_t mC = p3 | (p1 >> 1);              /* sigma_N[3] | sigma_NE[1] */
        p1 = p3 = 0;
        int carry = 0;
        for (int qx = 0; qx < QW; qx += 2) {
            const bool two = qx + 1 < QW;
            uint32_t o0, o1;
            const uint32_t sa = mA >> qx, sc = mC >> qx;
            vlc_pair_lane<false>(v, mel, tab + 1024, uvtab, two, sa & 1, sc & 1, (sa >> 1) & 1, (sc >> 1) & 1, carry,
                                 o0, o1);
            q[qx] = o0;
            if (two)
                q[qx + 1] = o1;
            p1 |= (((o0 >> 2) & 1) | (((o1 >> 2) & 1) << 1)) << qx;
            p3 |= (((o0 >> 4) & 1) | (((o1 >> 4) & 1) << 1)) << qx;
        }
    }
}

/* ---- K2 ------------------------------------------------------------------------------------------------ */
__device__ __forceinline__ void zero_block(uint32_t *dst, uint32_t dstride, int w, int h, int lane)
{
    for (int y = 0; y < h; y++)
        for (int x = lane; x < w; x += 32)
            dst[(size_t)y * dstride + x] = 0;
}

/* dequantisation variants selected per block (warp-uniform): the two common cases skip the ROI / step-size
 * branches of dequant() */
enum { DQ_GENERIC = 0, DQ_INT53 = 1, DQ_FLOAT97 = 2 };

template <int DQ>
__device__ __forceinline__ uint32_t dq(uint32_t mu, int sh53, float fscale, int mode, int M_b, int roi_shift, const BandDesc &bd)
{
    if (DQ == DQ_INT53) {                        /* dequantization_int with i_stepsize == 32768 (jpeg2000dec.c:2120-2151) */
        const int32_t v = (int32_t)((mu & 0x7FFFFFFFu) >> sh53);
        return (uint32_t)((int32_t)mu < 0 ? -v : v);
    } else if (DQ == DQ_FLOAT97) {               /* dequantization_float (:2098-2117) */
        const int32_t m = (int32_t)(mu & 0x7FFFFFFFu);
        return __float_as_uint(__fmul_rn((float)((int32_t)mu < 0 ? -m : m), fscale));
    }
    return dequant(mu, mode, M_b, roi_shift, bd);
}

template <int DQ>
__device__ __forceinline__ void magsgn_block(const CblkDesc &d, uint32_t bi, int Scup, const uint8_t *Dcup,
                                             const BandDesc &bd, const uint32_t *__restrict__ q, uint32_t *ms,
                                             uint32_t *__restrict__ dst, int32_t *__restrict__ status, int lane)
{
    const int w = d.w, h = d.h;
    const uint32_t dstride = d.dst_stride;
    const int mode = (d.flags >> 1) & 3;
    const int M_b = d.M_b;
    const int roi_shift = d.roi_shift;
    const int rem = d.npasses % 3;
    const int p0 = (rem ? d.npasses - rem : d.npasses - 3) / 3;
    const int S_blk = (p0 + d.zbp) & 0xFF;
    const int pLSB = 30 - S_blk;
    const int maxbp = S_blk + 1;
    const int Pcup = d.Lcup - Scup;
    const int sh53 = 31 - M_b;
    const float fscale = bd.fscale;

    /* un-stuff MagSgn: byte after 0xFF carries 7 bits (jpeg2000htdec.c:207-221); bit 7 of such a byte is
     * OR-ed into the next byte's bit 0 exactly as the reference's 64-bit accumulator does.  Four bytes per
     * lane, fetched as two aligned words. */
    const int ms_words = (Pcup + 3) / 4 + 2;
    for (int i = lane; i < ms_words; i += 32)
        ms[i] = 0;
    __syncwarp();
    uint32_t ms_bits = 0;
    {
        const uint32_t *wbase = reinterpret_cast<const uint32_t *>(reinterpret_cast<uintptr_t>(Dcup) & ~(uintptr_t)3);
        const int bsh = (int)(reinterpret_cast<uintptr_t>(Dcup) & 3) * 8;
        unsigned carry = 0;
        for (int chunk = 0; chunk < Pcup; chunk += 128) {
            const int idx = chunk + lane * 4;
            uint32_t b4 = 0;
            const int nvalid = min(4, Pcup - idx);
            if (nvalid > 0) {
                const uint32_t w0 = wbase[idx >> 2], w1 = wbase[(idx >> 2) + 1];
                b4 = __funnelshift_r(w0, w1, bsh);
                if (nvalid < 4)
                    b4 &= (1u << (8 * nvalid)) - 1;
            }
            unsigned before = __shfl_up_sync(FULL, b4 >> 24, 1);
            if (lane == 0)
                before = carry;
            uint64_t v = 0;
            int tot = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const unsigned bk = (b4 >> (8 * k)) & 0xFF;
                const int nb = k < nvalid ? (before == 0xFF ? 7 : 8) : 0;
                v |= (uint64_t)bk << tot;
                tot += nb;
                before = bk;
            }
            int incl = tot;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int t = __shfl_up_sync(FULL, incl, o);
                if (lane >= o)
                    incl += t;
            }
            const uint32_t off = ms_bits + (uint32_t)(incl - tot);
            if (tot) {
                const uint32_t wd = off >> 5, sft = off & 31;
                const uint64_t shv = v << sft;
                atomicOr(&ms[wd], (uint32_t)shv);
                if ((uint32_t)(shv >> 32))
                    atomicOr(&ms[wd + 1], (uint32_t)(shv >> 32));
            }
            ms_bits += (uint32_t)__shfl_sync(FULL, incl, 31);
            carry = __shfl_sync(FULL, b4 >> 24, 31);
        }
        __syncwarp();
        if (lane == 0) {                                 /* past the end the decoder reads ones */
            ms[ms_bits >> 5] |= 0xFFFFFFFFu << (ms_bits & 31);
            for (int i = (int)(ms_bits >> 5) + 1; i < ms_words; i++)
                ms[i] = 0xFFFFFFFFu;
        }
        __syncwarp();
    }

    const int QW = (w + 1) >> 1, QH = (h + 1) >> 1;
    const bool act = lane < QW;
    const bool vec = (((d.dst_off | dstride) & 1) == 0);      /* 8-byte aligned sample pairs */
    uint32_t nextq = act ? q[lane] : 0;
    unsigned pE1 = 0, pE3 = 0;                           /* exponents of samples 1 and 3 of the quad above */
    uint32_t ms_pos = 0;
    bool bad = false;
    const int x = 2 * lane;
    for (int qy = 0; qy < QH; qy++) {
        const unsigned mine = nextq;                     /* u_off | rho<<1 | emb_k<<5 | emb_1<<9 | u<<16; 0 in idle lanes */
        if (qy + 1 < QH)
            nextq = act ? q[(qy + 1) * QW + lane] : 0;
        const unsigned rho = (mine >> 1) & 15;
        const unsigned ek = (mine >> 5) & 15, e1 = (mine >> 9) & 15;
        const int u = (int)(mine >> 16);
        int kappa = 1;
        if (qy > 0) {                                    /* :855-885 */
            unsigned nw3 = __shfl_up_sync(FULL, pE3, 1), ne1 = __shfl_down_sync(FULL, pE1, 1);
            if (lane == 0)
                nw3 = 0;
            if (lane == 31)
                ne1 = 0;
            const int m = (int)max(max(pE1, pE3), max(nw3, ne1));
            const int gamma = (rho & (rho - 1)) != 0;
            kappa = max(1, gamma * (m - 1));
        }
        const int U = kappa + u;
        if (__any_sync(FULL, act && U > maxbp)) {        /* :715, :889 */
            bad = true;
            break;
        }
        int mlen[4];
        int tot = 0;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            mlen[i] = (((rho >> i) & 1) ? U : 0) - (int)((ek >> i) & 1);
            tot += max(mlen[i], 0);
        }
        int incl = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(FULL, incl, o);
            if (lane >= o)
                incl += t;
        }
        uint32_t pos = ms_pos + (uint32_t)(incl - tot);
        ms_pos += (uint32_t)__shfl_sync(FULL, incl, 31);
        uint32_t mu[4];
        unsigned E1 = 0, E3 = 0;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            uint32_t v = 0, val = 0;
            const int m = mlen[i];
            if (m > 0) {
                uint32_t p = min(pos, ms_bits);          /* past the end: ones */
                v = ms_read(ms, p, m) + (((e1 >> i) & 1u) << m);
                pos += m;
            }
            if (m != 0) {                                /* :419-425 */
                unsigned E = 32 - __clz(v | 1);
                if (i == 1) E1 = E;
                if (i == 3) E3 = E;
                val = (uint32_t)(((int32_t)v >> 1) + 1) << pLSB;
                val |= 1u << (pLSB - 1);
                val |= (v & 1u) << 31;
            }
            mu[i] = val;
        }
        pE1 = E1;
        pE3 = E3;
        /* samples of quad (qy, lane): i=0 (y,x) i=1 (y+1,x) i=2 (y,x+1) i=3 (y+1,x+1) */
        const int y = 2 * qy;
        if (act) {
            const uint32_t o0 = dq<DQ>(mu[0], sh53, fscale, mode, M_b, roi_shift, bd),
                           o1 = dq<DQ>(mu[1], sh53, fscale, mode, M_b, roi_shift, bd),
                           o2 = dq<DQ>(mu[2], sh53, fscale, mode, M_b, roi_shift, bd),
                           o3 = dq<DQ>(mu[3], sh53, fscale, mode, M_b, roi_shift, bd);
            uint32_t *r0 = dst + (size_t)y * dstride + x, *r1 = r0 + dstride;
            if (vec && x + 1 < w) {
                *reinterpret_cast<uint2 *>(r0) = make_uint2(o0, o2);
                if (y + 1 < h)
                    *reinterpret_cast<uint2 *>(r1) = make_uint2(o1, o3);
            } else {
                r0[0] = o0;
                if (y + 1 < h)
                    r1[0] = o1;
                if (x + 1 < w) {
                    r0[1] = o2;
                    if (y + 1 < h)
                        r1[1] = o3;
                }
            }
        }
    }
    if (bad) {
        __syncwarp();
        zero_block(dst, dstride, w, h, lane);
    }
    if (lane == 0 && status)
        status[bi] = bad ? B200J2K_KERR_INVALID : 1;
}

__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
ht_magsgn_kernel(const CblkDesc *__restrict__ descs, const uint32_t *__restrict__ order,
                 const uint32_t *__restrict__ scr_off, int nblocks, const uint8_t *__restrict__ bytes,
                 const BandDesc *__restrict__ bands, const uint32_t *__restrict__ scratch,
                 uint32_t *__restrict__ coef, int32_t *__restrict__ status, int stream_words)
{
    extern __shared__ __align__(16) uint32_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gw = blockIdx.x * WARPS_PER_CTA + warp;
    if (gw >= nblocks)
        return;
    const uint32_t bi = order[gw];
    const CblkDesc d = descs[bi];
    uint32_t *dst = coef + d.dst_off;
    const uint8_t *Dcup = bytes + d.data_off;
    const int Scup = ht_block_scup(d, Dcup);
    if (Scup <= 0) {
        /* nothing decoded or damaged header: the block is all zeros (quirk Q1) */
        zero_block(dst, d.dst_stride, d.w, d.h, lane);
        if (lane == 0 && status)
            status[bi] = Scup == 0 ? 0 : B200J2K_KERR_INVALID;
        return;
    }
    const BandDesc bd = bands[d.band];
    const int mode = (d.flags >> 1) & 3;
    const uint32_t *q = scratch + scr_off[bi] + 2 * ht_stream_words(Scup);
    uint32_t *ms = smem + (size_t)warp * stream_words;
    if (d.roi_shift == 0 && mode == J2K_DWT53 && bd.i_stepsize == 32768)
        magsgn_block<DQ_INT53>(d, bi, Scup, Dcup, bd, q, ms, dst, status, lane);
    else if (d.roi_shift == 0 && mode == J2K_DWT97)
        magsgn_block<DQ_FLOAT97>(d, bi, Scup, Dcup, bd, q, ms, dst, status, lane);
    else
        magsgn_block<DQ_GENERIC>(d, bi, Scup, Dcup, bd, q, ms, dst, status, lane);
}

}  // namespace

/* ---- host side ----------------------------------------------------------------------------------------- */
uint32_t ht_cleanup_scratch_words(const CblkDesc &d)
{
    if (d.npasses == 0 || d.Lcup < 2)
        return 0;
    const int S = d.Lcup < 4079 ? d.Lcup : 4079;
    const int QW = (d.w + 1) >> 1, QH = (d.h + 1) >> 1;
    return (uint32_t)(2 * ht_stream_words(S) + QW * QH);
}

int launch_ht_cleanup(cudaStream_t stream, const CblkDesc *descs, const uint32_t *order, const uint32_t *scr_off,
                      int n_coded, int n_all, const uint8_t *bytes, const BandDesc *bands, void *scratch, void *coef,
                      int32_t *status, int max_Lcup)
{
    if (n_all <= 0)
        return 0;
    int launches = 0;
    if (n_coded > 0) {
        const int warp_words = 2 * ht_stream_words(max_Lcup < 4079 ? max_Lcup : 4079);
        const size_t smem0 = (size_t)warp_words * 4 * WARPS_PER_CTA;
        if (smem0 > 48 * 1024 &&
            cudaFuncSetAttribute(ht_unstuff_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem0) != cudaSuccess)
            return -2;
        ht_unstuff_kernel<<<(n_coded + WARPS_PER_CTA - 1) / WARPS_PER_CTA, WARPS_PER_CTA * 32, smem0, stream>>>(
            descs, order, scr_off, n_coded, bytes, (uint32_t *)scratch, warp_words);
        ht_vlc_kernel<<<(n_coded + VLC_LANES_PER_CTA - 1) / VLC_LANES_PER_CTA, VLC_LANES_PER_CTA, 0, stream>>>(
            descs, order, scr_off, n_coded, bytes, (uint32_t *)scratch);
        launches += 2;
    }
    const int stream_words = (max_Lcup + 3) / 4 + 4;
    const size_t smem2 = (size_t)stream_words * 4 * WARPS_PER_CTA;
    if (smem2 > 227 * 1024)
        return -1;
    if (smem2 > 48 * 1024 &&
        cudaFuncSetAttribute(ht_magsgn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2) != cudaSuccess)
        return -2;
    ht_magsgn_kernel<<<(n_all + WARPS_PER_CTA - 1) / WARPS_PER_CTA, WARPS_PER_CTA * 32, smem2, stream>>>(
        descs, order, scr_off, n_all, bytes, bands, (const uint32_t *)scratch, (uint32_t *)coef, status, stream_words);
    launches++;
    return cudaGetLastError() == cudaSuccess ? launches : -3;
}

}  // namespace b200j2k
