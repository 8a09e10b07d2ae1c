/*
 * ht_cleanup.cu -- the HT cleanup pass (Rec. ITU-T T.814) as a three-kernel pipeline for sm_100a.
 *
 * Replaces, for code-blocks up to 64 samples wide, jpeg2000_decode_ht_cleanup_segment
 * (libavcodec/jpeg2000htdec.c:547-1014) with its bit readers (:145-221, :283-295, :429-440), the MEL
 * decoder (:462-495), CxtVLC (:300-331, :510-531) and U-VLC (:338-388), plus the dequantisation of
 * jpeg2000dec.c:2098-2181 fused into the coefficient write-back.  SURVEY.md section 8(a) rows H1-H6, D1-D3.
 *
 * Why three kernels.  ncu on the warp-per-block kernel (ht_block.cu, profiles/r1_ncu_full_v1_summary_warp_per_block_baseline.md)
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
/* One warp per CTA for the warp-per-block kernels of this file: the hardware scheduler then hands out code-blocks
 * one at a time (32 resident per SM either way), which shortens the tail of the un-stuff and MagSgn kernels --
 * measured 4 -> 1 warps per CTA: MagSgn 89.7 -> 86.0 us, un-stuff 26.9 -> 25.2 us, kernel-only 4K 5845 -> 6000 fps
 * (2 warps: 5908, 8 warps: 5684). */
#ifndef B200J2K_WARPS_PER_CTA
#define B200J2K_WARPS_PER_CTA 1
#endif
#include "ht_common.cuh"

namespace b200j2k {

namespace {

/* scratch layout of a block with refinement passes, word offsets from the block's scratch base.  The
 * cleanup part comes first: [VLC bits | MEL bits | one word per quad], sized for the largest possible suffix
 * so that the offsets only depend on the descriptor. */
struct RefLayout {
    int rw;             /* words of one un-stuffed Dref bit array */
    int nst;            /* 4-row stripes (the row below an odd-height block included, quirk Q2) */
    int sp, mr;         /* un-stuffed SigProp (forward) and MagRef (backward) bits */
    int sig, nsg, sgn;  /* nibble-layout maps, 8 words per stripe: cleanup significance, newly significant, signs */
    int total;
};

__host__ __device__ inline RefLayout ref_layout(const CblkDesc &d)
{
    RefLayout r;
    const int S = d.Lcup < 4079 ? d.Lcup : 4079;
    const int QW = (d.w + 1) >> 1, QH = (d.h + 1) >> 1;
    const int cleanup = (2 * ht_stream_words(S) + QW * QH + 1) & ~1;
    r.rw = ((d.Lref + 3) / 4 + 3 + 1) & ~1;
    r.nst = (QH + 1) >> 1;
    r.sp = cleanup;
    r.mr = r.sp + r.rw;
    r.sig = r.mr + r.rw;
    r.nsg = r.sig + 8 * r.nst;
    r.sgn = r.nsg + 8 * r.nst;
    r.total = r.sgn + 8 * r.nst;
    return r;
}

/* H1 (jpeg2000htdec.c:1237-1273): Scup (>= 2) when the cleanup segment is decodable, 0 when there is
 * nothing to decode, -1 when the segment header is damaged (block = zeros, quirk Q1) */
__device__ __forceinline__ int ht_block_scup(const CblkDesc &d, const uint8_t *Dcup)
{
    /* a Part-1 block only reaches these kernels when it carries no bytes (decode_cblk returns zeros for
     * !cblk->length, jpeg2000dec.c:2005): its Lcup / Lref fields hold other things */
    if (d.npasses == 0 || (d.flags & J2K_CBLK_MQ))
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
    J2K_BOUNDS(2 * vlw <= warp_words);
    for (int i = lane; i < 2 * vlw; i += 32)
        vl[i] = 0;
    __syncwarp();
    const uint32_t vbits = unstuff<US_VLC>(Dcup + Lcup - 2, Scup - 1, vl, lane);
    {
        /* base[0], base[1] of the reference's reader = Dcup[Pcup], Dcup[Pcup+1] after modDcup */
        unsigned b0 = Dcup[Pcup], b1 = Dcup[Pcup + 1];
        if (Scup == 2) {
            b0 |= 0x0F;
            b1 = 0xFF;
        } else if (Scup == 3) {
            b1 |= 0x0F;
        }
        bwd_tail_fill(vl, vlw, bwd_tail_make(vbits, Scup - 1, b0, b1), lane);
    }
    const uint32_t mbits = unstuff<US_MEL>(Dcup + Pcup, Scup, ml, lane);
    ones_behind(ml, mbits, vlw, lane);
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

/* the VLC reader: same window, but behind the array it continues with the reference's periodic tail */
struct VlcReader {
    uint64_t win;
    int avail;
    const uint32_t *base;
    int i, words;
    uint32_t nxt;
    BwdTail tail;
    __device__ __forceinline__ void init(const uint32_t *ptr, int w)
    {
        win = (uint64_t)ptr[0] | ((uint64_t)ptr[1] << 32);
        avail = 64;
        nxt = ptr[2];
        base = ptr;
        i = 3;
        words = w;
        tail = bwd_tail_unpack(ptr[w - 1]);
    }
    __device__ __forceinline__ void refill()
    {
        if (avail <= 32) {
            win |= (uint64_t)nxt << avail;
            avail += 32;
            nxt = bwd_word(base, words, tail, i);
            i++;
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
__device__ __forceinline__ void vlc_pair_lane(VlcReader &v, MelLane &mel, const uint16_t *table, const uint16_t *uvtab,
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
constexpr int VLC_ACTIVE_LANES = 32;     /* blocks per warp (B200J2K_VLC_ACTIVE overrides, for experiments).  Round 2,
                                          * kernel-only 4K frames/s with 12 frames in flight: 8 lanes 5471, 16 lanes
                                          * 6368, 24 lanes 6582, 32 lanes 6590 (round 1 favoured 16 when the kernels
                                          * around it were slower) */

__global__ void __launch_bounds__(VLC_LANES_PER_CTA)
ht_vlc_kernel(const CblkDesc *__restrict__ descs, const uint32_t *__restrict__ order,
              const uint32_t *__restrict__ scr_off, int nblocks, const uint8_t *__restrict__ bytes,
              uint32_t *__restrict__ scratch, int active)
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
    asm volatile("griddepcontrol.wait;" ::: "memory");   /* the tables were staged beside the un-stuff kernel's tail */
    /* `active` lanes of the warp own a block each, the others only helped with the tables: every lane walks its
     * own control flow (MEL events, U-VLC cases), so the warp pays for the union of its lanes' paths */
    const int gt = blockIdx.x * active + threadIdx.x;
    if ((int)threadIdx.x >= active || gt >= nblocks)
        return;
    const uint32_t bi = order[gt];
    const CblkDesc d = descs[bi];
    const int Scup = ht_block_scup(d, bytes + d.data_off);
    if (Scup <= 0)
        return;
    const int QW = (d.w + 1) >> 1, QH = (d.h + 1) >> 1; /* QW <= 32: the row's significance fits two words */
    const int vlw = ht_stream_words(Scup);
    uint32_t *base = scratch + scr_off[bi];
    VlcReader v;
    v.init(base, vlw);
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
enum { DQ_GENERIC = 0, DQ_INT53 = 1, DQ_FLOAT97 = 2, DQ_INT53_DN = 3 };    /* DQ_INT53(_DN): see magsgn_block */

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

/* REFINE: the block carries SigProp / MagRef passes.  The samples are stored RAW (sign-magnitude, not
 * dequantised; ht_refine_apply_kernel finishes them) with the reference's odd-edge masking (quirk Q2), and
 * the significance map goes to scratch in "stripe nibble" layout: per 4-row stripe 32 bytes, byte j =
 * sigma of columns 2j (low nibble) and 2j+1 (high nibble), bit i of a nibble = row i of the stripe -- bit
 * order == the stripe-oriented scan order of jpeg2000htdec.c:1034-1185. */
template <int DQ, bool REFINE>
__device__ __forceinline__ void magsgn_block(const CblkDesc &d, uint32_t bi, int Scup, const uint8_t *Dcup,
                                             const BandDesc &bd, const uint32_t *__restrict__ q, uint32_t *ms,
                                             uint32_t *__restrict__ dst, int32_t *__restrict__ status, int lane,
                                             uint8_t *__restrict__ sigb)
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
    /* DQ_INT53 / DQ_INT53_DN: (x << pLSB | half) >> sh53 with per-block constants, written as MULTIPLICATIONS.  ncu
     * (profiles/r2_ncu_full_summary.md): this kernel is bound by the integer ALU pipe (shifts, logic, compares, selects:
     * one warp instruction per two cycles and scheduler), the multiply-add pipe beside it is half empty.  So
     *   pLSB >= sh53 (DQ_INT53):    r = mag * 2^(pLSB - sh53) + half        one IMAD
     *   pLSB <  sh53 (DQ_INT53_DN): r = umulhi(mag, 2^(32 - (sh53 - pLSB)))  one IMAD.HI  (= mag >> (sh53 - pLSB))
     * and the sign and the "no MagSgn bits: zero" select become a factor of -1 / 0 / +1 (two's complement wraps
     * exactly like the shifts did: mag << pLSB of a valid sample stays below bit 31, U <= maxbp). */
    uint32_t dq_mul = pLSB >= sh53 ? 1u << (pLSB - sh53) : 1u << (32 - (sh53 - pLSB));
    const uint32_t dq_half = pLSB - 1 >= sh53 ? 1u << (pLSB - 1 - sh53) : 0;
    asm("" : "+r"(dq_mul));          /* opaque: else the compiler turns the multiplications back into shifts */

    /* un-stuffed MagSgn bits in shared memory; past the end the decoder reads ones */
    const int ms_words = (Pcup + 3) / 4 + 2;
    for (int i = lane; i < ms_words; i += 32)
        ms[i] = 0;
    __syncwarp();
    const uint32_t ms_bits = unstuff<US_MAGSGN>(Dcup, Pcup, ms, lane);
    ones_behind(ms, ms_bits, ms_words, lane);
    /* launched as a programmatic dependent of the VLC kernel: everything above reads only the packet bytes, the quad
     * words below are that kernel's output (a no-op in a plain launch) */
    asm volatile("griddepcontrol.wait;" ::: "memory");

    const int QW = (w + 1) >> 1, QH = (h + 1) >> 1;
    const bool act = lane < QW;
    const bool vec = (((d.dst_off | dstride) & 1) == 0);      /* 8-byte aligned sample pairs */
    /* the quad words come from HBM scratch: two rows ahead, so the load of a row is never waited for (ncu: a quarter
     * of this kernel's stall samples sat on the one-row-ahead load) */
    uint32_t nextq = act ? q[lane] : 0, next2 = act && QH > 1 ? q[QW + lane] : 0;
    unsigned pE1 = 0, pE3 = 0;                           /* exponents of samples 1 and 3 of the quad above */
    uint32_t ms_pos = 0;
    bool bad = false;
    const int x = 2 * lane;
    unsigned prev_sg = 0;
    uint32_t *rowp = dst + x;
    /* The four samples of a quad are handled as four BYTES of one register where that saves instructions (ncu: this
     * kernel is bound by the integer pipes, ~260 instructions per quad row before): bit i of rho / emb_k becomes byte
     * i (multiplication by 0x00204081 puts bit i at bit 8i without a carry), the MagSgn bit counts are then
     * m = rho * U - emb_k in all four bytes at once, their sum and their prefix sums one multiplication each.
     * That needs 0 <= m <= 32 per byte: emb_k is a subset of rho and emb_1 of emb_k in every CxtVLC table entry
     * (checked over both tables), U >= 1, and U is clamped to 32 -- a larger one exceeds maxbp <= 31 anyway, which
     * condemns the block (quirk Q1): the lanes note that and carry on; one vote after the last row decides. */
    for (int qy = 0; qy < QH; qy++, rowp += 2 * (size_t)dstride) {
        const unsigned mine = nextq;                     /* u_off | rho<<1 | emb_k<<5 | emb_1<<9 | u<<16; 0 in idle lanes */
        nextq = next2;
        if (qy + 2 < QH)
            next2 = act ? q[(qy + 2) * QW + lane] : 0;
        const unsigned rho = (mine >> 1) & 15;
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
        bad |= U > maxbp;                                /* :715, :889; idle lanes have U = 1 */
        const uint32_t rhoB = (rho * 0x00204081u) & 0x01010101u;
        const uint32_t ekB = (((mine >> 5) & 15) * 0x00204081u) & 0x01010101u;
        const uint32_t e1B = (((mine >> 9) & 15) * 0x00204081u) & 0x01010101u;
        const uint32_t mB = rhoB * (uint32_t)min(U, 32) - ekB;           /* MagSgn bits of the four samples */
        const int tot = (int)((mB * 0x01010101u) >> 24);
        const int incl = warp_scan_add(tot);
        const uint32_t pos0 = ms_pos + (uint32_t)(incl - tot);
        ms_pos += (uint32_t)__shfl_sync(FULL, incl, 31);
        const uint32_t preB = mB * 0x01010100u;                          /* bits in front of each sample, inside the quad */
        uint32_t mu[4];
        unsigned E1 = 0, E3 = 0;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            /* straight-line on purpose (no divergent regions in the row loop): a sample without MagSgn bits
             * reads a zero-length field */
            const int mm = (int)((mB >> (8 * i)) & 0xFF);
            const uint32_t p = min(pos0 + ((preB >> (8 * i)) & 0xFF), ms_bits);   /* past the end: ones */
            J2K_BOUNDS((int)(p >> 5) + 1 < ms_words);
            /* the mm-bit field at bit p, plus the implicit leading one (emb_1) on top of it: the mask and the placing
             * of that bit as arithmetic on 2^mm (multiply-add pipe) instead of two more shifts */
            uint32_t pow2 = 1u << mm;
            asm("" : "+r"(pow2));
            const uint32_t wi = p >> 5;
            const uint32_t field = __funnelshift_r(ms[wi], ms[wi + 1], p) & (pow2 - 1u);
            uint32_t v = ((e1B >> (8 * i)) & 1u) * pow2 + field;
            const unsigned E = 32 - __clz(v | 1);        /* :419-425 */
            if (i == 1) E1 = mm ? E : 0;
            if (i == 3) E3 = mm ? E : 0;
            uint32_t val;
            if ((DQ == DQ_INT53 || DQ == DQ_INT53_DN) && !REFINE) {
                const uint32_t mag = (v >> 1) + 1;
                const uint32_t r = DQ == DQ_INT53 ? mag * dq_mul + dq_half : __umulhi(mag, dq_mul);
                const uint32_t sgnz = (uint32_t)min(mm, 1) * (1u - 2u * (v & 1u));     /* -1, 0 or +1 */
                val = r * sgnz;
            } else {
                val = ((v >> 1) + 1) << pLSB;
                val |= 1u << (pLSB - 1);
                val |= (v & 1u) << 31;
                val = mm ? val : 0u;
            }
            mu[i] = val;
        }
        pE1 = E1;
        pE3 = E3;
        /* samples of quad (qy, lane): i=0 (y,x) i=1 (y+1,x) i=2 (y,x+1) i=3 (y+1,x+1) */
        const int y = 2 * qy;
        if (REFINE) {
            /* quad -> raster with the reference's odd-edge masks (:975-1007, quirk Q2) */
            const unsigned keep_row = qy != QH - 1 || !(h & 1), keep_col = lane != QW - 1 || !(w & 1);
            const unsigned km = 1u | (keep_row << 1) | (keep_col << 2) | ((keep_row | keep_col) << 3);
#pragma unroll
            for (int i = 0; i < 4; i++)
                if (!((km >> i) & 1))
                    mu[i] = 0;
            const unsigned sg = rho & km;
            if (qy & 1)
                sigb[(qy >> 1) * 32 + lane] = (uint8_t)((prev_sg & 3) | ((sg & 3) << 2) | (((prev_sg >> 2) & 3) << 4) |
                                                        (((sg >> 2) & 3) << 6));
            else
                prev_sg = sg;
        }
        if (act) {
            constexpr bool RAW = REFINE || DQ == DQ_INT53 || DQ == DQ_INT53_DN;   /* already final (or finished later) */
            const uint32_t o0 = RAW ? mu[0] : dq<DQ>(mu[0], sh53, fscale, mode, M_b, roi_shift, bd),
                           o1 = RAW ? mu[1] : dq<DQ>(mu[1], sh53, fscale, mode, M_b, roi_shift, bd),
                           o2 = RAW ? mu[2] : dq<DQ>(mu[2], sh53, fscale, mode, M_b, roi_shift, bd),
                           o3 = RAW ? mu[3] : dq<DQ>(mu[3], sh53, fscale, mode, M_b, roi_shift, bd);
            uint32_t *r0 = rowp, *r1 = r0 + dstride;
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
    bad = __any_sync(FULL, bad);
    if (REFINE && (QH & 1) && !bad)
        sigb[(QH >> 1) * 32 + lane] = (uint8_t)((prev_sg & 3) | (((prev_sg >> 2) & 3) << 4));
    if (bad) {
        __syncwarp();
        zero_block(dst, dstride, w, h, lane);
    }
    if (lane == 0)
        j2k_block_status(status, bi, bad ? B200J2K_KERR_INVALID : 1);
}

template <bool REFINE>
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
        if (lane == 0)
            j2k_block_status(status, bi, Scup == 0 ? 0 : B200J2K_KERR_INVALID);
        return;
    }
    const BandDesc bd = bands[d.band];
    const int mode = (d.flags >> 1) & 3;
    const uint32_t *q = scratch + scr_off[bi] + 2 * ht_stream_words(Scup);
    uint32_t *ms = smem + (size_t)warp * stream_words;
    J2K_BOUNDS((d.Lcup - Scup + 3) / 4 + 2 <= stream_words);
    if (REFINE) {
        const RefLayout rl = ref_layout(d);
        uint8_t *sigb = reinterpret_cast<uint8_t *>(const_cast<uint32_t *>(scratch) + scr_off[bi] + rl.sig);
        magsgn_block<DQ_GENERIC, true>(d, bi, Scup, Dcup, bd, q, ms, dst, status, lane, sigb);
    } else if (d.roi_shift == 0 && mode == J2K_DWT53 && bd.i_stepsize == 32768) {
        const int rem = d.npasses % 3;
        const int S_blk = ((rem ? d.npasses - rem : d.npasses - 3) / 3 + d.zbp) & 0xFF;
        if (30 - S_blk >= 31 - d.M_b)
            magsgn_block<DQ_INT53, false>(d, bi, Scup, Dcup, bd, q, ms, dst, status, lane, nullptr);
        else
            magsgn_block<DQ_INT53_DN, false>(d, bi, Scup, Dcup, bd, q, ms, dst, status, lane, nullptr);
    } else if (d.roi_shift == 0 && mode == J2K_DWT97) {
        magsgn_block<DQ_FLOAT97, false>(d, bi, Scup, Dcup, bd, q, ms, dst, status, lane, nullptr);
    } else {
        magsgn_block<DQ_GENERIC, false>(d, bi, Scup, Dcup, bd, q, ms, dst, status, lane, nullptr);
    }
}

/* ---- refinement passes: K0R un-stuff Dref, K3 SigProp (lane per block), K4 MagRef + apply + dequantise ---- */
__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
ht_unstuff_ref_kernel(const CblkDesc *__restrict__ descs, const uint32_t *__restrict__ order,
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
    /* a block whose cleanup segment cannot be decoded gets zeros and no refinement; with Lcup < 2 it does not even
     * own scratch (ht_cleanup_scratch_words): writing its streams would land in another block's words */
    if (ht_block_scup(d, bytes + d.data_off) <= 0)
        return;
    const RefLayout rl = ref_layout(d);
    const uint8_t *Dref = bytes + d.data_off + d.Lcup;
    const int Lref = d.Lref;
    uint32_t *sp = smem + (size_t)warp * warp_words, *mr = sp + rl.rw;
    J2K_BOUNDS(2 * rl.rw <= warp_words);
    for (int i = lane; i < 2 * rl.rw; i += 32)
        sp[i] = 0;
    __syncwarp();
    unstuff<US_SIGPROP>(Dref, Lref, sp, lane);                    /* zeros behind */
    {
        /* bytes at and behind Dref[Lref]: the 0xFF the decoder plants at cblk->length (jpeg2000htdec.c:1258),
         * then nothing */
        auto ref_byte = [&](int i) -> unsigned {
            if (i < Lref)
                return Dref[i];
            const uint32_t at = (uint32_t)d.Lcup + (uint32_t)i;
            return at == d.length ? 0xFFu : (at < d.length ? Dref[i] : 0u);
        };
        const uint32_t rbits = unstuff<US_MAGREF>(Dref + Lref - 1, Lref, mr, lane, ref_byte(Lref));
        bwd_tail_fill(mr, rl.rw, bwd_tail_make(rbits, Lref, ref_byte(0), ref_byte(1)), lane);
    }
    uint32_t *out = scratch + scr_off[bi] + rl.sp;
    for (int i = lane; i < 2 * rl.rw; i += 32)
        out[i] = sp[i];
}

constexpr uint64_t NIB0 = 0x1111111111111111ull;

/* horizontal 3-neighbourhood OR over a 64-column stripe in nibble layout (column x = bits 4x..4x+3 of v[x/16]) */
__device__ __forceinline__ void dilate_cols(const uint64_t (&v)[4], uint64_t (&o)[4])
{
#pragma unroll
    for (int k = 0; k < 4; k++) {
        uint64_t t = v[k] | (v[k] << 4) | (v[k] >> 4);
        if (k > 0)
            t |= v[k - 1] >> 60;
        if (k < 3)
            t |= v[k + 1] << 60;
        o[k] = t;
    }
}

/* H7 SigProp (jpeg2000htdec.c:1016-1131), one code-block per LANE, entirely on bit masks.
 * A sample is a member when it is insignificant and one of its 8 neighbours is significant: significant in
 * the cleanup pass, or made significant earlier in this pass's scan (stripes of 4 rows, column by column).
 * Static members come from dilating the cleanup significance map; the others are added as the scan sets
 * new bits: a newly significant sample can only recruit the sample below it and the three samples of the
 * next column (the next stripe sees it through its "row above").  In nibble layout the scan order IS the
 * bit order, so the scan is a find-first-set loop over the candidate mask. */
__global__ void __launch_bounds__(32)
ht_sigprop_kernel(const CblkDesc *__restrict__ descs, const uint32_t *__restrict__ order,
                  const uint32_t *__restrict__ scr_off, int nblocks, uint32_t *__restrict__ scratch,
                  const int32_t *__restrict__ status)
{
    const int gt = blockIdx.x * blockDim.x + threadIdx.x;
    if (gt >= nblocks)
        return;
    const uint32_t bi = order[gt];
    if (status[bi] != 1)
        return;
    const CblkDesc d = descs[bi];
    const RefLayout rl = ref_layout(d);
    const int w = d.w, h = d.h;
    const bool causal = d.flags & 1;
    uint32_t *base = scratch + scr_off[bi];
    LsbReader sp;
    sp.init(base + rl.sp, rl.rw, 0);
    const uint64_t *sig = reinterpret_cast<const uint64_t *>(base + rl.sig);
    uint64_t *nsg = reinterpret_cast<uint64_t *>(base + rl.nsg), *sgn = reinterpret_cast<uint64_t *>(base + rl.sgn);
    uint64_t CM[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int cols = min(max(w - 16 * k, 0), 16);
        CM[k] = cols >= 16 ? ~0ull : ((1ull << (4 * cols)) - 1);
    }
    uint64_t ab[4] = { 0, 0, 0, 0 };            /* (sigma | new) of the row above the stripe, at nibble bit 0 */
    uint64_t cur[4];
#pragma unroll
    for (int k = 0; k < 4; k++)
        cur[k] = sig[k];
    for (int s = 0; s < rl.nst; s++) {
        uint64_t nx[4];
#pragma unroll
        for (int k = 0; k < 4; k++)
            nx[k] = s + 1 < rl.nst ? sig[4 * (s + 1) + k] : 0;
        const int gh = min(4, h - 4 * s);
        uint64_t N[4] = { 0, 0, 0, 0 }, SG[4] = { 0, 0, 0, 0 };
        if (gh > 0) {
            const uint64_t RM = NIB0 * ((1u << gh) - 1);
            uint64_t D[4], AD[4], BD[4], bl[4], cand[4];
#pragma unroll
            for (int k = 0; k < 4; k++)
                bl[k] = (nx[k] & NIB0) << 3;
            dilate_cols(cur, D);
            dilate_cols(ab, AD);
            dilate_cols(bl, BD);
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const uint64_t up = ((D[k] << 1) & (NIB0 * 0xE)) | AD[k];
                uint64_t dn = ((D[k] >> 1) & (NIB0 * 0x7)) | BD[k];
                if (causal)                      /* the row below the stripe's last row is not consulted (:1045) */
                    dn &= ~(NIB0 << (gh - 1));
                cand[k] = (D[k] | up | dn) & ~cur[k] & RM & CM[k];
            }
            uint64_t carry = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const uint64_t okm = ~cur[k] & RM & CM[k];
                uint64_t c = cand[k] | (carry & okm);
                carry = 0;
                uint64_t n = 0, sg = 0;
#pragma unroll 1
                for (int g = 0; g < 4; g++) {
                    const uint64_t gm = 0xFFFFull << (16 * g);
                    uint64_t m;
                    while ((m = c & gm) != 0) {
                        const int b = __ffsll((long long)m) - 1;
                        c &= c - 1;
                        sp.refill();
                        const unsigned bit = (unsigned)sp.win & 1;
                        sp.drop(1);
                        if (bit) {
                            n |= 1ull << b;
                            const int i = b & 3;
                            uint64_t nc = i < 3 ? 1ull << (b + 1) : 0;
                            const uint64_t nb3 = ((7u << i) >> 1) & 0xF;       /* rows i-1 .. i+1 of the next column */
                            const int colpos = (b | 3) + 1;
                            if (colpos < 64)
                                nc |= nb3 << colpos;
                            else
                                carry |= nb3;
                            c |= nc & okm;
                        }
                    }
                    /* sign bits of the group's newly significant samples, in scan order (:1066-1076) */
                    uint64_t ng = n & gm;
                    while (ng) {
                        const int b = __ffsll((long long)ng) - 1;
                        ng &= ng - 1;
                        sp.refill();
                        sg |= (uint64_t)((unsigned)sp.win & 1) << b;
                        sp.drop(1);
                    }
                }
                N[k] = n;
                SG[k] = sg;
            }
        }
#pragma unroll
        for (int k = 0; k < 4; k++) {
            nsg[4 * s + k] = N[k];
            sgn[4 * s + k] = SG[k];
            ab[k] = ((cur[k] | N[k]) >> 3) & NIB0;
            cur[k] = nx[k];
        }
    }
}

/* H8 MagRef (:1136-1185) + the sample updates of both refinement passes + ROI shift + dequantisation.
 * Warp per block, lane j = columns 2j, 2j+1 of a stripe (one byte of the nibble-layout maps).  The k-th
 * cleanup-significant sample in scan order takes the k-th MagRef bit: its rank is a popcount over the
 * lane's byte plus a warp prefix sum. */
__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
ht_refine_apply_kernel(const CblkDesc *__restrict__ descs, const uint32_t *__restrict__ order,
                       const uint32_t *__restrict__ scr_off, int nblocks, const BandDesc *__restrict__ bands,
                       const uint32_t *__restrict__ scratch, uint32_t *__restrict__ coef,
                       const int32_t *__restrict__ status)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gw = blockIdx.x * WARPS_PER_CTA + warp;
    if (gw >= nblocks)
        return;
    const uint32_t bi = order[gw];
    if (status[bi] != 1)
        return;
    const CblkDesc d = descs[bi];
    const BandDesc bd = bands[d.band];
    const RefLayout rl = ref_layout(d);
    const int w = d.w, h = d.h;
    const int mode = (d.flags >> 1) & 3;
    const int rem = d.npasses % 3;
    const int plhd = rem ? d.npasses - rem : d.npasses - 3;
    const int z_blk = d.npasses - plhd;
    const int S_blk = (plhd / 3 + d.zbp) & 0xFF;
    const int p = 30 - S_blk - 1;
    const uint32_t *base = scratch + scr_off[bi];
    const uint32_t *mr = base + rl.mr;
    const BwdTail mr_tail = bwd_tail_unpack(mr[rl.rw - 1]);
    const uint8_t *sigb = reinterpret_cast<const uint8_t *>(base + rl.sig);
    const uint8_t *nsgb = reinterpret_cast<const uint8_t *>(base + rl.nsg);
    const uint8_t *sgnb = reinterpret_cast<const uint8_t *>(base + rl.sgn);
    uint32_t *dst = coef + d.dst_off;
    const uint32_t dstride = d.dst_stride;
    const int x = 2 * lane;
    uint32_t mr_pos = 0;
    for (int s = 0; s < rl.nst; s++) {
        const int gh = min(4, h - 4 * s);
        if (gh <= 0)
            break;
        const unsigned sb = sigb[s * 32 + lane], nb = z_blk > 1 ? nsgb[s * 32 + lane] : 0, gb = sgnb[s * 32 + lane];
        const unsigned rm = (1u << gh) - 1;
        const unsigned vm = ((x < w ? 0x0Fu : 0) | (x + 1 < w ? 0xF0u : 0)) & (rm | (rm << 4));
        const unsigned sv = z_blk > 2 ? sb & vm : 0;
        const int cnt = __popc(sv);
        const int incl = warp_scan_add(cnt);
        const uint32_t pos = mr_pos + (uint32_t)(incl - cnt);
        mr_pos += (uint32_t)__shfl_sync(FULL, incl, 31);
        const uint32_t win = cnt ? __funnelshift_r(bwd_word(mr, rl.rw, mr_tail, (int)(pos >> 5)),
                                                   bwd_word(mr, rl.rw, mr_tail, (int)(pos >> 5) + 1), pos & 31) : 0;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            if (i >= gh)
                break;
            const int y = 4 * s + i;
            uint32_t *row = dst + (size_t)y * dstride + x;
#pragma unroll
            for (int c = 0; c < 2; c++) {
                const int bp = 4 * c + i;
                if (!((vm >> bp) & 1))
                    continue;
                uint32_t mu = row[c];
                if ((nb >> bp) & 1)              /* SigProp: newly significant (:1057-1060, :1066-1076) */
                    mu |= (1u << p) | (1u << (p - 1)) | (((gb >> bp) & 1u) << 31);
                if ((sv >> bp) & 1) {            /* MagRef (:1161-1165) */
                    const uint32_t bit = (win >> __popc(sv & ((1u << bp) - 1))) & 1;
                    mu = (mu & ((0xFFFFFFFEu | bit) << p)) | (1u << (p - 1));
                }
                row[c] = dequant(mu, mode, d.M_b, d.roi_shift, bd);
            }
        }
    }
}


}  // namespace

/* ---- host side ----------------------------------------------------------------------------------------- */
int ht_cleanup_bounds_failures()
{
#ifdef B200J2K_DEBUG_BOUNDS
    int n = 0, zero = 0;
    if (cudaMemcpyFromSymbol(&n, g_bounds_failures, sizeof(n)) != cudaSuccess)
        return -1;
    cudaMemcpyToSymbol(g_bounds_failures, &zero, sizeof(zero));
    return n;
#else
    return 0;
#endif
}

uint32_t ht_cleanup_scratch_words(const CblkDesc &d)
{
    if (d.npasses == 0 || d.Lcup < 2)
        return 0;
    if ((d.npasses % 3) != 1)
        return (uint32_t)ref_layout(d).total;
    const int S = d.Lcup < 4079 ? d.Lcup : 4079;
    const int QW = (d.w + 1) >> 1, QH = (d.h + 1) >> 1;
    return (uint32_t)((2 * ht_stream_words(S) + QW * QH + 1) & ~1);
}

/* launch as a programmatic dependent of the kernel before it in the stream: the kernel may start while that one
 * drains and holds at its griddepcontrol.wait until the predecessor's memory is visible */
template <class... P, class... A>
static void launch_dependent(void (*kernel)(P...), int grid, int block, size_t smem, cudaStream_t stream, A... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr.val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kernel, static_cast<P>(args)...);
}

static int set_smem(const void *fn, size_t bytes)
{
    if (bytes > 227 * 1024)
        return -1;
    if (bytes > 48 * 1024 && cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess)
        return -2;
    return 0;
}

int launch_ht_cleanup(cudaStream_t stream, const CblkDesc *descs, const uint32_t *order, const uint32_t *scr_off,
                      int n_refine, int n_coded, int n_zero, const uint8_t *bytes, const BandDesc *bands, void *scratch,
                      void *coef, int32_t *status, int max_Lcup, int max_Lref, cudaEvent_t after_unstuff,
                      cudaEvent_t after_vlc)
{
    int launches = 0, r;
    const int n_vlc = n_refine + n_coded;
    uint32_t *scr = (uint32_t *)scratch;
    static const int skip = getenv("B200J2K_SKIP") ? atoi(getenv("B200J2K_SKIP")) : 0;   /* tuning: marginal cost of a kernel */
    static const int chain = getenv("B200J2K_HT_PDL") ? atoi(getenv("B200J2K_HT_PDL")) : 0;  /* tuning: 1 VLC, 2 MagSgn */
    if (n_vlc > 0 && !(skip & 1)) {
        const int warp_words = 2 * ht_stream_words(max_Lcup < 4079 ? max_Lcup : 4079);
        const size_t smem0 = (size_t)warp_words * 4 * WARPS_PER_CTA;
        if ((r = set_smem((const void *)ht_unstuff_kernel, smem0)) < 0)
            return r;
        ht_unstuff_kernel<<<(n_vlc + WARPS_PER_CTA - 1) / WARPS_PER_CTA, WARPS_PER_CTA * 32, smem0, stream>>>(
            descs, order, scr_off, n_vlc, bytes, scr, warp_words);
        launches++;
    }
    if (n_refine > 0) {
        const int rw = ((max_Lref + 3) / 4 + 3 + 1) & ~1;
        const size_t smemr = (size_t)2 * rw * 4 * WARPS_PER_CTA;
        if ((r = set_smem((const void *)ht_unstuff_ref_kernel, smemr)) < 0)
            return r;
        ht_unstuff_ref_kernel<<<(n_refine + WARPS_PER_CTA - 1) / WARPS_PER_CTA, WARPS_PER_CTA * 32, smemr, stream>>>(
            descs, order, scr_off, n_refine, bytes, scr, 2 * rw);
        launches++;
    }
    if (after_unstuff)
        cudaEventRecord(after_unstuff, stream);
    if (n_vlc > 0 && !(skip & 2)) {
        static const int active = [] {
            const char *e = getenv("B200J2K_VLC_ACTIVE");
            const int a = e ? atoi(e) : VLC_ACTIVE_LANES;
            return a >= 1 && a <= VLC_LANES_PER_CTA ? a : VLC_ACTIVE_LANES;
        }();
        if (!(chain & 1) || n_refine > 0 || after_unstuff || (skip & 1))   /* not directly behind the un-stuff kernel */
            ht_vlc_kernel<<<(n_vlc + active - 1) / active, VLC_LANES_PER_CTA, 0, stream>>>(
                descs, order, scr_off, n_vlc, bytes, scr, active);
        else
            launch_dependent(ht_vlc_kernel, (n_vlc + active - 1) / active, VLC_LANES_PER_CTA, 0, stream, descs, order,
                             scr_off, n_vlc, bytes, scr, active);
        launches++;
    }
    if (after_vlc)
        cudaEventRecord(after_vlc, stream);
    if (n_coded + n_zero > 0 && !(skip & 4)) {
        const int stream_words = (max_Lcup + 3) / 4 + 4;
        const size_t smem2 = (size_t)stream_words * 4 * WARPS_PER_CTA;
        if ((r = set_smem((const void *)ht_magsgn_kernel<false>, smem2)) < 0)
            return r;
        const int grid2 = (n_coded + n_zero + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
        if (!(chain & 2) || n_vlc == 0 || after_vlc || (skip & 2))          /* not directly behind the VLC kernel */
            ht_magsgn_kernel<false><<<grid2, WARPS_PER_CTA * 32, smem2, stream>>>(
                descs, order + n_refine, scr_off, n_coded + n_zero, bytes, bands, scr, (uint32_t *)coef, status, stream_words);
        else
            launch_dependent(ht_magsgn_kernel<false>, grid2, WARPS_PER_CTA * 32, smem2, stream, descs, order + n_refine,
                             scr_off, n_coded + n_zero, bytes, bands, (const uint32_t *)scr, (uint32_t *)coef, status,
                             stream_words);
        launches++;
    }
    return cudaGetLastError() == cudaSuccess ? launches : -3;
}

int launch_ht_refine(cudaStream_t stream, const CblkDesc *descs, const uint32_t *order, const uint32_t *scr_off,
                     int n_refine, const uint8_t *bytes, const BandDesc *bands, void *scratch, void *coef,
                     int32_t *status, int max_Lcup)
{
    if (n_refine <= 0)
        return 0;
    int r;
    uint32_t *scr = (uint32_t *)scratch;
    const int stream_words = (max_Lcup + 3) / 4 + 4;
    const size_t smem2 = (size_t)stream_words * 4 * WARPS_PER_CTA;
    if ((r = set_smem((const void *)ht_magsgn_kernel<true>, smem2)) < 0)
        return r;
    const int wgrid = (n_refine + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
    ht_magsgn_kernel<true><<<wgrid, WARPS_PER_CTA * 32, smem2, stream>>>(descs, order, scr_off, n_refine, bytes, bands, scr,
                                                                         (uint32_t *)coef, status, stream_words);
    ht_sigprop_kernel<<<(n_refine + 31) / 32, 32, 0, stream>>>(descs, order, scr_off, n_refine, scr, status);
    ht_refine_apply_kernel<<<wgrid, WARPS_PER_CTA * 32, 0, stream>>>(descs, order, scr_off, n_refine, bands, scr,
                                                                     (uint32_t *)coef, status);
    return cudaGetLastError() == cudaSuccess ? 3 : -3;
}

}  // namespace b200j2k
