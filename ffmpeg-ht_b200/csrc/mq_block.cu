/*
 * mq_block.cu -- Part-1 (EBCOT / MQ) code-block decode for sm_100a with the dequantisation fused into the
 * coefficient write-back.  Code-block styles with one codeword segment: 0, context
 * reset (0x02), vertically causal contexts (0x08), predictable termination (0x10), segmentation symbols (0x20).
 *
 * Replaces, for code-blocks that are not HT:
 *   decode_cblk + decode_sigpass / decode_refpass / decode_clnpass  libavcodec/jpeg2000dec.c:1872-2089
 *   the MQ decoder                                                     libavcodec/mqcdec.c:30-140, mqc.c:28-90
 *   context labels                                                     libavcodec/jpeg2000.c:90-170
 *   dequantization_{int,float,int_97}                                  libavcodec/jpeg2000dec.c:2098-2181
 *
 * First correct version, parity before speed: the arithmetic decoder is one serial chain per code-block (every
 * decision feeds the interval of the next and, through the significance state, its context), so one warp owns a
 * block and lane 0 walks the coding passes.  The state of the block lives in shared memory in scan order (a
 * word per stripe column: own state and significant-neighbour mask of its four samples, see Blk), so a column
 * with nothing to do costs one load and a context label comes from one byte instead of eight neighbour reads;
 * the sample values are built up in place in the block's own rectangle of the coefficient plane with plain
 * stores and fire-and-forget reductions (never read back by the decoder), which keeps shared memory at 9.5 KB for
 * 64x64 blocks (23 blocks resident per SM).  All 32 lanes share the clearing and the
 * final in-place ROI shift + dequantisation.  (Round 2: let the idle lanes form contexts ahead of the decoder.)
 *
 * Error behaviour of the reference kept: no data => the block stays zero; a bit-plane counter outside 0..29
 * ("bpno became invalid") stops the passes and what was decoded so far is dequantised WITHOUT the ROI step
 * (decode_cblk returns before it, tile_codeblocks dequantises on any non-zero return, :2264-2287).
 */
#include <cuda_runtime.h>
#include <stdint.h>

#include "ht_common.cuh"
#include "j2k_kernels.h"
#include "j2k_plan.h"

namespace b200j2k {

namespace {

/* ITU-T T.800 Table C.2 (normative): Qe, next state after an MPS / LPS, MPS switch */
__device__ __constant__ uint16_t c_mq_qe[47] = {
    0x5601, 0x3401, 0x1801, 0x0AC1, 0x0521, 0x0221, 0x5601, 0x5401, 0x4801, 0x3801, 0x3001, 0x2401,
    0x1C01, 0x1601, 0x5601, 0x5401, 0x5101, 0x4801, 0x3801, 0x3401, 0x3001, 0x2801, 0x2401, 0x2201,
    0x1C01, 0x1801, 0x1601, 0x1401, 0x1201, 0x1101, 0x0AC1, 0x09C1, 0x08A1, 0x0521, 0x0441, 0x02A1,
    0x0221, 0x0141, 0x0111, 0x0085, 0x0049, 0x0025, 0x0015, 0x0009, 0x0005, 0x0001, 0x5601 };
__device__ __constant__ uint8_t c_mq_nmps[47] = {
    1, 2, 3, 4, 5, 38, 7, 8, 9, 10, 11, 12, 13, 29, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24,
    25, 26, 27, 28, 29, 30, 31, 32, 33, 34, 35, 36, 37, 38, 39, 40, 41, 42, 43, 44, 45, 45, 46 };
__device__ __constant__ uint8_t c_mq_nlps[47] = {
    1, 6, 9, 12, 29, 33, 6, 14, 14, 14, 17, 18, 20, 21, 14, 14, 15, 16, 17, 18, 19, 19, 20, 21,
    22, 23, 24, 25, 26, 27, 28, 29, 30, 31, 32, 33, 34, 35, 36, 37, 38, 39, 40, 41, 42, 43, 46 };
__device__ __constant__ uint8_t c_mq_switch[47] = {
    1, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0,
    0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0 };

enum { CX_SIG = 0, CX_SGN = 9, CX_REF = 14, CX_UNI = 17, CX_RL = 18, CX_COUNT = 19 };
enum : uint8_t { S_SIG = 1, S_NEG = 2, S_VIS = 4, S_REF = 8 };


struct Mq {
    const uint8_t *data;
    int len, pos;                             /* pos = index of the byte last folded into C */
    unsigned cur, nxt;                        /* bytes at pos and pos + 1: the next one is fetched a byte ahead, so the
                                               * load is long done when the renormalisation needs it */
    uint32_t a, c;
    uint8_t *cx;                              /* CX_COUNT context states in shared memory: 2 * state + MPS */
};

/* the reference plants 0xFF 0xFF behind the segment (:2014-2015): a marker the decoder never steps over */
__device__ __forceinline__ unsigned mq_byte(const Mq &m, int i) { return i < m.len ? __ldg(m.data + i) : 0xFFu; }

__device__ __forceinline__ void mq_bytein(Mq &m)
{
    if (m.cur == 0xFF) {
        if (m.nxt > 0x8F) {
            m.c++;                             /* marker: feed ones, stay */
            return;
        }
        m.c += 2 + 0xFE00 - (m.nxt << 9);      /* stuffed byte: 7 bits */
    } else {
        m.c += 1 + 0xFF00 - (m.nxt << 8);
    }
    m.pos++;
    m.cur = m.nxt;
    m.nxt = mq_byte(m, m.pos + 1);
}

__device__ void mq_reset_contexts(Mq &m)      /* ff_mqc_init_contexts */
{
    for (int i = 0; i < CX_COUNT; i++)
        m.cx[i] = 0;
    m.cx[CX_UNI] = 2 * 46;
    m.cx[CX_RL] = 2 * 3;
    m.cx[0] = 2 * 4;
}

__device__ void mq_init(Mq &m, uint8_t *cx, const uint8_t *data, int len)
{
    m.cx = cx;
    mq_reset_contexts(m);
    m.data = data;
    m.len = len;
    m.pos = 0;
    m.cur = mq_byte(m, 0);
    m.nxt = mq_byte(m, 1);
    m.c = (m.cur ^ 0xFFu) << 16;
    mq_bytein(m);
    m.c <<= 7;
    m.a = 0x8000;
}

__device__ __forceinline__ int mq_decode(Mq &m, int ctx)
{
    const unsigned s = m.cx[ctx], idx = s >> 1, mps = s & 1, qe = c_mq_qe[idx];
    int d;
    m.a -= qe;
    if ((m.c >> 16) < m.a) {
        if (m.a & 0x8000)
            return (int)mps;
        if (m.a < qe) {                        /* conditional exchange on the MPS path */
            d = 1 - (int)mps;
            m.cx[ctx] = (uint8_t)(2 * c_mq_nlps[idx] + (mps ^ c_mq_switch[idx]));
        } else {
            d = (int)mps;
            m.cx[ctx] = (uint8_t)(2 * c_mq_nmps[idx] + mps);
        }
    } else {
        m.c -= m.a << 16;
        if (m.a < qe) {                        /* conditional exchange on the LPS path */
            d = (int)mps;
            m.cx[ctx] = (uint8_t)(2 * c_mq_nmps[idx] + mps);
        } else {
            d = 1 - (int)mps;
            m.cx[ctx] = (uint8_t)(2 * c_mq_nlps[idx] + (mps ^ c_mq_switch[idx]));
        }
        m.a = qe;
    }
    do {                                       /* RENORMD */
        if (!(m.c & 0xFF)) {
            m.c -= 0x100;
            mq_bytein(m);
        }
        m.a += m.a;
        m.c += m.c;
    } while (!(m.a & 0x8000));
    return d;
}

/* Sample state of the block in shared memory, laid out for the scan: per 4-row stripe and column one 32-bit word
 * whose byte r belongs to row 4 * stripe + r, so that one load answers "anything to do in this column?" for the
 * four samples a coding pass visits next.  Two such arrays:
 *   st4  S_SIG | S_NEG | S_VIS | S_REF of the sample itself
 *   nb4  which of its eight neighbours are significant (kept up to date when a sample becomes significant, as
 *        the reference keeps JPEG2000_T1_SIG_* in its flags, jpeg2000.c:172-200)
 * A frame of one column left / right and one stripe above / below the block stands in for the neighbours
 * outside.
 * Vertically causal contexts (JPEG2000_CBLK_VSC; the passes are instantiated for it, VSC): the fourth row of a
 * stripe reads its S / SW / SE neighbours as insignificant and the sign of S as absent (flags_mask,
 * jpeg2000dec.c:1882-1883, :1916-1917, :1966-1968) -- a mask on byte 3 of the nb4 word when it is READ. */
constexpr uint32_t VSC_BYTE3 = 0xFFu & ~(2u | 64u | 128u);          /* ~(NB_S | NB_SW | NB_SE) */
template <bool VSC> __device__ __forceinline__ uint32_t nb_word(uint32_t n4) { return VSC ? n4 & (0x00FFFFFFu | (VSC_BYTE3 << 24)) : n4; }
template <bool VSC> __device__ __forceinline__ unsigned nb_byte(unsigned n, int r) { return VSC && r == 3 ? n & VSC_BYTE3 : n; }
enum : uint8_t { NB_N = 1, NB_S = 2, NB_W = 4, NB_E = 8, NB_NW = 16, NB_NE = 32, NB_SW = 64, NB_SE = 128 };

struct Blk {
    int w, h, W2;                              /* W2 = w + 2 words per stripe */
    uint32_t *st4, *nb4;
    uint32_t *val;                             /* the block's rectangle in the coefficient plane ... */
    uint32_t vstride;                          /* ... and its row stride */
    int bandpos;
};

/* byte address of sample (x, y), -1 <= x <= w, -4 <= y < h + 4 */
__device__ __forceinline__ int byte_at(const Blk &b, int x, int y) { return ((((y + 4) >> 2) * b.W2 + x + 1) << 2) + ((y + 4) & 3); }

/* T.800 Table D.1 from the neighbour byte; bandpos 0 LL, 1 HL, 2 LH, 3 HH */
__device__ __forceinline__ int sig_context(const Blk &b, unsigned n)
{
    int h = ((n >> 2) & 1) + ((n >> 3) & 1), v = (n & 1) + ((n >> 1) & 1);
    const int d = __popc(n >> 4);
    if (b.bandpos == 3) {
        const int hv = h + v;
        if (d >= 3) return 8;
        if (d == 2) return hv >= 1 ? 7 : 6;
        if (d == 1) return hv >= 2 ? 5 : hv == 1 ? 4 : 3;
        return hv >= 2 ? 2 : hv;
    }
    if (b.bandpos == 1) { const int t = h; h = v; v = t; }
    if (h == 2) return 8;
    if (h == 1) return v >= 1 ? 7 : d >= 1 ? 6 : 5;
    if (v == 2) return 4;
    if (v == 1) return 3;
    return d >= 2 ? 2 : d;
}

/* Sample (x, y) becomes significant: decode its sign (Tables D.2 / D.3 from the states of the four direct
 * neighbours), store the value, mark it and tell the eight neighbours.  The byte addresses of the rows above and
 * below follow from the sample's own: one byte back / forth inside the stripe word, or over to the neighbouring
 * stripe's word. */
template <bool VSC>
__device__ __forceinline__ void become_significant(const Blk &b, Mq &m, int x, int y, uint32_t mask)
{
    uint8_t *st = reinterpret_cast<uint8_t *>(b.st4), *nb = reinterpret_cast<uint8_t *>(b.nb4);
    const int r = y & 3;                                   /* y + 4 has the same low bits */
    const int at = ((((y + 4) >> 2) * b.W2 + x + 1) << 2) + r;
    const int up = r ? at - 1 : at - (b.W2 << 2) + 3, dn = r != 3 ? at + 1 : at + (b.W2 << 2) - 3;
    auto contrib = [](uint8_t s) { return (s & S_SIG) ? ((s & S_NEG) ? -1 : 1) : 0; };
    int hc = contrib(st[at - 4]) + contrib(st[at + 4]);
    int vc = contrib(st[up]) + (VSC && r == 3 ? 0 : contrib(st[dn]));
    hc = max(-1, min(1, hc));
    vc = max(-1, min(1, vc));
    int xorbit = 0;
    if (hc < 0 || (hc == 0 && vc < 0)) {
        xorbit = 1;
        hc = -hc;
        vc = -vc;
    }
    const int ctx = hc == 1 ? CX_SGN + 3 + vc : CX_SGN + vc;
    const unsigned neg = (unsigned)(mq_decode(m, ctx) ^ xorbit);
    /* an insignificant sample is still zero: a plain store, no read of the plane on the decoder's critical path */
    b.val[(size_t)y * b.vstride + x] = (neg << 31) | mask;
    st[at] |= (uint8_t)(S_SIG | (neg ? S_NEG : 0));
    /* seen from the sample above, this one lies to its south, and so on */
    nb[up] |= NB_S;
    nb[dn] |= NB_N;
    nb[at - 4] |= NB_E;
    nb[at + 4] |= NB_W;
    nb[up - 4] |= NB_SE;
    nb[up + 4] |= NB_SW;
    nb[dn - 4] |= NB_NE;
    nb[dn + 4] |= NB_NW;
}

/* 0x80 in every byte of v that is not zero */
__device__ __forceinline__ uint32_t nonzero_bytes(uint32_t v)
{
    return (((v & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | v) & 0x80808080u;
}

/* decode_sigpass (:1872-1904) */
template <bool VSC>
__device__ void pass_sigprop(const Blk &b, Mq &m, int bp)
{
    const uint32_t mask = 3u << (bp - 1);
    uint8_t *st = reinterpret_cast<uint8_t *>(b.st4);
    const uint8_t *nb = reinterpret_cast<const uint8_t *>(b.nb4);
    for (int y0 = 0; y0 < b.h; y0 += 4) {
        const int rows = min(4, b.h - y0), base = ((y0 + 4) >> 2) * b.W2 + 1;
        for (int x = 0; x < b.w; x++) {
            /* candidates: a significant neighbour, not significant, not visited -- most columns have none */
            const uint32_t n4 = nb_word<VSC>(b.nb4[base + x]);
            if (!n4 || !(nonzero_bytes(n4) & ~nonzero_bytes(b.st4[base + x] & 0x05050505u)))
                continue;
            const int at = (base + x) << 2;
            for (int r = 0; r < rows; r++) {
                const uint8_t s = st[at + r];
                const unsigned n = nb_byte<VSC>(nb[at + r], r);
                if ((s & (S_SIG | S_VIS)) || !n)
                    continue;
                if (mq_decode(m, CX_SIG + sig_context(b, n)))
                    become_significant<VSC>(b, m, x, y0 + r, mask);
                st[at + r] |= S_VIS;
            }
        }
    }
}

/* decode_refpass (:1906-1931) */
template <bool VSC>
__device__ void pass_magref(const Blk &b, Mq &m, int bp)
{
    const uint32_t phalf = 1u << (bp - 1);
    uint8_t *st = reinterpret_cast<uint8_t *>(b.st4);
    const uint8_t *nb = reinterpret_cast<const uint8_t *>(b.nb4);
    for (int y0 = 0; y0 < b.h; y0 += 4) {
        const int rows = min(4, b.h - y0), base = ((y0 + 4) >> 2) * b.W2 + 1;
        for (int x = 0; x < b.w; x++) {
            const uint32_t s4 = b.st4[base + x];
            if (!(s4 & ~(s4 >> 2) & 0x01010101u))  /* nothing significant that was not just visited */
                continue;
            const int at = (base + x) << 2;
            for (int r = 0; r < rows; r++) {
                const uint8_t s = st[at + r];
                if ((s & (S_SIG | S_VIS)) != S_SIG)
                    continue;
                const int ctx = (s & S_REF) ? CX_REF + 2 : CX_REF + (nb_byte<VSC>(nb[at + r], r) ? 1 : 0);
                /* The reference does  v |= phalf;  v = bit ? v | 2 phalf : v & ~(2 phalf).  Bit 2 phalf is the
                 * half-LSB the previous coding step of this sample left set and the bits below it are clear, so
                 * that is "OR phalf" for a one and "XOR 3 phalf" for a zero: fire-and-forget reductions, no load
                 * of the plane (an L2 round trip per symbol when the value is read back). */
                uint32_t *p = b.val + (size_t)(y0 + r) * b.vstride + x;
                if (mq_decode(m, ctx))
                    atomicOr(p, phalf);
                else
                    atomicXor(p, 3u * phalf);
                st[at + r] = s | S_REF;
            }
        }
    }
}

/* decode_clnpass (:1933-1990); the segmentation symbol is read by the caller */
template <bool VSC>
__device__ void pass_cleanup(const Blk &b, Mq &m, int bp)
{
    const uint32_t mask = 3u << (bp - 1);
    uint8_t *st = reinterpret_cast<uint8_t *>(b.st4);
    const uint8_t *nb = reinterpret_cast<const uint8_t *>(b.nb4);
    for (int y0 = 0; y0 < b.h; y0 += 4) {
        const int rows = min(4, b.h - y0), base = ((y0 + 4) >> 2) * b.W2 + 1;
        for (int x = 0; x < b.w; x++) {
            int runlen = 0, dec = 0;
            const uint32_t s4 = b.st4[base + x];
            if (rows == 4 && nonzero_bytes(s4 & 0x05050505u) == 0x80808080u) {
                b.st4[base + x] = s4 & ~0x04040404u;       /* every sample significant or visited: only un-visit */
                continue;
            }
            /* a full column with nothing significant, visited or next to significance: run-length mode */
            if (rows == 4 && !(s4 & 0x05050505u) && !nb_word<VSC>(b.nb4[base + x])) {
                if (!mq_decode(m, CX_RL))
                    continue;
                runlen = mq_decode(m, CX_UNI);
                runlen = (runlen << 1) | mq_decode(m, CX_UNI);
                dec = 1;
            }
            const int at = (base + x) << 2;
            for (int r = runlen; r < rows; r++) {
                if (!dec && !(st[at + r] & (S_SIG | S_VIS)))
                    dec = mq_decode(m, CX_SIG + sig_context(b, nb_byte<VSC>(nb[at + r], r)));
                if (dec)
                    become_significant<VSC>(b, m, x, y0 + r, mask);
                dec = 0;
                st[at + r] &= (uint8_t)~S_VIS;
            }
        }
    }
}

/* the pass loop of decode_cblk (:2017-2059): clean-up first, then SigProp / MagRef / clean-up per bit-plane */
template <bool VSC>
__device__ int run_passes(const Blk &b, Mq &m, int npasses, int bpno, int sty)
{
    int pass_t = 2;
    while (npasses--) {
        if (bpno < 0 || bpno > 29)
            return B200J2K_KERR_INVALID;
        if (pass_t == 0) pass_sigprop<VSC>(b, m, bpno + 1);
        else if (pass_t == 1) pass_magref<VSC>(b, m, bpno + 1);
        else {
            pass_cleanup<VSC>(b, m, bpno + 1);
            if (sty & J2K_CBLK_MQ_SEGSYM)                   /* four UNIFORM decisions, 1010 expected; the reference
                                                            * only logs a mismatch (:1981-1989) */
                for (int i = 0; i < 4; i++)
                    mq_decode(m, CX_UNI);
        }
        if (sty & J2K_CBLK_MQ_RESET)                        /* JPEG2000_CBLK_RESET (:2038-2039) */
            mq_reset_contexts(m);
        if (++pass_t == 3) {
            bpno--;
            pass_t = 0;
        }
    }
    return 1;
}

/* Code-blocks (warps) per CTA.  One: a CTA's shared memory is free again the moment its block is done.  Two would
 * share the 1 KB the system reserves per CTA (22 instead of 21 blocks of 64x64 resident per SM), but a CTA then
 * lives as long as the slower of its two blocks: measured 13.6 ms instead of 11.7 ms for the 4K Part-1 frame. */
constexpr int MQ_WARPS_PER_CTA = 1;

__global__ void __launch_bounds__(32 * MQ_WARPS_PER_CTA)
mq_block_kernel(const CblkDesc *__restrict__ descs, const uint32_t *__restrict__ order, int nblocks,
                const uint8_t *__restrict__ bytes, const BandDesc *__restrict__ bands,
                uint32_t *__restrict__ coef, int32_t *__restrict__ status, int state_words)
{
    extern __shared__ __align__(16) uint32_t smem_all[]; /* per warp: st4 | nb4 (state_words in all) | 19 context states */
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int slot = (int)blockIdx.x * MQ_WARPS_PER_CTA + warp;
    if (slot >= nblocks)
        return;
    uint32_t *smem = smem_all + (size_t)warp * (state_words + 8);
    const uint32_t bi = order ? order[slot] : (uint32_t)slot;
    const CblkDesc d = descs[bi];
    const int w = d.w, h = d.h, W2 = w + 2, words = mq_state_words_of(w, h) / 2;
    uint32_t *dst = coef + d.dst_off;
    for (int y = 0; y < h; y++)
        for (int x = lane; x < w; x += 32)
            dst[(size_t)y * d.dst_stride + x] = 0;
    for (int i = lane; i < 2 * words; i += 32)
        smem[i] = 0;
    __syncwarp();
    int ret = 1;
    if (lane == 0) {
        Blk b = { w, h, W2, smem, smem + words, dst, d.dst_stride, (int)d.Lref };
        Mq m;
        mq_init(m, reinterpret_cast<uint8_t *>(smem + state_words), bytes + d.data_off, (int)d.length);
        const int bpno0 = (int)d.Lcup - 1 + 31 - d.M_b - 1 - d.roi_shift;      /* Lcup carries nonzerobits */
        const int sty = d.flags & (J2K_CBLK_MQ_RESET | J2K_CBLK_MQ_SEGSYM);
        ret = (d.flags & 1) ? run_passes<true>(b, m, d.npasses, bpno0, sty) : run_passes<false>(b, m, d.npasses, bpno0, sty);
        j2k_block_status(status, bi, ret);
    }
    ret = __shfl_sync(0xFFFFFFFFu, ret, 0);
    __syncwarp();
    const BandDesc bd = bands[d.band];
    const int mode = (d.flags >> 1) & 3;
    const int roi = ret == 1 ? d.roi_shift : 0;
    for (int y = 0; y < h; y++)
        for (int x = lane; x < w; x += 32) {
            uint32_t *p = dst + (size_t)y * d.dst_stride + x;
            *p = dequant(*p, mode, d.M_b, roi, bd);
        }
}

}  // namespace

int launch_mq_blocks(cudaStream_t stream, const CblkDesc *descs, const uint32_t *order, int nblocks, const uint8_t *bytes,
                     const BandDesc *bands, void *coef, int32_t *status, int state_words)
{
    if (nblocks <= 0)
        return 0;
    /* the two state arrays of the largest block of the frame: w * h <= 4096, so at most 2 x 12.3 KB (1024 x 4) */
    const int smem = (state_words * 4 + 32) * MQ_WARPS_PER_CTA;
    if (smem > 48 * 1024) {                     /* 1024 x 4 blocks: 2 x 24.6 KB */
        /* function attributes belong to the device: no caching across the devices a host thread may visit */
        if (cudaFuncSetAttribute((const void *)mq_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) !=
            cudaSuccess)
            return -1;
    }
    mq_block_kernel<<<(nblocks + MQ_WARPS_PER_CTA - 1) / MQ_WARPS_PER_CTA, 32 * MQ_WARPS_PER_CTA, smem, stream>>>(descs, order, nblocks, bytes, bands, (uint32_t *)coef, status,
                                                   state_words);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

}  // namespace b200j2k
