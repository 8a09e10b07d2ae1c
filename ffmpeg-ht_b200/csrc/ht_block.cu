/*
 * ht_block.cu -- HT (Rec. ITU-T T.814) code-block decoder for sm_100a, one code-block per warp, with the
 * dequantisation fused into the coefficient write-back.
 *
 * Replaces, for the HT path, ff_jpeg2000_decode_htj2k (libavcodec/jpeg2000htdec.c:1188-1336) and
 * dequantization_{int,float,int_97} (libavcodec/jpeg2000dec.c:2098-2181) as called from tile_codeblocks
 * (:2212-2299).  SURVEY.md section 8(a) rows H1-H9, D1-D3.
 *
 * Work split inside a warp (codeblock-per-warp, BASELINE.json north_star):
 *   1. all lanes: copy the cleanup segment's suffix (MEL + VLC bytes) to shared memory and UN-STUFF the
 *      MagSgn byte stream in parallel (warp prefix sum over 7/8-bit byte lengths) into a contiguous
 *      bit array in shared memory;
 *   2. per quad row, in segments of 32 quads: the context-adaptive part (MEL run decoder, CxtVLC table
 *      look-ups from shared memory, U-VLC) is inherently serial -- every lane executes it redundantly on
 *      warp-uniform data (shared-memory broadcasts, no divergence) and lane j keeps quad j's result;
 *   3. then lane j owns quad j: exponent predictor from the row above, warp prefix sum over the quads'
 *      MagSgn bit counts gives each lane its bit offset, four samples are extracted, dequantised and
 *      stored (two coalesced row segments per quad row);
 *   4. SigProp / MagRef refinement (only for blocks that carry them) runs on a shared-memory copy.
 * The CxtVLC tables (2 x 1024 x u16, cxtvlc_tables.h) are staged in shared memory per CTA.
 * Integer/bit work throughout: no tensor cores (nothing here is a contraction).
 */
/* Two warps per CTA: this kernel also takes the blocks whose segment lengths are out of any sane range (up to
 * 65535 bytes each, straight from damaged packet headers); even those must fit 227 KB of shared memory, so that a
 * crafted block costs its own decode and not the frame */
#ifndef B200J2K_WARPS_PER_CTA
#define B200J2K_WARPS_PER_CTA 2
#endif
#include "ht_common.cuh"

namespace b200j2k {


/*
 * Shared memory per CTA: [tables 4 KiB][per-warp region] x WARPS_PER_CTA.
 * Per-warp region (32-bit words):
 *   ms   [ms_words]          un-stuffed MagSgn bits (+2 words of ones)
 *   vl   [vl_words]          un-stuffed VLC bits, forward order (+2 words of zeros)
 *   row  [2][maxQW + 2]      per-quad (rho | E1<<8 | E3<<16) of the previous / current quad row
 *   qinfo[32]                the serial decoder's per-quad results of the current segment
 *   REFINE only:
 *   samp [maxw * maxh]       sign-magnitude samples
 *   st   [(maxh+3)*(maxw+3)] bytes, sample states with a one-sample apron
 * ms and vl share one allocation of `stream_words`: ms takes what Pcup needs, vl what Scup needs.
 */
template <bool REFINE>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
ht_decode_kernel(const CblkDesc *__restrict__ descs, const uint32_t *__restrict__ order, int nblocks,
                 const uint8_t *__restrict__ bytes, const BandDesc *__restrict__ bands,
                 uint32_t *__restrict__ coef, int32_t *__restrict__ status,
                 int stream_words, int maxQW, int samp_words, int st_words)
{
    extern __shared__ __align__(16) uint32_t smem[];
    uint16_t *tab = reinterpret_cast<uint16_t *>(smem);
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) {
        tab[i] = CXTVLC_FIRST_ROW[i];
        tab[1024 + i] = CXTVLC_OTHER_ROWS[i];
    }
    uint16_t *uvtab = tab + 2048;
    for (int i = threadIdx.x; i < 256; i += blockDim.x)
        uvtab[i] = uvlc_pair_entry(i);
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int row_words = 2 * (maxQW + 2) + 32;
    int per_warp = stream_words + row_words;
    if (REFINE)
        per_warp += samp_words + st_words;
    uint32_t *wbase = smem + TABLE_WORDS + (size_t)warp * per_warp;
    uint32_t *streams = wbase;
    uint32_t *rowbuf = wbase + stream_words;
    uint32_t *samp = rowbuf + row_words;
    uint8_t *st = reinterpret_cast<uint8_t *>(samp + samp_words);

    const int gw = blockIdx.x * WARPS_PER_CTA + warp;
    if (gw >= nblocks)
        return;
    const uint32_t bi = order ? order[gw] : (uint32_t)gw;
    const CblkDesc d = descs[bi];
    const BandDesc bd = bands[d.band];
    const int w = d.w, h = d.h;
    const int mode = (d.flags >> 1) & 3;
    const int M_b = d.M_b;
    uint32_t *dst = coef + d.dst_off;
    const uint32_t dstride = d.dst_stride;

    /* ---- H1: pass bookkeeping (jpeg2000htdec.c:1237-1278) ---- */
    int result = 1;
    int z_blk = 0, p0 = 0;
    if (d.npasses == 0) {
        result = 0;
    } else {
        int rem = d.npasses % 3;
        int plhd = rem ? d.npasses - rem : d.npasses - 3;
        p0 = plhd / 3;
        z_blk = d.npasses - plhd;
        if (z_blk <= 0)
            result = 0;
    }
    const int Lcup = d.Lcup, Lref = d.Lref;
    const uint8_t *Dcup = bytes + d.data_off;
    int Scup = 0, Pcup = 0;
    const int S_blk = (p0 + d.zbp) & 0xFF;
    const int pLSB = 30 - S_blk;
    const int maxbp = S_blk + 1;
    if (result > 0) {
        if (Lcup < 2) {
            result = B200J2K_KERR_INVALID;
        } else {
            Scup = (Dcup[Lcup - 1] << 4) + (Dcup[Lcup - 2] & 0x0F);
            if (Scup < 2 || Scup > Lcup || Scup > 4079)
                result = B200J2K_KERR_INVALID;
            else if (maxbp >= 32 || pLSB < 1)
                result = B200J2K_KERR_INVALID;
            else if (stream_words < (Lcup - Scup + 3) / 4 + (Scup + 3) / 4 + 6)
                result = B200J2K_KERR_INVALID;          /* launch sized too small: host bug guard */
        }
        Pcup = Lcup - Scup;
    }

    if (result <= 0) {
        /* nothing decoded or damaged: the block is all zeros (quirk Q1) */
        for (int y = 0; y < h; y++)
            for (int x = lane; x < w; x += 32)
                dst[(size_t)y * dstride + x] = 0;
        if (lane == 0)
            j2k_block_status(status, bi, result);
        return;
    }

    /* ---- stage the streams ---- */
    const int ms_words = (Pcup + 3) / 4 + 2;
    const int vl_words = (Scup + 3) / 4 + 4;
    uint32_t *ms = streams;
    uint32_t *vl = streams + ms_words;
    for (int i = lane; i < ms_words + vl_words; i += 32)
        streams[i] = 0;
    __syncwarp();
    /* un-stuffed VLC bits (backward reader, with the reference's tail behind them) and MagSgn bits (forward,
     * ones behind): see unstuff() in ht_common.cuh */
    BwdTail vtail;
    {
        const uint32_t vbits = unstuff<US_VLC>(Dcup + Lcup - 2, Scup - 1, vl, lane);
        unsigned b0 = Dcup[Pcup], b1 = Dcup[Pcup + 1];
        if (Scup == 2) {
            b0 |= 0x0F;
            b1 = 0xFF;
        } else if (Scup == 3) {
            b1 |= 0x0F;
        }
        vtail = bwd_tail_make(vbits, Scup - 1, b0, b1);
        bwd_tail_fill(vl, vl_words, vtail, lane);
    }
    const uint32_t ms_bits = unstuff<US_MAGSGN>(Dcup, Pcup, ms, lane);
    ones_behind(ms, ms_bits, ms_words, lane);

    const int QW = (w + 1) >> 1, QH = (h + 1) >> 1;
    uint32_t *rowprev = rowbuf, *rowcur = rowbuf + (maxQW + 2);
    for (int i = lane; i < 2 * (maxQW + 2); i += 32)
        rowbuf[i] = 0;
    if (REFINE) {
        for (int i = lane; i < (h + 3) * (w + 3); i += 32)
            st[i] = 0;
    }
    __syncwarp();

    Mel mel = { Dcup + Pcup, Scup, 0, 0, 0, 0, 0, 0 };
    uint32_t vpos = 4;                                   /* the Scup nibble is skipped (:283-295) */
    const VlcArr varr = { vl, vl_words, vtail };
    uint32_t *qinfo = rowbuf + 2 * (maxQW + 2);
    uint32_t ms_pos = 0;
    bool bad = false;
    const int SW = w + 3;                                /* state row stride (REFINE) */

    for (int qy = 0; qy < QH && !bad; qy++) {
        int carry = 0;
        for (int qb = 0; qb < QW; qb += 32) {
            const int nq = min(32, QW - qb);
            const int qx = qb + lane;
            if (qy == 0) {
                decode_segment<true>(mel, varr, vpos, tab, uvtab, nq, 0, 0, carry, qinfo);
            } else {
                /* row[] is stored with a +1 offset and zero sentinels at both ends */
                const unsigned n = qx < QW ? rowprev[qx + 1] : 0, nw = qx < QW ? rowprev[qx] : 0,
                               ne = qx < QW ? rowprev[qx + 2] : 0;
                const unsigned mA = __ballot_sync(FULL, ((n >> 1) | (nw >> 3)) & 1);   /* sigma_N[1] | sigma_NW[3] */
                const unsigned mC = __ballot_sync(FULL, ((n >> 3) | (ne >> 1)) & 1);   /* sigma_N[3] | sigma_NE[1] */
                decode_segment<false>(mel, varr, vpos, tab + 1024, uvtab, nq, mA, mC, carry, qinfo);
            }
            __syncwarp();
            const unsigned mine = lane < nq ? qinfo[lane] : 0;   /* u_off | rho<<1 | emb_k<<5 | emb_1<<9 | u<<16 */
            __syncwarp();
            /* ---- parallel: lane = quad.  exponent predictor, MagSgn, write-back ---- */
            const bool act = lane < nq;
            const unsigned rho = act ? (mine >> 1) & 15 : 0;
            const unsigned ek = (mine >> 5) & 15, e1 = (mine >> 9) & 15;
            const int u = (int)(mine >> 16);
            int kappa = 1;
            if (qy > 0 && act) {                         /* :855-885 */
                unsigned n = rowprev[qx + 1], nw = rowprev[qx], ne = rowprev[qx + 2];
                int m = max(max((int)((n >> 8) & 63), (int)((n >> 16) & 63)),
                            max((int)((nw >> 16) & 63), (int)((ne >> 8) & 63)));
                int gamma = (rho & (rho - 1)) != 0;
                kappa = max(1, gamma * (m - 1));
            }
            const int U = kappa + u;
            if (__any_sync(FULL, act && U > maxbp)) {    /* :715, :889 */
                bad = true;
                break;
            }
            int mlen[4];
            int tot = 0;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                mlen[i] = act ? (int)((rho >> i) & 1) * U - (int)((ek >> i) & 1) : 0;
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
                    uint32_t p = min(pos, ms_bits);      /* past the end: ones */
                    v = ms_read(ms, p, m) + (((e1 >> i) & 1u) << m);
                    pos += m;
                }
                if (m != 0) {                            /* :419-425 */
                    unsigned E = 32 - __clz(v | 1);
                    if (i == 1) E1 = E;
                    if (i == 3) E3 = E;
                    val = (uint32_t)(((int32_t)v >> 1) + 1) << pLSB;
                    val |= 1u << (pLSB - 1);
                    val |= (v & 1u) << 31;
                }
                mu[i] = val;
            }
            if (act)
                rowcur[qx + 1] = rho | (E1 << 8) | (E3 << 16);
            /* samples of quad (qy, qx): i=0 (y,x) i=1 (y+1,x) i=2 (y,x+1) i=3 (y+1,x+1) */
            const int y = 2 * qy, x = 2 * qx;
            if (!REFINE) {
                if (act) {
                    if (x < w) {
                        dst[(size_t)y * dstride + x] = dequant(mu[0], mode, M_b, d.roi_shift, bd);
                        if (y + 1 < h)
                            dst[(size_t)(y + 1) * dstride + x] = dequant(mu[1], mode, M_b, d.roi_shift, bd);
                    }
                    if (x + 1 < w) {
                        dst[(size_t)y * dstride + x + 1] = dequant(mu[2], mode, M_b, d.roi_shift, bd);
                        if (y + 1 < h)
                            dst[(size_t)(y + 1) * dstride + x + 1] = dequant(mu[3], mode, M_b, d.roi_shift, bd);
                    }
                }
            } else if (act) {
                /* quad -> raster with the reference's odd-edge masks (:975-1007, quirk Q2) */
                const int keep_row = qy != QH - 1 || !(h & 1);
                const int keep_col = qx != QW - 1 || !(w & 1);
                const int keep[4] = { 1, keep_row, keep_col, keep_row | keep_col };
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const int yy = y + (i & 1), xx = x + (i >> 1);
                    if (keep[i]) {
                        st[(yy + 1) * SW + xx + 1] |= (rho >> i) & 1;
                        if (yy < h && xx < w)
                            samp[yy * w + xx] = mu[i];
                    } else if (yy < h && xx < w) {
                        samp[yy * w + xx] = 0;
                    }
                }
            }
        }
        __syncwarp();
        uint32_t *t = rowprev;
        rowprev = rowcur;
        rowcur = t;
    }

    if (bad) {
        __syncwarp();
        for (int y = 0; y < h; y++)
            for (int x = lane; x < w; x += 32)
                dst[(size_t)y * dstride + x] = 0;
        if (lane == 0)
            j2k_block_status(status, bi, B200J2K_KERR_INVALID);
        return;
    }

    if (REFINE) {
        __syncwarp();
        const int p = pLSB - 1;
        const uint8_t *Dref = Dcup + Lcup;
        if (lane == 0) {
            if (z_blk > 1) {
                /* H7 SigProp (:1034-1131): stripes of 4 rows, 4-column groups, column-major */
                SpBits sp = { Dref, Lref, 0, 0, 0, 0 };
                const int causal = d.flags & 1;
                for (int y0 = 0; y0 < h; y0 += 4) {
                    const int sh = min(4, h - y0);
                    for (int x0 = 0; x0 < w; x0 += 4) {
                        const int sw = min(4, w - x0);
                        for (int x = x0; x < x0 + sw; x++)
                            for (int y = y0; y < y0 + sh; y++) {
                                uint8_t *c = &st[(y + 1) * SW + x + 1];
                                int member = 0;
                                if (!(c[0] & ST_SIGMA)) {
                                    const int below_ok = !causal || y != y0 + sh - 1;
                                    unsigned up = c[-SW - 1] | c[-SW] | c[-SW + 1];
                                    unsigned mid = c[-1] | c[1];
                                    unsigned dn = c[SW - 1] | c[SW] | c[SW + 1];
                                    unsigned m = up | mid | (below_ok ? dn : 0);
                                    m |= ((up >> 3) & (up >> 4)) | ((mid >> 3) & (mid >> 4));
                                    if (below_ok)
                                        m |= (dn >> 3) & (dn >> 4);
                                    member = m & 1;
                                }
                                unsigned s = c[0] | ST_SCANNED;
                                if (member) {
                                    const unsigned bit = sp_bit(sp);
                                    if (bit)
                                        s |= ST_NEWSIG;
                                    samp[y * w + x] |= (bit << p) | (bit << (p - 1));
                                }
                                c[0] = (uint8_t)s;
                            }
                        for (int x = x0; x < x0 + sw; x++)
                            for (int y = y0; y < y0 + sh; y++)
                                if (st[(y + 1) * SW + x + 1] & ST_NEWSIG)
                                    samp[y * w + x] |= (uint32_t)sp_bit(sp) << 31;
                    }
                }
            }
            if (z_blk > 2) {
                /* H8 MagRef (:1136-1185): bits read backwards from the end of Dref with the reference's group
                 * reader (see BwdTail): byte k of its read order and the byte it compares against */
                auto ref_byte = [&](int i) -> unsigned {
                    if (i < Lref)
                        return Dref[i];
                    const uint32_t at = (uint32_t)Lcup + (uint32_t)i;
                    return at == d.length ? 0xFFu : (at < d.length ? Dref[i] : 0u);
                };
                const int mpad = 4 - (Lref & 3);
                int mk = 0;
                unsigned prev = ref_byte(Lref);
                uint32_t acc = 0;
                int n = 0;
                for (int y0 = 0; y0 < h; y0 += 4) {
                    const int sh = min(4, h - y0);
                    for (int x = 0; x < w; x++)
                        for (int y = y0; y < y0 + sh; y++) {
                            if (!(st[(y + 1) * SW + x + 1] & ST_SIGMA))
                                continue;
                            if (n == 0) {
                                unsigned b = 0;
                                if (mk < Lref) {
                                    b = Dref[Lref - 1 - mk];
                                } else if (mk >= Lref + mpad && ((mk - Lref - mpad) & 3) == 0) {
                                    b = ref_byte(0);             /* parked reader: base[0] again, compared with base[1] */
                                    prev = ref_byte(1);
                                }
                                mk++;
                                n = ((b & 0x7F) == 0x7F && prev > 0x8F) ? 7 : 8;
                                acc = b & ((1u << n) - 1);
                                prev = b;
                            }
                            const uint32_t bit = acc & 1;
                            acc >>= 1;
                            n--;
                            const uint32_t keepm = (0xFFFFFFFEu | bit) << p;
                            samp[y * w + x] = (samp[y * w + x] & keepm) | (1u << (p - 1));
                        }
                }
            }
        }
        __syncwarp();
        for (int y = 0; y < h; y++)
            for (int x = lane; x < w; x += 32)
                dst[(size_t)y * dstride + x] = dequant(samp[y * w + x], mode, M_b, d.roi_shift, bd);
    }
    if (lane == 0 && status)
        status[bi] = 1;
}

/* ---- host-side launcher ------------------------------------------------------------------------------ */
size_t ht_smem_bytes(bool refine, int stream_words, int maxQW, int samp_words, int st_words)
{
    size_t per_warp = (size_t)stream_words + 2 * (size_t)(maxQW + 2) + 32;
    if (refine)
        per_warp += (size_t)samp_words + st_words;
    return (size_t)TABLE_WORDS * 4 + per_warp * 4 * WARPS_PER_CTA;
}

int launch_ht_decode(cudaStream_t stream, bool refine, const CblkDesc *descs, const uint32_t *order, int nblocks,
                     const uint8_t *bytes, const BandDesc *bands, void *coef, int32_t *status,
                     int max_Lcup, int maxw, int max_area, int max_state)
{
    if (nblocks <= 0)
        return 0;
    const int stream_words = (max_Lcup + 3) / 4 + 10;
    const int maxQW = (maxw + 1) / 2;
    const int samp_words = max_area, st_words = (max_state + 3) / 4;
    const size_t smem = ht_smem_bytes(refine, stream_words, maxQW, samp_words, st_words);
    if (smem > 227 * 1024)
        return -1;
    const int grid = (nblocks + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
    cudaError_t err;
    if (refine) {
        err = cudaFuncSetAttribute(ht_decode_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess)
            return -2;
        ht_decode_kernel<true><<<grid, WARPS_PER_CTA * 32, smem, stream>>>(
            descs, order, nblocks, bytes, bands, (uint32_t *)coef, status, stream_words, maxQW, samp_words, st_words);
    } else {
        err = cudaFuncSetAttribute(ht_decode_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess)
            return -2;
        ht_decode_kernel<false><<<grid, WARPS_PER_CTA * 32, smem, stream>>>(
            descs, order, nblocks, bytes, bands, (uint32_t *)coef, status, stream_words, maxQW, samp_words, st_words);
    }
    return cudaGetLastError() == cudaSuccess ? 1 : -3;
}

}  // namespace b200j2k
