/*
 * idwt.cu -- inverse DWT (5/3 integer, 9/7 float, 9/7 fixed point), inverse RCT/ICT, DC level shift,
 * clip and pixel-format pack for sm_100a.
 *
 * Replaces ff_dwt_decode (libavcodec/jpeg2000dwt.c:309-537, :601-620), mct_decode /
 * rct_int / ict_float / ict_int (libavcodec/jpeg2000dec.c:2183-2209, libavcodec/jpeg2000dsp.c:43-91; rows
 * M1-M3) and write_frame_8/16 (libavcodec/jpeg2000dec.c:2301-2366; row P1).
 *
 * One launch per decomposition level, all tile-components of the frame in the grid.  A CTA reconstructs a
 * TX x TY tile of the level's output: it gathers the four sub-band quadrants (LL from the previous
 * level's output, HL/LH/HH from the Mallat-layout coefficient plane the HT kernel filled) with coalesced
 * row-segment loads into an interleaved shared-memory tile with a HALO-sample apron, runs the row
 * lifting steps then the column lifting steps in shared memory (rows before columns at every level,
 * which is what makes the 5/3 path bit-exact, quirk Q7), and writes the tile once.  Image borders use
 * whole-sample symmetric extension by index reflection, bit-identical to the reference's sequential
 * copies.  The last level is fused with the component transform, level shift, clip
 * and pack, so the full-resolution planes are never written to HBM: per sample the stage reads 4 B and
 * writes 2 B (1 B for 8-bit formats) -- the algorithmic traffic SURVEY.md section 8(d) names.
 *
 * Float path: every operation is an explicit __fmul_rn / __fadd_rn / __fsub_rn in the reference's order
 * (a - c*(b+d) as add, multiply, subtract), so no FMA contraction can change a bit; lrintf is
 * __float2int_rn (round to nearest even).
 */
#include <cstdlib>

#include "idwt_common.cuh"

namespace b200j2k {

namespace {

/* Tile geometry.  A warp reconstructs 32 sample pairs (even position, odd position) of a row at a time with
 * shuffles; HP pairs at each end of the warp are apron, so a warp iteration yields VP = 32 - 2*HP valid
 * pairs and a tile row is two iterations wide. */
template <int TYPE> struct Geo {
    static constexpr int HALO = Lift<TYPE>::HALO;
    static constexpr int HP = HALO / 2;
    static constexpr int VP = 32 - 2 * HP;
    static constexpr int ITERS = 2;
    static constexpr int TX = 2 * VP * ITERS;       /* 120 samples for 5/3, 112 for 9/7 */
};

/*
 * HOR_SD for SH rows of a tile, result into tile[SH][TX] (shared memory).
 * XA: even absolute X of tile column 0; YS: even absolute Y of tile row 0 (apron included).
 * Lane = one (low, high) coefficient pair of the row: the low-pass sample sits at the even absolute
 * position 2i, the high-pass one at 2i+1; the lifting steps exchange neighbours with shuffles
 * (jpeg2000dwt.c:309-325, :376-401, :453-481).  Rows and columns outside the resolution are fetched by
 * index reflection, which reproduces the reference's symmetric extension.
 */
/* row lifting of one warp-wide group of 32 coefficient pairs (jpeg2000dwt.c:309-325, :376-401, :453-481) */
template <int TYPE>
__device__ __forceinline__ void hlift(uint32_t Lc, uint32_t Hc, uint32_t &xe, uint32_t &xo)
{
    using L = Lift<TYPE>;
    if (TYPE == J2K_DWT53) {
        const uint32_t Hl = __shfl_up_sync(FULLMASK, Hc, 1);
        xe = L::apply(0, Lc, Hl, Hc);
        const uint32_t er = __shfl_down_sync(FULLMASK, xe, 1);
        xo = L::apply(1, Hc, xe, er);
    } else {
        const uint32_t Hl = __shfl_up_sync(FULLMASK, Hc, 1);
        const uint32_t e1 = L::apply(0, Lc, Hl, Hc);
        const uint32_t e1r = __shfl_down_sync(FULLMASK, e1, 1);
        const uint32_t o1 = L::apply(1, Hc, e1, e1r);
        const uint32_t o1l = __shfl_up_sync(FULLMASK, o1, 1);
        xe = L::apply(2, e1, o1l, o1);
        const uint32_t er = __shfl_down_sync(FULLMASK, xe, 1);
        xo = L::apply(3, o1, xe, er);
    }
}

/* Per-tile row table (shared memory): where the low-pass and high-pass coefficients of tile row r live.
 * offL >= 0: element offset into the LL source; offL < 0: ~offL is an element offset into the plane
 * (LH band).  offH: element offset into the plane (HL or HH band). */
struct RowTab {
    int offL, offH;
};

template <int SH>
__device__ __forceinline__ void build_rowtab(RowTab *tab, const TileCompDesc &tc, int j, int YS)
{
    const LevelDesc o = tc.lev[j + 1];
    const int u0 = o.u0, u1 = o.u1, v0 = o.v0, v1 = o.v1;
    const int nLh = ((u1 + 1) >> 1) - ((u0 + 1) >> 1), nLv = ((v1 + 1) >> 1) - ((v0 + 1) >> 1);
    const int W = tc.width;
    const int LLstride = j == 0 ? W : nLh;
    for (int r = threadIdx.x; r < SH; r += THREADS) {
        const int Y = YS + r;
        const int Yr = (Y >= v0 && Y < v1) ? Y : reflect(Y, v0, v1);
        const int ky = band_index(Yr, v0);
        RowTab t;
        if (Yr & 1) {
            t.offL = ~((nLv + ky) * W);
            t.offH = (nLv + ky) * W + nLh;
        } else {
            t.offL = ky * LLstride;
            t.offH = ky * W + nLh;
        }
        tab[r] = t;
    }
}

/*
 * HOR_SD for the SH rows of NC tiles (one per component of a pack group), results into tiles[c][SH][TX]
 * (shared memory).  XA: even absolute X of tile column 0.  Lane = one (low, high) coefficient pair of a
 * row: the low-pass sample sits at the even absolute position 2i, the high-pass one at 2i+1; the lifting
 * steps exchange neighbours with shuffles.  Rows and columns outside the resolution are fetched by index
 * reflection, which reproduces the reference's symmetric extension.  Each warp first
 * issues the loads of a batch of rows, then lifts them: the batch keeps several kilobytes per warp in
 * flight (the kernel is bound by HBM latency x parallelism, not by arithmetic).
 */
template <int TYPE, int SH, int BATCH>
__device__ __forceinline__ void hpass(uint32_t *tiles, const RowTab *tab, const uint32_t *const *Ap,
                                      const uint32_t *const *LLp, int ncomp, int u0, int u1, int XA)
{
    using L = Lift<TYPE>;
    using G = Geo<TYPE>;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int lh = u1 - u0;
    const int eoff = (u0 + 1) >> 1, ooff = u0 >> 1;
    /* column indices of this lane's pairs: identical for every row */
    int ie[G::ITERS], io[G::ITERS];
#pragma unroll
    for (int it = 0; it < G::ITERS; it++) {
        const int pair = (XA >> 1) + it * G::VP - G::HP + lane;
        int Xe = 2 * pair, Xo = Xe + 1;
        if (lh > 1) {
            if (Xe < u0 || Xe >= u1) Xe = reflect(Xe, u0, u1);
            if (Xo < u0 || Xo >= u1) Xo = reflect(Xo, u0, u1);
            ie[it] = (Xe >> 1) - eoff;
            io[it] = (Xo >> 1) - ooff;
        } else {
            ie[it] = io[it] = 0;
        }
    }
    const int nitems = ncomp * SH;
    for (int base = warp; base < nitems; base += NWARPS * BATCH) {
        uint32_t Lc[BATCH][G::ITERS], Hc[BATCH][G::ITERS];
#pragma unroll
        for (int b = 0; b < BATCH; b++) {
            const int item = base + b * NWARPS;
            if (item < nitems) {
                const int c = item / SH, r = item - c * SH;
                const RowTab t = tab[r];
                const uint32_t *rowL = t.offL >= 0 ? LLp[c] + t.offL : Ap[c] + ~t.offL;
                const uint32_t *rowH = Ap[c] + t.offH;
#pragma unroll
                for (int it = 0; it < G::ITERS; it++) {
                    if (lh > 1) {
                        Lc[b][it] = rowL[ie[it]];
                        Hc[b][it] = rowH[io[it]];
                    } else {
                        Lc[b][it] = Hc[b][it] = (u0 & 1) ? rowH[0] : rowL[0];
                    }
                }
            }
        }
#pragma unroll
        for (int b = 0; b < BATCH; b++) {
            const int item = base + b * NWARPS;
            if (item < nitems) {
                const int c = item / SH, r = item - c * SH;
                uint32_t *trow = tiles + ((size_t)c * SH + r) * G::TX;
#pragma unroll
                for (int it = 0; it < G::ITERS; it++) {
                    uint32_t xe, xo;
                    if (lh > 1)
                        hlift<TYPE>(Lc[b][it], Hc[b][it], xe, xo);
                    else     /* one-sample line (jpeg2000dwt.c:313-317, :380-386, :457-463) */
                        xe = xo = L::single(Lc[b][it], u0 & 1);
                    if (lane >= G::HP && lane < 32 - G::HP)
                        *reinterpret_cast<uint2 *>(trow + 2 * (it * G::VP + lane - G::HP)) = make_uint2(xe, xo);
                }
            }
        }
    }
}

/* VER_SD in registers: v[0..R) is one column, v[k] at absolute row (even base + k); the centre
 * v[HALO .. R-HALO) is valid afterwards. */
template <int TYPE, int R>
__device__ __forceinline__ void vlift(uint32_t (&v)[R], int lv, int v0)
{
    using L = Lift<TYPE>;
    if (lv > 1) {
#pragma unroll
        for (int s = 0; s < L::STEPS; s++) {
#pragma unroll
            for (int k = s + 1; k < R - s - 1; k++)
                if ((k & 1) == (s & 1))
                    v[k] = L::apply(s, v[k], v[k - 1], v[k + 1]);
        }
    } else {
#pragma unroll
        for (int k = 0; k < R; k++)
            v[k] = L::single(v[k], v0 & 1);
    }
}

/* ---- intermediate level: write the reconstructed resolution to the compact ping-pong buffer ---------- */
template <int TYPE, int TY>
__global__ void __launch_bounds__(THREADS)
idwt_level_kernel(const TileCompDesc *__restrict__ tcs, int j, const uint32_t *__restrict__ coefA,
                  const uint32_t *__restrict__ coefIn, uint32_t *__restrict__ coefOut, int to_plane)
{
    using G = Geo<TYPE>;
    constexpr int HALO = G::HALO, TX = G::TX, SH = TY + 2 * HALO;
    constexpr int NHALF = 2, RH = TY / NHALF, R = RH + 2 * HALO;
    static_assert(TX * NHALF <= THREADS, "one thread per (column, half)");
    __shared__ __align__(16) uint32_t tile[SH * TX];
    const TileCompDesc &tc = tcs[blockIdx.z];
    if (tc.transform != TYPE || j >= tc.nlevels)
        return;
    /* the last level of a plane is done by the fused pack kernel, unless the caller wants the plane */
    if (!to_plane && j == tc.nlevels - 1)
        return;
    const LevelDesc o = tc.lev[j + 1];
    const int lh = o.u1 - o.u0, lv = o.v1 - o.v0;
    const int XA = (o.u0 & ~1) + blockIdx.x * TX, YA = (o.v0 & ~1) + blockIdx.y * TY;
    if (XA >= o.u1 || YA >= o.v1)
        return;
    __shared__ RowTab rowtab[SH];
    build_rowtab<SH>(rowtab, tc, j, YA - HALO);
    __syncthreads();
    const uint32_t *Ap[1] = { coefA + tc.coef_off };
    const uint32_t *LLp[1] = { j == 0 ? Ap[0] : coefIn + tc.coef_off };
    hpass<TYPE, SH, 3>(tile, rowtab, Ap, LLp, 1, o.u0, o.u1, XA);
    __syncthreads();
    const int col = threadIdx.x % TX, half = threadIdx.x / TX;
    if (half >= NHALF)
        return;
    uint32_t v[R];
#pragma unroll
    for (int k = 0; k < R; k++)
        v[k] = tile[(half * RH + k) * TX + col];
    vlift<TYPE, R>(v, lv, o.v0);
    uint32_t *out = coefOut + tc.coef_off;
    const int ostride = (to_plane && j == tc.nlevels - 1) ? tc.width : lh;
    const int X = XA + col;
    if (X >= o.u0 && X < o.u1) {
#pragma unroll
        for (int k = 0; k < RH; k++) {
            const int Y = YA + half * RH + k;
            if (Y >= o.v0 && Y < o.v1)
                out[(size_t)(Y - o.v0) * ostride + (X - o.u0)] = v[HALO + k];
        }
    }
}

/* final fixed-point rescale of the 9/7 integer path when the plane itself is the output (jpeg2000dwt.c:534-536) */
__global__ void rescale97i_kernel(uint32_t *p, size_t n)
{
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n)
        p[i] = (uint32_t)((int32_t)((int64_t)(int32_t)p[i] + 128) >> 8);
}

/* ---- last level fused with MCT + level shift + clip + pack -------------------------------------------- */
template <int TYPE, int TY>
__global__ void __launch_bounds__(THREADS, 4)
idwt_final_pack_kernel(const TileCompDesc *__restrict__ tcs, const PackGroup *__restrict__ groups,
                       const uint32_t *__restrict__ coefA, const uint32_t *__restrict__ coefIn0,
                       const uint32_t *__restrict__ coefIn1, uint8_t *__restrict__ picture,
                       uint64_t po0, uint64_t po1, uint64_t po2, uint64_t po3, int rb0, int rb1, int rb2, int rb3)
{
    using G = Geo<TYPE>;
    constexpr int HALO = G::HALO, TX = G::TX, SH = TY + 2 * HALO;
    constexpr int NHALF = 2, RH = TY / NHALF, R = RH + 2 * HALO;
    extern __shared__ __align__(16) uint32_t dyn[];
    const PackGroup &g = groups[blockIdx.z];
    const TileCompDesc &t0 = tcs[g.tc[0]];
    if (t0.transform != TYPE)
        return;
    const int nl = t0.nlevels;
    const LevelDesc o = t0.lev[nl];
    const int XA = (o.u0 & ~1) + blockIdx.x * TX, YA = (o.v0 & ~1) + blockIdx.y * TY;
    if (XA >= o.u1 || YA >= o.v1)
        return;
    const int ncomp = g.ncomp;
    if (nl > 0) {
        /* LL of the last level is the output of level nl-2: ping-pong parity */
        const uint32_t *in = ((nl - 1) & 1) ? coefIn0 : coefIn1;
        __shared__ RowTab rowtab[SH];
        build_rowtab<SH>(rowtab, t0, nl - 1, YA - HALO);
        __shared__ const uint32_t *Ap[J2K_MAX_COMP], *LLp[J2K_MAX_COMP];
        if (threadIdx.x < J2K_MAX_COMP) {
            const int c = threadIdx.x;
            const uint32_t off = tcs[g.tc[c < ncomp ? c : 0]].coef_off;
            Ap[c] = coefA + off;
            LLp[c] = nl == 1 ? coefA + off : in + off;
        }
        __syncthreads();
        hpass<TYPE, SH, 4>(dyn, rowtab, Ap, LLp, ncomp, o.u0, o.u1, XA);
        __syncthreads();
        const int col = threadIdx.x % TX, half = threadIdx.x / TX;
        const int lv = o.v1 - o.v0;
        for (int c = 0; c < ncomp; c++) {
            uint32_t *tile = dyn + (size_t)c * SH * TX;
            uint32_t v[R];
            if (half < NHALF) {
#pragma unroll
                for (int k = 0; k < R; k++)
                    v[k] = tile[(half * RH + k) * TX + col];
                vlift<TYPE, R>(v, lv, o.v0);
            }
            __syncthreads();                      /* every column has been read: rows can be overwritten */
            if (half < NHALF) {
#pragma unroll
                for (int k = 0; k < RH; k++)
                    tile[(HALO + half * RH + k) * TX + col] = v[HALO + k];
            }
        }
        __syncthreads();
    } else {
        for (int c = 0; c < ncomp; c++) {
            const TileCompDesc &tc = tcs[g.tc[c]];
            const uint32_t *A = coefA + tc.coef_off;
            uint32_t *tile = dyn + (size_t)c * SH * TX;
            for (int idx = threadIdx.x; idx < TX * TY; idx += THREADS) {
                const int r = idx / TX, cx = idx - r * TX;
                const int X = XA + cx, Y = YA + r;
                if (X >= o.u0 && X < o.u1 && Y >= o.v0 && Y < o.v1)
                    tile[(r + HALO) * TX + cx] = A[(size_t)(Y - o.v0) * tc.width + (X - o.u0)];
            }
        }
        __syncthreads();
    }
    /* ---- MCT + level shift + clip + pack: one thread per pixel pair ---- */
    const uint64_t poff = g.plane == 0 ? po0 : g.plane == 1 ? po1 : g.plane == 2 ? po2 : po3;
    const int rb = g.plane == 0 ? rb0 : g.plane == 1 ? rb1 : g.plane == 2 ? rb2 : rb3;
    uint8_t *plane = picture + poff;
    int cbps[J2K_MAX_COMP], shift[J2K_MAX_COMP], soff[J2K_MAX_COMP];
#pragma unroll
    for (int c = 0; c < J2K_MAX_COMP; c++) {
        cbps[c] = c < ncomp ? tcs[g.tc[c]].cbps : 1;
        shift[c] = g.shift[c];
        soff[c] = g.sample_off[c];
    }
    const int mct_on = g.mct, wide = g.wide, pstride = g.pixel_stride;
    constexpr int PPR = TX / 2;                   /* pixel pairs per tile row */
    constexpr int RGROUPS = THREADS / PPR;        /* row groups: thread = (pixel pair, row group) */
    const int pp = threadIdx.x % PPR, rg = threadIdx.x / PPR;
    const int X = XA + 2 * pp;
    const bool okA = X >= o.u0 && X < o.u1, okB = X + 1 >= o.u0 && X + 1 < o.u1;
    const long px = (long)(t0.out_x + X - o.u0) * pstride;          /* sample index of pixel A, component 0 */
    for (int r = rg; r < TY && rg < RGROUPS && (okA || okB); r += RGROUPS) {
        const int Y = YA + r;
        if (Y < o.v0 || Y >= o.v1)
            continue;
        uint32_t a[J2K_MAX_COMP], b[J2K_MAX_COMP];
#pragma unroll
        for (int c = 0; c < J2K_MAX_COMP; c++) {
            if (c < ncomp) {
                const uint2 t = *reinterpret_cast<const uint2 *>(dyn + (size_t)c * SH * TX + (r + HALO) * TX + 2 * pp);
                a[c] = t.x;
                b[c] = t.y;
            } else {
                a[c] = b[c] = 0;
            }
        }
        if (TYPE == J2K_DWT97_INT && nl > 0) {
            /* undo I_PRESHIFT: the last thing dwt_decode97_int does, i.e. before the component transform
             * (jpeg2000dwt.c:534-536).  ff_dwt_decode returns before it when there is no level to undo
             * (jpeg2000dwt.c:603-604): a plane without decomposition keeps its 2^8 pre-scale, as in the
             * reference (reachable with lowres == levels under AV_CODEC_FLAG_BITEXACT) */
#pragma unroll
            for (int c = 0; c < J2K_MAX_COMP; c++) {
                a[c] = (uint32_t)((int32_t)((int64_t)(int32_t)a[c] + 128) >> 8);
                b[c] = (uint32_t)((int32_t)((int64_t)(int32_t)b[c] + 128) >> 8);
            }
        }
        if (mct_on) {
            mct<TYPE>(a[0], a[1], a[2]);
            mct<TYPE>(b[0], b[1], b[2]);
        }
        int pa[J2K_MAX_COMP], pb[J2K_MAX_COMP];
#pragma unroll
        for (int c = 0; c < J2K_MAX_COMP; c++) {
            pa[c] = to_pixel<TYPE>(a[c], cbps[c], shift[c]);
            pb[c] = to_pixel<TYPE>(b[c], cbps[c], shift[c]);
        }
        uint8_t *row = plane + (size_t)(t0.out_y + Y - o.v0) * rb;
        if (wide && ncomp == 3 && pstride == 3 && okA && okB && soff[1] == 1 && soff[2] == 2 &&
            ((reinterpret_cast<uintptr_t>(row) + 2 * px) & 3) == 0) {
            /* rgb48: two pixels = three aligned 32-bit words */
            uint32_t *w32 = reinterpret_cast<uint32_t *>(row + 2 * px);
            w32[0] = (uint32_t)pa[0] | ((uint32_t)pa[1] << 16);
            w32[1] = (uint32_t)pa[2] | ((uint32_t)pb[0] << 16);
            w32[2] = (uint32_t)pb[1] | ((uint32_t)pb[2] << 16);
        } else if (wide && ncomp == 1 && pstride == 1 && okA && okB &&
                   ((reinterpret_cast<uintptr_t>(row) + 2 * px) & 3) == 0) {
            *reinterpret_cast<uint32_t *>(row + 2 * px) = (uint32_t)pa[0] | ((uint32_t)pb[0] << 16);
        } else {
#pragma unroll
            for (int c = 0; c < J2K_MAX_COMP; c++) {
                if (c < ncomp) {
                    /* planar groups may send every component to its own plane (comp_plane) */
                    const int pl = g.comp_plane[c];
                    const uint64_t pof = pl == 0 ? po0 : pl == 1 ? po1 : pl == 2 ? po2 : po3;
                    const int prb = pl == 0 ? rb0 : pl == 1 ? rb1 : pl == 2 ? rb2 : rb3;
                    uint8_t *rowc = picture + pof + (size_t)(t0.out_y + Y - o.v0) * prb;
                    if (wide) {
                        uint16_t *r16 = reinterpret_cast<uint16_t *>(rowc);
                        if (okA) r16[px + soff[c]] = (uint16_t)pa[c];
                        if (okB) r16[px + pstride + soff[c]] = (uint16_t)pb[c];
                    } else {
                        if (okA) rowc[px + soff[c]] = (uint8_t)pa[c];
                        if (okB) rowc[px + pstride + soff[c]] = (uint8_t)pb[c];
                    }
                }
            }
        }
    }
}

template <int TYPE, int TY> constexpr size_t final_smem(int ncomp)
{
    return (size_t)ncomp * (TY + 2 * Lift<TYPE>::HALO) * Geo<TYPE>::TX * 4;
}

constexpr int LTY = 64;    /* tile rows, intermediate levels */
constexpr int FTY = 32;    /* tile rows, fused last level */

template <int TYPE>
int run_levels(cudaStream_t stream, const IdwtFrame &f, int max_levels, bool any_type, int max_w[], int max_h[],
               int to_plane)
{
    int launches = 0;
    if (!any_type)
        return 0;
    const bool fast = !to_plane && idwt_stream_levels_ok(f.h_tcs, f.ntcs, TYPE);
    for (int j = 0; j < max_levels; j++) {
        if (max_w[j] <= 0)
            continue;
        if (fast) {
            const int r = launch_idwt_stream_level(stream, TYPE, f, j, max_w[j], max_h[j], (j & 1) ? f.coefB : f.coefC,
                                                   (j & 1) ? f.coefC : f.coefB);
            if (r < 0)
                return r;
            launches += r;
            continue;
        }
        constexpr int LTX = Geo<TYPE>::TX;
        /* tiles start at even absolute coordinates: one extra sample may precede the resolution */
        dim3 grid((max_w[j] + 1 + LTX - 1) / LTX, (max_h[j] + 1 + LTY - 1) / LTY, f.ntcs);
        const void *in = (j & 1) ? f.coefB : f.coefC;        /* level j reads what level j-1 wrote */
        void *out = (j & 1) ? f.coefC : f.coefB;
        idwt_level_kernel<TYPE, LTY><<<grid, THREADS, 0, stream>>>(
            f.d_tcs, j, (const uint32_t *)f.coefA, (const uint32_t *)in, (uint32_t *)out, to_plane);
        launches++;
    }
    return launches;
}

}  // namespace

int launch_idwt_frame(cudaStream_t stream, const IdwtFrame &f)
{
    int launches = 0;
    static const int skip = getenv("B200J2K_SKIP") ? atoi(getenv("B200J2K_SKIP")) : 0;   /* tuning: marginal cost of the stage */
    if (skip & 8)
        return 0;
    idwt_stream_cache_begin(f.stream_cache, f.geom_serial, f);
    /* pass 1: intermediate levels of every transform type present (host copies of the tables size the grids) */
    for (int type = 0; type < 3; type++) {
        int mw[J2K_MAX_LEVELS] = { 0 }, mh[J2K_MAX_LEVELS] = { 0 };
        int maxlev = 0;
        bool any = false;
        for (int i = 0; i < f.ntcs; i++) {
            const TileCompDesc &tc = f.h_tcs[i];
            if (tc.transform != type)
                continue;
            any = true;
            /* levels 0 .. nlevels-2 go through the level kernel */
            for (int j = 0; j + 1 < tc.nlevels; j++) {
                const LevelDesc &o = tc.lev[j + 1];
                if (o.u1 - o.u0 > mw[j]) mw[j] = o.u1 - o.u0;
                if (o.v1 - o.v0 > mh[j]) mh[j] = o.v1 - o.v0;
                if (j + 1 > maxlev) maxlev = j + 1;
            }
        }
        int r = 0;
        if (type == J2K_DWT53) r = run_levels<J2K_DWT53>(stream, f, maxlev, any, mw, mh, 0);
        if (type == J2K_DWT97) r = run_levels<J2K_DWT97>(stream, f, maxlev, any, mw, mh, 0);
        if (type == J2K_DWT97_INT) r = run_levels<J2K_DWT97_INT>(stream, f, maxlev, any, mw, mh, 0);
        if (r < 0)
            return -1;
        launches += r;
    }
    if (f.before_final)
        cudaEventRecord(f.before_final, stream);
    /* pass 2: last level fused with MCT + level shift + clip + pack */
    for (int type = 0; type < 3; type++) {
        int gw = 0, gh = 0, maxc = 0;
        bool anyg = false;
        for (int i = 0; i < f.ngroups; i++) {
            const PackGroup &g = f.h_groups[i];
            const TileCompDesc &tc = f.h_tcs[g.tc[0]];
            if (tc.transform != type)
                continue;
            anyg = true;
            const LevelDesc &o = tc.lev[tc.nlevels];
            if (o.u1 - o.u0 > gw) gw = o.u1 - o.u0;
            if (o.v1 - o.v0 > gh) gh = o.v1 - o.v0;
            if (g.ncomp > maxc) maxc = g.ncomp;
        }
        int snc = 0;
        if (anyg && gw > 0 && gh > 0 && idwt_stream_final_ok(f, type, &snc)) {
            const int r = launch_idwt_stream_final(stream, type, snc, f, gw, gh);
            if (r < 0)
                return -1;
            launches += r;
        } else if (anyg && gw > 0 && gh > 0) {
#define LAUNCH_FINAL(T)                                                                                     \
    do {                                                                                                    \
        constexpr int FTX = Geo<T>::TX;                                                                     \
        dim3 grid((gw + 1 + FTX - 1) / FTX, (gh + 1 + FTY - 1) / FTY, f.ngroups);                           \
        size_t sm = final_smem<T, FTY>(maxc);                                                               \
        cudaFuncSetAttribute(idwt_final_pack_kernel<T, FTY>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                             (int)sm);                                                                      \
        idwt_final_pack_kernel<T, FTY><<<grid, THREADS, sm, stream>>>(                                      \
            f.d_tcs, f.d_groups, (const uint32_t *)f.coefA, (const uint32_t *)f.coefB,                      \
            (const uint32_t *)f.coefC, f.picture, f.plane_off[0], f.plane_off[1], f.plane_off[2],           \
            f.plane_off[3], f.plane_row_bytes[0], f.plane_row_bytes[1], f.plane_row_bytes[2],               \
            f.plane_row_bytes[3]);                                                                          \
    } while (0)
            if (type == J2K_DWT53) LAUNCH_FINAL(J2K_DWT53);
            if (type == J2K_DWT97) LAUNCH_FINAL(J2K_DWT97);
            if (type == J2K_DWT97_INT) LAUNCH_FINAL(J2K_DWT97_INT);
#undef LAUNCH_FINAL
            launches++;
        }
    }
    return cudaGetLastError() == cudaSuccess ? launches : -1;
}

/* stage-level entry: one plane, result written back into the plane (Mallat -> image), used by tests */
int launch_idwt_plane(cudaStream_t stream, int type, const int32_t *border, int levels, void *plane,
                      void *scratchB, void *scratchC, TileCompDesc *d_tc)
{
    if (levels <= 0)
        return 0;
    TileCompDesc tc;
    memset(&tc, 0, sizeof(tc));
    tc.coef_off = 0;
    tc.width = border[1] - border[0];
    tc.height = border[3] - border[2];
    tc.x0 = border[0];
    tc.y0 = border[2];
    tc.nlevels = levels;
    tc.transform = type;
    int b[4] = { border[0], border[1], border[2], border[3] };
    for (int k = levels; k >= 0; k--) {
        tc.lev[k] = LevelDesc{ b[0], b[1], b[2], b[3] };
        for (int i = 0; i < 4; i++)
            b[i] = (b[i] + 1) >> 1;
    }
    if (cudaMemcpyAsync(d_tc, &tc, sizeof(tc), cudaMemcpyHostToDevice, stream) != cudaSuccess)
        return -1;
    IdwtFrame f;
    memset(&f, 0, sizeof(f));
    f.d_tcs = d_tc;
    f.ntcs = 1;
    f.coefA = plane;
    f.coefB = scratchB;
    f.coefC = scratchC;
    int launches = 0;
    for (int j = 0; j < levels; j++) {
        const LevelDesc &o = tc.lev[j + 1];
        const int lw = o.u1 - o.u0, lhh = o.v1 - o.v0;
        if (lw <= 0 || lhh <= 0)
            continue;
        const int ltx = type == J2K_DWT53 ? Geo<J2K_DWT53>::TX : Geo<J2K_DWT97>::TX;
        dim3 grid((lw + 1 + ltx - 1) / ltx, (lhh + 1 + LTY - 1) / LTY, 1);
        const void *in = (j & 1) ? f.coefB : f.coefC;
        /* the last level cannot go to the plane in place (other CTAs still read its quadrants): use the
         * ping-pong buffer with the plane's stride and copy back */
        void *out = (j & 1) ? f.coefC : f.coefB;
#define LAUNCH_LEVEL(T)                                                                                     \
    idwt_level_kernel<T, LTY><<<grid, THREADS, 0, stream>>>(d_tc, j, (const uint32_t *)plane,               \
                                                                 (const uint32_t *)in, (uint32_t *)out, 1)
        if (type == J2K_DWT53) LAUNCH_LEVEL(J2K_DWT53);
        else if (type == J2K_DWT97) LAUNCH_LEVEL(J2K_DWT97);
        else LAUNCH_LEVEL(J2K_DWT97_INT);
#undef LAUNCH_LEVEL
        launches++;
    }
    const void *last = ((levels - 1) & 1) ? f.coefC : f.coefB;
    const size_t n = (size_t)tc.width * tc.height;
    if (cudaMemcpyAsync(plane, last, n * 4, cudaMemcpyDeviceToDevice, stream) != cudaSuccess)
        return -1;
    if (type == J2K_DWT97_INT) {
        rescale97i_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>((uint32_t *)plane, n);
        launches++;
    }
    return cudaGetLastError() == cudaSuccess ? launches : -1;
}

}  // namespace b200j2k
