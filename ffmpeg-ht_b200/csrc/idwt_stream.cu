/*
 * idwt_stream.cu -- register-streaming inverse DWT for sm_100a: the fast path of ff_dwt_decode
 * (libavcodec/jpeg2000dwt.c:309-537, :601-620), fused at the last level with mct_decode
 * (libavcodec/jpeg2000dec.c:2183-2209, jpeg2000dsp.c:43-91) and write_frame_16 (jpeg2000dec.c:2301-2366).
 * SURVEY.md section 8(a) rows W1-W4, M1-M3, P1.
 *
 * ncu on the shared-memory tiled kernel (idwt.cu; profiles/r1_ncu_full_v1_summary_warp_per_block_baseline.md) showed it bound by
 * instruction issue (60 % of issue slots, ~52 instructions per sample), not by HBM.  This kernel removes
 * the shared-memory round trips and most of the per-sample addressing:
 *   - a warp owns a strip of 32 x 2NP columns and walks down SEG output rows;
 *   - the rows of the four Mallat quadrants reach the warp through TMA: one cp.async.bulk.tensor box copy per
 *     quadrant and row pair (all components in one box), descriptor and coordinates in uniform registers,
 *     completion counted by an mbarrier per stage; each lane reads its NP coefficients back with one LDS;
 *   - horizontal synthesis runs in registers, neighbours across lanes come from one shuffle per lifting
 *     step; the first and last lane of the warp are apron;
 *   - vertical synthesis is a sliding window over the rows: 2 (5/3) or 4 (9/7) state values per column
 *     stay in registers, so every coefficient is read once and no intermediate ever leaves the SM;
 *   - the last level applies RCT / ICT, level shift, clip and packs rgb48 / gray16.
 * Rows before columns at every level (quirk Q7), uint32 wrap-around 5/3 arithmetic and un-contracted fp32
 * 9/7 arithmetic exactly as idwt.cu.  Geometry the fast path does not cover (odd origins, unaligned
 * strides, one-sample lines, 8-bit or planar output) stays on idwt.cu; launch_idwt_frame picks per launch.
 */
#include <cuda.h>
#include <cudaTypedefs.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "idwt_common.cuh"

namespace b200j2k {

namespace {

constexpr int SWARPS = 4;                 /* warps (independent strips) per CTA */
constexpr int MAXDOM = 4;                 /* DWT domains (tile-components / pack groups) per launch */

/* One DWT domain of a streaming launch -- a tile-component (intermediate level) or a pack group (last level) --
 * as the kernel sees it: two TMA tensor maps and the geometry, all in the kernel's PARAMETER space
 * (__grid_constant__), so that descriptor addresses, box coordinates and the row walk live in uniform
 * registers.  (Round 1 staged rows with 1-D cp.async.bulk from vector-register addresses: ptxas wrapped each
 * of the twelve copies of a row pair in an ELECT / R2UR loop, 37 % of the loop's instructions.)
 *   mapL  rank 3 (column, row, component) over the LL source: the compact buffer the previous level wrote,
 *         or the top-left quadrant of the plane itself at the first level
 *   mapA  rank 3 over the Mallat-layout coefficient planes; HL, LH and HH are boxes of the same map
 * A box is one row of 32 x NP columns of every component: four copies per row pair instead of twelve. */
struct alignas(64) StreamDom {
    CUtensorMap mapL, mapA;
    const uint32_t *ll;       /* LL source (border strips read it through pointers) */
    int32_t u0, u1, v0, v1;   /* bounds of the resolution this level reconstructs */
    int32_t W, LLstride;      /* row strides in elements: plane, LL source */
    int32_t tma;              /* both maps were encoded (else every strip takes the pointer path) */
    uint32_t off[3];          /* element offsets of the component planes (same in the ping-pong buffers) */
    /* last level only */
    int32_t out_x, out_y, rb, mct, shift;
    int32_t cbps[3];
    uint64_t plane_off;
};
struct StreamLaunch {
    StreamDom dom[MAXDOM];
    const uint32_t *coefA;
    void *out;                /* intermediate level: compact ping-pong buffer; last level: the picture */
    int32_t strips, SEG;
};

#ifndef STREAM_MINB
#define STREAM_MINB 5
#endif
#ifndef STREAM_FILL
#define STREAM_FILL 100                   /* percent of the one-wave capacity the grid may use */
#endif
#ifndef STREAM_STAGES
#define STREAM_STAGES 2
#endif
#ifndef STREAM_STAGES_LEVEL
#define STREAM_STAGES_LEVEL 6
#endif
/* staging slots per warp (STAGES - 1 row pairs in flight).  The last level runs at full occupancy and is fastest
 * with two; the intermediate levels are small grids whose warps walk few rows, each step a full memory latency:
 * they keep a whole segment's rows in flight */
constexpr int STAGES_FINAL = STREAM_STAGES, STAGES_LEVEL = STREAM_STAGES_LEVEL;

template <int NP> __device__ __forceinline__ void cp_async_vec(uint32_t dst, const uint32_t *src)
{
    if (NP == 4)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
    else
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_4(uint32_t dst, const uint32_t *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
template <int NP> __device__ __forceinline__ void lds_vec(uint32_t addr, uint32_t (&v)[NP]);
template <> __device__ __forceinline__ void lds_vec<4>(uint32_t addr, uint32_t (&v)[4])
{
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(addr) : "memory");
}
template <> __device__ __forceinline__ void lds_vec<2>(uint32_t addr, uint32_t (&v)[2])
{
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v[0]), "=r"(v[1]) : "r"(addr) : "memory");
}
/* TMA staging (interior strips).  ncu + tools/membench.cu: per-lane loads of this row pattern -- LDG or
 * cp.async alike -- are bounded by the L1 miss path (113 MB arrive in ~35 us), while the same rows fetched by
 * the TMA engine arrive in ~22 us whatever the carve-out.  One elected lane fetches, per quadrant, the warp's
 * 32 x NP columns of every component with ONE tensor box copy into the warp's private slots; an mbarrier per
 * stage counts the bytes in.  Stage layout: [quadrant LL, HL, LH, HH][component][32 x NP columns]. */
/* A box must start at a multiple of 16 bytes in its innermost dimension (a start 8 bytes off raises "illegal
 * instruction" -- tools/tma_probe2.cu).  With NP = 2 the strip's first column sits at 2 mod 4 elements, so the
 * box starts PADE = 2 columns early and is 4 columns wider. */
template <int NP, int NC> struct Box {
    static constexpr int PADE = (4 - NP % 4) % 4;                          /* columns in front of lane 0's */
    static constexpr uint32_t BOXW = 32 * NP + (PADE ? 4 : 0);             /* columns of a box */
    static constexpr uint32_t ROWB = BOXW * 4;                             /* bytes of one row of the strip in one band */
    static constexpr uint32_t QSTRIDE = (NC * ROWB + 127) / 128 * 128;     /* one quadrant: NC rows, 128-byte aligned */
    static constexpr uint32_t SSTRIDE = 4 * QSTRIDE;                       /* one stage */
    static constexpr uint32_t TXBYTES = 4 * NC * ROWB;                     /* what the four copies of a stage deliver */
};
__device__ __forceinline__ void mbar_init(uint32_t bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
/* Programmatic dependent launch: the kernels of a frame follow each other on one stream, each a few microseconds
 * long.  A kernel launched with the programmatic-serialisation attribute may start while its predecessor drains;
 * everything up to here (geometry, barrier set-up) overlaps with that tail, the first read of the predecessor's
 * output waits for its completion.  Dependents of this kernel may be scheduled from now on. */
__device__ __forceinline__ void pdl_prologue_done()
{
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}
__device__ __forceinline__ bool elect_one()
{
    uint32_t p;
    asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}" : "=r"(p));
    return p != 0;
}
/* the coefficients are read exactly once: let them leave L2 first, the picture written here is read next
 * (the constant is what createpolicy.fractional.L2::evict_first.b64 with fraction 1.0 returns) */
constexpr uint64_t L2_EVICT_FIRST = 0x12F0000000000000ull;
#ifndef STREAM_HINT
#define STREAM_HINT 1
#endif
#ifndef STREAM_ELECT
#define STREAM_ELECT 1
#endif
__device__ __forceinline__ void tma_box(uint32_t dst, const CUtensorMap *map, int x, int y, uint32_t bar)
{
#if STREAM_HINT
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
                 " [%0], [%1, {%2, %3, %4}], [%5], %6;"
                 ::"r"(dst), "l"(map), "r"(x), "r"(y), "r"(0), "r"(bar), "l"(L2_EVICT_FIRST) : "memory");
#else
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
                 " [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(dst), "l"(map), "r"(x), "r"(y), "r"(0), "r"(bar) : "memory");
#endif
}
/* bytes of staging shared memory per warp and per CTA (one 128-byte line of mbarriers in front) */
template <int NP, int NC, int STAGES> __host__ __device__ constexpr int stage_bytes_per_warp() { return STAGES * (int)Box<NP, NC>::SSTRIDE; }
template <int NP, int NC, int STAGES> __host__ __device__ constexpr int bar_bytes_per_cta() { return (SWARPS * STAGES * 8 + 127) / 128 * 128; }
template <int NP, int NC, int STAGES> __host__ __device__ constexpr int stage_bytes_per_cta()
{
    return bar_bytes_per_cta<NP, NC, STAGES>() + SWARPS * stage_bytes_per_warp<NP, NC, STAGES>();
}

/* horizontal synthesis of NP (low, high) pairs per lane; x[2k] = even sample, x[2k+1] = odd sample */
template <int TYPE, int NP>
__device__ __forceinline__ void hsyn(const uint32_t (&Lc)[NP], const uint32_t (&Hc)[NP], uint32_t (&x)[2 * NP])
{
    using L = Lift<TYPE>;
    if (TYPE == J2K_DWT53) {
        const uint32_t hl = __shfl_up_sync(FULLMASK, Hc[NP - 1], 1);
        uint32_t e[NP + 1];
#pragma unroll
        for (int k = 0; k < NP; k++)
            e[k] = L::apply(0, Lc[k], k ? Hc[k ? k - 1 : 0] : hl, Hc[k]);
        e[NP] = __shfl_down_sync(FULLMASK, e[0], 1);
#pragma unroll
        for (int k = 0; k < NP; k++) {
            x[2 * k] = e[k];
            x[2 * k + 1] = L::apply(1, Hc[k], e[k], e[k + 1]);
        }
    } else {
        const uint32_t hl = __shfl_up_sync(FULLMASK, Hc[NP - 1], 1);
        uint32_t e1[NP + 1], o1[NP], e2[NP + 1];
#pragma unroll
        for (int k = 0; k < NP; k++)
            e1[k] = L::apply(0, Lc[k], k ? Hc[k ? k - 1 : 0] : hl, Hc[k]);
        e1[NP] = __shfl_down_sync(FULLMASK, e1[0], 1);
#pragma unroll
        for (int k = 0; k < NP; k++)
            o1[k] = L::apply(1, Hc[k], e1[k], e1[k + 1]);
        const uint32_t ol = __shfl_up_sync(FULLMASK, o1[NP - 1], 1);
#pragma unroll
        for (int k = 0; k < NP; k++)
            e2[k] = L::apply(2, e1[k], k ? o1[k ? k - 1 : 0] : ol, o1[k]);
        e2[NP] = __shfl_down_sync(FULLMASK, e2[0], 1);
#pragma unroll
        for (int k = 0; k < NP; k++) {
            x[2 * k] = e2[k];
            x[2 * k + 1] = L::apply(3, o1[k], e2[k], e2[k + 1]);
        }
    }
}

/* geometry of one (tile-component, level) seen by a warp */
struct StripGeo {
    int u0, u1, v0, v1;
    int nLh, nLv, W, LLstride;
    int eoff, ooff;     /* ceil(u0/2), floor(u0/2) */
    int eoffv, ooffv;   /* ceil(v0/2), floor(v0/2) */
};

__device__ __forceinline__ StripGeo make_geo(const StreamDom &d)
{
    StripGeo g;
    g.u0 = d.u0; g.u1 = d.u1; g.v0 = d.v0; g.v1 = d.v1;
    g.nLh = ((d.u1 + 1) >> 1) - ((d.u0 + 1) >> 1);
    g.nLv = ((d.v1 + 1) >> 1) - ((d.v0 + 1) >> 1);
    g.W = d.W;
    g.LLstride = d.LLstride;
    g.eoff = (d.u0 + 1) >> 1;
    g.ooff = d.u0 >> 1;
    g.eoffv = (d.v0 + 1) >> 1;
    g.ooffv = d.v0 >> 1;
    return g;
}

/* per-lane column bookkeeping: fixed for all rows of the strip */
template <int NP> struct LaneCols {
    bool inside;        /* all 2NP samples inside [u0,u1): vector loads */
    int vcolL, vcolH;   /* first column inside the low / high band */
    int ie[NP], io[NP]; /* reflected columns (border lanes) */
};

template <int NP>
__device__ __forceinline__ LaneCols<NP> make_cols(const StripGeo &g, int p0)
{
    LaneCols<NP> lc;
    lc.inside = 2 * p0 >= g.u0 && 2 * (p0 + NP - 1) + 1 < g.u1;
    lc.vcolL = p0 - g.eoff;
    lc.vcolH = p0 - g.ooff;
#pragma unroll
    for (int k = 0; k < NP; k++) {
        int Xe = 2 * (p0 + k), Xo = Xe + 1;
        if (Xe < g.u0 || Xe >= g.u1) Xe = reflect(Xe, g.u0, g.u1);
        if (Xo < g.u0 || Xo >= g.u1) Xo = reflect(Xo, g.u0, g.u1);
        /* reflection preserves parity: an even position stays in the low band, an odd one in the high band */
        lc.ie[k] = (Xe >> 1) - g.eoff;
        lc.io[k] = (Xo >> 1) - g.ooff;
    }
    return lc;
}

/* vertical lifting state of one lane: per component and column the values the next row pair needs */
template <int TYPE, int NC, int C2> struct VState {
    uint32_t vh[NC][C2], e1[NC][C2];
    uint32_t o1[TYPE == J2K_DWT53 ? 1 : NC][TYPE == J2K_DWT53 ? 1 : C2];
    uint32_t e2[TYPE == J2K_DWT53 ? 1 : NC][TYPE == J2K_DWT53 ? 1 : C2];
};

/* band_index of the (reflected) absolute row Y.  Only the first and last row pairs of a plane come here, at most
 * HALO + 1 rows outside [v0, v1) with v1 - v0 >= 2 (level_streamable): folding back and forth gives what
 * reflect() computes with an integer division, in a fraction of the code */
__device__ __forceinline__ int row_ky(const StripGeo &g, int Y)
{
    while (Y < g.v0 || Y >= g.v1)
        Y = Y < g.v0 ? 2 * g.v0 - Y : 2 * (g.v1 - 1) - Y;
    return band_index(Y, g.v0);
}

/* The streaming core: walks row pairs (even row = vertical low-pass, odd row = vertical high-pass), keeps the
 * vertical lifting state in registers and calls emit(allin, Y, rows[NC][2NP]) for every finished output row.
 * The loop is unrolled by two with the state ping-ponging between two register sets, so the sliding window
 * never costs a register move, and the staging slot of a step is a compile-time constant (STAGES == 2).
 * ALLIN: every lane of the warp lies inside the line (no column needs reflection) -- the common case, fed by
 * TMA and free of per-lane branches; the other instantiation serves the first and last strip of a line with
 * per-lane cp.async / synchronous loads through pointers.
 * cs: the strip's first column inside the bands -- warp-uniform, like Ya / Yb and everything derived from them,
 * so the box coordinates of the TMA copies are computed on the uniform datapath. */
template <int TYPE, int NP, int NC, int STAGES, bool ALLIN, class Emit>
__device__ __forceinline__ void stream_rows(const StreamDom &D, const StripGeo &g, const uint32_t *__restrict__ baseA,
                                            const LaneCols<NP> &lc, int cs, int Ya, int Yb, uint32_t warp_base,
                                            uint32_t bar_base, Emit &&emit)
{
    using L = Lift<TYPE>;
    constexpr int HP = L::HALO / 2;
    constexpr int C2 = 2 * NP;
    using B = Box<NP, NC>;
    constexpr uint32_t ROWB = B::ROWB, QSTRIDE = B::QSTRIDE, SSTRIDE = B::SSTRIDE;
    constexpr int PADE = B::PADE;
    const int lane = threadIdx.x & 31;
    const uint32_t *__restrict__ baseL = D.ll;
    VState<TYPE, NC, C2> S0, S1;
#pragma unroll
    for (int c = 0; c < NC; c++)
#pragma unroll
        for (int k = 0; k < C2; k++) {
            S0.vh[c][k] = S0.e1[c][k] = 0;
            if (TYPE != J2K_DWT53)
                S0.o1[c][k] = S0.e2[c][k] = 0;
        }
    const int n0 = (Ya >> 1) - HP, n1 = (Yb >> 1) + HP;     /* iterations [n0, n1) */
    const int Ylo = max(Ya, g.v0), Yhi = min(Yb, g.v1);     /* rows this warp emits */
    /* border strips: element offsets (from the kernel-uniform bases) of the lane's first column in the LL source
     * and in the HL band of the plane */
    uint32_t eL[NC], eA[NC];
#pragma unroll
    for (int c = 0; c < NC; c++) {
        eL[c] = D.off[c] + (uint32_t)lc.vcolL;
        eA[c] = D.off[c] + (uint32_t)(g.nLh + lc.vcolH);
    }
    const uint32_t dLH = (uint32_t)(lc.vcolL - lc.vcolH - g.nLh);   /* LH column relative to the HH column */
    const uint32_t lane_off = (uint32_t)(PADE + lane * NP) * 4;           /* the lane's columns inside a staged row */
    const uint32_t lane_base = warp_base + lane_off;
    /* Asynchronous staging, one row pair ahead, so the bytes in flight do not cost registers.  Interior
     * strips: four TMA box copies by one elected lane, completion through the stage's mbarrier.  Border strips:
     * every inside lane fills its own vector slots with cp.async; lanes whose columns need reflection
     * load synchronously when the row pair is consumed.  A lane only ever reads back its own bytes. */
    auto issue = [&](int n, uint32_t st, uint32_t bar) {
        if (n < n1) {
            /* band_index() of rows 2n / 2n+1 is n - ceil(v0/2) / n - floor(v0/2) unless a row needs reflection */
            int kyE = n - g.eoffv, kyO = n - g.ooffv;
            if (2 * n < g.v0 || 2 * n + 1 >= g.v1) {
                kyE = row_ky(g, 2 * n);
                kyO = row_ky(g, 2 * n + 1);
            }
            if (ALLIN) {
                __syncwarp();                    /* every lane is done with the slots this refills */
                if (STREAM_ELECT ? elect_one() : lane == 0) {
                    mbar_expect_tx(bar, B::TXBYTES);
                    tma_box(st, &D.mapL, cs - PADE, kyE, bar);
                    tma_box(st + QSTRIDE, &D.mapA, g.nLh + cs - PADE, kyE, bar);
                    tma_box(st + 2 * QSTRIDE, &D.mapA, cs - PADE, g.nLv + kyO, bar);
                    tma_box(st + 3 * QSTRIDE, &D.mapA, g.nLh + cs - PADE, g.nLv + kyO, bar);
                }
            } else {
                const uint32_t oLL = (uint32_t)(kyE * g.LLstride), oHL = (uint32_t)(kyE * g.W);
                const uint32_t oHH = (uint32_t)((g.nLv + kyO) * g.W), oLH = oHH + dLH;
                const uint32_t sl = st + lane_off;
                if (lc.inside) {
#pragma unroll
                    for (int c = 0; c < NC; c++) {
                        cp_async_vec<NP>(sl + c * ROWB, baseL + (eL[c] + oLL));
                        cp_async_vec<NP>(sl + QSTRIDE + c * ROWB, baseA + (eA[c] + oHL));
                        cp_async_vec<NP>(sl + 2 * QSTRIDE + c * ROWB, baseA + (eA[c] + oLH));
                        cp_async_vec<NP>(sl + 3 * QSTRIDE + c * ROWB, baseA + (eA[c] + oHH));
                    }
                } else {
                    /* columns that need reflection: element by element, still asynchronous */
                    const uint32_t rHL = oHL + (uint32_t)g.nLh, rLH = oHH, rHH = oHH + (uint32_t)g.nLh;
#pragma unroll
                    for (int c = 0; c < NC; c++)
#pragma unroll
                        for (int k = 0; k < NP; k++) {
                            cp_async_4(sl + c * ROWB + 4 * k, baseL + (D.off[c] + oLL + (uint32_t)lc.ie[k]));
                            cp_async_4(sl + QSTRIDE + c * ROWB + 4 * k, baseA + (D.off[c] + rHL + (uint32_t)lc.io[k]));
                            cp_async_4(sl + 2 * QSTRIDE + c * ROWB + 4 * k, baseA + (D.off[c] + rLH + (uint32_t)lc.ie[k]));
                            cp_async_4(sl + 3 * QSTRIDE + c * ROWB + 4 * k, baseA + (D.off[c] + rHH + (uint32_t)lc.io[k]));
                        }
                }
            }
        }
        if (!ALLIN)
            asm volatile("cp.async.commit_group;" ::: "memory");
    };
    int slot = 0;                                /* staging slot of the row pair consumed next, and the parity */
    uint32_t phase = 0;                          /* of its mbarrier: flips every time the ring wraps */
    auto step = [&](int n, const VState<TYPE, NC, C2> &in, VState<TYPE, NC, C2> &out) {
        const uint32_t st = lane_base + (uint32_t)slot * SSTRIDE;
        const int prev = slot == 0 ? STAGES - 1 : slot - 1;             /* consumed by the previous step */
        issue(n + STAGES - 1, warp_base + (uint32_t)prev * SSTRIDE, bar_base + (uint32_t)prev * 8);
        if (ALLIN)
            mbar_wait(bar_base + (uint32_t)slot * 8, phase);
        else
            asm volatile("cp.async.wait_group %0;" ::"n"(STAGES - 1) : "memory");
        if (++slot == STAGES) {
            slot = 0;
            phase ^= 1;
        }
        uint32_t cLe[NC][NP], cHe[NC][NP], cLo[NC][NP], cHo[NC][NP];
#pragma unroll
        for (int c = 0; c < NC; c++) {
            lds_vec<NP>(st + c * ROWB, cLe[c]);
            lds_vec<NP>(st + QSTRIDE + c * ROWB, cHe[c]);
            lds_vec<NP>(st + 2 * QSTRIDE + c * ROWB, cLo[c]);
            lds_vec<NP>(st + 3 * QSTRIDE + c * ROWB, cHo[c]);
        }
        uint32_t rowA[NC][C2];                   /* finished odd row 2n-(2HP-1); the even row 2n-(2HP-2) is out.e1 / out.e2 */
#pragma unroll
        for (int c = 0; c < NC; c++) {
            uint32_t lo[C2];
            hsyn<TYPE, NP>(cLe[c], cHe[c], lo);
            hsyn<TYPE, NP>(cLo[c], cHo[c], out.vh[c]);
#pragma unroll
            for (int k = 0; k < C2; k++) {
                if (TYPE == J2K_DWT53) {
                    out.e1[c][k] = L::apply(0, lo[k], in.vh[c][k], out.vh[c][k]);
                    rowA[c][k] = L::apply(1, in.vh[c][k], in.e1[c][k], out.e1[c][k]);
                } else {
                    out.e1[c][k] = L::apply(0, lo[k], in.vh[c][k], out.vh[c][k]);
                    out.o1[c][k] = L::apply(1, in.vh[c][k], in.e1[c][k], out.e1[c][k]);
                    out.e2[c][k] = L::apply(2, in.e1[c][k], in.o1[c][k], out.o1[c][k]);
                    rowA[c][k] = L::apply(3, in.o1[c][k], in.e2[c][k], out.e2[c][k]);
                }
            }
        }
        const int YA = 2 * n - (2 * HP - 1), YB = YA + 1;
        if (YA >= Ylo && YA < Yhi)
            emit(ALLIN, YA, rowA);
        if (YB >= Ylo && YB < Yhi) {
            if constexpr (TYPE == J2K_DWT53)
                emit(ALLIN, YB, out.e1);
            else
                emit(ALLIN, YB, out.e2);
        }
    };
#pragma unroll
    for (int s = 0; s < STAGES - 1; s++)
        issue(n0 + s, warp_base + (uint32_t)s * SSTRIDE, bar_base + (uint32_t)s * 8);
    int n = n0;
    for (; n + 1 < n1; n += 2) {
        step(n, S0, S1);
        step(n + 1, S1, S0);
    }
    if (n < n1)
        step(n, S0, S1);
}

/* what a warp works on: its index inside the CTA is made provably warp-uniform with a broadcast (the compiler
 * cannot know that threadIdx.x >> 5 is), so the strip / segment arithmetic stays in uniform registers */
struct WarpUnit {
    int warp, strip, seg;
};
__device__ __forceinline__ WarpUnit warp_unit(int strips)
{
    WarpUnit u;
    u.warp = __shfl_sync(FULLMASK, (int)(threadIdx.x >> 5), 0);
    const int unit = (int)blockIdx.x * SWARPS + u.warp;
    u.seg = unit / strips;
    u.strip = unit - u.seg * strips;
    return u;
}

/* ---- intermediate level: reconstructed resolution -> compact ping-pong buffer -------------------------- */
template <int TYPE, int NP>
__global__ void __launch_bounds__(SWARPS * 32)
idwt_stream_level_kernel(const __grid_constant__ StreamLaunch P)
{
    extern __shared__ __align__(128) uint32_t stage_smem[];
    constexpr int C2 = 2 * NP, VALID = 30 * C2;
    const StreamDom &D = P.dom[blockIdx.y];
    const int lane = threadIdx.x & 31;
    const WarpUnit wu = warp_unit(P.strips);
    const StripGeo g = make_geo(D);
    const int XA = (g.u0 & ~1) + wu.strip * VALID - C2;       /* first column of lane 0 (apron lane) */
    const int Ya = (g.v0 & ~1) + wu.seg * P.SEG, Yb = min(Ya + P.SEG, (g.v1 + 1) & ~1);
    if (XA + C2 >= g.u1 || Ya >= g.v1)
        return;
    const int p0 = (XA >> 1) + lane * NP;
    const LaneCols<NP> lc = make_cols<NP>(g, p0);
    uint32_t *out = (uint32_t *)P.out + D.off[0];
    const int lh = g.u1 - g.u0;
    const int X0 = 2 * p0;
    const bool lane_ok = lane >= 1 && lane <= 30;
    uint32_t *out_base = out + (X0 - g.u0) - (long)g.v0 * lh;
    auto emit = [&](bool allin, int Y, const uint32_t (&row)[1][C2]) {
        if (!lane_ok)
            return;
        uint32_t *o = out_base + (long)Y * lh;
        if (allin || lc.inside) {
            if (NP == 4) {
                reinterpret_cast<uint4 *>(o)[0] = make_uint4(row[0][0], row[0][1], row[0][2], row[0][3]);
                reinterpret_cast<uint4 *>(o)[1] = make_uint4(row[0][4], row[0][5], row[0][6], row[0][7]);
            } else {
                reinterpret_cast<uint4 *>(o)[0] = make_uint4(row[0][0], row[0][1], row[0][2], row[0][3]);
            }
        } else {
#pragma unroll
            for (int k = 0; k < C2; k++)
                if (X0 + k >= g.u0 && X0 + k < g.u1)
                    o[k] = row[0][k];
        }
    };
    const uint32_t smem0 = (uint32_t)__cvta_generic_to_shared(stage_smem);
    constexpr int STAGES = STAGES_LEVEL;
    const uint32_t bar_base = smem0 + wu.warp * (STAGES * 8);
    const uint32_t warp_base = smem0 + bar_bytes_per_cta<NP, 1, STAGES>() + wu.warp * stage_bytes_per_warp<NP, 1, STAGES>();
    const bool allin = __all_sync(FULLMASK, lc.inside) && D.tma;
    if (allin && lane == 0) {
#pragma unroll
        for (int i = 0; i < STAGES; i++)
            mbar_init(bar_base + 8 * i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    pdl_prologue_done();
    const int cs = (XA >> 1) - g.eoff;
    if (allin)
        stream_rows<TYPE, NP, 1, STAGES, true>(D, g, P.coefA, lc, cs, Ya, Yb, warp_base, bar_base, emit);
    else
        stream_rows<TYPE, NP, 1, STAGES, false>(D, g, P.coefA, lc, cs, Ya, Yb, warp_base, bar_base, emit);
}

/* ---- last level fused with MCT + level shift + clip + pack (packed 16-bit output) ---------------------- */
template <int TYPE, int NP, int NC>
__global__ void __launch_bounds__(SWARPS * 32, NP == 4 ? 2 : (TYPE == J2K_DWT53 ? STREAM_MINB : 3))
idwt_stream_final_kernel(const __grid_constant__ StreamLaunch P)
{
    extern __shared__ __align__(128) uint32_t stage_smem[];
    constexpr int C2 = 2 * NP, VALID = 30 * C2;
    const StreamDom &D = P.dom[blockIdx.y];
    const int lane = threadIdx.x & 31;
    const WarpUnit wu = warp_unit(P.strips);
    const StripGeo g = make_geo(D);
    const int XA = (g.u0 & ~1) + wu.strip * VALID - C2;
    const int Ya = (g.v0 & ~1) + wu.seg * P.SEG, Yb = min(Ya + P.SEG, (g.v1 + 1) & ~1);
    if (XA + C2 >= g.u1 || Ya >= g.v1)
        return;
    const int p0 = (XA >> 1) + lane * NP;
    const LaneCols<NP> lc = make_cols<NP>(g, p0);
    const int shift = D.shift;
    const int rb = D.rb;
    uint8_t *plane = (uint8_t *)P.out + D.plane_off;
    const int mct_on = D.mct;
    const int X0 = 2 * p0;
    const bool lane_ok = lane >= 1 && lane <= 30;
    const long px = (long)(D.out_x + X0 - g.u0) * NC;        /* sample index of the lane's first pixel */
    int half[NC], vmax[NC];
#pragma unroll
    for (int c = 0; c < NC; c++) {
        half[c] = 1 << (D.cbps[c] - 1);
        vmax[c] = (1 << D.cbps[c]) - 1;
    }
    /* byte address of this lane's first sample in output row Y = out_base + Y * rb */
    uint8_t *out_base = plane + ((long)D.out_y - g.v0) * rb + 2 * px;
    auto emit = [&](bool allin, int Y, const uint32_t (&row)[NC][C2]) {
        if (!lane_ok)
            return;
        uint32_t pix[NC][C2];                    /* level-shifted and clipped, not yet left-aligned */
#pragma unroll
        for (int k = 0; k < C2; k++) {
            uint32_t v[NC];
#pragma unroll
            for (int c = 0; c < NC; c++) {
                v[c] = row[c][k];
                if (TYPE == J2K_DWT97_INT)       /* undo I_PRESHIFT before the component transform (jpeg2000dwt.c:534-536) */
                    v[c] = (uint32_t)((int32_t)((int64_t)(int32_t)v[c] + 128) >> 8);
            }
            if (NC == 3 && mct_on)
                mct<TYPE>(v[0], v[NC > 1 ? 1 : 0], v[NC > 2 ? 2 : 0]);
#pragma unroll
            for (int c = 0; c < NC; c++) {
                const int val = TYPE == J2K_DWT97 ? __float2int_rn(__uint_as_float(v[c])) : (int32_t)v[c];
                pix[c][k] = (uint32_t)__viaddmin_s32_relu(val, half[c], vmax[c]);    /* clip(val + half, 0, max) */
            }
        }
        uint16_t *r16 = reinterpret_cast<uint16_t *>(out_base + (long)Y * rb);
        if (allin || lc.inside) {
            /* 2NP pixels x NC samples x 16 bit: 16-byte (NP = 4) or 8-byte (NP = 2) aligned by construction */
            uint32_t w[NC * NP];
#pragma unroll
            for (int i = 0; i < NC * NP; i++) {
                const int s0 = 2 * i, s1 = 2 * i + 1;            /* sample indices inside the lane's run */
                /* all components share one shift on this path (idwt_stream_final_ok) */
                w[i] = __byte_perm(pix[s0 % NC][s0 / NC], pix[s1 % NC][s1 / NC], 0x5410) << shift;
            }
            if (NP == 4) {
                uint4 *o = reinterpret_cast<uint4 *>(r16);
#pragma unroll
                for (int i = 0; i < NC * NP / 4; i++)
                    o[i] = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
            } else {                                              /* a lane's run is 8 or 24 bytes: 8-byte aligned */
                uint2 *o = reinterpret_cast<uint2 *>(r16);
#pragma unroll
                for (int i = 0; i < NC * NP / 2; i++)
                    o[i] = make_uint2(w[2 * i], w[2 * i + 1]);
            }
        } else {
#pragma unroll
            for (int k = 0; k < C2; k++)
                if (X0 + k >= g.u0 && X0 + k < g.u1)
#pragma unroll
                    for (int c = 0; c < NC; c++)
                        r16[k * NC + c] = (uint16_t)(pix[c][k] << shift);
        }
    };
    const uint32_t smem0 = (uint32_t)__cvta_generic_to_shared(stage_smem);
    constexpr int STAGES = STAGES_FINAL;
    const uint32_t bar_base = smem0 + wu.warp * (STAGES * 8);
    const uint32_t warp_base = smem0 + bar_bytes_per_cta<NP, NC, STAGES>() + wu.warp * stage_bytes_per_warp<NP, NC, STAGES>();
    const bool allin = __all_sync(FULLMASK, lc.inside) && D.tma;
    if (allin && lane == 0) {
#pragma unroll
        for (int i = 0; i < STAGES; i++)
            mbar_init(bar_base + 8 * i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    pdl_prologue_done();
    const int cs = (XA >> 1) - g.eoff;
    if (allin)
        stream_rows<TYPE, NP, NC, STAGES, true>(D, g, P.coefA, lc, cs, Ya, Yb, warp_base, bar_base, emit);
    else
        stream_rows<TYPE, NP, NC, STAGES, false>(D, g, P.coefA, lc, cs, Ya, Yb, warp_base, bar_base, emit);
}

template <int TYPE> struct StreamCfg {
#ifndef STREAM_NP53
#define STREAM_NP53 2
#endif
    static constexpr int NP = TYPE == J2K_DWT53 ? STREAM_NP53 : 2;
};
/* Rows per warp.  These kernels run for tens of microseconds: a second, nearly empty wave of CTAs costs as much
 * as the first, so the segment height is the smallest one (least serial work per warp) whose grid still fits
 * the GPU in ONE wave -- `capacity` = SMs x resident CTAs per SM of the kernel as the occupancy calculator
 * reports it -- and never so small that the HALO warm-up rows dominate. */
int pick_seg(int strips, int rows, int capacity)
{
    const long fill = (long)capacity * SWARPS * STREAM_FILL / 100;      /* warps */
    long max_segs = fill / (strips > 0 ? strips : 1);
    if (max_segs < 1)
        max_segs = 1;
    long seg = (rows + max_segs - 1) / max_segs;
    seg = (seg + 1) & ~1L;
    if (seg < 8) seg = 8;
    return (int)seg;
}
/* resident CTAs of a kernel on the current device (function attributes are per device, so is this cache) */
struct PerDevice {
    int cap[64] = { 0 };
    template <class K> int get(K kernel, int smem)
    {
        int dev = 0, sms = 0, per_sm = 0;
        cudaGetDevice(&dev);
        if (dev >= 0 && dev < 64 && cap[dev])
            return cap[dev];
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, SWARPS * 32, smem);
        const int c = sms * (per_sm > 0 ? per_sm : 1);
        if (dev >= 0 && dev < 64)
            cap[dev] = c;
        return c;
    }
};

/* can the fast path reconstruct level j of this tile-component?  (even origin, lines longer than one
 * sample, 16-byte aligned band rows) */
bool level_streamable(const TileCompDesc &tc, int j)
{
    const LevelDesc &o = tc.lev[j + 1];
    const int lh = o.u1 - o.u0, lv = o.v1 - o.v0;
    if (lh <= 1 || lv <= 1)
        return false;
    if (o.u0 & 1)
        return false;
    const int nLh = ((o.u1 + 1) >> 1) - ((o.u0 + 1) >> 1);
    const int LLstride = j == 0 ? tc.width : nLh;
    if ((tc.width & 3) || (nLh & 3) || (LLstride & 3) || (tc.coef_off & 3))
        return false;
    return true;
}

/* ---- host: tensor maps --------------------------------------------------------------------------------- */
PFN_cuTensorMapEncodeTiled_v12000 tensor_map_encoder()
{
    static PFN_cuTensorMapEncodeTiled_v12000 fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            p = nullptr;
        }
        return (PFN_cuTensorMapEncodeTiled_v12000)p;
    }();
    return fn;
}

/* rank-3 map (column, row, component) of uint32 elements; box = one row of `boxw` columns of all components */
bool encode_rows(CUtensorMap *m, const void *base, uint64_t cols, uint64_t rows, uint64_t nc, uint64_t row_stride_elems,
                 uint64_t comp_stride_elems, uint32_t boxw)
{
    PFN_cuTensorMapEncodeTiled_v12000 enc = tensor_map_encoder();
    if (!enc)
        return false;
    const cuuint64_t dims[3] = { cols, rows, nc };
    const cuuint64_t strides[2] = { row_stride_elems * 4, (nc > 1 ? comp_stride_elems : row_stride_elems * rows) * 4 };
    const cuuint32_t box[3] = { boxw, 1, (cuuint32_t)nc };
    const cuuint32_t estr[3] = { 1, 1, 1 };
    if ((strides[0] & 15) || (strides[1] & 15) || ((uintptr_t)base & 15))
        return false;
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, const_cast<void *>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

/* geometry + maps of level j of the nc tile-components tcs[0..nc) (equal geometry, equally spaced planes) */
bool fill_dom(StreamDom &d, const TileCompDesc *const *tcs, int nc, int j, const void *coefA, const void *ll_buf, int NP)
{
    const TileCompDesc &t0 = *tcs[0];
    const LevelDesc &o = t0.lev[j + 1];
    memset(&d, 0, sizeof(d));
    d.u0 = o.u0; d.u1 = o.u1; d.v0 = o.v0; d.v1 = o.v1;
    d.W = t0.width;
    const int nLh = ((o.u1 + 1) >> 1) - ((o.u0 + 1) >> 1), nLv = ((o.v1 + 1) >> 1) - ((o.v0 + 1) >> 1);
    d.LLstride = j == 0 ? t0.width : nLh;
    for (int c = 0; c < 3; c++)
        d.off[c] = tcs[c < nc ? c : 0]->coef_off;
    uint64_t cstride = 0;
    if (nc > 1) {
        if (tcs[1]->coef_off <= t0.coef_off)
            return false;
        cstride = (uint64_t)tcs[1]->coef_off - t0.coef_off;
        for (int c = 1; c < nc; c++)
            if ((uint64_t)tcs[c]->coef_off != t0.coef_off + (uint64_t)c * cstride)
                return false;
    }
    const uint32_t *A = (const uint32_t *)coefA + t0.coef_off;
    const uint32_t *Lb = j == 0 ? A : (const uint32_t *)ll_buf + t0.coef_off;
    d.ll = j == 0 ? (const uint32_t *)coefA : (const uint32_t *)ll_buf;
    const uint32_t boxw = NP == 2 ? Box<2, 1>::BOXW : Box<4, 1>::BOXW;
    /* a line with an interior strip is at least two boxes long; shorter ones only have border strips */
    d.tma = nLh >= (int)boxw && nLv >= 1 &&
            encode_rows(&d.mapA, A, (uint64_t)t0.width, (uint64_t)t0.height, nc, (uint64_t)t0.width, cstride, boxw) &&
            encode_rows(&d.mapL, Lb, (uint64_t)nLh, (uint64_t)nLv, nc, (uint64_t)d.LLstride, cstride, boxw);
    return true;
}

}  // namespace

/* What launch_idwt_frame memoises per slot: the parameter blocks (tensor maps included) of the streaming
 * launches of one frame geometry on one set of device buffers.  Encoding a tensor map costs about a
 * microsecond of host time and a 5-level 4K frame needs 26 of them; a sequence re-uses them frame after frame. */
struct IdwtStreamCache {
    struct Item {
        StreamLaunch p;
        dim3 grid;
        int last;             /* last launch of the call that recorded it */
    };
    uint64_t serial = ~0ull;
    const void *a = nullptr, *b = nullptr, *c = nullptr, *pic = nullptr;
    std::vector<Item> items;
    size_t cursor = 0;
    bool replay = false;
};
IdwtStreamCache *idwt_stream_cache_create() { return new IdwtStreamCache(); }
void idwt_stream_cache_destroy(IdwtStreamCache *c) { delete c; }
void idwt_stream_cache_begin(IdwtStreamCache *c, uint64_t serial, const IdwtFrame &f)
{
    if (!c)
        return;
    c->replay = c->serial == serial && c->a == f.coefA && c->b == f.coefB && c->c == f.coefC && c->pic == f.picture &&
                !c->items.empty();
    if (!c->replay) {
        c->items.clear();
        c->serial = serial;
        c->a = f.coefA; c->b = f.coefB; c->c = f.coefC; c->pic = f.picture;
    }
    c->cursor = 0;
}

bool idwt_stream_levels_ok(const TileCompDesc *h_tcs, int ntcs, int type)
{
    if (!tensor_map_encoder())
        return false;
    for (int i = 0; i < ntcs; i++) {
        const TileCompDesc &tc = h_tcs[i];
        if (tc.transform != type)
            continue;
        for (int j = 0; j + 1 < tc.nlevels; j++) {
            if (!level_streamable(tc, j))
                return false;
            /* the level's output is the next level's LL source: its rows must stay 16-byte aligned */
            if ((tc.lev[j + 1].u1 - tc.lev[j + 1].u0) & 3)
                return false;
        }
    }
    return true;
}

bool idwt_stream_final_ok(const IdwtFrame &f, int type, int *ncomp)
{
    int nc = 0;
    if (!tensor_map_encoder())
        return false;
    for (int i = 0; i < f.ngroups; i++) {
        const PackGroup &g = f.h_groups[i];
        const TileCompDesc &t0 = f.h_tcs[g.tc[0]];
        if (t0.transform != type)
            continue;
        if (t0.nlevels < 1 || !g.wide || (g.ncomp != 1 && g.ncomp != 3) || g.pixel_stride != g.ncomp)
            return false;
        if (nc && nc != g.ncomp)
            return false;
        nc = g.ncomp;
        for (int c = 0; c < g.ncomp; c++) {
            const TileCompDesc &tc = f.h_tcs[g.tc[c]];
            if (g.sample_off[c] != c || tc.nlevels != t0.nlevels || !level_streamable(tc, tc.nlevels - 1))
                return false;
            if (g.shift[c] != g.shift[0] || tc.width != t0.width || tc.height != t0.height)
                return false;
            const LevelDesc &a = tc.lev[tc.nlevels], &b = t0.lev[t0.nlevels];
            if (a.u0 != b.u0 || a.u1 != b.u1 || a.v0 != b.v0 || a.v1 != b.v1)
                return false;
            /* one box covers the components: their planes must be equally spaced (they are allocated back to back) */
            if (c > 0) {
                const TileCompDesc &t1 = f.h_tcs[g.tc[1]];
                if (t1.coef_off <= t0.coef_off ||
                    (uint64_t)tc.coef_off != t0.coef_off + (uint64_t)c * ((uint64_t)t1.coef_off - t0.coef_off))
                    return false;
            }
        }
        if ((t0.out_x & 7) || (f.plane_row_bytes[g.plane] & 15) || (f.plane_off[g.plane] & 15))
            return false;
    }
    *ncomp = nc;
    return nc != 0;
}

namespace {

template <class K> void launch_item(cudaStream_t stream, K kernel, int smem, const IdwtStreamCache::Item &it, bool first = false)
{
    static const int pdl = getenv("B200J2K_IDWT_PDL") ? atoi(getenv("B200J2K_IDWT_PDL")) : 1;   /* tuning: 0 plain launches */
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = it.grid;
    cfg.blockDim = dim3(SWARPS * 32);
    cfg.dynamicSmemBytes = (size_t)smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr.val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = pdl == 1 || (pdl == 2 && !first) ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, it.p);
}

template <class K>
int run_items(cudaStream_t stream, K kernel, int smem, IdwtStreamCache *cache, std::vector<IdwtStreamCache::Item> &fresh)
{
    int launches = 0;
    if (cache && cache->replay) {
        /* same geometry, same buffers: the items of this call were recorded in the same order */
        while (cache->cursor < cache->items.size()) {
            IdwtStreamCache::Item &it = cache->items[cache->cursor++];
            launch_item(stream, kernel, smem, it, cache->cursor == 1);
            launches++;
            if (it.last)
                break;
        }
        return launches;
    }
    for (size_t i = 0; i < fresh.size(); i++) {
        fresh[i].last = i + 1 == fresh.size();
        launch_item(stream, kernel, smem, fresh[i]);
        launches++;
        if (cache)
            cache->items.push_back(fresh[i]);
    }
    return launches;
}

int cache_fail(IdwtStreamCache *cache)
{
    if (cache) {
        cache->items.clear();
        cache->serial = ~0ull;
    }
    return -1;
}

void finish_item(std::vector<IdwtStreamCache::Item> &fresh, IdwtStreamCache::Item &it, int nd, const IdwtFrame &f,
                 void *out, int strips, int seg, int segs)
{
    it.p.coefA = (const uint32_t *)f.coefA;
    it.p.out = out;
    it.p.strips = strips;
    it.p.SEG = seg;
    it.grid = dim3((strips * segs + SWARPS - 1) / SWARPS, nd);
    fresh.push_back(it);
}

template <int T>
int stream_level(cudaStream_t stream, const IdwtFrame &f, int j, int max_w, int max_h, const void *in, void *out)
{
    constexpr int NP = StreamCfg<T>::NP, VALID = 60 * NP;
    constexpr int sm = stage_bytes_per_cta<NP, 1, STAGES_LEVEL>();
    static PerDevice capacity;
    const int cap = capacity.get(idwt_stream_level_kernel<T, NP>, sm);
    IdwtStreamCache *cache = f.stream_cache;
    std::vector<IdwtStreamCache::Item> fresh;
    if (!(cache && cache->replay)) {
        const int strips = (max_w + 1 + VALID - 1) / VALID;
        int ndom = 0;
        for (int i = 0; i < f.ntcs; i++)
            ndom += f.h_tcs[i].transform == T && j < f.h_tcs[i].nlevels - 1;
        const int seg = pick_seg(strips * ndom, max_h, cap), segs = (max_h + 1 + seg - 1) / seg;
        IdwtStreamCache::Item it = {};
        int nd = 0;
        for (int i = 0; i < f.ntcs; i++) {
            const TileCompDesc *tc = &f.h_tcs[i];
            if (tc->transform != T || j >= tc->nlevels - 1)
                continue;
            if (!fill_dom(it.p.dom[nd], &tc, 1, j, f.coefA, in, NP))
                return cache_fail(cache);
            if (++nd == MAXDOM) {
                finish_item(fresh, it, nd, f, out, strips, seg, segs);
                nd = 0;
            }
        }
        if (nd)
            finish_item(fresh, it, nd, f, out, strips, seg, segs);
    }
    return run_items(stream, idwt_stream_level_kernel<T, NP>, sm, cache, fresh);
}

template <int T, int NC>
int stream_final(cudaStream_t stream, const IdwtFrame &f, int max_w, int max_h)
{
    constexpr int NP = StreamCfg<T>::NP, VALID = 60 * NP;
    constexpr int sm = stage_bytes_per_cta<NP, NC, STAGES_FINAL>();
    static PerDevice capacity;
    const int cap = capacity.get(idwt_stream_final_kernel<T, NP, NC>, sm);
    IdwtStreamCache *cache = f.stream_cache;
    std::vector<IdwtStreamCache::Item> fresh;
    if (!(cache && cache->replay)) {
        const int strips = (max_w + 1 + VALID - 1) / VALID;
        int ndom = 0;
        for (int i = 0; i < f.ngroups; i++)
            ndom += f.h_tcs[f.h_groups[i].tc[0]].transform == T;
        const int seg = pick_seg(strips * ndom, max_h, cap), segs = (max_h + 1 + seg - 1) / seg;
        IdwtStreamCache::Item it = {};
        int nd = 0;
        for (int i = 0; i < f.ngroups; i++) {
            const PackGroup &g = f.h_groups[i];
            const TileCompDesc *tcs[3] = { &f.h_tcs[g.tc[0]], &f.h_tcs[g.tc[NC > 1 ? 1 : 0]], &f.h_tcs[g.tc[NC > 2 ? 2 : 0]] };
            if (tcs[0]->transform != T)
                continue;
            const int nl = tcs[0]->nlevels;
            StreamDom &d = it.p.dom[nd];
            if (!fill_dom(d, tcs, NC, nl - 1, f.coefA, ((nl - 1) & 1) ? f.coefB : f.coefC, NP))
                return cache_fail(cache);
            d.out_x = tcs[0]->out_x;
            d.out_y = tcs[0]->out_y;
            d.rb = f.plane_row_bytes[g.plane];
            d.plane_off = f.plane_off[g.plane];
            d.mct = g.mct;
            d.shift = g.shift[0];
            for (int c = 0; c < 3; c++)
                d.cbps[c] = tcs[c]->cbps;
            if (++nd == MAXDOM) {
                finish_item(fresh, it, nd, f, f.picture, strips, seg, segs);
                nd = 0;
            }
        }
        if (nd)
            finish_item(fresh, it, nd, f, f.picture, strips, seg, segs);
    }
    return run_items(stream, idwt_stream_final_kernel<T, NP, NC>, sm, cache, fresh);
}

}  // namespace

int launch_idwt_stream_level(cudaStream_t stream, int type, const IdwtFrame &f, int j, int max_w, int max_h,
                             const void *in, void *out)
{
    if (type == J2K_DWT53) return stream_level<J2K_DWT53>(stream, f, j, max_w, max_h, in, out);
    if (type == J2K_DWT97) return stream_level<J2K_DWT97>(stream, f, j, max_w, max_h, in, out);
    return stream_level<J2K_DWT97_INT>(stream, f, j, max_w, max_h, in, out);
}

int launch_idwt_stream_final(cudaStream_t stream, int type, int ncomp, const IdwtFrame &f, int max_w, int max_h)
{
    if (ncomp == 3) {
        if (type == J2K_DWT53) return stream_final<J2K_DWT53, 3>(stream, f, max_w, max_h);
        if (type == J2K_DWT97) return stream_final<J2K_DWT97, 3>(stream, f, max_w, max_h);
        return stream_final<J2K_DWT97_INT, 3>(stream, f, max_w, max_h);
    }
    if (type == J2K_DWT53) return stream_final<J2K_DWT53, 1>(stream, f, max_w, max_h);
    if (type == J2K_DWT97) return stream_final<J2K_DWT97, 1>(stream, f, max_w, max_h);
    return stream_final<J2K_DWT97_INT, 1>(stream, f, max_w, max_h);
}

}  // namespace b200j2k
