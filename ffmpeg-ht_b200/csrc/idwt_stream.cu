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
 *   - per row each lane loads NP low-pass and NP high-pass coefficients per component with ONE 128-bit
 *     (NP = 4) or 64-bit (NP = 2) load each -- the Mallat quadrants are read as coalesced row segments;
 *   - horizontal synthesis runs in registers, neighbours across lanes come from one shuffle per lifting
 *     step; the first and last lane of the warp are apron;
 *   - vertical synthesis is a sliding window over the rows: 2 (5/3) or 4 (9/7) state values per column
 *     stay in registers, so every coefficient is read once and no intermediate ever leaves the SM;
 *   - the last level applies RCT / ICT, level shift, clip and packs rgb48 / gray16 with 128-bit stores.
 * Rows before columns at every level (quirk Q7), uint32 wrap-around 5/3 arithmetic and un-contracted fp32
 * 9/7 arithmetic exactly as idwt.cu.  Geometry the fast path does not cover (odd origins, unaligned
 * strides, one-sample lines, 8-bit or planar output) stays on idwt.cu; launch_idwt_frame picks per launch.
 */
#include "idwt_common.cuh"

namespace b200j2k {

namespace {

constexpr int SWARPS = 4;                 /* warps (independent strips) per CTA */

template <int NP> struct VecIO;
template <> struct VecIO<4> {
    static __device__ __forceinline__ void load(const uint32_t *p, uint32_t (&v)[4])
    {
        const uint4 t = __ldg(reinterpret_cast<const uint4 *>(p));
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
};
template <> struct VecIO<2> {
    static __device__ __forceinline__ void load(const uint32_t *p, uint32_t (&v)[2])
    {
        const uint2 t = __ldg(reinterpret_cast<const uint2 *>(p));
        v[0] = t.x; v[1] = t.y;
    }
};

#ifndef STREAM_STAGES
#define STREAM_STAGES 2
#endif
#ifndef STREAM_MINB
#define STREAM_MINB 5
#endif
#ifndef STREAM_EXP
#define STREAM_EXP 0
#endif
#ifndef STREAM_FILL
#define STREAM_FILL 100                   /* percent of the one-wave capacity the grid may use */
#endif
constexpr int STAGES = STREAM_STAGES;     /* row pairs in flight per warp (cp.async groups) */

template <int NP> __device__ __forceinline__ void cp_async_vec(uint32_t dst, const uint32_t *src)
{
    if (NP == 4)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
    else
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
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
/* Bulk-copy staging (interior strips).  ncu + tools/membench.cu: per-lane loads of this row pattern -- LDG or
 * cp.async alike -- are bounded by the L1 miss path (the kernel slows from 44 to 59 us when the shared-memory
 * carve-out leaves L1 small; 113 MB arrive in ~35 us), while the same rows fetched by the TMA engine with
 * cp.async.bulk arrive in ~22 us whatever the carve-out.  So lane 0 of a warp fetches the warp's 32 x NP
 * columns of each of the 4 NC row streams with ONE bulk copy per stream into the warp's private slots and
 * an mbarrier per stage counts the bytes in.  Bulk copies need 16-byte aligned addresses: with NP = 2 the
 * warp's first column sits at 8 mod 16 bytes, so the copy starts PADE = 2 elements early. */
template <int NP> struct Bulk {
    static constexpr int PADE = (4 - NP % 4) % 4;                         /* elements in front of lane 0's */
    static constexpr uint32_t SLOT = 32 * NP * 4 + (PADE ? 16 : 0);       /* bytes per stream and stage */
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
#ifndef STREAM_HINT
#define STREAM_HINT 1
#endif
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
#if STREAM_HINT
    /* the coefficients are read exactly once: let them leave L2 first, the picture written here is read next */
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(pol) : "memory");
#else
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
#endif
}
/* the four row streams of one component in one statement: one elected-thread wrapper and one set of uniform
 * operands for the destination base, the size and the barrier instead of four */
template <uint32_t STRIDE>
__device__ __forceinline__ void bulk_g2s_x4(uint32_t dst, const void *s0, const void *s1, const void *s2, const void *s3, uint32_t bar)
{
#if STREAM_HINT
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %5, [%6], %7;\n"
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0+%5], [%2], %5, [%6], %7;\n"
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0+%8], [%3], %5, [%6], %7;\n"
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0+%9], [%4], %5, [%6], %7;"
        ::"r"(dst), "l"(s0), "l"(s1), "l"(s2), "l"(s3), "n"(STRIDE), "r"(bar), "l"(pol), "n"(2 * STRIDE), "n"(3 * STRIDE) : "memory");
#else
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %5, [%6];\n"
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0+%5], [%2], %5, [%6];\n"
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0+%7], [%3], %5, [%6];\n"
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0+%8], [%4], %5, [%6];"
        ::"r"(dst), "l"(s0), "l"(s1), "l"(s2), "l"(s3), "n"(STRIDE), "r"(bar), "n"(2 * STRIDE), "n"(3 * STRIDE) : "memory");
#endif
}
/* bytes of staging shared memory per warp (the bulk layout is the larger one; border strips use per-lane
 * cp.async slots inside the same region), and of the CTA: SWARPS regions + two mbarriers per warp */
template <int NP, int NC> __host__ __device__ constexpr int stage_bytes_per_warp() { return STAGES * 4 * NC * (int)Bulk<NP>::SLOT; }
template <int NP, int NC> __host__ __device__ constexpr int stage_bytes_per_cta() { return SWARPS * stage_bytes_per_warp<NP, NC>() + SWARPS * STAGES * 8; }

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

__device__ __forceinline__ StripGeo make_geo(const TileCompDesc &tc, int j)
{
    const LevelDesc o = tc.lev[j + 1];
    StripGeo g;
    g.u0 = o.u0; g.u1 = o.u1; g.v0 = o.v0; g.v1 = o.v1;
    g.nLh = ((o.u1 + 1) >> 1) - ((o.u0 + 1) >> 1);
    g.nLv = ((o.v1 + 1) >> 1) - ((o.v0 + 1) >> 1);
    g.W = tc.width;
    g.LLstride = j == 0 ? tc.width : g.nLh;
    g.eoff = (o.u0 + 1) >> 1;
    g.ooff = o.u0 >> 1;
    g.eoffv = (o.v0 + 1) >> 1;
    g.ooffv = o.v0 >> 1;
    return g;
}

/* element offsets of the low-pass / high-pass halves of absolute row Y: offL < 0 means ~offL into the
 * plane (LH band), else into the LL source; offH into the plane (HL or HH band) */
__device__ __forceinline__ void row_offsets(const StripGeo &g, int Y, int &offL, int &offH)
{
    const int Yr = (Y >= g.v0 && Y < g.v1) ? Y : reflect(Y, g.v0, g.v1);
    const int ky = band_index(Yr, g.v0);
    if (Yr & 1) {
        offL = ~((g.nLv + ky) * g.W);
        offH = (g.nLv + ky) * g.W + g.nLh;
    } else {
        offL = ky * g.LLstride;
        offH = ky * g.W + g.nLh;
    }
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

template <int NP>
__device__ __forceinline__ void load_row(const uint32_t *rowL, const uint32_t *rowH, const LaneCols<NP> &lc,
                                         uint32_t (&Lc)[NP], uint32_t (&Hc)[NP])
{
    if (lc.inside) {
        VecIO<NP>::load(rowL + lc.vcolL, Lc);
        VecIO<NP>::load(rowH + lc.vcolH, Hc);
    } else {
#pragma unroll
        for (int k = 0; k < NP; k++) {
            Lc[k] = rowL[lc.ie[k]];
            Hc[k] = rowH[lc.io[k]];
        }
    }
}

/* vertical lifting state of one lane: per component and column the values the next row pair needs */
template <int TYPE, int NC, int C2> struct VState {
    uint32_t vh[NC][C2], e1[NC][C2];
    uint32_t o1[TYPE == J2K_DWT53 ? 1 : NC][TYPE == J2K_DWT53 ? 1 : C2];
    uint32_t e2[TYPE == J2K_DWT53 ? 1 : NC][TYPE == J2K_DWT53 ? 1 : C2];
};

/* band_index of the (reflected) absolute row Y */
__device__ __forceinline__ int row_ky(const StripGeo &g, int Y)
{
    const int Yr = (Y >= g.v0 && Y < g.v1) ? Y : reflect(Y, g.v0, g.v1);
    return band_index(Yr, g.v0);
}

/* The streaming core: walks row pairs (even row = vertical low-pass, odd row = vertical high-pass), keeps the
 * vertical lifting state in registers and calls emit(allin, Y, rows[NC][2NP]) for every finished output row.
 * The loop is unrolled by two with the state ping-ponging between two register sets, so the sliding window
 * never costs a register move, and the staging slot of a step is a compile-time constant (STAGES == 2).
 * ALLIN: every lane of the warp lies inside the line (no column needs reflection) -- the common case, kept
 * free of per-lane branches; the other instantiation serves the first and last strip of a line.
 * Addressing (ncu: the first version spent ~190 of its ~700 instructions per row pair re-deriving twelve
 * 64-bit addresses): the lane keeps one pointer per component into the LL source and one into the plane
 * (at its column of the HL band); a row pair contributes four element offsets shared by all components. */
template <int TYPE, int NP, int NC, bool ALLIN, class Emit>
__device__ __forceinline__ void stream_rows(const StripGeo &g, const uint32_t *__restrict__ baseA,
                                            const uint32_t *__restrict__ baseL, const uint32_t (&offA)[NC],
                                            const uint32_t (&offL)[NC], const LaneCols<NP> &lc, int Ya, int Yb,
                                            uint32_t warp_base, uint32_t bar_base, Emit &&emit)
{
    using L = Lift<TYPE>;
    constexpr int HP = L::HALO / 2;
    constexpr int C2 = 2 * NP;
    constexpr int NV = 4 * NC;
    constexpr uint32_t VSTRIDE = ALLIN ? Bulk<NP>::SLOT : 32 * NP * 4, SSTRIDE = NV * VSTRIDE;
    constexpr int PADE = Bulk<NP>::PADE;
    const int lane = threadIdx.x & 31;
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
    const bool fast = ALLIN || lc.inside;
    /* element offsets (from the kernel-uniform bases) of the lane's first column in the LL source and in the
     * HL band of the plane; 32-bit on purpose: twelve 64-bit lane pointers do not fit the register budget */
    uint32_t eL[NC], eA[NC];
#pragma unroll
    for (int c = 0; c < NC; c++) {
        eL[c] = offL[c] + (uint32_t)lc.vcolL;
        eA[c] = offA[c] + (uint32_t)(g.nLh + lc.vcolH);
    }
    const uint32_t dLH = (uint32_t)(lc.vcolL - lc.vcolH - g.nLh);   /* LH column relative to the HH column */
    /* where copies land (stage_base) and where the lane reads its NP elements of a stream back (lane_base) */
    const uint32_t stage_base = ALLIN ? warp_base : warp_base + (uint32_t)lane * (NP * 4);
    const uint32_t lane_base = warp_base + (uint32_t)lane * (NP * 4) + (ALLIN ? PADE * 4 : 0);
    /* Asynchronous staging, one row pair ahead, so the bytes in flight do not cost registers.  Interior
     * strips: bulk copies by lane 0 (see Bulk<>), completion through the stage's mbarrier.  Border strips:
     * every inside lane fills its own NV vector slots with cp.async; lanes whose columns need reflection
     * load synchronously when the row pair is consumed.  A lane only ever reads back its own bytes. */
    auto issue = [&](int n, uint32_t st, uint32_t bar) {
        if (n < n1 && fast) {
            /* band_index() of rows 2n / 2n+1 is n - ceil(v0/2) / n - floor(v0/2) unless a row needs reflection */
            int kyE = n - g.eoffv, kyO = n - g.ooffv;
            if (2 * n < g.v0 || 2 * n + 1 >= g.v1) {
                kyE = row_ky(g, 2 * n);
                kyO = row_ky(g, 2 * n + 1);
            }
            const uint32_t oLL = (uint32_t)(kyE * g.LLstride), oHL = (uint32_t)(kyE * g.W);
            const uint32_t oHH = (uint32_t)((g.nLv + kyO) * g.W), oLH = oHH + dLH;
            if (ALLIN) {
                __syncwarp();                    /* every lane is done with the slots this refills */
                if (lane == 0) {
                    mbar_expect_tx(bar, NV * VSTRIDE);
#pragma unroll
                    for (int c = 0; c < NC; c++)
                        bulk_g2s_x4<VSTRIDE>(st + 4 * c * VSTRIDE, baseL + (eL[c] + oLL - PADE), baseA + (eA[c] + oHL - PADE),
                                             baseA + (eA[c] + oLH - PADE), baseA + (eA[c] + oHH - PADE), bar);
                }
            } else {
#pragma unroll
                for (int c = 0; c < NC; c++) {
                    cp_async_vec<NP>(st + (4 * c + 0) * VSTRIDE, baseL + (eL[c] + oLL));
                    cp_async_vec<NP>(st + (4 * c + 1) * VSTRIDE, baseA + (eA[c] + oHL));
                    cp_async_vec<NP>(st + (4 * c + 2) * VSTRIDE, baseA + (eA[c] + oLH));
                    cp_async_vec<NP>(st + (4 * c + 3) * VSTRIDE, baseA + (eA[c] + oHH));
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
        issue(n + STAGES - 1, stage_base + (uint32_t)prev * SSTRIDE, bar_base + (uint32_t)prev * 8);
        if (ALLIN)
            mbar_wait(bar_base + (uint32_t)slot * 8, phase);
        else
            asm volatile("cp.async.wait_group %0;" ::"n"(STAGES - 1) : "memory");
        if (++slot == STAGES) {
            slot = 0;
            phase ^= 1;
        }
        uint32_t cLe[NC][NP], cHe[NC][NP], cLo[NC][NP], cHo[NC][NP];
        if (fast) {
#pragma unroll
            for (int c = 0; c < NC; c++) {
                lds_vec<NP>(st + (4 * c + 0) * VSTRIDE, cLe[c]);
                lds_vec<NP>(st + (4 * c + 1) * VSTRIDE, cHe[c]);
                lds_vec<NP>(st + (4 * c + 2) * VSTRIDE, cLo[c]);
                lds_vec<NP>(st + (4 * c + 3) * VSTRIDE, cHo[c]);
            }
        } else {
            int oL, oH;
            row_offsets(g, 2 * n, oL, oH);
#pragma unroll
            for (int c = 0; c < NC; c++)
                load_row<NP>(oL >= 0 ? baseL + offL[c] + oL : baseA + offA[c] + ~oL, baseA + offA[c] + oH, lc, cLe[c], cHe[c]);
            row_offsets(g, 2 * n + 1, oL, oH);
#pragma unroll
            for (int c = 0; c < NC; c++)
                load_row<NP>(oL >= 0 ? baseL + offL[c] + oL : baseA + offA[c] + ~oL, baseA + offA[c] + oH, lc, cLo[c], cHo[c]);
        }
#if STREAM_EXP == 1                              /* tuning experiment: loads only */
        {
            uint32_t acc = 0;
#pragma unroll
            for (int c = 0; c < NC; c++)
#pragma unroll
                for (int k = 0; k < NP; k++)
                    acc ^= cLe[c][k] + cHe[c][k] + cLo[c][k] + cHo[c][k];
            out.vh[0][0] = in.vh[0][0] ^ acc;
            if (n == n1 - 1 && out.vh[0][0] == 0x12345678u)
                emit(ALLIN, Ylo, out.e1);
            return;
        }
#endif
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
#if STREAM_EXP == 2                              /* tuning experiment: loads + arithmetic, no stores */
        {
            uint32_t acc = 0;
#pragma unroll
            for (int c = 0; c < NC; c++)
#pragma unroll
                for (int k = 0; k < C2; k++)
                    acc ^= rowA[c][k];
            if (acc == 0x12345678u && n == n1 - 1)
                emit(ALLIN, Ylo, rowA);
            return;
        }
#endif
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
        issue(n0 + s, stage_base + (uint32_t)s * SSTRIDE, bar_base + (uint32_t)s * 8);
    int n = n0;
    for (; n + 1 < n1; n += 2) {
        step(n, S0, S1);
        step(n + 1, S1, S0);
    }
    if (n < n1)
        step(n, S0, S1);
}

/* ---- intermediate level: reconstructed resolution -> compact ping-pong buffer -------------------------- */
template <int TYPE, int NP>
__global__ void __launch_bounds__(SWARPS * 32)
idwt_stream_level_kernel(const TileCompDesc *__restrict__ tcs, int j, const uint32_t *__restrict__ coefA,
                         const uint32_t *__restrict__ coefIn, uint32_t *__restrict__ coefOut, int strips, int SEG)
{
    extern __shared__ __align__(16) uint32_t stage_smem[];
    constexpr int C2 = 2 * NP, VALID = 30 * C2;
    const TileCompDesc &tc = tcs[blockIdx.y];
    if (tc.transform != TYPE || j >= tc.nlevels - 1)
        return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int unit = blockIdx.x * SWARPS + warp;
    const int strip = unit % strips, seg = unit / strips;
    const StripGeo g = make_geo(tc, j);
    const int XA = (g.u0 & ~1) + strip * VALID - C2;          /* first column of lane 0 (apron lane) */
    const int Ya = (g.v0 & ~1) + seg * SEG, Yb = min(Ya + SEG, (g.v1 + 1) & ~1);
    if (XA + C2 >= g.u1 || Ya >= g.v1)
        return;
    const int p0 = (XA >> 1) + lane * NP;
    const LaneCols<NP> lc = make_cols<NP>(g, p0);
    const uint32_t offA[1] = { (uint32_t)tc.coef_off }, offL[1] = { (uint32_t)tc.coef_off };
    const uint32_t *baseL = j == 0 ? coefA : coefIn;
    uint32_t *out = coefOut + tc.coef_off;
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
    const uint32_t bar_base = smem0 + warp * (STAGES * 8);
    const uint32_t warp_base = smem0 + SWARPS * STAGES * 8 + warp * stage_bytes_per_warp<NP, 1>();
    const bool allin = __all_sync(FULLMASK, lc.inside);
    if (allin && lane == 0) {
#pragma unroll
        for (int i = 0; i < STAGES; i++)
            mbar_init(bar_base + 8 * i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    if (allin)
        stream_rows<TYPE, NP, 1, true>(g, coefA, baseL, offA, offL, lc, Ya, Yb, warp_base, bar_base, emit);
    else
        stream_rows<TYPE, NP, 1, false>(g, coefA, baseL, offA, offL, lc, Ya, Yb, warp_base, bar_base, emit);
}

/* ---- last level fused with MCT + level shift + clip + pack (packed 16-bit output) ---------------------- */
template <int TYPE, int NP, int NC>
__global__ void __launch_bounds__(SWARPS * 32, NP == 4 ? 2 : (TYPE == J2K_DWT53 ? STREAM_MINB : 3))
idwt_stream_final_kernel(const TileCompDesc *__restrict__ tcs, const PackGroup *__restrict__ groups,
                         const uint32_t *__restrict__ coefA, const uint32_t *__restrict__ coefIn0,
                         const uint32_t *__restrict__ coefIn1, uint8_t *__restrict__ picture,
                         uint64_t po0, uint64_t po1, uint64_t po2, uint64_t po3, int rb0, int rb1, int rb2, int rb3,
                         int strips, int SEG)
{
    extern __shared__ __align__(16) uint32_t stage_smem[];
    constexpr int C2 = 2 * NP, VALID = 30 * C2;
    const PackGroup &gr = groups[blockIdx.y];
    const TileCompDesc &t0 = tcs[gr.tc[0]];
    if (t0.transform != TYPE || gr.ncomp != NC)
        return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int unit = blockIdx.x * SWARPS + warp;
    const int strip = unit % strips, seg = unit / strips;
    const int nl = t0.nlevels;
    const StripGeo g = make_geo(t0, nl - 1);
    const int XA = (g.u0 & ~1) + strip * VALID - C2;
    const int Ya = (g.v0 & ~1) + seg * SEG, Yb = min(Ya + SEG, (g.v1 + 1) & ~1);
    if (XA + C2 >= g.u1 || Ya >= g.v1)
        return;
    const int p0 = (XA >> 1) + lane * NP;
    const LaneCols<NP> lc = make_cols<NP>(g, p0);
    const uint32_t *baseL = nl == 1 ? coefA : ((nl - 1) & 1) ? coefIn0 : coefIn1;
    uint32_t offA[NC], offL[NC];
    int cbps[NC], shift[NC];
#pragma unroll
    for (int c = 0; c < NC; c++) {
        const TileCompDesc &tc = tcs[gr.tc[c]];
        offA[c] = offL[c] = (uint32_t)tc.coef_off;
        cbps[c] = tc.cbps;
        shift[c] = gr.shift[c];
    }
    const uint64_t poff = gr.plane == 0 ? po0 : gr.plane == 1 ? po1 : gr.plane == 2 ? po2 : po3;
    const int rb = gr.plane == 0 ? rb0 : gr.plane == 1 ? rb1 : gr.plane == 2 ? rb2 : rb3;
    uint8_t *plane = picture + poff;
    const int mct_on = gr.mct;
    const int X0 = 2 * p0;
    const bool lane_ok = lane >= 1 && lane <= 30;
    const long px = (long)(t0.out_x + X0 - g.u0) * NC;       /* sample index of the lane's first pixel */
    int half[NC], vmax[NC];
#pragma unroll
    for (int c = 0; c < NC; c++) {
        half[c] = 1 << (cbps[c] - 1);
        vmax[c] = (1 << cbps[c]) - 1;
    }
    /* byte address of this lane's first sample in output row Y = out_base + Y * rb */
    uint8_t *out_base = plane + ((long)t0.out_y - g.v0) * rb + 2 * px;
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
                w[i] = __byte_perm(pix[s0 % NC][s0 / NC], pix[s1 % NC][s1 / NC], 0x5410) << shift[0];
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
                        r16[k * NC + c] = (uint16_t)(pix[c][k] << shift[c]);
        }
    };
    const uint32_t smem0 = (uint32_t)__cvta_generic_to_shared(stage_smem);
    const uint32_t bar_base = smem0 + warp * (STAGES * 8);
    const uint32_t warp_base = smem0 + SWARPS * STAGES * 8 + warp * stage_bytes_per_warp<NP, NC>();
    const bool allin = __all_sync(FULLMASK, lc.inside);
    if (allin && lane == 0) {
#pragma unroll
        for (int i = 0; i < STAGES; i++)
            mbar_init(bar_base + 8 * i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    if (allin)
        stream_rows<TYPE, NP, NC, true>(g, coefA, baseL, offA, offL, lc, Ya, Yb, warp_base, bar_base, emit);
    else
        stream_rows<TYPE, NP, NC, false>(g, coefA, baseL, offA, offL, lc, Ya, Yb, warp_base, bar_base, emit);
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
template <class K> int resident_ctas(K kernel, int smem)
{
    int dev = 0, sms = 0, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, SWARPS * 32, smem);
    return sms * (per_sm > 0 ? per_sm : 1);
}

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

}  // namespace

bool idwt_stream_levels_ok(const TileCompDesc *h_tcs, int ntcs, int type)
{
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
            if (g.shift[c] != g.shift[0] || tc.width != t0.width)
                return false;
            const LevelDesc &a = tc.lev[tc.nlevels], &b = t0.lev[t0.nlevels];
            if (a.u0 != b.u0 || a.u1 != b.u1 || a.v0 != b.v0 || a.v1 != b.v1)
                return false;
        }
        if ((t0.out_x & 7) || (f.plane_row_bytes[g.plane] & 15) || (f.plane_off[g.plane] & 15))
            return false;
    }
    *ncomp = nc;
    return nc != 0;
}

int launch_idwt_stream_level(cudaStream_t stream, int type, const IdwtFrame &f, int j, int max_w, int max_h,
                             const void *in, void *out)
{
#define LAUNCH(T)                                                                                              \
    do {                                                                                                       \
        constexpr int NP = StreamCfg<T>::NP, VALID = 60 * NP;                                                  \
        constexpr int sm = stage_bytes_per_cta<NP, 1>();                                                       \
        static int capacity = 0;                                                                               \
        if (!capacity) {                                                                                       \
            cudaFuncSetAttribute(idwt_stream_level_kernel<T, NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm); \
            capacity = resident_ctas(idwt_stream_level_kernel<T, NP>, sm);                                     \
        }                                                                                                      \
        const int strips = (max_w + 1 + VALID - 1) / VALID;                                                    \
        const int seg = pick_seg(strips * f.ntcs, max_h, capacity), segs = (max_h + 1 + seg - 1) / seg;        \
        dim3 grid((strips * segs + SWARPS - 1) / SWARPS, f.ntcs);                                              \
        idwt_stream_level_kernel<T, NP><<<grid, SWARPS * 32, sm, stream>>>(                                    \
            f.d_tcs, j, (const uint32_t *)f.coefA, (const uint32_t *)in, (uint32_t *)out, strips, seg);        \
    } while (0)
    if (type == J2K_DWT53) LAUNCH(J2K_DWT53);
    else if (type == J2K_DWT97) LAUNCH(J2K_DWT97);
    else LAUNCH(J2K_DWT97_INT);
#undef LAUNCH
    return 1;
}

int launch_idwt_stream_final(cudaStream_t stream, int type, int ncomp, const IdwtFrame &f, int max_w, int max_h)
{
#define LAUNCH(T, NC)                                                                                          \
    do {                                                                                                       \
        constexpr int NP = StreamCfg<T>::NP, VALID = 60 * NP;                                                  \
        constexpr int sm = stage_bytes_per_cta<NP, NC>();                                                      \
        static int capacity = 0;                                                                               \
        if (!capacity) {                                                                                       \
            cudaFuncSetAttribute(idwt_stream_final_kernel<T, NP, NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm); \
            capacity = resident_ctas(idwt_stream_final_kernel<T, NP, NC>, sm);                                 \
        }                                                                                                      \
        const int strips = (max_w + 1 + VALID - 1) / VALID;                                                    \
        const int seg = pick_seg(strips * f.ngroups, max_h, capacity), segs = (max_h + 1 + seg - 1) / seg;     \
        dim3 grid((strips * segs + SWARPS - 1) / SWARPS, f.ngroups);                                           \
        idwt_stream_final_kernel<T, NP, NC><<<grid, SWARPS * 32, sm, stream>>>(                                \
            f.d_tcs, f.d_groups, (const uint32_t *)f.coefA, (const uint32_t *)f.coefB, (const uint32_t *)f.coefC, \
            f.picture, f.plane_off[0], f.plane_off[1], f.plane_off[2], f.plane_off[3], f.plane_row_bytes[0],   \
            f.plane_row_bytes[1], f.plane_row_bytes[2], f.plane_row_bytes[3], strips, seg);                    \
    } while (0)
    if (ncomp == 3) {
        if (type == J2K_DWT53) LAUNCH(J2K_DWT53, 3);
        else if (type == J2K_DWT97) LAUNCH(J2K_DWT97, 3);
        else LAUNCH(J2K_DWT97_INT, 3);
    } else {
        if (type == J2K_DWT53) LAUNCH(J2K_DWT53, 1);
        else if (type == J2K_DWT97) LAUNCH(J2K_DWT97, 1);
        else LAUNCH(J2K_DWT97_INT, 1);
    }
#undef LAUNCH
    return 1;
}

}  // namespace b200j2k
