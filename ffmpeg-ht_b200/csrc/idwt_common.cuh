/*
 * idwt_common.cuh -- lifting steps, symmetric-extension index helpers, component transforms and the
 * sample -> pixel conversion shared by the inverse-DWT kernels (idwt.cu: shared-memory tiled, general;
 * idwt_stream.cu: register-streaming fast path).
 * References: libavcodec/jpeg2000dwt.c:309-537 (lifting), libavcodec/jpeg2000dsp.c:43-91 (RCT / ICT),
 * libavcodec/jpeg2000dec.c:2301-2366 (level shift, clip, pack).
 */
#ifndef B200J2K_IDWT_COMMON_CUH
#define B200J2K_IDWT_COMMON_CUH

#include <cuda_runtime.h>
#include <stdint.h>

#include "j2k_kernels.h"
#include "j2k_plan.h"

namespace b200j2k {
namespace {

constexpr int THREADS = 256;
constexpr int NWARPS = THREADS / 32;
constexpr unsigned FULLMASK = 0xFFFFFFFFu;

__device__ __forceinline__ int reflect(int i, int i0, int i1)
{
    const int n = i1 - i0;
    if (n == 1)
        return i0;
    const int period = 2 * (n - 1);
    int k = (i - i0) % period;
    if (k < 0)
        k += period;
    return i0 + (k < n ? k : period - k);
}

/* index of absolute position X inside its (low or high) band of a line that starts at absolute u0 */
__device__ __forceinline__ int band_index(int X, int u0)
{
    return (X & 1) ? ((X - 1) >> 1) - (u0 >> 1) : (X >> 1) - ((u0 + 1) >> 1);
}

template <int TYPE> struct Lift;

/* 5/3: uint32 wrap-around adds, arithmetic shift of the int cast (jpeg2000dwt.c:321-324) */
template <> struct Lift<J2K_DWT53> {
    static constexpr int HALO = 2, STEPS = 2;
    /* step s updates positions of parity (s & 1): step 0 evens, step 1 odds */
    static __device__ __forceinline__ uint32_t apply(int s, uint32_t c, uint32_t a, uint32_t b)
    {
        if (s == 0)
            return c - (uint32_t)((int32_t)(a + b + 2) >> 2);
        return c + (uint32_t)((int32_t)(a + b) >> 1);
    }
    static __device__ __forceinline__ uint32_t single(uint32_t c, int odd)
    {
        return odd ? (uint32_t)((int32_t)c >> 1) : c;
    }
};

/* 9/7 float (jpeg2000dwt.c:376-401) */
template <> struct Lift<J2K_DWT97> {
    static constexpr int HALO = 4, STEPS = 4;
    static __device__ __forceinline__ uint32_t apply(int s, uint32_t c, uint32_t a, uint32_t b)
    {
        const float fc = __uint_as_float(c);
        const float sum = __fadd_rn(__uint_as_float(a), __uint_as_float(b));
        float r;
        if (s == 0) r = __fsub_rn(fc, __fmul_rn(0.443506852043971f, sum));
        else if (s == 1) r = __fsub_rn(fc, __fmul_rn(0.882911075530934f, sum));
        else if (s == 2) r = __fadd_rn(fc, __fmul_rn(0.052980118572961f, sum));
        else r = __fadd_rn(fc, __fmul_rn(1.586134342059924f, sum));
        return __float_as_uint(r);
    }
    static __device__ __forceinline__ uint32_t single(uint32_t c, int odd)
    {
        const float f = __uint_as_float(c);
        return __float_as_uint(odd ? __fmul_rn(f, 1.230174104914001f / 2) : __fmul_rn(f, 0.812893066115961f));
    }
};

/* 9/7 16.16 fixed point (jpeg2000dwt.c:453-481) */
template <> struct Lift<J2K_DWT97_INT> {
    static constexpr int HALO = 4, STEPS = 4;
    static __device__ __forceinline__ uint32_t apply(int s, uint32_t c, uint32_t a, uint32_t b)
    {
        const int64_t sum = (int64_t)(int32_t)a + (int64_t)(int32_t)b;
        int32_t v = (int32_t)c;
        if (s == 0) v -= (int32_t)((29066ll * sum + (1 << 15)) >> 16);
        else if (s == 1) v -= (int32_t)((57862ll * sum + (1 << 15)) >> 16);
        else if (s == 2) v += (int32_t)((3472ll * sum + (1 << 15)) >> 16);
        else {
            v += (int32_t)sum;
            v += (int32_t)((38413ll * sum + (1 << 15)) >> 16);
        }
        return (uint32_t)v;
    }
    static __device__ __forceinline__ uint32_t single(uint32_t c, int odd)
    {
        const int32_t v = (int32_t)c;
        return odd ? (uint32_t)(int32_t)((v * 80621ll + (1 << 16)) >> 17)
                   : (uint32_t)(int32_t)((v * 53274ll + (1 << 15)) >> 16);
    }
};

template <int TYPE> __device__ __forceinline__ void mct(uint32_t &a, uint32_t &b, uint32_t &c)
{
    if (TYPE == J2K_DWT53) {                     /* rct_int, jpeg2000dsp.c:78-91 */
        const uint32_t g = a - (uint32_t)((int32_t)(c + b) >> 2);
        const uint32_t r = g + c, bl = g + b;
        a = r; b = g; c = bl;
    } else if (TYPE == J2K_DWT97) {              /* ict_float, :43-59 */
        const float y = __uint_as_float(a), cb = __uint_as_float(b), cr = __uint_as_float(c);
        const float r = __fadd_rn(y, __fmul_rn(1.402f, cr));
        const float g = __fsub_rn(__fsub_rn(y, __fmul_rn(0.34413f, cb)), __fmul_rn(0.71414f, cr));
        const float bl = __fadd_rn(y, __fmul_rn(1.772f, cb));
        a = __float_as_uint(r); b = __float_as_uint(g); c = __float_as_uint(bl);
    } else {                                     /* ict_int, :61-76 */
        const int32_t y = (int32_t)a, cb = (int32_t)b, cr = (int32_t)c;
        const int32_t r = y + cr + ((int32_t)(26345u * (uint32_t)cr + (1u << 15)) >> 16);
        const int32_t g = y - ((int32_t)(22553u * (uint32_t)cb + (1u << 15)) >> 16)
                            - ((int32_t)(46802u * (uint32_t)cr + (1u << 15)) >> 16);
        const int32_t bl = y + 2 * cb + ((int32_t)((uint32_t)-14942 * (uint32_t)cb + (1u << 15)) >> 16);
        a = (uint32_t)r; b = (uint32_t)g; c = (uint32_t)bl;
    }
}

template <int TYPE> __device__ __forceinline__ int to_pixel(uint32_t raw, int cbps, int shift)
{
    int val;
    if (TYPE == J2K_DWT97)
        val = __float2int_rn(__uint_as_float(raw));
    else
        val = (int32_t)raw;
    val += 1 << (cbps - 1);
    val = min(max(val, 0), (1 << cbps) - 1);
    return val << shift;
}


}  // namespace
}  // namespace b200j2k
#endif
