// Probe 2: the streaming kernels' shared-memory layout (4 warps, mbarriers at smem0 + 16*warp, slots behind a 128-byte line).
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cstdlib>
#include <vector>
struct alignas(64) Dom { CUtensorMap mL, mA; const uint32_t *ll; int u0, u1, v0, v1, W, LLs, tma; uint32_t off[3]; int o[5]; int cb[3]; uint64_t po; };
struct Launch { Dom d[4]; const uint32_t *a; uint32_t *out; int strips, seg; };
template <int EXITW>
__global__ void __launch_bounds__(128) k(const __grid_constant__ Launch P)
{
    extern __shared__ __align__(128) uint32_t sm[];
    const Dom &D = P.d[blockIdx.y];
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    if (EXITW && warp == 0) return;
    const uint32_t s0 = (uint32_t)__cvta_generic_to_shared(sm);
    const uint32_t bar = s0 + warp * 16, dst = s0 + 128 + warp * 2048;
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar + 8) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    uint32_t el;
    asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}" : "=r"(el));
    if (el) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(2 * 256) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4}], [%5], %6;"
                     ::"r"(dst), "l"(&D.mL), "r"(D.u0 + 64 * warp), "r"(D.v0), "r"(0), "r"(bar), "l"(0x12F0000000000000ull) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4}], [%5], %6;"
                     ::"r"(dst + 256), "l"(&D.mA), "r"(D.u1 + 64 * warp), "r"(D.v1), "r"(0), "r"(bar), "l"(0x12F0000000000000ull) : "memory");
    }
    uint32_t ok;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(bar), "r"(0) : "memory");
    } while (!ok);
    for (int i = lane; i < 128; i += 32)
        P.out[(blockIdx.y * 4 + warp) * 128 + i] = sm[32 + warp * 512 + i];
}
int main(int argc, char **argv)
{
    const int X0 = argc > 1 ? atoi(argv[1]) : 2, X1 = argc > 2 ? atoi(argv[2]) : 58;
    const int W = 3840, H = 2160;
    std::vector<uint32_t> h((size_t)W * H);
    for (size_t i = 0; i < h.size(); i++) h[i] = (uint32_t)i;
    uint32_t *d, *o;
    cudaMalloc(&d, h.size() * 4); cudaMalloc(&o, 16 * 128 * 4);
    cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    void *fp = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
    auto enc = (PFN_cuTensorMapEncodeTiled_v12000)fp;
    Launch L; memset(&L, 0, sizeof(L));
    for (int i = 0; i < 4; i++) {
        cuuint64_t dims[3] = { W, H, 1 }, str[2] = { W * 4, (cuuint64_t)W * H * 4 };
        cuuint64_t dimsL[3] = { 240, 135, 1 }, strL[2] = { 240 * 4, (cuuint64_t)240 * 135 * 4 };
        cuuint32_t box[3] = { 64, 1, 1 }, es[3] = { 1, 1, 1 };
        CUresult r = enc(&L.d[i].mA, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        CUresult r2 = enc(&L.d[i].mL, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, d, dimsL, strL, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r || r2) printf("encode %d %d\n", (int)r, (int)r2);
        L.d[i].u0 = X0; L.d[i].v0 = 3 + i; L.d[i].u1 = 240 + X1; L.d[i].v1 = 135 + i;
    }
    L.out = o;
    printf("sizeof Dom %zu Launch %zu\n", sizeof(Dom), sizeof(Launch));
    for (int v = 0; v < 2; v++) {
        cudaMemset(o, 0, 16 * 128 * 4);
        if (v == 0) k<0><<<dim3(1, 4), 128, 8320>>>(L);
        if (v == 1) k<1><<<dim3(1, 4), 128, 8320>>>(L);
        cudaError_t e = cudaDeviceSynchronize();
        uint32_t r[16 * 128];
        cudaMemcpy(r, o, sizeof(r), cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int b = 0; b < 4; b++) for (int w = v; w < 4; w++) for (int x = 0; x < 64; x++) {
            bad += r[(b * 4 + w) * 128 + x] != (uint32_t)((3 + b) * 240 + X0 + 64 * w + x);
            bad += r[(b * 4 + w) * 128 + 64 + x] != (uint32_t)((135 + b) * W + 240 + X1 + 64 * w + x);
        }
        printf("variant exit-warp0=%d: %s, mismatches %d\n", v, cudaGetErrorString(e), bad);
        if (e) return 1;
    }
    return 0;
}
