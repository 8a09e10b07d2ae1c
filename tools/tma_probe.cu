// Probe: rank-3 TMA box loads with the tensor map inside an indexed __grid_constant__ struct (tuning/debug helper).
#ifndef WW
#define WW 512
#define HH 64
#define NCC 3
#endif
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <vector>
struct alignas(64) Dom { CUtensorMap m; int x, y; int pad[14]; };
struct Launch { Dom d[4]; uint32_t *out; };
template <int HINT, int IDX>
__global__ void k(const __grid_constant__ Launch P)
{
    extern __shared__ __align__(128) uint32_t sm[];
    const Dom &D = P.d[IDX ? blockIdx.y : 0];
    const uint32_t s0 = (uint32_t)__cvta_generic_to_shared(sm);
    const uint32_t bar = s0, dst = s0 + 128;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    uint32_t el;
    asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}" : "=r"(el));
    if (el) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(NCC * 256) : "memory");
        if (HINT)
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4}], [%5], %6;"
                         ::"r"(dst), "l"(&D.m), "r"(D.x), "r"(D.y), "r"(0), "r"(bar), "l"(0x12F0000000000000ull) : "memory");
        else
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                         ::"r"(dst), "l"(&D.m), "r"(D.x), "r"(D.y), "r"(0), "r"(bar) : "memory");
    }
    uint32_t ok;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(bar), "r"(0) : "memory");
    } while (!ok);
    for (int i = threadIdx.x; i < NCC * 64; i += 32)
        P.out[blockIdx.y * 192 + i] = sm[32 + i];
}
int main(int argc, char **argv)
{
    const int W = WW, H = HH, NC = NCC;
    std::vector<uint32_t> h((size_t)W * H * NC);
    for (size_t i = 0; i < h.size(); i++) h[i] = (uint32_t)i;
    uint32_t *d, *o;
    cudaMalloc(&d, h.size() * 4); cudaMalloc(&o, 4 * 192 * 4);
    cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    void *fp = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
    auto enc = (PFN_cuTensorMapEncodeTiled_v12000)fp;
    Launch L; memset(&L, 0, sizeof(L));
    for (int i = 0; i < 4; i++) {
        cuuint64_t dims[3] = { W, H, NC }, str[2] = { W * 4, (cuuint64_t)W * H * 4 };
        cuuint32_t box[3] = { 64, 1, NC }, es[3] = { 1, 1, 1 };
        CUresult r = enc(&L.d[i].m, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r) printf("encode %d\n", (int)r);
        L.d[i].x = 8 * i + 4; L.d[i].y = 3 + i;
    }
    L.out = o;
    for (int v = 0; v < 4; v++) {
        cudaMemset(o, 0, 4 * 192 * 4);
        if (v == 0) k<0, 0><<<dim3(1, 4), 32, 2048>>>(L);
        if (v == 1) k<1, 0><<<dim3(1, 4), 32, 2048>>>(L);
        if (v == 2) k<0, 1><<<dim3(1, 4), 32, 2048>>>(L);
        if (v == 3) k<1, 1><<<dim3(1, 4), 32, 2048>>>(L);
        cudaError_t e = cudaDeviceSynchronize();
        uint32_t r[4 * 192];
        cudaMemcpy(r, o, sizeof(r), cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int b = 0; b < 4; b++) {
            const int di = (v & 2) ? b : 0;
            for (int c = 0; c < NC; c++) for (int x = 0; x < 64; x++)
                bad += r[b * 192 + c * 64 + x] != (uint32_t)(c * W * H + (3 + di) * W + 8 * di + 4 + x);
        }
        printf("variant hint=%d indexed=%d: %s, mismatches %d\n", v & 1, v >> 1, cudaGetErrorString(e), bad);
        if (e) return 1;
    }
    return 0;
}
