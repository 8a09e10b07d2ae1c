// Multi-GPU end-to-end ceiling probe (tuning / evidence helper, not product code).
// One process per GPU, all started for the same wall-clock second: each moves 50 MB "pictures" device -> host for a
// fixed window and reports its own GB/s, so that the SUM over processes is what the platform lets N GPUs write into
// host memory at once.  Variants of the host destination and of who does the writing:
//   pinned      cudaHostAlloc (what the decoder's frame pools use), copy engine
//   wc          cudaHostAllocWriteCombined, copy engine (diagnostic: is CPU-cache snooping the cap?)
//   thp         mmap + MADV_HUGEPAGE + cudaHostRegister (2 MB pages: fewer IOMMU / IOTLB entries), copy engine
//   sm          pinned + mapped, written by a kernel (SM stores over PCIe) instead of the copy engine
//   h2d         pinned -> device, copy engine (the packets' direction)
//   both        pinned D2H and H2D at the same time on two streams
// usage: pcie_probe <device> <start_unix_seconds> <window_seconds> [variant ...]
#include <cuda_runtime.h>
#include <sys/mman.h>
#include <sys/time.h>
#include <unistd.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

static double now()
{
    timeval tv;
    gettimeofday(&tv, nullptr);
    return tv.tv_sec + 1e-6 * tv.tv_usec;
}
__global__ void sm_copy(const uint4 *__restrict__ src, uint4 *__restrict__ dst, size_t n)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        dst[i] = src[i];
}
int main(int argc, char **argv)
{
    if (argc < 4)
        return 2;
    const int dev = atoi(argv[1]);
    const double start = atof(argv[2]), window = atof(argv[3]);
    cudaSetDevice(dev);
    const size_t bytes = 49766400;                          /* one 4K rgb48 picture */
    const int NB = 4;
    void *d[NB], *h[NB];
    cudaStream_t st[2];
    cudaStreamCreate(&st[0]);
    cudaStreamCreate(&st[1]);
    for (int i = 0; i < NB; i++) {
        cudaMalloc(&d[i], bytes);
        cudaMemset(d[i], i + 1, bytes);
    }
    double t_next = start;
    for (int a = 4; a < argc; a++) {
        const char *v = argv[a];
        bool ok = true;
        for (int i = 0; i < NB; i++) {
            h[i] = nullptr;
            if (!strcmp(v, "wc"))
                ok &= cudaHostAlloc(&h[i], bytes, cudaHostAllocWriteCombined) == cudaSuccess;
            else if (!strcmp(v, "thp")) {
                const size_t len = (bytes + (2u << 20) - 1) & ~((size_t)(2u << 20) - 1);
                void *p = mmap(nullptr, len + (2u << 20), PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
                if (p == MAP_FAILED) { ok = false; break; }
                p = (void *)(((uintptr_t)p + (2u << 20) - 1) & ~((uintptr_t)(2u << 20) - 1));
                madvise(p, len, MADV_HUGEPAGE);
                memset(p, 0, len);
                ok &= cudaHostRegister(p, len, cudaHostRegisterDefault) == cudaSuccess;
                h[i] = p;
            } else if (!strcmp(v, "sm"))
                ok &= cudaHostAlloc(&h[i], bytes, cudaHostAllocMapped) == cudaSuccess;
            else
                ok &= cudaHostAlloc(&h[i], bytes, cudaHostAllocDefault) == cudaSuccess;
        }
        if (!ok) {
            printf("dev %d %-6s allocation failed: %s\n", dev, v, cudaGetErrorString(cudaGetLastError()));
            t_next += window + 1.0;
            continue;
        }
        void *hd[NB];
        if (!strcmp(v, "sm"))
            for (int i = 0; i < NB; i++)
                cudaHostGetDevicePointer(&hd[i], h[i], 0);
        /* warm up, then wait for the common start */
        for (int i = 0; i < NB; i++)
            cudaMemcpyAsync(h[i], d[i], bytes, cudaMemcpyDeviceToHost, st[0]);
        cudaDeviceSynchronize();
        while (now() < t_next)
            usleep(200);
        const double t0 = now();
        size_t moved = 0, moved2 = 0;
        int k = 0;
        while (now() - t0 < window) {
            for (int r = 0; r < 4; r++, k++) {
                if (!strcmp(v, "sm"))
                    sm_copy<<<148 * 4, 256, 0, st[0]>>>((const uint4 *)d[k % NB], (uint4 *)hd[k % NB], bytes / 16);
                else if (!strcmp(v, "h2d"))
                    cudaMemcpyAsync(d[k % NB], h[k % NB], bytes, cudaMemcpyHostToDevice, st[0]);
                else
                    cudaMemcpyAsync(h[k % NB], d[k % NB], bytes, cudaMemcpyDeviceToHost, st[0]);
                moved += bytes;
                if (!strcmp(v, "both")) {
                    cudaMemcpyAsync(d[(k + 2) % NB], h[(k + 2) % NB], bytes / 2, cudaMemcpyHostToDevice, st[1]);
                    moved2 += bytes / 2;
                }
            }
            cudaStreamSynchronize(st[0]);
        }
        cudaDeviceSynchronize();
        const double dt = now() - t0;
        printf("dev %d %-6s %7.2f GB/s", dev, v, moved / dt / 1e9);
        if (moved2)
            printf("  (+ %6.2f GB/s host->device at the same time)", moved2 / dt / 1e9);
        printf("\n");
        fflush(stdout);
        for (int i = 0; i < NB; i++) {
            if (!strcmp(v, "thp"))
                cudaHostUnregister(h[i]);
            else
                cudaFreeHost(h[i]);
        }
        t_next += window + 1.0;
    }
    return 0;
}
