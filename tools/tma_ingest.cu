// Per-SM ingest rate of cp.async.bulk (global -> shared) on B200: one CTA per SM, one thread keeps R copies of `piece` bytes in
// flight into a ring and re-issues a stage as soon as it lands; nothing reads the data. Answers: is ~50 GB/s per SM (what k_gram_ws
// reaches on 120 SMs) a limit of the copy engine / fabric per SM, or of bytes in flight x latency?
//   nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/tma_ingest tools/tma_ingest.cu && /tmp/tma_ingest
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(32, 1) ingest(const char* __restrict__ src, size_t bytes_per_cta, int R, uint32_t piece, int* sink) {
    extern __shared__ __align__(128) char ring[];
    __shared__ __align__(8) uint64_t bar[16];
    if (threadIdx.x != 0) return;
    for (int i = 0; i < R; i++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const char* p = src + (size_t)blockIdx.x * bytes_per_cta;
    const long long n = (long long)(bytes_per_cta / piece);
    auto issue = [&](long long i) {
        const int st = (int)(i % R);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar[st])), "r"(piece) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(s32(ring + (size_t)st * piece)), "l"(p + i * piece), "r"(piece), "r"(s32(&bar[st])) : "memory");
    };
    for (long long i = 0; i < R && i < n; i++) issue(i);
    for (long long i = 0; i < n; i++) {
        const int st = (int)(i % R);
        const uint32_t par = (uint32_t)((i / R) & 1);
        uint32_t ok = 0;
        while (!ok)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(s32(&bar[st])), "r"(par) : "memory");
        if (i + R < n) issue(i + R);
    }
    if (sink) sink[blockIdx.x] = ring[0];
}
int main() {
    const size_t total = (size_t)17 << 30;
    char* src; cudaMalloc(&src, total); cudaMemset(src, 1, total);
    cudaFuncSetAttribute(ingest, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    struct { int R; uint32_t piece; } cfgs[] = {{4, 40960}, {8, 20480}, {2, 40960}, {3, 40960}, {1, 40960}, {4, 20480}, {2, 20480}, {16, 10240}, {9, 20480}};
    for (int grid : {148, 120, 74, 32, 8, 1})
        for (auto& c : cfgs) {
            const size_t per = (total / grid) / c.piece * c.piece > ((size_t)1 << 28) ? ((size_t)1 << 28) / c.piece * c.piece : (total / grid) / c.piece * c.piece;
            const size_t smem = (size_t)c.R * c.piece;
            float best = 1e30f;
            for (int rep = 0; rep < 3; rep++) {
                cudaEventRecord(e0);
                ingest<<<grid, 32, smem>>>(src, per, c.R, c.piece, nullptr);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
            }
            cudaError_t e = cudaGetLastError();
            printf("grid %3d  R %2d x %6u B (%3zu kB in flight): %8.1f GB/s total, %6.1f GB/s per SM  %s\n", grid, c.R, c.piece, smem >> 10,
                   per * grid / best / 1e6, per / best / 1e6, e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
    return 0;
}
