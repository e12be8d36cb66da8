// Follow-up of tools/tma_ingest.cu: what is the ~250 ns per bulk copy that caps one SM's ingest when ONE thread issues the copies
// (R = 4 x 40 kB: 158 GB/s, R = 8 x 20 kB: 81 GB/s, R = 16 x 10 kB: 40 GB/s — the same copies per second whatever their size)?
//   mode 0: W issuing warps (lane 0 of each), every warp with its own ring of R stages and its own slice of the CTA's bytes
//   mode 1: one issuing thread, 2-D tensor copies (cp.async.bulk.tensor.3d over a {16, N/16, M} view of a column-major matrix):
//           `cols` column pieces of `rows` rows per copy — what k_gram's step would fetch with ONE instruction instead of `cols`
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/_bin/tma_ingest2 tools/tma_ingest2.cu && tools/_bin/tma_ingest2
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void wait_par(uint32_t bar, uint32_t par) {
    uint32_t ok = 0;
    while (!ok)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(par) : "memory");
}

__global__ void __launch_bounds__(256, 1) ingest_w(const char* __restrict__ src, size_t bytes_per_warp, int R, uint32_t piece, long long* cyc) {
    extern __shared__ __align__(128) char ring[];
    __shared__ __align__(8) uint64_t bar[8][16];
    const int w = threadIdx.x >> 5, W = blockDim.x >> 5;
    if ((threadIdx.x & 31) != 0) return;
    for (int i = 0; i < R; i++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar[w][i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const char* p = src + ((size_t)blockIdx.x * W + w) * bytes_per_warp;
    char* myring = ring + (size_t)w * R * piece;
    const long long n = (long long)(bytes_per_warp / piece);
    auto issue = [&](long long i) {
        const int st = (int)(i % R);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar[w][st])), "r"(piece) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(s32(myring + (size_t)st * piece)), "l"(p + i * piece), "r"(piece), "r"(s32(&bar[w][st])) : "memory");
    };
    const long long t0 = clock64();
    for (long long i = 0; i < R && i < n; i++) issue(i);
    for (long long i = 0; i < n; i++) {
        wait_par(s32(&bar[w][i % R]), (uint32_t)((i / R) & 1));
        if (i + R < n) issue(i + R);
    }
    if (cyc && blockIdx.x == 0 && w == 0) *cyc = clock64() - t0;
}

__global__ void __launch_bounds__(32, 1) ingest_t(const __grid_constant__ CUtensorMap tm, int rows16, int cols, long long ncopies, int R, int tiles, long long* cyc) {
    extern __shared__ __align__(128) char ring[];
    __shared__ __align__(8) uint64_t bar[16];
    if (threadIdx.x != 0) return;
    for (int i = 0; i < R; i++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const uint32_t piece = (uint32_t)rows16 * 16u * 8u * (uint32_t)cols;
    const int tile = blockIdx.x % tiles;                          // row tile of this CTA (as the cluster rank in k_gram)
    const long long col0 = (long long)(blockIdx.x / tiles) * ncopies * cols;
    auto issue = [&](long long i) {
        const int st = (int)(i % R);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar[st])), "r"(piece) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     ::"r"(s32(ring + (size_t)st * piece)), "l"(&tm), "r"(0), "r"(tile * rows16), "r"((int)(col0 + i * cols)), "r"(s32(&bar[st])) : "memory");
    };
    const long long t0 = clock64();
    for (long long i = 0; i < R && i < ncopies; i++) issue(i);
    for (long long i = 0; i < ncopies; i++) {
        wait_par(s32(&bar[i % R]), (uint32_t)((i / R) & 1));
        if (i + R < ncopies) issue(i + R);
    }
    if (cyc && blockIdx.x == 0) *cyc = clock64() - t0;
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                             CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    const size_t N = 20000, M = 106250, total = N * M * 8;
    char* src; cudaMalloc(&src, total); cudaMemset(src, 1, total);
    long long* dcyc; cudaMalloc(&dcyc, 8);
    cudaFuncSetAttribute(ingest_w, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(ingest_t, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    printf("# mode 0: W issuing warps, each its own ring\n");
    struct { int W, R; uint32_t piece; } cfgs[] = {{1, 8, 20480}, {2, 4, 20480}, {4, 2, 20480}, {4, 4, 10240}, {8, 2, 10240}, {2, 2, 40960}, {1, 4, 40960}, {1, 5, 40960},
                                                    {1, 2, 81920}, {1, 32, 5120}, {4, 8, 5120}, {2, 4, 10240}, {1, 1, 163840}};
    for (int grid : {120, 8})
        for (auto& c : cfgs) {
            size_t per = (total / grid / c.W) / c.piece * c.piece;
            if (per > ((size_t)1 << 27)) per = ((size_t)1 << 27) / c.piece * c.piece;
            const size_t smem = (size_t)c.W * c.R * c.piece;
            float best = 1e30f; long long cyc = 0;
            for (int rep = 0; rep < 3; rep++) {
                cudaEventRecord(e0);
                ingest_w<<<grid, c.W * 32, smem>>>(src, per, c.R, c.piece, dcyc);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
            }
            cudaMemcpy(&cyc, dcyc, 8, cudaMemcpyDeviceToHost);
            cudaError_t e = cudaGetLastError();
            const double copies = (double)(per / c.piece) * c.W;
            printf("grid %3d  W %d x R %2d x %6u B (%3zu kB ring): %8.1f GB/s total, %6.1f GB/s per SM, %6.0f ns and %6.0f cycles per copy per SM  %s\n", grid, c.W, c.R, c.piece,
                   smem >> 10, per * c.W * grid / best / 1e6, per * c.W / best / 1e6, best * 1e6 / copies, cyc / copies, e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
    printf("# mode 1: one issuing thread, 3-D tensor copies of `cols` column pieces x (rows16 * 16) rows\n");
    EncodeFn encode = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres) != cudaSuccess || !encode) { printf("no cuTensorMapEncodeTiled\n"); return 0; }
    struct { int rows16, cols, R; } tc[] = {{157, 2, 4}, {157, 1, 8}, {157, 4, 2}, {157, 2, 3}, {157, 2, 2}, {157, 1, 4}, {78, 4, 4}, {157, 3, 3}};
    for (int grid : {120, 8})
        for (auto& c : tc) {
            CUtensorMap tm;
            cuuint64_t dims[3] = {16, N / 16, M};
            cuuint64_t strides[2] = {16 * 8, N * 8};
            cuuint32_t box[3] = {16, (cuuint32_t)c.rows16, (cuuint32_t)c.cols};
            cuuint32_t estr[3] = {1, 1, 1};
            CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, src, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { printf("encode failed %d for rows16 %d cols %d\n", (int)r, c.rows16, c.cols); continue; }
            const int tiles = 8;
            const int groups = grid / tiles;
            long long ncopies = (long long)(M / groups) / c.cols;
            if (ncopies > 3000) ncopies = 3000;
            const size_t piece = (size_t)c.rows16 * 128 * c.cols, smem = piece * c.R;
            float best = 1e30f; long long cyc = 0;
            for (int rep = 0; rep < 3; rep++) {
                cudaEventRecord(e0);
                ingest_t<<<grid, 32, smem>>>(tm, c.rows16, c.cols, ncopies, c.R, tiles, dcyc);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
            }
            cudaMemcpy(&cyc, dcyc, 8, cudaMemcpyDeviceToHost);
            cudaError_t e = cudaGetLastError();
            printf("grid %3d  box 16 x %3d x %d (%6zu B) R %d (%3zu kB ring): %8.1f GB/s total, %6.1f GB/s per SM, %6.0f ns and %6.0f cycles per copy  %s\n", grid, c.rows16, c.cols,
                   piece, c.R, smem >> 10, (double)piece * ncopies * grid / best / 1e6, (double)piece * ncopies / best / 1e6, best * 1e6 / ncopies, (double)cyc / ncopies,
                   e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
    return 0;
}
