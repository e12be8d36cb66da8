// What does it cost one SM to bring 40 kB per step into shared memory AND read it back with LDS.128 — the data path of the fused pass
// (kernels_gram.cu) without its arithmetic and without its exchange?
//   mode 0: producer warp + bulk-copy ring (R stages of `step` bytes, full / empty mbarriers), 10 consumer warps read their rows
//   mode 1: no producer, no barriers: every consumer thread brings ITS OWN 16-byte pieces with cp.async.cg (LDGSTS) R steps ahead,
//           waits for its own groups and reads them back — the ring is private to the thread
//   mode 2: as 1, reading nothing back (the copy path alone);  mode 3: as 0, reading nothing back
// Prints GB/s per SM and SM cycles per 40 kB step for 8 CTAs (no HBM limit) and 120 CTAs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/_bin/smem_path_probe tools/smem_path_probe.cu && tools/_bin/smem_path_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void wait_par(uint32_t bar, uint32_t par) {
    uint32_t ok = 0;
    while (!ok)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(par) : "memory");
}
constexpr int NCW = 10, RP = 4, C = 2, R = 4, CT = NCW * 32, PIECE = CT * RP * 2;   // doubles per column piece: 2560
constexpr int STEP_BYTES = C * PIECE * 8;                                           // 40960

__global__ void __launch_bounds__((NCW + 1) * 32, 1) probe(const double* __restrict__ src, size_t col_stride, int nsteps, int mode, long long* cyc, double* sink) {
    extern __shared__ __align__(128) double ring[];              // [R][C][PIECE]
    __shared__ __align__(8) uint64_t full[R], empty[R];
    const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
    const double* base = src + (size_t)blockIdx.x * nsteps * C * col_stride;        // this CTA's columns
    if (tid == 0) {
        for (int i = 0; i < R; i++) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&full[i])));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&empty[i])), "r"(NCW));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const long long t0 = clock64();
    double acc = 0.0;
    if (mode == 0 || mode == 3) {
        if (wid == NCW) {
            if (lane == 0) {
                for (int s = 0; s < nsteps; s++) {
                    const int st = s % R;
                    if (s >= R) wait_par(s32(&empty[st]), (uint32_t)(((s / R) - 1) & 1));
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full[st])), "r"(STEP_BYTES) : "memory");
                    for (int cc = 0; cc < C; cc++)
                        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                     ::"r"(s32(ring + (size_t)(st * C + cc) * PIECE)), "l"(base + (size_t)(s * C + cc) * col_stride), "r"(PIECE * 8), "r"(s32(&full[st])) : "memory");
                }
            }
        } else {
            for (int s = 0; s < nsteps; s++) {
                const int st = s % R;
                wait_par(s32(&full[st]), (uint32_t)((s / R) & 1));
                if (mode == 0) {
#pragma unroll
                    for (int cc = 0; cc < C; cc++)
#pragma unroll
                        for (int i = 0; i < RP; i++) {
                            const double2 v = *reinterpret_cast<const double2*>(ring + (size_t)(st * C + cc) * PIECE + (i * CT + tid) * 2);
                            acc += v.x * v.y;
                        }
                }
                __syncwarp();
                if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&empty[st])) : "memory");
            }
        }
    } else if (wid < NCW) {
        auto issue = [&](int s) {
            const int st = s % R;
#pragma unroll
            for (int cc = 0; cc < C; cc++)
#pragma unroll
                for (int i = 0; i < RP; i++)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s32(ring + (size_t)(st * C + cc) * PIECE + (i * CT + tid) * 2)),
                                 "l"(base + (size_t)(s * C + cc) * col_stride + (i * CT + tid) * 2) : "memory");
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        for (int s = 0; s < R - 1 && s < nsteps; s++) issue(s);
        for (int s = 0; s < nsteps; s++) {
            if (s + R - 1 < nsteps) issue(s + R - 1); else asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group %0;" ::"n"(R - 1) : "memory");
            if (mode == 1) {
                const int st = s % R;
#pragma unroll
                for (int cc = 0; cc < C; cc++)
#pragma unroll
                    for (int i = 0; i < RP; i++) {
                        const double2 v = *reinterpret_cast<const double2*>(ring + (size_t)(st * C + cc) * PIECE + (i * CT + tid) * 2);
                        acc += v.x * v.y;
                    }
            }
        }
    }
    const long long t1 = clock64();
    if (blockIdx.x == 0 && tid == 0) *cyc = t1 - t0;
    if (acc == 1234.5678) sink[0] = acc;
}

int main() {
    const size_t N = 20000, M = 106250;
    double* src; cudaMalloc(&src, N * M * 8); cudaMemset(src, 0, N * M * 8);
    long long* dcyc; cudaMalloc(&dcyc, 8);
    double* sink; cudaMalloc(&sink, 8);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, R * STEP_BYTES);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int grid : {8, 120})
        for (int mode : {0, 1, 2, 3}) {
            int nsteps = (int)(M / C / grid);
            if (nsteps > 3000) nsteps = 3000;
            float best = 1e30f;
            for (int rep = 0; rep < 3; rep++) {
                cudaEventRecord(e0);
                probe<<<grid, (NCW + 1) * 32, R * STEP_BYTES>>>(src, N, nsteps, mode, dcyc, sink);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
            }
            long long cyc = 0;
            cudaMemcpy(&cyc, dcyc, 8, cudaMemcpyDeviceToHost);
            cudaError_t e = cudaGetLastError();
            printf("grid %3d mode %d: %7.1f GB/s per SM, %8.1f GB/s total, %7.1f cycles per 40 kB step (CTA 0)  %s\n", grid, mode, (double)nsteps * STEP_BYTES / best / 1e6,
                   (double)nsteps * STEP_BYTES * grid / best / 1e6, (double)cyc / nsteps, e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
    return 0;
}
