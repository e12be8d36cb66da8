// Round trip of the fused pass's exchange with nothing else in the way: a cluster of CS CTAs, one warp each; per step every CTA sends
// CK 8-byte values to every CTA of the cluster with st.async (data + 8 bytes on the receiver's mbarrier), waits until its own
// CS * CK values of that step have arrived, reads them and re-arms the slot — the four-slot protocol of k_gram_wsx's communication
// warp without the sums before and the weights after. Prints SM cycles per step: the floor of the exchange latency E.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/_bin/stas_probe tools/stas_probe.cu && tools/_bin/stas_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa(uint32_t a, uint32_t r) { uint32_t o; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(o) : "r"(a), "r"(r)); return o; }

template <int CS, int CK>
__global__ void __launch_bounds__(32, 1) xchg(int steps, int mode, long long* cyc, double* sink) {
    __shared__ __align__(16) double xbuf[4][CS][CK];
    __shared__ __align__(8) uint64_t full[4];
    const int lane = threadIdx.x;
    uint32_t crank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
    if (lane == 0) {
        for (int i = 0; i < 4; i++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&full[i])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int i = 0; i < 4; i++) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full[i])), "r"(CS * CK * 8) : "memory");
    }
    __syncwarp();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    constexpr int LPV = 32 / CK;
    const int ck = lane / LPV, dst = lane % LPV;
    const uint32_t rx = mapa(s32(&xbuf[0][crank][ck]), dst < CS ? dst : 0), rf = mapa(s32(&full[0]), dst < CS ? dst : 0);
    double acc = 0.0, v = 1.0 + lane;
    const long long t0 = clock64();
    uint32_t ph = 0;
    for (int s0 = 0; s0 < steps; s0 += 4, ph ^= 1u) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
            if (mode == 1) {                                     // a butterfly over the lanes before sending, as warp_sum_multi does
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            }
            if (dst < CS)
                asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];"
                             ::"r"(rx + u * (uint32_t)sizeof(xbuf[0])), "l"(__double_as_longlong(v)), "r"(rf + 8 * u) : "memory");
            uint32_t ok = 0;
            while (!ok)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(ok) : "r"(s32(&full[u])), "r"(ph) : "memory");
            double t = 0.0;
#pragma unroll
            for (int r = 0; r < CS; r++) t += xbuf[u][r][ck];
            acc += t;
            v = t * 1e-3 + 1.0;
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full[u])), "r"(CS * CK * 8) : "memory");
        }
    }
    const long long t1 = clock64();
    if (blockIdx.x == 0 && lane == 0) *cyc = t1 - t0;
    if (sink && acc == 12345.678) sink[0] = acc;
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int CS, int CK>
void run(int clusters, int mode, long long* dcyc) {
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.gridDim = dim3(CS * clusters); cfg.blockDim = dim3(32); cfg.attrs = attr; cfg.numAttrs = 1;
    const int steps = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0);
        cudaLaunchKernelEx(&cfg, xchg<CS, CK>, steps, mode, dcyc, (double*)nullptr);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    long long cyc = 0;
    cudaMemcpy(&cyc, dcyc, 8, cudaMemcpyDeviceToHost);
    cudaError_t e = cudaGetLastError();
    printf("cluster %2d, %d values per CTA and step, %2d clusters, mode %d: %7.1f ns and %7.1f cycles per step  %s\n", CS, CK, clusters, mode, best * 1e6 / steps,
           (double)cyc / steps, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
    long long* dcyc; cudaMalloc(&dcyc, 8);
    for (int mode : {0, 1}) {
        run<8, 4>(1, mode, dcyc); run<8, 4>(15, mode, dcyc); run<8, 2>(15, mode, dcyc); run<4, 4>(15, mode, dcyc); run<2, 4>(15, mode, dcyc); run<1, 4>(15, mode, dcyc);
    }
    return 0;
}
