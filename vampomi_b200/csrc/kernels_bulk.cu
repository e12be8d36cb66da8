// Warp-specialised variants of the two matrix passes that move A through shared memory with the bulk-copy engine
// (cp.async.bulk global -> shared, completion on an mbarrier: SASS UBLKCP + SYNCS) instead of per-thread LDG:
//
//     one producer lane per CTA keeps a ring of STAGES shared-memory stages full; each stage is a handful of contiguous
//     8-16 KB pieces of columns, so the bytes in flight per SM (3 x 64 KB / 2 x 96 KB) no longer depend on how many
//     registers ptxas is willing to spend on outstanding loads; 8 consumer warps read the stage with conflict-free
//     128-bit shared loads, standardise on the fly and accumulate in registers; full/empty mbarriers hand the
//     stages back and forth. A is tagged L2::evict_first (touched once per pass) so the small reused vectors stay in L2.
//
// Selected with vampomi_set_tuning("ax_impl" / "atx_impl", 1); results are bitwise independent of scheduling
// (fixed work split, fixed reduction order) but differ from the LDG variants in summation order (~1e-16).
#include "common.h"

namespace vampomi {

namespace {

constexpr int CONSUMERS = 256;                 // 8 consumer warps
constexpr int THREADS = CONSUMERS + 32;        // + 1 producer warp

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    unsigned long long spins = 0;
    while (!mbar_try_wait(bar, parity))
        if (++spins > (1ull << 31)) __trap();      // a protocol bug must fault, not hang the GPU
}
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
}
__device__ __forceinline__ void bulk_g2s_keep(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ double warp_sum_b(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ void consumer_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(CONSUMERS) : "memory"); }

}  // namespace

// ---------------------------------------------------------------------------------------------------------------
// A^T p.  Stage = rows [s*SEG, (s+1)*SEG) of C columns + the same rows of p.  CTA b owns column groups [g0, g1).
// ---------------------------------------------------------------------------------------------------------------
template <int C, int SEG, int STAGES>
__global__ void __launch_bounds__(THREADS) k_atx_bulk(const double* __restrict__ A, size_t ld, const double* __restrict__ mave,
                                                      const double* __restrict__ msig, const double* __restrict__ p, long long M,
                                                      double scale, double* __restrict__ out, const int* __restrict__ done) {
    if (done != nullptr && *done != 0) return;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* stage_base = reinterpret_cast<double*>(smem_raw);               // [STAGES][(C+1)*SEG]
    __shared__ uint64_t full[STAGES], empty[STAGES];
    __shared__ double red[2][CONSUMERS / 32][C];
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], CONSUMERS / 32); }
        fence_mbar_init();
    }
    __syncthreads();

    const long long ngroups = (M + C - 1) / C;
    const long long per = (ngroups + gridDim.x - 1) / gridDim.x;
    const long long g0 = (long long)blockIdx.x * per;
    long long g1 = g0 + per;
    if (g1 > ngroups) g1 = ngroups;
    const int nseg = (int)((ld + SEG - 1) / SEG);

    if (tid >= CONSUMERS) {
        // ------------------------------------------------ producer ------------------------------------------------
        if (tid == CONSUMERS) {
            const uint64_t pol = policy_evict_first();
            int st = 0;
            uint32_t ph = 0;
            for (long long g = g0; g < g1; g++) {
                for (int s = 0; s < nseg; s++) {
                    mbar_wait(&empty[st], ph ^ 1);
                    const size_t r0 = (size_t)s * SEG;
                    const uint32_t rows = (uint32_t)((ld - r0) < (size_t)SEG ? (ld - r0) : (size_t)SEG);
                    const uint32_t bytes = rows * 8u;
                    double* sb = stage_base + (size_t)st * (C + 1) * SEG;
                    mbar_expect_tx(&full[st], (C + 1) * bytes);
#pragma unroll
                    for (int cc = 0; cc < C; cc++) {
                        long long j = g * C + cc < M ? g * C + cc : M - 1;
                        bulk_g2s(sb + (size_t)cc * SEG, A + (size_t)j * ld + r0, bytes, &full[st], pol);
                    }
                    bulk_g2s_keep(sb + (size_t)C * SEG, p + r0, bytes, &full[st]);
                    if (++st == STAGES) { st = 0; ph ^= 1; }
                }
            }
        }
        return;
    }
    // ---------------------------------------------------- consumers ----------------------------------------------------
    const int lane = tid & 31, wid = tid >> 5;
    int st = 0;
    uint32_t ph = 0;
    int par = 0;
    for (long long g = g0; g < g1; g++) {
        double m[C], acc[C][2];
#pragma unroll
        for (int cc = 0; cc < C; cc++) {
            long long j = g * C + cc < M ? g * C + cc : M - 1;
            m[cc] = __ldg(mave + j);
            acc[cc][0] = acc[cc][1] = 0.0;
        }
        for (int s = 0; s < nseg; s++) {
            mbar_wait(&full[st], ph);
            const size_t r0 = (size_t)s * SEG;
            const int nd2 = (int)(((ld - r0) < (size_t)SEG ? (ld - r0) : (size_t)SEG) >> 1);
            const double2* sb = reinterpret_cast<const double2*>(stage_base + (size_t)st * (C + 1) * SEG);
            const double2* sp = sb + (size_t)C * (SEG / 2);
#pragma unroll
            for (int k = 0; k < SEG / 2 / CONSUMERS; k++) {
                const int i2 = k * CONSUMERS + tid;                         // consecutive lanes -> consecutive 16 B: no bank conflict
                if (i2 < nd2) {
                    const double2 pv = sp[i2];
#pragma unroll
                    for (int cc = 0; cc < C; cc++) {
                        const double2 a = sb[(size_t)cc * (SEG / 2) + i2];
                        acc[cc][0] = fma(a.x - m[cc], pv.x, acc[cc][0]);       // (meth[i] - mu) * phen[i], src/data.cpp:304
                        acc[cc][1] = fma(a.y - m[cc], pv.y, acc[cc][1]);
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[st]);
            if (++st == STAGES) { st = 0; ph ^= 1; }
        }
#pragma unroll
        for (int cc = 0; cc < C; cc++) {
            const double sws = warp_sum_b(acc[cc][0] + acc[cc][1]);
            if (lane == 0) red[par][wid][cc] = sws;
        }
        consumer_bar_sync();
        if (tid < C) {
            double t = red[par][0][tid];
#pragma unroll
            for (int w = 1; w < CONSUMERS / 32; w++) t += red[par][w][tid];
            const long long j = g * C + tid;
            if (j < M) out[j] = (__ldg(msig + j) * t) * scale;                 // sigma_inv * dpa (:306), then * scale (:330)
        }
        par ^= 1;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// A x partials.  CTA (tile, chunk): rows [rbase, rbase + tile_rows) (tile_rows <= TR), columns [c0, c1).
// Stage = G consecutive columns of the tile.  Thread t owns the 16-byte pieces t, t+256, ... of the tile.
// ---------------------------------------------------------------------------------------------------------------
template <int TR, int G, int STAGES>
__global__ void __launch_bounds__(THREADS) k_ax_bulk(const double* __restrict__ A, size_t ld, const double* __restrict__ mave,
                                                     const double* __restrict__ msig, const double* __restrict__ x, int tile_rows,
                                                     int cols_per_chunk, long long M, double* __restrict__ partial,
                                                     const int* __restrict__ done) {
    if (done != nullptr && *done != 0) return;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* stage_base = reinterpret_cast<double*>(smem_raw);               // [STAGES][G][TR]
    __shared__ uint64_t full[STAGES], empty[STAGES];
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], CONSUMERS / 32); }
        fence_mbar_init();
    }
    __syncthreads();

    const size_t rbase = (size_t)blockIdx.x * tile_rows;
    if (rbase >= ld) return;
    size_t rows_sz = ld - rbase < (size_t)tile_rows ? ld - rbase : (size_t)tile_rows;
    const int rows = (int)rows_sz;                                          // multiple of 16
    const long long c0 = (long long)blockIdx.y * cols_per_chunk;
    long long c1 = c0 + cols_per_chunk;
    if (c1 > M) c1 = M;
    const long long nst = (c1 - c0 + G - 1) / G;

    if (tid >= CONSUMERS) {
        if (tid == CONSUMERS) {
            const uint64_t pol = policy_evict_first();
            int st = 0;
            uint32_t ph = 0;
            const uint32_t bytes = (uint32_t)rows * 8u;
            for (long long q = 0; q < nst; q++) {
                mbar_wait(&empty[st], ph ^ 1);
                const long long j0 = c0 + q * G;
                const int nc = (int)(c1 - j0 < G ? c1 - j0 : G);
                double* sb = stage_base + (size_t)st * G * TR;
                mbar_expect_tx(&full[st], (uint32_t)nc * bytes);
                for (int cc = 0; cc < nc; cc++)
                    bulk_g2s(sb + (size_t)cc * TR, A + (size_t)(j0 + cc) * ld + rbase, bytes, &full[st], pol);
                if (++st == STAGES) { st = 0; ph ^= 1; }
            }
        }
        return;
    }
    const int lane = tid & 31;
    constexpr int KP = TR / 2 / CONSUMERS;                                  // 16-byte pieces per thread per column
    double2 acc[KP];
#pragma unroll
    for (int k = 0; k < KP; k++) acc[k] = make_double2(0.0, 0.0);
    const int nd2 = rows >> 1;
    int st = 0;
    uint32_t ph = 0;
    for (long long q = 0; q < nst; q++) {
        const long long j0 = c0 + q * G;
        const int nc = (int)(c1 - j0 < G ? c1 - j0 : G);
        double m[G], w[G];
#pragma unroll
        for (int cc = 0; cc < G; cc++) {
            const long long j = j0 + cc < c1 ? j0 + cc : c1 - 1;
            m[cc] = __ldg(mave + j);
            w[cc] = __ldg(msig + j) * __ldg(x + j);                         // sig_phen_i, src/data.cpp:354
        }
        mbar_wait(&full[st], ph);
        const double2* sb = reinterpret_cast<const double2*>(stage_base + (size_t)st * G * TR);
#pragma unroll
        for (int cc = 0; cc < G; cc++) {
            if (cc < nc) {
#pragma unroll
                for (int k = 0; k < KP; k++) {
                    const int i2 = k * CONSUMERS + tid;
                    if (i2 < nd2) {
                        const double2 a = sb[(size_t)cc * (TR / 2) + i2];
                        acc[k].x = fma(a.x - m[cc], w[cc], acc[k].x);       // (meth[j] - ave) * sig_phen_i, src/data.cpp:360
                        acc[k].y = fma(a.y - m[cc], w[cc], acc[k].y);
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[st]);
        if (++st == STAGES) { st = 0; ph ^= 1; }
    }
    double2* prow = reinterpret_cast<double2*>(partial + (size_t)blockIdx.y * ld + rbase);
#pragma unroll
    for (int k = 0; k < KP; k++) {
        const int i2 = k * CONSUMERS + tid;
        if (i2 < nd2) prow[i2] = acc[k];
    }
}

// ---------------------------------------------------------------------------------------------------------------
// launch helpers (called from kernels_matrix.cu when the *_impl knobs select this path)
// ---------------------------------------------------------------------------------------------------------------
constexpr int ATX_C = 2, ATX_SEG = 1024, ATX_STAGES = 4;                    // 4 x 24 KB = 96 KB per CTA -> 2 CTAs / SM
constexpr int AX_TR = 1024, AX_G = 2, AX_STAGES = 4;                        // 4 x 16 KB = 64 KB per CTA -> 3 CTAs / SM

static int bulk_resident(const void* kernel, size_t smem) {
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, THREADS, smem) != cudaSuccess || n < 1) n = 1;
    return n;
}

int launch_atx_bulk(vampomi_ctx* c, const double* p_dev, double* out_dev, const int* done_flag) {
    if (c->storage != VAMPOMI_STORE_F64) { set_error("the bulk-copy kernels exist for FP64 storage only"); return VAMPOMI_ERR_STATE; }
    auto kern = k_atx_bulk<ATX_C, ATX_SEG, ATX_STAGES>;
    const size_t smem = (size_t)ATX_STAGES * (ATX_C + 1) * ATX_SEG * sizeof(double);
    if (!c->bulk_attr_atx) {                                                // per device: the context owns the flag
        VO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        c->bulk_attr_atx = true;
    }
    int per_sm = c->tune.atx_ctas_per_sm > 0 ? c->tune.atx_ctas_per_sm : bulk_resident((const void*)kern, smem);
    long long blocks = (long long)c->num_sms * per_sm;
    const long long ngroups = (c->M + ATX_C - 1) / ATX_C;
    if (blocks > ngroups) blocks = ngroups;
    const double scale = 1.0 / sqrt((double)c->N);
    kern<<<(unsigned)blocks, THREADS, smem, c->stream>>>(c->A, c->ld, c->mave, c->msig, p_dev, c->M, scale, out_dev, done_flag);
    VO_CUDA(cudaGetLastError());
    return VAMPOMI_OK;
}

// plans the (tile, chunk) grid for the bulk Ax kernel and launches it; returns the number of chunks written
int launch_ax_bulk(vampomi_ctx* c, const double* x_dev, const int* done_flag, int* nchunks_out) {
    if (c->storage != VAMPOMI_STORE_F64) { set_error("the bulk-copy kernels exist for FP64 storage only"); return VAMPOMI_ERR_STATE; }
    auto kern = k_ax_bulk<AX_TR, AX_G, AX_STAGES>;
    const size_t smem = (size_t)AX_STAGES * AX_G * AX_TR * sizeof(double);
    if (!c->bulk_attr_ax) {
        VO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        c->bulk_attr_ax = true;
    }
    const int ntiles = (int)((c->ld + AX_TR - 1) / AX_TR);
    size_t tr = (c->ld + ntiles - 1) / ntiles;
    const int tile_rows = (int)((tr + 15) / 16 * 16);
    int per_sm = c->tune.ax_ctas_per_sm > 0 ? c->tune.ax_ctas_per_sm : bulk_resident((const void*)kern, smem);
    long long slots = (long long)c->num_sms * per_sm;
    long long nch = slots / ntiles;
    if (nch < 1) nch = 1;
    if (nch > c->M) nch = c->M;
    const int cols_per_chunk = (int)((c->M + nch - 1) / nch);
    const int nchunks = (int)((c->M + cols_per_chunk - 1) / cols_per_chunk);
    size_t need = (size_t)nchunks * c->ld;
    if (need > c->ax_partial_elems) {
        VO_CUDA(cudaStreamSynchronize(c->stream));
        if (c->ax_partial) VO_CUDA(cudaFree(c->ax_partial));
        c->ax_partial = nullptr;
        VO_CUDA(cudaMalloc(&c->ax_partial, need * sizeof(double)));
        c->ax_partial_elems = need;
    }
    dim3 grid(ntiles, nchunks);
    kern<<<grid, THREADS, smem, c->stream>>>(c->A, c->ld, c->mave, c->msig, x_dev, tile_rows, cols_per_chunk, c->M, c->ax_partial,
                                             done_flag);
    VO_CUDA(cudaGetLastError());
    *nchunks_out = nchunks;
    return VAMPOMI_OK;
}

}  // namespace vampomi
