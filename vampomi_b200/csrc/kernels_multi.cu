// Multi-right-hand-side passes over the marker block: ONE read of A serves K vectors.
//
// The VAMP iteration contains pairs of matrix passes that are independent of each other: the LMMSE solve and the Onsager
// (trace) solve are two CG solves with the same operator tau*A^T A + gam2*I (src/vamp.cpp:308-311 and :494-501), and
// A x1_hat / A mu_start (:232,:681) as well as A x2_hat / A Q^-1 u (:508,:518) apply A to two known vectors. The
// reference runs them one after the other; since every pass is bound by streaming A from HBM, running them side by side on
// the same stream of A halves the bytes of those passes while each right-hand side keeps exactly its own arithmetic.
//
//   k_ax_multi        out_k[i] partials = sum_j (A[i,j]-mave[j]) * (msig[j] * x_k[j]),  k < K <= 4   (mirror of k_ax_partial)
//   k_ax_reduce_multi out_k[i] = (sum of the chunk partials [+ the other GPUs' over peer memory]) / sqrt(N)
//                                                                                   (mirror of k_ax_reduce[_xchg])
//   k_atx_smem        out_k[j] partials over a row tile = sum_i (A[i,j]-mave[j]) * p_k[i], K <= 2: the tile of the K vectors
//                     is staged in shared memory once per CTA and every warp streams its own columns down the tile —
//                     the default; 6.9 TB/s for two vectors, the same as the one-vector kernel
//   k_atx_tiled       the same with the K tiles of p held in registers (162 registers for K = 2, one CTA per SM, 4.7 TB/s:
//                     kept as knob multi_atx_impl = 0 for A/B)
//   k_atx_reduce      out_k[j] = msig[j] * (sum over row tiles) * (1/sqrt(N))
//
// A slot whose `done` flag is set (a CG solve that has already stopped) is skipped by every kernel, consistently on all GPUs.
// Grids are (row tile x column chunk) with the chunk count chosen so that whole waves of resident CTAs are filled
// (balanced_chunks, common.h).
#include "common.h"
#include "vec32.cuh"

namespace vampomi {

// ---------------------------------------------------------------------------------------------------------------
// OCC: CTAs per SM the register allocation is held to (0 = 2, or 1 for the largest tiles). OCC = 3 (80 registers, no spills
// for two vectors) trades registers for a third resident CTA: 96 instead of 64 KB of loads in flight per SM.
template <typename T, int K, int RV, int U, int OCC = 0>
__global__ void __launch_bounds__(256, (OCC > 0 ? OCC : (K * RV >= 4 ? 1 : 2))) k_ax_multi(const T* __restrict__ A, size_t ld, const double* __restrict__ mave,
                                                  const double* __restrict__ msig, MultiVec mv, int tile_rows, int cols_per_chunk,
                                                  long long M, double* __restrict__ partial, int nchunks) {
    constexpr int VE = V32<T>::VE;
    bool active[K];
    bool any = false;
#pragma unroll
    for (int k = 0; k < K; k++) { active[k] = mv.done[k] == nullptr || *mv.done[k] == 0; any |= active[k]; }
    if (!any) return;
    const int tid = threadIdx.x;
    const size_t rbase = (size_t)blockIdx.x * tile_rows;
    const long long c0 = (long long)blockIdx.y * cols_per_chunk;
    long long c1 = c0 + cols_per_chunk;
    if (c1 > M) c1 = M;
    const T* ap[RV];
    bool valid[RV];
    double acc[K][RV][VE];
#pragma unroll
    for (int rv = 0; rv < RV; rv++) {
        const int off = (rv * 256 + tid) * VE;
        valid[rv] = off < tile_rows && rbase + off < ld;
        ap[rv] = A + rbase + (valid[rv] ? off : 0);
#pragma unroll
        for (int k = 0; k < K; k++)
#pragma unroll
            for (int e = 0; e < VE; e++) acc[k][rv][e] = 0.0;
    }

    long long j = c0;
    for (; j + U <= c1; j += U) {
        V32<T> a[U][RV];
#pragma unroll
        for (int u = 0; u < U; u++)
#pragma unroll
            for (int rv = 0; rv < RV; rv++)
                if (valid[rv]) a[u][rv] = V32<T>::stream(ap[rv] + (size_t)(j + u) * ld);
#pragma unroll
        for (int u = 0; u < U; u++) {
            const double m = __ldg(mave + j + u), sg = __ldg(msig + j + u);
            double w[K];
#pragma unroll
            for (int k = 0; k < K; k++) w[k] = active[k] ? sg * __ldg(mv.in[k] + j + u) : 0.0;       // sig_phen_i, src/data.cpp:354
#pragma unroll
            for (int rv = 0; rv < RV; rv++) {
                if (valid[rv]) {
#pragma unroll
                    for (int e = 0; e < VE; e++) {
                        const double d = a[u][rv].val(e) - m;                                         // meth[j] - ave, src/data.cpp:360
#pragma unroll
                        for (int k = 0; k < K; k++) acc[k][rv][e] = fma(d, w[k], acc[k][rv][e]);
                    }
                }
            }
        }
    }
    for (; j < c1; j++) {
        const double m = __ldg(mave + j), sg = __ldg(msig + j);
        double w[K];
#pragma unroll
        for (int k = 0; k < K; k++) w[k] = active[k] ? sg * __ldg(mv.in[k] + j) : 0.0;
#pragma unroll
        for (int rv = 0; rv < RV; rv++) {
            if (valid[rv]) {
                V32<T> a = V32<T>::stream(ap[rv] + (size_t)j * ld);
#pragma unroll
                for (int e = 0; e < VE; e++) {
                    const double d = a.val(e) - m;
#pragma unroll
                    for (int k = 0; k < K; k++) acc[k][rv][e] = fma(d, w[k], acc[k][rv][e]);
                }
            }
        }
    }
#pragma unroll
    for (int k = 0; k < K; k++) {
        if (!active[k]) continue;
        double* prow = partial + ((size_t)k * nchunks + blockIdx.y) * ld + rbase;
#pragma unroll
        for (int rv = 0; rv < RV; rv++) {
            if (!valid[rv]) continue;
#pragma unroll
            for (int q = 0; q < VE / 4; q++)
                st256(prow + (rv * 256 + tid) * VE + 4 * q, d4{acc[k][rv][4 * q], acc[k][rv][4 * q + 1], acc[k][rv][4 * q + 2], acc[k][rv][4 * q + 3]});
        }
    }
}

// out_k[i] = (sum_c partial[k][c][i]) / divisor; blockIdx.y = k. With x.enabled the cross-GPU sum happens here (xchg.cuh).
template <int SL>
__global__ void __launch_bounds__(256) k_ax_reduce_multi(const double* __restrict__ partial, size_t ld, int nchunks, int N,
                                                         double divisor, const __grid_constant__ MultiVec mv, const __grid_constant__ Xchg x,
                                                         int use_xchg) {      // __grid_constant__: indexed straight from the constant bank, no stack copy
    __shared__ double sm[SL][256 / SL];
    __shared__ unsigned int s_seq;
    constexpr int ROWS = 256 / SL;
    static_assert(ROWS == 32, "one warp owns the CTA's rows in the exchange");
    const int k = blockIdx.y;
    // nothing to do for any vector (look-ahead launch of a finished solve): leave the exchange sequence untouched, exactly
    // like the single-vector kernel — the sequence number advances only for launches in which an exchange really runs
    bool any = false;
    for (int q = 0; q < mv.K; q++) any |= mv.done[q] == nullptr || *mv.done[q] == 0;
    if (!any) return;
    const bool active = mv.done[k] == nullptr || *mv.done[k] == 0;
    const int r = threadIdx.x % ROWS, s = threadIdx.x / ROWS;
    const int i = blockIdx.x * ROWS + r;
    if (use_xchg && threadIdx.x == 0) s_seq = ld_volatile_u32(x.seq) + 1u;
    double acc = 0.0;
    if (active && i < N) {
        const double* base = partial + (size_t)k * nchunks * ld;
        for (int cidx = s; cidx < nchunks; cidx += SL) acc += __ldcg(base + (size_t)cidx * ld + i);
    }
    sm[s][r] = acc;
    __syncthreads();
    if (s != 0) return;
    double t = sm[0][r];
#pragma unroll
    for (int q = 1; q < SL; q++) t += sm[q][r];
    double* out = mv.out[k];
    if (!use_xchg) {
        if (active && i < N) out[i] = t / divisor;
        return;
    }
    // slot k of this exchange: rows of rank g land at recv[slot][g][k*ld_vec + i]; flags are indexed by (k, CTA)
    const unsigned int seq = s_seq, slot = seq & 1u;
    const size_t koff = (size_t)k * (x.ld / XCHG_KMAX);
    const int cta = k * gridDim.x + blockIdx.x;
    if (active && x.ll) {                                     // tagged words: the data is its own arrival signal (xchg.cuh)
        if (i < N) {
            for (int g = 0; g < x.G; g++) xchg_ll_store(xchg_recv_ll(x, g, slot, x.rank) + 2 * (koff + i), t, seq);
            double tot = 0.0;
            for (int g = 0; g < x.G; g++) tot += xchg_ll_load(x, xchg_recv_ll(x, x.rank, slot, g) + 2 * (koff + i), seq);
            out[i] = tot / divisor;
        }
        __syncwarp();
    } else if (active) {
        if (i < N)
            for (int g = 0; g < x.G; g++) xchg_recv_vec(x, g, slot, x.rank)[koff + i] = t;
        __threadfence_system();
        __syncwarp();
        if (r < x.G) {
            st_release_sys(xchg_flag_vec(x, r, x.rank, cta), seq);
            xchg_wait_flag(x, xchg_flag_vec(x, x.rank, r, cta), seq);
        }
        __syncwarp();
        if (i < N) {
            double tot = 0.0;
            for (int g = 0; g < x.G; g++) tot += __ldcg(xchg_recv_vec(x, x.rank, slot, g) + koff + i);
            out[i] = tot / divisor;
        }
        __syncwarp();
    }
    if (r == 0) {                                             // the last CTA of the whole (rows x K) grid publishes the sequence number
        __threadfence();
        if (atomicAdd(x.ticket, 1u) == gridDim.x * gridDim.y - 1) { x.seq[0] = seq; *x.ticket = 0u; }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// A^T [p_0 .. p_{K-1}] over a row tile: thread t keeps its rows of all K vectors in registers; CB columns per step.
// ---------------------------------------------------------------------------------------------------------------
template <typename T, int K, int RV, int CB>
__global__ void __launch_bounds__(256) k_atx_tiled(const T* __restrict__ A, size_t ld, const double* __restrict__ mave, MultiVec mv,
                                                   int tile_rows, int cols_per_chunk, long long M, double* __restrict__ partial) {
    constexpr int VE = V32<T>::VE;
    bool active[K];
    bool any = false;
#pragma unroll
    for (int k = 0; k < K; k++) { active[k] = mv.done[k] == nullptr || *mv.done[k] == 0; any |= active[k]; }
    if (!any) return;
    __shared__ double red[2][8][K * CB];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const size_t rbase = (size_t)blockIdx.x * tile_rows;
    const long long c0 = (long long)blockIdx.y * cols_per_chunk;
    long long c1 = c0 + cols_per_chunk;
    if (c1 > M) c1 = M;

    const T* ap[RV];
    bool valid[RV];
    double pr[K][RV][VE];
#pragma unroll
    for (int rv = 0; rv < RV; rv++) {
        const int off = (rv * 256 + tid) * VE;
        valid[rv] = off < tile_rows && rbase + off < ld;
        ap[rv] = A + rbase + (valid[rv] ? off : 0);
#pragma unroll
        for (int k = 0; k < K; k++) {
#pragma unroll
            for (int e = 0; e < VE; e++) pr[k][rv][e] = 0.0;
            if (valid[rv] && active[k]) {
                PV<VE> pv = PV<VE>::load(mv.in[k] + rbase + off);          // pad rows of p are zero
#pragma unroll
                for (int e = 0; e < VE; e++) pr[k][rv][e] = pv.v[e];
            }
        }
    }
    // partial layout: [tile][k][M]
    double* pout = partial + (size_t)blockIdx.x * K * M;
    int par = 0;
    for (long long j = c0; j < c1; j += CB) {
        const int ncol = (int)(c1 - j < CB ? c1 - j : CB);
        V32<T> a[CB][RV];
        double m[CB];
#pragma unroll
        for (int cc = 0; cc < CB; cc++) {
            const long long jj = cc < ncol ? j + cc : j;
            m[cc] = __ldg(mave + jj);
#pragma unroll
            for (int rv = 0; rv < RV; rv++)
                if (valid[rv]) a[cc][rv] = V32<T>::stream(ap[rv] + (size_t)jj * ld);
        }
        double acc[K][CB];
#pragma unroll
        for (int k = 0; k < K; k++)
#pragma unroll
            for (int cc = 0; cc < CB; cc++) acc[k][cc] = 0.0;
#pragma unroll
        for (int cc = 0; cc < CB; cc++) {
#pragma unroll
            for (int rv = 0; rv < RV; rv++) {
                if (valid[rv]) {
#pragma unroll
                    for (int e = 0; e < VE; e++) {
                        const double d = a[cc][rv].val(e) - m[cc];                   // meth[i] - mu, src/data.cpp:304
#pragma unroll
                        for (int k = 0; k < K; k++) acc[k][cc] = fma(d, pr[k][rv][e], acc[k][cc]);
                    }
                }
            }
        }
#pragma unroll
        for (int k = 0; k < K; k++)
#pragma unroll
            for (int cc = 0; cc < CB; cc++) {
                const double sw = warp_sum(acc[k][cc]);
                if (lane == 0) red[par][wid][k * CB + cc] = sw;
            }
        __syncthreads();
        if (tid < K * CB) {
            const int k = tid / CB, cc = tid % CB;
            if (cc < ncol && active[k]) {
                double t = red[par][0][tid];
#pragma unroll
                for (int w = 1; w < 8; w++) t += red[par][w][tid];
                pout[(size_t)k * M + j + cc] = t;
            }
        }
        par ^= 1;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// A^T [p_0 .. p_{K-1}] over a row tile with the K tiles of p staged in SHARED memory: every warp walks its own C columns
// down the tile (U steps in flight), reads p with conflict-free LDS.128 — each piece of p serves C columns, so shared
// memory carries K/C bytes per byte of A — and needs one warp reduction per (vector, column) at the end of the tile:
// no block barrier in the loop, few registers (two or three CTAs per SM).
// ---------------------------------------------------------------------------------------------------------------
// Position (in doubles) of the q-th pair of the VE values that belong to 32-byte vector `v` of the tile: the 32 lanes of a warp
// read 32 consecutive vectors, so pair q of lane l sits at ((q*32 + l)*2) inside the block — consecutive lanes, consecutive
// 16-byte words, no bank conflicts (the natural layout would put lanes 32 or 64 bytes apart).
template <int VE>
__device__ __forceinline__ int sp_pos(int v, int q) { return (v >> 5) * (32 * VE) + ((q << 5) + (v & 31)) * 2; }

template <typename T, int K, int C, int U>
__global__ void __launch_bounds__(256, (C * U >= 8 ? 2 : C * U >= 4 ? 3 : 4)) k_atx_smem(const T* __restrict__ A, size_t ld, const double* __restrict__ mave, MultiVec mv,
                                                  int tile_rows, int cols_per_chunk, long long M, double* __restrict__ partial) {
    constexpr int VE = V32<T>::VE;
    extern __shared__ __align__(32) double sp[];                 // [K][sp_stride], swizzled (sp_pos)
    bool active[K];
    bool any = false;
#pragma unroll
    for (int k = 0; k < K; k++) { active[k] = mv.done[k] == nullptr || *mv.done[k] == 0; any |= active[k]; }
    if (!any) return;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const size_t rbase = (size_t)blockIdx.x * tile_rows;
    int rows = tile_rows;
    if (rbase + rows > ld) rows = (int)(ld - rbase);             // ld is a multiple of 16, so is tile_rows -> whole vectors
    const int sp_stride = (tile_rows + 32 * VE - 1) / (32 * VE) * (32 * VE);      // one vector of p, padded to whole 32-lane blocks
#pragma unroll
    for (int k = 0; k < K; k++)
        for (int i = tid * 4; i < rows; i += 256 * 4) {
            double2 lo = make_double2(0.0, 0.0), hi = lo;
            if (active[k]) { PV<4> pv = PV<4>::load(mv.in[k] + rbase + i); lo = make_double2(pv.v[0], pv.v[1]); hi = make_double2(pv.v[2], pv.v[3]); }
            const int v = i / VE, q0 = (i % VE) / 2;
            *reinterpret_cast<double2*>(sp + (size_t)k * sp_stride + sp_pos<VE>(v, q0)) = lo;
            *reinterpret_cast<double2*>(sp + (size_t)k * sp_stride + sp_pos<VE>(v, q0 + 1)) = hi;
        }
    __syncthreads();
    const long long c0 = (long long)blockIdx.y * cols_per_chunk;
    long long c1 = c0 + cols_per_chunk;
    if (c1 > M) c1 = M;
    const int nvec = rows / VE;
    const T* abase = A + rbase;
    double* pout = partial + (size_t)blockIdx.x * K * M;         // [tile][k][M]
    for (long long j0 = c0 + (long long)wid * C; j0 < c1; j0 += 8 * C) {
        const T* col[C];
        double m[C], acc[K][C][2];
#pragma unroll
        for (int cc = 0; cc < C; cc++) {
            const long long j = j0 + cc < c1 ? j0 + cc : c1 - 1;
            col[cc] = abase + (size_t)j * ld;
            m[cc] = __ldg(mave + j);
#pragma unroll
            for (int k = 0; k < K; k++) acc[k][cc][0] = acc[k][cc][1] = 0.0;
        }
        // U steps of 32 lanes x 32 bytes per column in flight; the last, partial group of steps is predicated rather than
        // peeled into a serial tail (a tile of 4000 rows is 31.25 steps: a peeled tail would expose the memory latency twice)
        for (int v0 = lane; v0 < nvec; v0 += 32 * U) {
            V32<T> a[U][C];
#pragma unroll
            for (int u = 0; u < U; u++) {
                if (v0 + 32 * u < nvec) {
#pragma unroll
                    for (int cc = 0; cc < C; cc++) a[u][cc] = V32<T>::stream(col[cc] + (size_t)VE * (v0 + 32 * u));
                }
            }
#pragma unroll
            for (int u = 0; u < U; u++) {
                if (v0 + 32 * u < nvec) {
#pragma unroll
                    for (int k = 0; k < K; k++) {
                        double pv[VE];
#pragma unroll
                        for (int q = 0; q < VE / 2; q++) {
                            const double2 t = *reinterpret_cast<const double2*>(sp + (size_t)k * sp_stride + sp_pos<VE>(v0 + 32 * u, q));
                            pv[2 * q] = t.x; pv[2 * q + 1] = t.y;
                        }
#pragma unroll
                        for (int cc = 0; cc < C; cc++)
#pragma unroll
                            for (int e = 0; e < VE; e++)              // (meth[i] - mu) * phen[i], src/data.cpp:304
                                acc[k][cc][e & 1] = fma(a[u][cc].val(e) - m[cc], pv[e], acc[k][cc][e & 1]);
                    }
                }
            }
        }
#pragma unroll
        for (int k = 0; k < K; k++)
#pragma unroll
            for (int cc = 0; cc < C; cc++) {
                const double sw = warp_sum(acc[k][cc][0] + acc[k][cc][1]);
                if (lane == 0 && active[k] && j0 + cc < c1) pout[(size_t)k * M + j0 + cc] = sw;
            }
    }
}

// out_k[j] = (msig[j] * sum_tiles partial[tile][k][j]) * scale; blockIdx.y = k
__global__ void __launch_bounds__(256) k_atx_reduce(const double* __restrict__ partial, int ntiles, int K, long long M,
                                                    const double* __restrict__ msig, double scale, MultiVec mv) {
    const int k = blockIdx.y;
    if (mv.done[k] != nullptr && *mv.done[k] != 0) return;
    double* out = mv.out[k];
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < M; j += (long long)gridDim.x * blockDim.x) {
        double t = 0.0;
        for (int tile = 0; tile < ntiles; tile++) t += __ldcg(partial + ((size_t)tile * K + k) * M + j);
        out[j] = (__ldg(msig + j) * t) * scale;                                      // sigma_inv * dpa (:306), then * scale (:330)
    }
}

// ---------------------------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------------------------
static int resident(const void* kernel) {
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, 256, 0) != cudaSuccess || n < 1) n = 1;
    return n;
}
static int ensure_buf(vampomi_ctx* c, double** buf, size_t* cap, size_t need) {
    if (need <= *cap) return VAMPOMI_OK;
    VO_CUDA(cudaStreamSynchronize(c->stream));
    if (*buf) VO_CUDA(cudaFree(*buf));
    *buf = nullptr; *cap = 0;
    VO_CUDA(cudaMalloc(buf, need * sizeof(double)));
    *cap = need;
    return VAMPOMI_OK;
}

int ensure_ax_partial(vampomi_ctx* c, size_t elems) { return ensure_buf(c, &c->ax_partial, &c->ax_partial_elems, elems); }

// out_k = (sum of the chunk partials [+ the other GPUs' over peer memory, or NCCL]) / sqrt(N) for the mv.K vectors
int launch_ax_reduce_multi(vampomi_ctx* c, int nchunks, const MultiVec& mv) {
    const int K = mv.K;
    int sp = prof_begin(c, 1, 0.0);
    const double sqrtN = sqrt((double)c->N);
    constexpr int SL = 8;
    const int rblocks = (c->N + (256 / SL) - 1) / (256 / SL);
    const bool fused = c->nranks > 1 && c->xchg.enabled;
    k_ax_reduce_multi<SL><<<dim3(rblocks, K), 256, 0, c->stream>>>(c->ax_partial, c->ld, nchunks, c->N,
                                                                    (c->nranks == 1 || fused) ? sqrtN : 1.0, mv, c->xchg, fused ? 1 : 0);
    VO_CUDA(cudaGetLastError());
    c->counters[0] += 1;
    if (c->nranks > 1 && !fused) {
        for (int k = 0; k < K; k++) {
            VO_CHECK(allreduce_inplace(c, mv.out[k], (size_t)c->N));
            VO_CHECK(launch_scale_div(c, mv.out[k], mv.out[k], sqrtN, c->N, mv.done[k]));
        }
    }
    prof_end(c, sp);
    return VAMPOMI_OK;
}

template <typename T, int K, int RV, int U, int OCC = 0>
static int ax_multi_launch(vampomi_ctx* c, const T* A, const MultiVec& mv) {
    constexpr int VE = V32<T>::VE;
    auto kern = k_ax_multi<T, K, RV, U, OCC>;
    const int cap = 256 * VE * RV;
    const int ntiles = (int)((c->ld + cap - 1) / cap);
    const size_t tr = (c->ld + ntiles - 1) / ntiles;
    const int tile_rows = (int)((tr + 15) / 16 * 16);
    const int per_sm = c->tune.ax_ctas_per_sm > 0 ? c->tune.ax_ctas_per_sm : resident((const void*)kern);
    const long long nch = balanced_chunks((long long)c->num_sms * per_sm, ntiles, c->M, 4 * U, c->tune.grid_balance != 0 && c->tune.ax_ctas_per_sm == 0);
    const int cols_per_chunk = (int)((c->M + nch - 1) / nch);
    const int nchunks = (int)((c->M + cols_per_chunk - 1) / cols_per_chunk);
    VO_CHECK(ensure_buf(c, &c->ax_partial, &c->ax_partial_elems, (size_t)K * nchunks * c->ld));
    if (c->prof_pending.size() > 8192) VO_CHECK(prof_resolve(c));
    int sp = prof_begin(c, 0, (double)c->M * c->N * (double)c->elem_bytes);
    kern<<<dim3(ntiles, nchunks), 256, 0, c->stream>>>(A, c->ld, c->mave, c->msig, mv, tile_rows, cols_per_chunk, c->M, c->ax_partial, nchunks);
    prof_end(c, sp);
    VO_CUDA(cudaGetLastError());
    c->counters[0] += 1; c->counters[1]++; c->counters[2] += (long long)c->M * c->N * c->elem_bytes;
    return launch_ax_reduce_multi(c, nchunks, mv);
}

// tile shape (32-byte vectors per thread per column, columns in flight): knobs multi_ax_rv / multi_ax_unroll, 0 = default
template <typename T, int K>
static int ax_multi_t(vampomi_ctx* c, const T* A, const MultiVec& mv) {
    int rv = c->tune.multi_ax_rv, u = c->tune.multi_ax_unroll;
    if (rv == 0) rv = 1;
    if (u == 0) u = sizeof(T) == 8 ? 4 : 2;
    if (rv == 1 && u == 4 && c->tune.multi_ax_occ == 3) return ax_multi_launch<T, K, 1, 4, 3>(c, A, mv);
    switch (rv * 10 + u) {
        case 12: return ax_multi_launch<T, K, 1, 2>(c, A, mv);
        case 14: return ax_multi_launch<T, K, 1, 4>(c, A, mv);
        case 18: return ax_multi_launch<T, K, 1, 8>(c, A, mv);
        case 22: return ax_multi_launch<T, K, 2, 2>(c, A, mv);
        case 24: return ax_multi_launch<T, K, 2, 4>(c, A, mv);
        default: set_error("ax_multi: unsupported tile shape rv=%d unroll=%d", rv, u); return VAMPOMI_ERR_ARG;
    }
}

int launch_ax_multi(vampomi_ctx* c, const MultiVec& mv) {
    if (mv.K < 1 || mv.K > XCHG_KMAX) { set_error("ax_multi: 1..%d vectors", XCHG_KMAX); return VAMPOMI_ERR_ARG; }
    if (c->storage == 1) {
        switch (mv.K) { case 1: return ax_multi_t<float, 1>(c, c->A32, mv); case 2: return ax_multi_t<float, 2>(c, c->A32, mv);
                        case 3: return ax_multi_t<float, 3>(c, c->A32, mv); default: return ax_multi_t<float, 4>(c, c->A32, mv); }
    }
    switch (mv.K) { case 1: return ax_multi_t<double, 1>(c, c->A, mv); case 2: return ax_multi_t<double, 2>(c, c->A, mv);
                    case 3: return ax_multi_t<double, 3>(c, c->A, mv); default: return ax_multi_t<double, 4>(c, c->A, mv); }
}

static int atx_reduce_launch(vampomi_ctx* c, int ntiles, int K, const MultiVec& mv) {
    long long rb = (c->M + 255) / 256;
    if (rb > 4 * c->num_sms) rb = 4 * c->num_sms;
    k_atx_reduce<<<dim3((unsigned)rb, K), 256, 0, c->stream>>>(c->atx_partial, ntiles, K, c->M, c->msig, 1.0 / sqrt((double)c->N), mv);
    VO_CUDA(cudaGetLastError());
    c->counters[0] += 2; c->counters[1]++; c->counters[2] += (long long)c->M * c->N * c->elem_bytes;
    return VAMPOMI_OK;
}

// register-tiled form (knob multi_atx_impl = 0)
template <typename T, int K>
static int atx_tiled_launch(vampomi_ctx* c, const T* A, const MultiVec& mv) {
    constexpr int VE = V32<T>::VE;
    constexpr int RV = sizeof(T) == 8 ? 2 : 1;
    constexpr int CB = 4;
    auto kern = k_atx_tiled<T, K, RV, CB>;
    const int cap = 256 * VE * RV;
    const int ntiles = (int)((c->ld + cap - 1) / cap);
    const size_t tr = (c->ld + ntiles - 1) / ntiles;
    const int tile_rows = (int)((tr + 15) / 16 * 16);
    const int per_sm = c->tune.atx_ctas_per_sm > 0 ? c->tune.atx_ctas_per_sm : resident((const void*)kern);
    long long nch = (long long)c->num_sms * per_sm / ntiles;
    if (nch < 1) nch = 1;
    if (nch > c->M) nch = c->M;
    int cols_per_chunk = (int)((c->M + nch - 1) / nch);
    cols_per_chunk = (cols_per_chunk + CB - 1) / CB * CB;
    const int nchunks = (int)((c->M + cols_per_chunk - 1) / cols_per_chunk);
    VO_CHECK(ensure_buf(c, &c->atx_partial, &c->atx_partial_elems, (size_t)ntiles * K * c->M));
    if (c->prof_pending.size() > 8192) VO_CHECK(prof_resolve(c));
    int sp = prof_begin(c, 2, (double)c->M * c->N * (double)c->elem_bytes);
    kern<<<dim3(ntiles, nchunks), 256, 0, c->stream>>>(A, c->ld, c->mave, mv, tile_rows, cols_per_chunk, c->M, c->atx_partial);
    VO_CUDA(cudaGetLastError());
    int rc = atx_reduce_launch(c, ntiles, K, mv);
    prof_end(c, sp);
    return rc;
}

// shared-memory form (knob multi_atx_impl = 1, the default): rows per tile from knob multi_atx_tile (0 = 2048)
template <typename T, int K, int C, int U>
static int atx_smem_launch(vampomi_ctx* c, const T* A, const MultiVec& mv) {
    constexpr int VE = V32<T>::VE;
    auto kern = k_atx_smem<T, K, C, U>;
    int want = c->tune.multi_atx_tile > 0 ? c->tune.multi_atx_tile : 2048;
    want = (want + 32 * VE - 1) / (32 * VE) * (32 * VE);
    const int ntiles = (int)((c->ld + want - 1) / want);
    const size_t tr = (c->ld + ntiles - 1) / ntiles;
    const int tile_rows = (int)((tr + 15) / 16 * 16);
    const int sp_stride = (tile_rows + 32 * VE - 1) / (32 * VE) * (32 * VE);
    const size_t smem = (size_t)K * sp_stride * sizeof(double);
    if (smem > 48 * 1024) {   // above the default limit the opt-in attribute is needed: set once per (device, instantiation, size)
        static thread_local int attr_dev[16];
        static thread_local const void* attr_fn[16] = {};
        static thread_local size_t attr_bytes[16] = {};
        int slot = -1;
        for (int i = 0; i < 16; i++) if (attr_fn[i] == (const void*)kern && attr_dev[i] == c->device) { slot = i; break; }
        if (slot < 0) for (int i = 0; i < 16; i++) if (attr_fn[i] == nullptr) { slot = i; break; }
        if (slot < 0) slot = 0;
        if (attr_fn[slot] != (const void*)kern || attr_dev[slot] != c->device || attr_bytes[slot] < smem) {
            VO_CUDA(cudaFuncSetAttribute((const void*)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            attr_fn[slot] = (const void*)kern; attr_dev[slot] = c->device; attr_bytes[slot] = smem;
        }
    }
    int per_sm = c->tune.atx_ctas_per_sm;
    if (per_sm <= 0) {
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void*)kern, 256, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    }
    const long long nch = balanced_chunks((long long)c->num_sms * per_sm, ntiles, c->M, 8 * C, c->tune.grid_balance != 0 && c->tune.atx_ctas_per_sm == 0);
    int cols_per_chunk = (int)((c->M + nch - 1) / nch);
    const int nchunks = (int)((c->M + cols_per_chunk - 1) / cols_per_chunk);
    VO_CHECK(ensure_buf(c, &c->atx_partial, &c->atx_partial_elems, (size_t)ntiles * K * c->M));
    if (c->prof_pending.size() > 8192) VO_CHECK(prof_resolve(c));
    int sp = prof_begin(c, 2, (double)c->M * c->N * (double)c->elem_bytes);
    kern<<<dim3(ntiles, nchunks), 256, smem, c->stream>>>(A, c->ld, c->mave, mv, tile_rows, cols_per_chunk, c->M, c->atx_partial);
    VO_CUDA(cudaGetLastError());
    int rc = atx_reduce_launch(c, ntiles, K, mv);
    prof_end(c, sp);
    return rc;
}

template <typename T, int K>
static int atx_multi_t(vampomi_ctx* c, const T* A, const MultiVec& mv) {
    if (c->tune.multi_atx_impl == 0) return atx_tiled_launch<T, K>(c, A, mv);
    int cc = c->tune.multi_atx_cols, u = c->tune.multi_atx_unroll;
    if (cc == 0) cc = K >= 2 ? 2 : 1;                   // measured defaults (profiles/r01_sweep_multi_vector_kernels*.jsonl)
    if (u == 0) u = 4;
    switch (cc * 10 + u) {
        case 12: return atx_smem_launch<T, K, 1, 2>(c, A, mv);
        case 14: return atx_smem_launch<T, K, 1, 4>(c, A, mv);
        case 22: return atx_smem_launch<T, K, 2, 2>(c, A, mv);
        case 24: return atx_smem_launch<T, K, 2, 4>(c, A, mv);
        case 42: return atx_smem_launch<T, K, 4, 2>(c, A, mv);
        default: set_error("atx_multi: unsupported shape cols=%d unroll=%d", cc, u); return VAMPOMI_ERR_ARG;
    }
}

int launch_atx_multi(vampomi_ctx* c, const MultiVec& mv) {
    if (mv.K < 1 || mv.K > 2) { set_error("atx_multi: 1 or 2 vectors"); return VAMPOMI_ERR_ARG; }
    if (c->storage == 1) return mv.K == 1 ? atx_multi_t<float, 1>(c, c->A32, mv) : atx_multi_t<float, 2>(c, c->A32, mv);
    return mv.K == 1 ? atx_multi_t<double, 1>(c, c->A, mv) : atx_multi_t<double, 2>(c, c->A, mv);
}

}  // namespace vampomi
