// One-shot all-reduce over NVLink peer memory, fused into the kernels that produce the values to be summed.
//
// The reference sums its rank-local partial results with MPI_Allreduce: N doubles after every data::Ax
// (src/data.cpp:367) and single doubles after every inner product of the CG loop (src/utilities.cpp:151). At N = 20 000
// the vector is 160 kB and the scalars are 8-24 bytes: pure latency. Instead of a library collective after the kernel,
// the producing kernel itself exchanges the values:
//
//   push   every rank stores its contribution into slot [seq & 1][my_rank] of EVERY rank's receive area (peer-mapped
//          pointers, plain st.global over NVLink), fences at system scope and then raises a per-rank flag to `seq`
//   wait   it polls its own, LOCAL flags until all ranks have reached `seq` (ld.acquire.sys)
//   sum    it adds the G contributions from its local receive area in rank order 0..G-1 — every rank adds the same
//          values in the same order, so results are bitwise identical on all GPUs and independent of timing
//
// Vector exchanges (the N doubles after A x) use a flag-less form of the same idea by default (Xchg::ll): every double travels
// as two 8-byte words {low half | seq << 32} and {high half | seq << 32}. An 8-byte store is indivisible, so a word whose tag
// equals `seq` IS the data of this exchange: the receiver polls the words themselves and needs neither a system-scope fence
// behind the payload nor a flag — the per-CTA fence + flag round trip (~70 us for 1250 CTAs) becomes one NVLink store
// latency. Tags of the previous use of a slot (seq - 2) never match. Twice the bytes (320 kB instead of 160 kB): immaterial.
//
// Slot reuse is safe with two slots: a rank can only start exchange s+2 after it has passed the wait of s+1, and a peer
// raises its flag for s+1 only after it has finished reading slot (s & 1) of exchange s (stream order on that GPU).
// `seq` lives in device memory and advances only when an exchange really ran, so launches that return early at the CG
// done flag (identical on all ranks) do not desynchronise it. A peer that never arrives makes the wait give up after a
// wall-clock time-out and raise an error flag (XchgDeadline below) instead of hanging or trapping.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace vampomi {

constexpr int XCHG_MAX_RANKS = 8;
constexpr int XCHG_SCALARS = 64;        // doubles per scalar exchange (>= MAX_SUMS)
constexpr int XCHG_KMAX = 4;            // vectors one exchange can carry (multi-right-hand-side A x)

struct Xchg {                           // passed by value to kernels; all offsets identical on every rank
    int enabled;
    int G, rank;
    int ll;                             // 1 = tagged-word vector exchange (no fence, no flags), 0 = payload + fence + per-CTA flags
    int maxb;                           // CTA slots per rank in the vector flag array
    unsigned long long ld;              // stride between the contributions of two ranks (doubles): XCHG_KMAX vectors of length ld_vec
    unsigned long long off_flag_vec, off_flag_sc, off_recv_vec, off_recv_sc;   // byte offsets inside a region
    unsigned char* peer[XCHG_MAX_RANKS];   // base of every rank's region as mapped into THIS rank's address space
    unsigned int* seq;                  // local: [0] vector exchanges done, [1] scalar exchanges done
    unsigned int* ticket;               // local: CTA completion counter of the vector exchange kernel
    unsigned long long timeout_ns;      // how long a wait for a peer may last (by %globaltimer) before it gives up
    int* err;                           // host-mapped flag: set to 1 by a wait that gave up (the host turns it into an error code)
};

__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_volatile_u32(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// A peer that never arrives (a rank that died, or one that is more than the time-out behind) must not hang the GPU and must
// not poison the context either: the wait is bounded in WALL time (%globaltimer, checked every 1024 polls; the limit is
// Xchg::timeout_ns, 120 s unless VAMPOMI_XCHG_TIMEOUT_S says otherwise), and a wait that gives up raises the host-mapped
// error flag and returns — the kernel finishes with meaningless values, and the next host synchronisation of the library
// reports VAMPOMI_ERR_STATE instead of handing them out (capi.cu: xchg_check).
struct XchgDeadline {
    unsigned long long start = 0;
    unsigned int polls = 0;
    __device__ __forceinline__ bool expired(const Xchg& x) {
        if ((++polls & 1023u) != 0) return false;
        const unsigned long long now = globaltimer_ns();
        if (start == 0) { start = now; return false; }
        if (now - start < x.timeout_ns) return false;
        *reinterpret_cast<volatile int*>(x.err) = 1;
        return true;
    }
};
__device__ __forceinline__ void xchg_wait_flag(const Xchg& x, const unsigned int* flag, unsigned int seq) {
    XchgDeadline dl;
    // sequence numbers are monotonic; (int) difference tolerates 32-bit wrap
    while ((int)(ld_acquire_sys(flag) - seq) < 0)
        if (dl.expired(x)) return;
}

__device__ __forceinline__ double* xchg_recv_vec(const Xchg& x, int on_rank, unsigned slot, int from_rank) {
    return reinterpret_cast<double*>(x.peer[on_rank] + x.off_recv_vec) + ((size_t)slot * x.G + from_rank) * x.ld;
}
__device__ __forceinline__ unsigned int* xchg_flag_vec(const Xchg& x, int on_rank, int from_rank, int cta) {
    return reinterpret_cast<unsigned int*>(x.peer[on_rank] + x.off_flag_vec) + (size_t)from_rank * x.maxb + cta;
}
// tagged-word form: value i of rank `from_rank` occupies two u64 at the same place, i.e. 16 bytes per value
__device__ __forceinline__ unsigned long long* xchg_recv_ll(const Xchg& x, int on_rank, unsigned slot, int from_rank) {
    return reinterpret_cast<unsigned long long*>(x.peer[on_rank] + x.off_recv_vec) + ((size_t)slot * x.G + from_rank) * x.ld * 2;
}
__device__ __forceinline__ void xchg_ll_store(unsigned long long* dst, double v, unsigned int seq) {
    const unsigned long long bits = (unsigned long long)__double_as_longlong(v), tag = (unsigned long long)seq << 32;
    const unsigned long long lo = (bits & 0xffffffffull) | tag, hi = (bits >> 32) | tag;
    asm volatile("st.volatile.global.v2.u64 [%0], {%1,%2};" ::"l"(dst), "l"(lo), "l"(hi) : "memory");
}
__device__ __forceinline__ double xchg_ll_load(const Xchg& x, const unsigned long long* src, unsigned int seq) {
    unsigned long long lo, hi;
    XchgDeadline dl;
    for (;;) {
        asm volatile("ld.volatile.global.v2.u64 {%0,%1}, [%2];" : "=l"(lo), "=l"(hi) : "l"(src) : "memory");
        if ((unsigned int)(lo >> 32) == seq && (unsigned int)(hi >> 32) == seq) break;
        if (dl.expired(x)) break;                    // a rank is gone: flag raised, value meaningless
    }
    return __longlong_as_double((long long)((hi << 32) | (lo & 0xffffffffull)));
}
__device__ __forceinline__ double* xchg_recv_sc(const Xchg& x, int on_rank, unsigned slot, int from_rank) {
    return reinterpret_cast<double*>(x.peer[on_rank] + x.off_recv_sc) + ((size_t)slot * x.G + from_rank) * XCHG_SCALARS;
}
__device__ __forceinline__ unsigned long long* xchg_recv_sc_ll(const Xchg& x, int on_rank, unsigned slot, int from_rank) {
    return reinterpret_cast<unsigned long long*>(x.peer[on_rank] + x.off_recv_sc) + ((size_t)slot * x.G + from_rank) * XCHG_SCALARS * 2;
}
__device__ __forceinline__ unsigned int* xchg_flag_sc(const Xchg& x, int on_rank, int from_rank) {
    return reinterpret_cast<unsigned int*>(x.peer[on_rank] + x.off_flag_sc) + (size_t)from_rank * 32;   // one 128-byte line each
}

// Scalar all-reduce executed by ONE thread block (the block that finished a grid reduction last). `vals` holds K <= 64
// local sums in shared memory; on return out[k] = sum over ranks, identical on every rank. All threads of the block call.
__device__ inline void xchg_allreduce_scalars(const Xchg& x, const double* vals, int K, double* out) {
    const int tid = threadIdx.x;
    __shared__ unsigned int s_seq;
    if (tid == 0) s_seq = ld_volatile_u32(x.seq + 1) + 1u;
    __syncthreads();
    const unsigned int seq = s_seq, slot = seq & 1u;
    if (x.ll) {                                                             // tagged words: no fence, no flags
        if (tid < K) {
            const double v = vals[tid];
            for (int g = 0; g < x.G; g++) xchg_ll_store(xchg_recv_sc_ll(x, g, slot, x.rank) + 2 * tid, v, seq);
            double t = 0.0;
            for (int g = 0; g < x.G; g++) t += xchg_ll_load(x, xchg_recv_sc_ll(x, x.rank, slot, g) + 2 * tid, seq);
            out[tid] = t;
        }
        __syncthreads();
        if (tid == 0) x.seq[1] = seq;
        return;
    }
    if (tid < K) {
        const double v = vals[tid];
        for (int g = 0; g < x.G; g++) xchg_recv_sc(x, g, slot, x.rank)[tid] = v;
        __threadfence_system();
    }
    __syncthreads();
    if (tid < x.G) {
        st_release_sys(xchg_flag_sc(x, tid, x.rank), seq);                  // tell rank `tid` that my K values have landed
        xchg_wait_flag(x, xchg_flag_sc(x, x.rank, tid), seq);                  // and wait for rank `tid`'s values here
    }
    __syncthreads();
    if (tid < K) {
        double t = 0.0;
        for (int g = 0; g < x.G; g++) t += __ldcg(xchg_recv_sc(x, x.rank, slot, g) + tid);
        out[tid] = t;
    }
    if (tid == 0) x.seq[1] = seq;
}

}  // namespace vampomi
