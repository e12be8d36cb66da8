// Counter-based randomness of the VAMP path, shared by host and device code.
// Replaces the reference's std::random_device draws (src/vamp.hpp:51, src/vamp.cpp:295-296,
// src/vamp_probit.cpp:53,297-298) with a stateless hash of (seed, stream, a, b), so that results are reproducible
// and independent of how markers are sharded over GPUs. The same functions are restated in
// oracle/ref_shims/oracle_hooks.h (for the compiled reference) and oracle/vamp_oracle.py.
#pragma once
#include <stdint.h>
#include <math.h>

#ifdef __CUDACC__
#define VO_HD __host__ __device__ __forceinline__
#else
#define VO_HD static inline
#endif

namespace vampomi {

enum : uint64_t { STREAM_PROBE = 1, STREAM_P1 = 2, STREAM_MATRIX = 3 };

VO_HD uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

VO_HD uint64_t hash3(uint64_t seed, uint64_t stream, uint64_t a, uint64_t b) {
    uint64_t h = splitmix64(seed + 0x632BE59BD9B4E019ULL * stream);
    h = splitmix64(h ^ a);
    h = splitmix64(h ^ b);
    return h;
}

// +1 / -1 Hutchinson probe sign for VAMP iteration `it` and GLOBAL marker index g.
VO_HD double probe_sign(uint64_t seed, int it, uint64_t g) {
    return (hash3(seed, STREAM_PROBE, (uint64_t)it, g) >> 63) ? 1.0 : -1.0;
}

// N(0,1) by Box-Muller from one 64-bit hash (second uniform from one more splitmix step).
VO_HD double normal_from_hash(uint64_t h1) {
    uint64_t h2 = splitmix64(h1);
    double u1 = (double)((h1 >> 11) + 1) * 0x1.0p-53;    // (0, 1]
    double u2 = (double)(h2 >> 11) * 0x1.0p-53;          // [0, 1)
    return sqrt(-2.0 * log(u1)) * cos(6.283185307179586476925 * u2);
}

}  // namespace vampomi
