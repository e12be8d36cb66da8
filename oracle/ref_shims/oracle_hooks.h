// TEST INFRASTRUCTURE ONLY — deterministic, sharding-invariant replacements for the reference's
// std::random_device draws (SURVEY.md §8c patches P2/P3). Force-included (-include) when building oracle/_ref.
//   P2: Hutchinson probe sign for (VAMP iteration it, global marker index g)   [vamp.cpp:296, vamp_probit.cpp:298]
//   P3: probit start vector p1 ~ N(0,1) from (seed, sample index)             [vamp_probit.cpp:53]
// The same counter hash is implemented by the product (vampomi_b200/csrc/rng.h) and by oracle/vamp_oracle.py.
#pragma once
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <vector>
static inline uint64_t vo_splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
static inline uint64_t vo_hash3(uint64_t seed, uint64_t stream, uint64_t a, uint64_t b) {
    uint64_t h = vo_splitmix64(seed + 0x632BE59BD9B4E019ULL * stream);
    h = vo_splitmix64(h ^ a);
    h = vo_splitmix64(h ^ b);
    return h;
}
static inline uint64_t vo_seed() {
    const char* s = getenv("VAMPOMI_SEED");
    return s ? strtoull(s, nullptr, 10) : 0ULL;
}
static inline double vampomi_oracle_probe(int it, long g) {          // returns +1 or -1
    return (vo_hash3(vo_seed(), 1, (uint64_t)it, (uint64_t)g) >> 63) ? 1.0 : -1.0;
}
static inline std::vector<double> vampomi_oracle_p1(int N) {
    std::vector<double> p(N);
    uint64_t seed = vo_seed();
    for (int i = 0; i < N; i++) {
        uint64_t h1 = vo_hash3(seed, 2, (uint64_t)i, 0), h2 = vo_splitmix64(h1);
        double u1 = (double)((h1 >> 11) + 1) * 0x1.0p-53, u2 = (double)(h2 >> 11) * 0x1.0p-53;
        p[i] = std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586476925 * u2);
    }
    return p;
}
