// TEST INFRASTRUCTURE ONLY — the `class data` adapter of INTEGRATION.md §2, compiled for real.
//
// The reference's own main_meth.cpp, vamp.cpp (+ vamp_probit.cpp), utilities.cpp and options.cpp are compiled unchanged (apart
// from the scripted oracle patches P1-P3, oracle/build_ref.py) and linked against THIS file instead of src/data.cpp: every
// member of `class data` (src/data.hpp:47-90) is defined here over the C ABI of libvampomi_cuda.so (include/vampomi.h), so the
// reference's VAMP loop runs on our operators — data::Ax (src/data.cpp:340-373), data::ATx (:315-333), data::pvals_loo
// (:385-417) — through the very seam its call sites use (src/vamp.cpp:232,303,508,518,519,653,654). Built by
// `python oracle/build_ref.py --gpu-data` into oracle/_ref/main_meth_ref_gpudata; needs a GPU to run (tests/test_gpu_adapter.py).
#include <cmath>
#include <cstdlib>
#include <iostream>
#include <map>
#include <string>
#include <vector>
#include <mpi.h>
#include "data.hpp"
#include "utilities.hpp"
#include "vampomi.h"
#include "vampomi_host.h"

namespace {
std::map<const data*, vampomi_ctx*> g_ctx;            // `class data` has no member to hang the context on: keyed by object
vampomi_ctx* ctx_of(const data* d) { return g_ctx.at(d); }
void ok(int rc, const char* what) {
    if (rc != VAMPOMI_OK) {
        std::cout << "FATAL: " << what << ": " << vampomi_last_error() << std::endl;     // abort-style, like check_mpi (src/utilities.cpp:21-46)
        exit(EXIT_FAILURE);
    }
}
}  // namespace

data::data(std::string phenfp, std::string methfp, std::string data_class, const int N, const int M, const int Mt, const int S, const int rank,
           double alpha_scale)
    : Mt(Mt), N(N), M(M), S(S), rank(rank), phenfp(phenfp), methfp(methfp), data_class(data_class), alpha_scale(alpha_scale) {
    read_phen(data_class != "bin_class");                                                // src/data.cpp:40-43
    read_methylation_data();
    compute_markers_statistics();
}

void data::read_phen(bool standardize) {                                                 // src/data.cpp:58-110
    phen_data.assign((size_t)N, 0.0);
    const long long n = vampomi_host_read_phen(phenfp.c_str(), standardize ? 1 : 0, phen_data.data(), N);
    if (n == -1) { std::cout << "FATAL: could not open phenotype file: " << phenfp << std::endl; exit(EXIT_FAILURE); }
    if (n == -2) throw "NAN in data!";
    if (n != N) { std::cout << "FATAL: phenotype file has " << n << " rows, N = " << N << std::endl; exit(EXIT_FAILURE); }
    nonas = (int)n; nas = 0;
}

void data::read_covariates(std::string, int) {}                                          // never called by the reference's main

void data::read_methylation_data() {                                                     // src/data.cpp:116-153
    int nranks = 1;                                                                      // the oracle build is single-rank (mpi.h stand-in)
    vampomi_ctx* c = nullptr;
    ok(vampomi_create(0, N, Mt, nranks, rank, &c), "vampomi_create");
    g_ctx[this] = c;
    ok(vampomi_load_file(c, methfp.c_str()), "vampomi_load_file");
}

void data::compute_markers_statistics() {                                                // src/data.cpp:233-283
    ok(vampomi_compute_stats(ctx_of(this), alpha_scale), "vampomi_compute_stats");
}

void data::compute_people_statistics() {}

std::vector<double> data::Ax(double* __restrict__ phen) {                                // src/data.cpp:340-373
    std::vector<double> out((size_t)N);
    ok(vampomi_ax(ctx_of(this), phen, out.data()), "vampomi_ax");
    return out;
}

std::vector<double> data::ATx(double* __restrict__ phen) {                               // src/data.cpp:315-333
    std::vector<double> out((size_t)M);
    ok(vampomi_atx(ctx_of(this), phen, out.data()), "vampomi_atx");
    return out;
}

double data::dot_product(const int, double* __restrict__, const double, const double) { return 0.0; }   // only data::ATx called it
std::vector<double> data::Zx(std::vector<double> phen) { return phen; }                                 // unused

std::vector<double> data::pvals_loo(std::vector<double> z1, std::vector<double> y, std::vector<double> x1_hat) {
    // src/data.cpp:385-417: per-marker sums on the GPU, t statistic and p-value with the reference's own helper
    std::vector<double> w((size_t)N), sums(3 * (size_t)M), p((size_t)M);
    double sw = 0, sww = 0;
    for (int i = 0; i < N; i++) { w[i] = y[i] - z1[i]; sw += w[i]; sww += w[i] * w[i]; }
    ok(vampomi_vec_set(ctx_of(this), VAMPOMI_V_USER_N1, w.data()), "vampomi_vec_set");
    ok(vampomi_loo_sums(ctx_of(this), VAMPOMI_V_USER_N1, sums.data()), "vampomi_loo_sums");
    for (int j = 0; j < M; j++) {
        const double c = x1_hat[j] / sqrt((double)N), sx = sums[3 * j], sxx = sums[3 * j + 1], sxw = sums[3 * j + 2];
        p[j] = linear_reg1d_pvals(sx, sxx, sxw + c * sxx, sw + c * sx, sww + 2 * c * sxw + c * c * sxx, N);
    }
    return p;
}
