// Host-side VAMP iteration logic — the counterpart of the reference's `class vamp` (src/vamp.hpp, src/vamp.cpp,
// src/vamp_probit.cpp). Everything that touches M- or N-length data is a call into the kernel ABI (vampomi.h);
// what stays here is the scalar algebra between those calls, the EM bookkeeping / component merging, and output.
#pragma once
#include <string>
#include <vector>
#include "../../../include/vampomi_host.h"

namespace vampomi_host {

// src/vamp.cpp:627-642 — merges mixture components with close variances, in place.
void merge_components(std::vector<double>& probs, std::vector<double>& vars, double thr);

class Vamp {
public:
    Vamp(vampomi_ctx* ctx, const vampomi_solver_config& cfg);
    // y: phenotype (length N). true_signal / x1hat_init: this shard's M values or nullptr.
    int init(const double* y, const double* true_signal, const double* x1hat_init);
    int step(vampomi_iter_result* res, double* x1_scaled, double* r1_scaled);
    // Covariates: Z is the standardised N x C matrix of read_covariates, row-major. Call before the first step:
    // iteration 1 then fits the covariate effects (Newton_method_cov) and removes them (linear: y -= Z cov_eff, src/vamp.cpp:155-169;
    // probit: offset m_cov = Z cov_eff inside the z-channel denoiser, src/vamp_probit.cpp:213-232).
    int set_covariates(int C, const double* Z);
    const std::vector<double>& cov_eff() const { return cov_eff_; }
    // checkpoint / resume (include/vampomi_host.h: vampomi_solver_save_state / _load_state)
    int save_state(const char* path);
    int load_state(const char* path);
    int iteration() const { return it_; }
    const std::vector<double>& probs() const { return probs_; }
    const std::vector<double>& vars() const { return vars_; }
    bool verbose = false;      // print the reference's per-iteration stdout lines (rank 0 of the CLI)
    int verbosity = 0;         // --verbosity 1 extras

private:
    int step_linear(vampomi_iter_result* res, double* x1_scaled, double* r1_scaled);
    int step_probit(vampomi_iter_result* res, double* x1_scaled, double* r1_scaled);
    int update_prior();
    int dump(double* x1_scaled, double* r1_scaled);
    int collect_dump();
    void fill_prior(vampomi_iter_result* res) const;

    vampomi_ctx* ctx_;
    vampomi_solver_config cfg_;
    int N_ = 0;
    long long M_ = 0, S_ = 0, Mt_ = 0;
    int it_ = 0;
    double gam1_, gamw_, gam2_ = 0, eta1_ = 0, eta2_ = 0, alpha1_ = 0, alpha2_ = 0, tau1_ = 0;
    std::vector<double> probs_, vars_;     // vars_ are the internal ones (x N, src/vamp.cpp:87-88)
    std::vector<double> y_host_, zbuf_;
    int C_ = 0;
    std::vector<double> Z_, cov_eff_;
    int fit_covariates();
    int apply_covariates();          // device vectors that follow from cov_eff_ (linear: adjusted y; probit: m_cov)
    int rank_ = 0, nranks_ = 1;
    double* pending_x1_ = nullptr;   // host buffers of read-outs begun by dump() and not yet collected
    double* pending_r1_ = nullptr;
    bool aty_ready_ = false;
    bool ata_x2_ready_ = false;   // VAMPOMI_V_ATA_X2 holds A^T A x2_hat of the previous iteration (fused schedule)
    long long passes_at_start_ = 0;
};

}  // namespace vampomi_host
