// Covariates of the probit / linear VAMP loops: data::read_covariates (src/data.cpp:159-227) and
// vamp::Newton_method_cov with its helpers (src/vamp_probit.cpp:490-617). N x C host work, once per run (iteration 1).
#pragma once
#include <string>
#include <vector>

namespace vampomi_host {

// Reads the covariate file (header line, then per individual: two ids and C values, whitespace separated) and standardises
// every covariate with the POPULATION standard deviation (a constant covariate becomes all zeros). Z is row-major N x C.
// Returns false and fills `err` with the reference's FATAL text when a row does not hold C values or the row count is not N.
bool read_covariates(const std::string& path, int C, int N, std::vector<double>* Z, std::string* err);

// erfcx with the reference's clamps (src/utilities.cpp:293-298): x < -10 -> +inf, x > 10 -> lowest().
double erfcx_ref(double x);

// Probit regression of y on the covariates by Newton-Raphson with backtracking (src/vamp_probit.cpp:525-617); gg = genetic
// predictor (all zeros at iteration 1, where the loops call it), eta = start. Prints the reference's progress lines when verbose.
std::vector<double> newton_method_cov(const std::vector<double>& y, const std::vector<double>& gg, const std::vector<double>& Z, int N, int C,
                                      std::vector<double> eta, bool verbose, int verbosity);

}  // namespace vampomi_host
