#include "cov.h"
#include <cmath>
#include <fstream>
#include <iostream>
#include <limits>
#include <sstream>
#include "io.h"

namespace vampomi_host {

bool read_covariates(const std::string& path, int C, int N, std::vector<double>* Z, std::string* err) {
    Z->clear();
    if (C == 0) return true;                                                     // src/data.cpp:161-162
    std::ifstream covf(path);
    std::string line;
    int i = 0, rows = 0;
    while (std::getline(covf, line)) {
        i++;
        if (i == 1) continue;                                                    // header, :176
        // tokens between runs of whitespace; like the reference's regex split (:181), a line that STARTS with whitespace has an
        // empty first token, which then counts as the first of the two ids that are skipped (:184-185)
        std::vector<std::string> tok;
        size_t p = 0;
        if (!line.empty() && isspace((unsigned char)line[0])) tok.emplace_back();
        while (p < line.size()) {
            while (p < line.size() && isspace((unsigned char)line[p])) p++;
            size_t q = p;
            while (q < line.size() && !isspace((unsigned char)line[q])) q++;
            if (q > p) tok.push_back(line.substr(p, q - p));
            p = q;
        }
        int Cobs = 0;
        for (size_t t = 2; t < tok.size(); t++) {
            Z->push_back(std::stod(tok[t]));
            Cobs++;
        }
        if (Cobs != C) {                                                         // :192-195
            std::ostringstream ss;
            ss << "number of covariates = " << Cobs << " does not match to the specified number of covariates = " << C;
            *err = ss.str();
            return false;
        }
        rows++;
    }
    if (rows != N) {                                                             // the reference indexes covs[i] for i < N (:208)
        std::ostringstream ss;
        ss << "covariate file " << path << " has " << rows << " rows but --N is " << N;
        *err = ss.str();
        return false;
    }
    for (int c = 0; c < C; c++) {                                                // :205-226, long double sums
        long double cavg = 0.0, csig = 0.0;
        for (int r = 0; r < N; r++) cavg += (*Z)[(size_t)r * C + c];
        cavg = cavg / double(N);
        for (int r = 0; r < N; r++) csig += (((*Z)[(size_t)r * C + c] - cavg) * ((*Z)[(size_t)r * C + c] - cavg));
        csig = sqrtl(csig / double(N));
        for (int r = 0; r < N; r++) {
            double& v = (*Z)[(size_t)r * C + c];
            if (csig < 0.00000001) v = 0;
            else v = (double)((v - cavg) / csig);
        }
    }
    return true;
}

double erfcx_ref(double x) {
    if (x < -10.0) return std::numeric_limits<double>::infinity();
    if (x > 10.0) return std::numeric_limits<double>::lowest();
    return std::exp(x * x) * std::erfc(x);              // |x| <= 10: erfc(10) = 2e-45 is a normal number, the product is accurate
}

namespace {
double zdot(const std::vector<double>& Z, int C, int i, const std::vector<double>& eta) {
    double s = 0;
    for (int j = 0; j < C; j++) s += Z[(size_t)i * C + j] * eta[j];
    return s;
}
// src/vamp_probit.cpp:490-502 (probit_var = 1, src/vamp.hpp:35)
double mlogL_probit(const std::vector<double>& y, const std::vector<double>& gg, const std::vector<double>& Z, int N, int C, const std::vector<double>& eta) {
    double mlogL = 0;
    for (int i = 0; i < N; i++) {
        const double g_i = gg[i] + zdot(Z, C, i, eta);
        const double arg = (2 * y[i] - 1) / std::sqrt(1.0) * g_i;
        mlogL -= std::log(normal_cdf(arg));
    }
    return mlogL / N;
}
// :504-523
std::vector<double> grad_cov(const std::vector<double>& y, const std::vector<double>& gg, const std::vector<double>& Z, int N, int C, const std::vector<double>& eta) {
    std::vector<double> grad((size_t)C, 0.0);
    for (int j = 0; j < C; j++)
        for (int i = 0; i < N; i++) {
            const double g_i = gg[i] + zdot(Z, C, i, eta);
            const double arg = (2 * y[i] - 1) / std::sqrt(1.0) * g_i;
            const double ratio = 2.0 / std::sqrt(2 * M_PI) / erfcx_ref(-arg / std::sqrt(2.0));
            grad[j] += (-1) * ratio * (2 * y[i] - 1) / std::sqrt(1.0) * Z[(size_t)i * C + j];
        }
    for (int j = 0; j < C; j++) grad[j] /= N;
    return grad;
}
// partial-pivot LU in place; returns 0 unless a pivot is exactly zero (what the reference asks of uBLAS lu_factorize, :551-556)
int lu_factorize(std::vector<double>& m, std::vector<int>& pm, int n) {
    int singular = 0;
    for (int k = 0; k < n; k++) {
        int piv = k;
        double best = std::fabs(m[(size_t)k * n + k]);
        for (int i = k + 1; i < n; i++)
            if (std::fabs(m[(size_t)i * n + k]) > best) { best = std::fabs(m[(size_t)i * n + k]); piv = i; }
        pm[k] = piv;
        if (best == 0.0) { if (!singular) singular = k + 1; continue; }
        if (piv != k) for (int j = 0; j < n; j++) std::swap(m[(size_t)k * n + j], m[(size_t)piv * n + j]);
        for (int i = k + 1; i < n; i++) {
            m[(size_t)i * n + k] /= m[(size_t)k * n + k];
            for (int j = k + 1; j < n; j++) m[(size_t)i * n + j] -= m[(size_t)i * n + k] * m[(size_t)k * n + j];
        }
    }
    return singular;
}
void lu_substitute(const std::vector<double>& m, const std::vector<int>& pm, std::vector<double>& b, int n) {
    for (int k = 0; k < n; k++) if (pm[k] != k) std::swap(b[k], b[pm[k]]);
    for (int i = 0; i < n; i++) for (int j = 0; j < i; j++) b[i] -= m[(size_t)i * n + j] * b[j];
    for (int ii = n; ii-- > 0;) {
        for (int j = ii + 1; j < n; j++) b[ii] -= m[(size_t)ii * n + j] * b[j];
        b[ii] /= m[(size_t)ii * n + ii];
    }
}
double norm2(const std::vector<double>& v) { double s = 0; for (double x : v) s += x * x; return s; }
}  // namespace

std::vector<double> newton_method_cov(const std::vector<double>& y, const std::vector<double>& gg, const std::vector<double>& Z, int N, int C,
                                      std::vector<double> eta, bool verbose, int verbosity) {
    std::vector<double> eta_new;
    for (int it = 0; it <= 500; it++) {                                          // :531
        std::vector<double> lambda((size_t)N), W((size_t)N);
        for (int i = 0; i < N; i++) {                                            // :536-548
            const double g_i = gg[i] + zdot(Z, C, i, eta);
            const double arg = (2 * y[i] - 1) * g_i;
            const double ratio = 2.0 / std::sqrt(2 * M_PI) / erfcx_ref(-arg / std::sqrt(2.0));
            lambda[i] = ratio * (2 * y[i] - 1);
            W[i] = lambda[i] * (lambda[i] + g_i);
        }
        std::vector<double> XtWX((size_t)C * C, 0.0), RHS((size_t)C, 0.0);       // prod(Xtm, WXm), prod(Xtm, lambda): sums over i in order
        for (int a = 0; a < C; a++) {
            for (int b = 0; b < C; b++) {
                double s = 0;
                for (int i = 0; i < N; i++) s += Z[(size_t)i * C + a] * (Z[(size_t)i * C + b] * W[i]);
                XtWX[(size_t)a * C + b] = s;
            }
            double s = 0;
            for (int i = 0; i < N; i++) s += Z[(size_t)i * C + a] * lambda[i];
            RHS[a] = s;
        }
        std::vector<int> pm((size_t)C);
        if (lu_factorize(XtWX, pm, C) == 0) lu_substitute(XtWX, pm, RHS, C);     // :553-558
        else RHS.assign((size_t)C, 0.0);

        eta_new = eta;
        std::vector<double> displ((size_t)C, 0.0);
        const std::vector<double> grad = grad_cov(y, gg, Z, N, C, eta);
        double scale = 1;
        double init_val = mlogL_probit(y, gg, Z, N, C, eta);
        for (int i = 1; i < 300; i++) {                                          // backtracking, :567-582
            double dg = 0;
            for (int j = 0; j < C; j++) { displ[j] = scale * RHS[j]; eta_new[j] = eta[j] + displ[j]; }
            for (int j = 0; j < C; j++) dg += displ[j] * grad[j];
            const double curr_val = mlogL_probit(y, gg, Z, N, C, eta_new);
            if (curr_val <= init_val + dg / 2) {
                if (verbose) std::cout << "scale = " << scale << std::endl;
                break;
            }
            scale *= 0.9;
        }
        std::vector<double> diff = eta;
        for (int j = 0; j < C; j++) diff[j] -= eta_new[j];
        const double norm_eta = std::sqrt(norm2(eta));
        const double rel_err = norm_eta == 0 ? 1 : std::sqrt(norm2(diff)) / norm_eta;      // :587-593
        if (verbose && verbosity == 1) std::cout << "[Newton_cov] it = " << it << ", relative err = " << rel_err << std::endl;
        if (rel_err < 1e-4) {
            if (verbose) std::cout << "[Newton_cov] relative error <= 1e-4 - stoping criteria satisfied" << std::endl;
            break;
        }
        init_val = mlogL_probit(y, gg, Z, N, C, eta);                            // :604-615
        eta = eta_new;
        const double curr_val = mlogL_probit(y, gg, Z, N, C, eta);
        if (curr_val > init_val) {
            if (verbose)
                std::cout << "previous mlogL = " << init_val << ", current mlogL = " << curr_val << std::endl
                          << "likelihood value is not increasing -> terminating Newton-Raphson mehod" << std::endl;
            break;
        }
    }
    return eta;
}

}  // namespace vampomi_host
