// TEST INFRASTRUCTURE ONLY — partial-pivot LU standing in for ublas lu_factorize / lu_substitute
// (used only by the covariate Newton solver, which is outside the hot path).
#pragma once
#include <cmath>
#include <utility>
#include "matrix.hpp"
namespace boost { namespace numeric { namespace ublas {
template <class T> class permutation_matrix {
public:
    explicit permutation_matrix(std::size_t n) : p_(n) { for (std::size_t i = 0; i < n; i++) p_[i] = i; }
    std::size_t size() const { return p_.size(); }
    std::size_t& operator()(std::size_t i) { return p_[i]; }
    const std::size_t& operator()(std::size_t i) const { return p_[i]; }
private:
    std::vector<std::size_t> p_;
};
template <class T, class P> int lu_factorize(matrix<T>& m, permutation_matrix<P>& pm) {
    int singular = 0;
    std::size_t n = m.size1();
    for (std::size_t k = 0; k < n; k++) {
        std::size_t piv = k; T best = std::fabs(m(k, k));
        for (std::size_t i = k + 1; i < n; i++) if (std::fabs(m(i, k)) > best) { best = std::fabs(m(i, k)); piv = i; }
        pm(k) = piv;
        if (best == T()) { if (!singular) singular = (int)k + 1; continue; }
        if (piv != k) for (std::size_t j = 0; j < n; j++) std::swap(m(k, j), m(piv, j));
        for (std::size_t i = k + 1; i < n; i++) {
            m(i, k) /= m(k, k);
            for (std::size_t j = k + 1; j < n; j++) m(i, j) -= m(i, k) * m(k, j);
        }
    }
    return singular;
}
template <class T, class P> void lu_substitute(const matrix<T>& m, const permutation_matrix<P>& pm, vector<T>& b) {
    std::size_t n = m.size1();
    for (std::size_t k = 0; k < n; k++) if (pm(k) != k) std::swap(b(k), b(pm(k)));
    for (std::size_t i = 0; i < n; i++) for (std::size_t j = 0; j < i; j++) b(i) -= m(i, j) * b(j);
    for (std::size_t ii = n; ii-- > 0;) {
        for (std::size_t j = ii + 1; j < n; j++) b(ii) -= m(ii, j) * b(j);
        b(ii) /= m(ii, ii);
    }
}
}}}
