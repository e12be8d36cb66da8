// TEST INFRASTRUCTURE ONLY — minimal dense row-major matrix standing in for boost::numeric::ublas::matrix.
#pragma once
#include <vector>
#include <cstddef>
#include "vector.hpp"
namespace boost { namespace numeric { namespace ublas {
template <class T> class matrix {
public:
    matrix() : r_(0), c_(0) {}
    matrix(std::size_t r, std::size_t c) : r_(r), c_(c), d_(r * c, T()) {}
    std::size_t size1() const { return r_; }
    std::size_t size2() const { return c_; }
    T& operator()(std::size_t i, std::size_t j) { return d_[i * c_ + j]; }
    const T& operator()(std::size_t i, std::size_t j) const { return d_[i * c_ + j]; }
private:
    std::size_t r_, c_;
    std::vector<T> d_;
};
template <class T> matrix<T> prod(const matrix<T>& a, const matrix<T>& b) {
    matrix<T> c(a.size1(), b.size2());
    for (std::size_t i = 0; i < a.size1(); i++)
        for (std::size_t k = 0; k < a.size2(); k++) {
            T aik = a(i, k);
            for (std::size_t j = 0; j < b.size2(); j++) c(i, j) += aik * b(k, j);
        }
    return c;
}
template <class T> vector<T> prod(const matrix<T>& a, const vector<T>& x) {
    vector<T> y(a.size1());
    for (std::size_t i = 0; i < a.size1(); i++) {
        T s = T();
        for (std::size_t k = 0; k < a.size2(); k++) s += a(i, k) * x(k);
        y(i) = s;
    }
    return y;
}
}}}
