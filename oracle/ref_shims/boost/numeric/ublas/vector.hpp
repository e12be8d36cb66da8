// TEST INFRASTRUCTURE ONLY — minimal dense vector standing in for boost::numeric::ublas::vector.
#pragma once
#include <vector>
#include <cstddef>
namespace boost { namespace numeric { namespace ublas {
template <class T> class vector {
public:
    vector() {}
    explicit vector(std::size_t n) : d_(n, T()) {}
    std::size_t size() const { return d_.size(); }
    T& operator()(std::size_t i) { return d_[i]; }
    const T& operator()(std::size_t i) const { return d_[i]; }
private:
    std::vector<T> d_;
};
}}}
