// TEST INFRASTRUCTURE ONLY — stand-in for boost::math::normal (Boost is absent from this image).
// Boost's normal cdf is erfc(-(x-mean)/(sd*sqrt(2)))/2; restated here from its documentation.
#pragma once
#include <cmath>
namespace boost { namespace math {
class normal {
public:
    normal(double mean = 0.0, double sd = 1.0) : m_(mean), s_(sd) {}
    double mean() const { return m_; }
    double standard_deviation() const { return s_; }
private:
    double m_, s_;
};
inline double cdf(const normal& d, double x) {
    return 0.5 * std::erfc(-(x - d.mean()) / (d.standard_deviation() * std::sqrt(2.0)));
}
}}
