// TEST INFRASTRUCTURE ONLY — stand-in for boost::math::students_t (Boost is absent from this image).
// Upper tail Q(t; v) = I_{v/(v+t^2)}(v/2, 1/2) / 2 for t >= 0, regularised incomplete beta evaluated with
// the modified Lentz continued fraction (DLMF 8.17.22); pinned against scipy.stats.t.sf in tests/.
#pragma once
#include <cmath>
namespace boost { namespace math {
namespace vo_detail {
inline double betacf(double a, double b, double x) {
    const double tiny = 1e-300, eps = 1e-16;
    double qab = a + b, qap = a + 1.0, qam = a - 1.0;
    double c = 1.0, d = 1.0 - qab * x / qap;
    if (std::fabs(d) < tiny) d = tiny;
    d = 1.0 / d;
    double h = d;
    for (int m = 1; m <= 10000; m++) {
        int m2 = 2 * m;
        double aa = m * (b - m) * x / ((qam + m2) * (a + m2));
        d = 1.0 + aa * d; if (std::fabs(d) < tiny) d = tiny;
        c = 1.0 + aa / c; if (std::fabs(c) < tiny) c = tiny;
        d = 1.0 / d; h *= d * c;
        aa = -(a + m) * (qab + m) * x / ((a + m2) * (qap + m2));
        d = 1.0 + aa * d; if (std::fabs(d) < tiny) d = tiny;
        c = 1.0 + aa / c; if (std::fabs(c) < tiny) c = tiny;
        d = 1.0 / d;
        double del = d * c;
        h *= del;
        if (std::fabs(del - 1.0) < eps) break;
    }
    return h;
}
inline double ibeta(double a, double b, double x) {
    if (x <= 0.0) return 0.0;
    if (x >= 1.0) return 1.0;
    double lnbt = std::lgamma(a + b) - std::lgamma(a) - std::lgamma(b) + a * std::log(x) + b * std::log1p(-x);
    double bt = std::exp(lnbt);
    if (x < (a + 1.0) / (a + b + 2.0)) return bt * betacf(a, b, x) / a;
    return 1.0 - bt * betacf(b, a, 1.0 - x) / b;
}
}
class students_t {
public:
    explicit students_t(double v) : v_(v) {}
    double degrees_of_freedom() const { return v_; }
private:
    double v_;
};
template <class D, class T> struct vo_complement2 { const D& dist; T param; };
template <class D, class T> inline vo_complement2<D, T> complement(const D& d, const T& t) { return vo_complement2<D, T>{d, t}; }
inline double cdf(const vo_complement2<students_t, double>& c) {
    double v = c.dist.degrees_of_freedom(), t = c.param;
    double x = v / (v + t * t);
    double tail = 0.5 * vo_detail::ibeta(0.5 * v, 0.5, x);
    return t >= 0 ? tail : 1.0 - tail;
}
}}
