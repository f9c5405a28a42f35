// stand-in for boost::math::digamma (test infrastructure; see ../../../RcppArmadillo.h):
// recurrence to x >= 10 + asymptotic series through B12 -- the same scheme as
// oracle_digamma() so that the reference sources and the oracle evaluate psi identically.
#ifndef NBMF_REF_SHIM_DIGAMMA_HPP
#define NBMF_REF_SHIM_DIGAMMA_HPP
#include <cmath>
namespace boost { namespace math {
inline double digamma(double x)
{
    double r = 0.0;
    if (!(x > 0.0)) {
        if (x == 0.0) return -INFINITY;
        if (x == std::floor(x)) return NAN;
        return digamma(1.0 - x) - M_PI / std::tan(M_PI * x);
    }
    while (x < 10.0) { r -= 1.0 / x; x += 1.0; }
    double inv = 1.0 / x, inv2 = inv * inv;
    double s = inv2 * (1.0 / 12.0 - inv2 * (1.0 / 120.0 - inv2 * (1.0 / 252.0 -
               inv2 * (1.0 / 240.0 - inv2 * (1.0 / 132.0 - inv2 * (691.0 / 32760.0 -
               inv2 * (1.0 / 12.0)))))));
    r += std::log(x) - 0.5 * inv - s;
    return r;
}
}}
#endif
