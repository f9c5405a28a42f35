// stand-in for boost::math::lgamma: libm lgamma (test infrastructure)
#ifndef NBMF_REF_SHIM_GAMMA_HPP
#define NBMF_REF_SHIM_GAMMA_HPP
#include <cmath>
namespace boost { namespace math {
inline double lgamma(double x) { return ::lgamma(x); }
}}
#endif
