// TEST INFRASTRUCTURE ONLY (oracle build). Stand-in for <boost/math/constants/constants.hpp>:
// the reference uses boost::math::double_constants::pi (Lsh.cpp:241), which is the double
// nearest to pi.
#ifndef EM2_ORACLE_SHIM_MATH_CONSTANTS_HPP
#define EM2_ORACLE_SHIM_MATH_CONSTANTS_HPP
namespace boost { namespace math { namespace double_constants {
static const double pi = 3.141592653589793238462643383279502884;
}}}
#endif
