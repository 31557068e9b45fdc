// TEST INFRASTRUCTURE ONLY (oracle build). Stand-in for boost::normal_distribution<>.
//
// Boost is an un-vendored, un-pinned dependency of the reference (doc/Programming.html:20 says
// Boost 1.58 on Linux Mint 18; src/ShortStaticString.hpp:6-8 says the CentOS 7 build uses
// Boost 1.53).  Those two versions use DIFFERENT samplers (Box-Muller up to 1.55, ziggurat
// from 1.56), so the reference's own releases do not agree on the hyperplanes for one seed.
// This shim restates the published Boost <= 1.55 algorithm (polar-free Box-Muller over
// uniform_01, one cached value), which is what the reference's CentOS 7 build runs:
//     r1 = U01, r2 = U01, rho = sqrt(-2 log(1-r2));  x0 = rho cos(2 pi r1), x1 = rho sin(2 pi r1)
// with U01 = mt19937() * 2^-32 (boost::uniform_01 over a 32-bit integer engine).
// The product's host generator (expressionmatrix2_b200/host) implements the same recipe, so the
// oracle and the GPU path see identical hyperplanes for identical seeds.  Parity of the
// hyperplane VALUES against a real Boost build is unpinned (Boost cannot be had offline).
#ifndef EM2_ORACLE_SHIM_NORMAL_HPP
#define EM2_ORACLE_SHIM_NORMAL_HPP
#include <cmath>
namespace boost {
template <class RealType = double> class normal_distribution {
public:
    typedef RealType input_type;
    typedef RealType result_type;
    explicit normal_distribution(RealType mean = 0, RealType sigma = 1)
        : mean_(mean), sigma_(sigma), r1_(0), r2_(0), rho_(0), valid_(false) {}
    void reset() { valid_ = false; }
    template <class Engine> result_type operator()(Engine& eng)
    {
        const RealType twoPi = RealType(2) * RealType(3.14159265358979323846264338327950288L);
        if (!valid_) {
            r1_ = uniform01(eng);
            r2_ = uniform01(eng);
            rho_ = std::sqrt(-RealType(2) * std::log(RealType(1) - r2_));
            valid_ = true;
        } else {
            valid_ = false;
        }
        return rho_ * (valid_ ? std::cos(twoPi * r1_) : std::sin(twoPi * r1_)) * sigma_ + mean_;
    }
private:
    template <class Engine> static RealType uniform01(Engine& eng)
    {
        // boost::uniform_01 on an integer engine: (x - min) / (max - min + 1), retry on 1.0.
        const RealType factor = RealType(1) / (RealType((eng.max)() - (eng.min)()) + RealType(1));
        for (;;) {
            const RealType r = RealType(eng() - (eng.min)()) * factor;
            if (r < RealType(1)) return r;
        }
    }
    RealType mean_, sigma_, r1_, r2_, rho_;
    bool valid_;
};
}
#endif
