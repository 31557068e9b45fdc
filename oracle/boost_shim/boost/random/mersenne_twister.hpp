// TEST INFRASTRUCTURE ONLY (oracle build). boost::mt19937 and std::mt19937 are the same
// generator by specification (MT19937, 32-bit, default tempering), so the shim aliases it.
#ifndef EM2_ORACLE_SHIM_MT_HPP
#define EM2_ORACLE_SHIM_MT_HPP
#include <random>
namespace boost {
typedef std::mt19937 mt19937;
namespace random { typedef std::mt19937 mt19937; }
}
#endif
