// TEST INFRASTRUCTURE ONLY (oracle build). Minimal stand-in for <boost/lexical_cast.hpp>.
// Boost is not installed in this image; the reference's hot-path sources only use
// lexical_cast<string>(integer) / lexical_cast<number>(string), BOOST_CURRENT_FUNCTION and
// BOOST_STATIC_ASSERT (CZI_ASSERT.hpp:25, MemoryMappedVector.hpp:197).  None of this is
// arithmetic on the LSH path.
#ifndef EM2_ORACLE_SHIM_LEXICAL_CAST_HPP
#define EM2_ORACLE_SHIM_LEXICAL_CAST_HPP
#include <sstream>
#include <string>
#include <typeinfo>
#include <cstdint>
#include <cmath>
#include <unistd.h>
namespace boost {
class bad_lexical_cast : public std::bad_cast {
public:
    const char* what() const noexcept override { return "bad lexical cast"; }
};
template <class Target, class Source> inline Target lexical_cast(const Source& s)
{
    std::stringstream ss;
    Target t;
    if (!(ss << s) || !(ss >> t) || !(ss >> std::ws).eof()) throw bad_lexical_cast();
    return t;
}
template <> inline std::string lexical_cast<std::string, std::string>(const std::string& s) { return s; }
}  // namespace boost
#ifndef BOOST_CURRENT_FUNCTION
#define BOOST_CURRENT_FUNCTION __PRETTY_FUNCTION__
#endif
#ifndef BOOST_STATIC_ASSERT
#define BOOST_STATIC_ASSERT(x) static_assert(x, #x)
#endif
#endif
