// TEST INFRASTRUCTURE ONLY -- stand-in for <boost/algorithm/string.hpp>: split + is_any_of (src/CellGraph.cpp:241).
#pragma once
#include <string>
namespace boost {
namespace algorithm {
struct shim_any_of {
    std::string chars;
    bool operator()(char c) const { return chars.find(c) != std::string::npos; }
};
inline shim_any_of is_any_of(const std::string& s) { return shim_any_of{s}; }
template <class Container, class Pred> Container& split(Container& out, const std::string& in, Pred isSeparator)
{
    out.clear();
    std::string token;
    for (char c : in) {
        if (isSeparator(c)) {
            out.push_back(token);
            token.clear();
        } else {
            token.push_back(c);
        }
    }
    out.push_back(token);
    return out;
}
}  // namespace algorithm
}  // namespace boost
