// Ids and small fixed-layout types shared by the host classes of the hot path.
#pragma once
#include <cstdint>
#include <cstring>
#include <limits>
#include <stdexcept>
#include <string>

namespace ChanZuckerberg {
namespace ExpressionMatrix2 {

// reference src/Ids.hpp:12-15
using GeneId = uint32_t;
using CellId = uint32_t;
static const GeneId invalidGeneId = std::numeric_limits<GeneId>::max();
static const CellId invalidCellId = std::numeric_limits<CellId>::max();

// Layout of the reference's ShortStaticString<255> (src/ShortStaticString.hpp:27-43): one length byte
// followed by 255 chars, 256 bytes, no heap -- it lives inside memory-mapped Info objects.
struct StaticString255 {
    uint8_t n = 0;
    char s[255];
    StaticString255() { std::memset(s, 0, sizeof(s)); }
    StaticString255(const std::string& x) { *this = x; }
    StaticString255& operator=(const std::string& x)
    {
        if (x.size() > 255) throw std::runtime_error("String is too long for a StaticString255: " + x);
        n = uint8_t(x.size());
        std::memset(s, 0, sizeof(s));
        std::memcpy(s, x.data(), x.size());
        return *this;
    }
    operator std::string() const { return std::string(s, s + n); }
    size_t size() const { return n; }
};
static_assert(sizeof(StaticString255) == 256, "StaticString255 must be 256 bytes");

}  // namespace ExpressionMatrix2
}  // namespace ChanZuckerberg
