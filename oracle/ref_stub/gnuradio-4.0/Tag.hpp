// ORACLE — TEST INFRASTRUCTURE.  Stand-in for <gnuradio-4.0/Tag.hpp> (GR/Tag.hpp:47-75): property_map, Tag,
// TagPropagationPolicy and a pmtv value type able to hold what the RX-sync blocks put in tags.
#pragma once
#include <complex>
#include <cstdint>
#include <map>
#include <string>
#include <sys/types.h>
#include <variant>
#include <vector>

#include "reflection.hpp"

namespace pmtv {
using pmt = std::variant<std::monostate, bool, int, unsigned, long, unsigned long, float, double, std::string,
                         std::vector<float>>;
using map_t = std::map<std::string, pmt, std::less<>>;
template <typename T, typename V>
inline T cast(const V& v)
{
    return std::visit(
        [](auto&& a) -> T {
            using A = std::decay_t<decltype(a)>;
            if constexpr (std::is_arithmetic_v<A> && std::is_arithmetic_v<T>) return static_cast<T>(a);
            else throw std::bad_variant_access();
        },
        v);
}
}  // namespace pmtv

namespace gr {
using property_map = pmtv::map_t;
enum class TagPropagationPolicy { TPP_DONT = 0, TPP_ALL_TO_ALL = 1, TPP_ONE_TO_ONE = 2, TPP_CUSTOM = 3 };
struct Tag {
    ssize_t index{ 0 };
    property_map map{};
};
}  // namespace gr
