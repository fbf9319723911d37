// ORACLE — TEST INFRASTRUCTURE.  Stand-in for magic_enum, for the one enum the RX-sync blocks reflect on
// (gr::packet_modem::Constellation { PILOT, BPSK, QPSK }, PM/constellation.hpp:6): names come from a
// specialisable table instead of compiler reflection.
#pragma once
#include <algorithm>
#include <array>
#include <cctype>
#include <optional>
#include <string_view>

namespace magic_enum {
template <typename E>
struct names;  // specialise: static constexpr std::array<std::string_view, N> value
struct case_insensitive_t {};
inline constexpr case_insensitive_t case_insensitive{};
template <typename E>
constexpr std::string_view enum_name(E e) { return names<E>::value[static_cast<size_t>(e)]; }
template <typename E>
std::optional<E> enum_cast(std::string_view s, case_insensitive_t)
{
    for (size_t i = 0; i < names<E>::value.size(); ++i) {
        const auto n = names<E>::value[i];
        if (n.size() == s.size() && std::equal(n.begin(), n.end(), s.begin(), [](char a, char b) {
                return std::toupper(static_cast<unsigned char>(a)) == std::toupper(static_cast<unsigned char>(b));
            }))
            return static_cast<E>(i);
    }
    return std::nullopt;
}
}  // namespace magic_enum
