// ORACLE — TEST INFRASTRUCTURE.  Stand-in for {fmt}: the reference sources only format text for exceptions and
// trace output; neither affects the data path that oracle/_ref pins.
#pragma once
#include <string>
namespace fmt {
template <typename... A>
inline std::string format(const char* f, A&&...) { return std::string(f); }
template <typename... A>
inline void println(A&&...) {}
template <typename... A>
inline void print(A&&...) {}
}  // namespace fmt
