// ORACLE — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
// Minimal stand-in for <gnuradio-4.0/Block.hpp>: just enough of the GR4 block API for the reference's RX-sync
// block headers (blocks/include/gnuradio-4.0/packet-modem/*.hpp) to compile UNMODIFIED, from where they lie
// under /root/reference, into oracle/_ref/librefblocks.so (recipe: oracle/Makefile, target refblocks).  The
// real runtime needs pmtv, fmt, vir-simd, ... which are fetched from the network; what the
// blocks' data paths use of it is small: ports with min_samples / publishTag, spans with consume / publish,
// the merged input tag, property_map, gr::exception, work::Status.  oracle/ref_blocks.cpp plays the scheduler:
// it offers chunks cut at tags (GR/Block.hpp:1501-1508) and honours consume()/publish().
#pragma once
#include <algorithm>
#include <bit>
#include <cassert>
#include <cmath>
#include <complex>
#include <cstddef>
#include <cstdint>
#include <expected>
#include <numeric>
#include <ranges>
#include <span>
#include <stdexcept>
#include <string>
#include <vector>

#include <fmt/format.h>

#include "Tag.hpp"

namespace gr {

namespace meta {
template <size_t N>
struct fixed_string {
    char data[N]{};
    constexpr fixed_string(const char (&s)[N]) { std::copy_n(s, N, data); }
};
}  // namespace meta
template <meta::fixed_string S>
struct Doc {};

struct exception : std::runtime_error {
    using std::runtime_error::runtime_error;
};

namespace work {
enum class Status { ERROR = -100, INSUFFICIENT_OUTPUT_ITEMS = -3, INSUFFICIENT_INPUT_ITEMS = -2, DONE = -1, OK = 0 };
}

struct Async {};
template <size_t = 1, size_t = 1, bool = false>
struct Resampling {};

template <typename T>
concept ConsumableSpan = requires(const T& s) { s.size(); s.consume(size_t{}); };
template <typename T>
concept PublishableSpan = requires(T& s) { s.size(); s.publish(size_t{}); };

// what workInternal hands to processBulk (GR/Block.hpp:1549-1578)
template <typename T>
class InSpan
{
    std::span<const T> _s;
    mutable size_t _consumed = static_cast<size_t>(-1);

public:
    using value_type = T;
    explicit InSpan(std::span<const T> s) : _s(s) {}
    size_t size() const { return _s.size(); }
    bool empty() const { return _s.empty(); }
    const T& operator[](size_t i) const { return _s[i]; }
    auto begin() const { return _s.begin(); }
    auto end() const { return _s.end(); }
    const T* data() const { return _s.data(); }
    bool consume(size_t n) const
    {
        if (n > _s.size()) return false;
        _consumed = n;
        return true;
    }
    size_t consumed() const { return _consumed == static_cast<size_t>(-1) ? _s.size() : _consumed; }  // default: all
};
template <typename T>
class OutSpan
{
    std::span<T> _s;
    size_t _published = static_cast<size_t>(-1);

public:
    using value_type = T;
    explicit OutSpan(std::span<T> s) : _s(s) {}
    size_t size() const { return _s.size(); }
    T& operator[](size_t i) { return _s[i]; }
    auto begin() { return _s.begin(); }
    auto end() { return _s.end(); }
    T* data() { return _s.data(); }
    void publish(size_t n) { _published = n; }
    size_t published() const { return _published == static_cast<size_t>(-1) ? _s.size() : _published; }
};

template <typename T, typename... Attr>
struct PortIn {
    using value_type = T;
    size_t min_samples = 1;
    size_t max_samples = static_cast<size_t>(-1);
};
template <typename T, typename... Attr>
struct PortOut {
    using value_type = T;
    size_t min_samples = 1;
    size_t max_samples = static_cast<size_t>(-1);
    std::vector<Tag> published_tags;  // chunk-relative (GR/Port.hpp:654-712)
    void publishTag(const property_map& map, ssize_t offset) { published_tags.push_back(Tag{ offset, map }); }
};

struct Error {
    std::string message;
};
struct Message {  // GR/Message.hpp:99-110: only the body is read by the RX-sync blocks
    std::expected<property_map, Error> data;
};

template <typename Derived, typename... Args>
class Block
{
public:
    Tag _mergedInputTag{};  // GR/Block.hpp:616-618; blocks may clear it themselves
    std::string name = "ref";
    size_t input_chunk_size = 1;   // gr::Resampling members (GR/Block.hpp)
    size_t output_chunk_size = 1;
    bool input_tags_present() const { return !_mergedInputTag.map.empty(); }
    const Tag& mergedInputTag() const { return _mergedInputTag; }
    // scheduler side
    void offer_input_tag(const property_map& m) { _mergedInputTag = Tag{ 0, m }; }
    void clear_input_tag() { _mergedInputTag = Tag{}; }
    template <typename... A>
    void emitErrorMessage(A&&...) {}
    void requestStop() {}
};

}  // namespace gr
