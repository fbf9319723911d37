// gr4_compat.hpp — the few GNU Radio 4.0 names the block shells use, for builds where
// <gnuradio-4.0/Block.hpp> is not available (GR4's dependencies are fetched from the network at
// configure time and cannot be installed in the build container — DESIGN.md §7).
// With real GR4 on the include path this header is NOT used: the shells include the real
// headers and derive from gr::Block<T> (see syncword_detection_b200.hpp).
//
// Stand-ins follow the contract of SURVEY §8(b):
//   ConsumableSpan:  contiguous view, consume(n) -> bool      (GR/Block.hpp:1549-1578)
//   PublishableSpan: contiguous view, publish(n)              (GR/Block.hpp:1636-1651)
//   PortOut::publishTag(map, offset)                          (GR/Port.hpp:654-712)
//   gr::exception, gr::work::Status, gr::property_map         (GR/Message.hpp:21-33, GR/Tag.hpp:54)
#pragma once
#include <complex>
#include <cstddef>
#include <cstdint>
#include <map>
#include <optional>
#include <span>
#include <stdexcept>
#include <string>
#include <sys/types.h>
#include <variant>
#include <vector>

namespace gr {

struct exception : std::runtime_error {
    using std::runtime_error::runtime_error;
};

namespace work {
enum class Status { ERROR = -100, INSUFFICIENT_OUTPUT_ITEMS = -3, INSUFFICIENT_INPUT_ITEMS = -2, DONE = -1, OK = 0 };
}

using pmt_value = std::variant<float, double, int, std::uint64_t, std::string>;
using property_map = std::map<std::string, pmt_value>;

struct Tag {
    ssize_t index;
    property_map map;
};

template <typename T>
class ConsumableSpanShim
{
    std::span<const T> _s;
    mutable std::size_t _consumed = static_cast<std::size_t>(-1);

public:
    explicit ConsumableSpanShim(std::span<const T> s) : _s(s) {}
    std::size_t size() const { return _s.size(); }
    const T* data() const { return _s.data(); }
    const T& operator[](std::size_t i) const { return _s[i]; }
    auto begin() const { return _s.begin(); }
    auto end() const { return _s.end(); }
    bool consume(std::size_t n) const
    {
        if (n > _s.size()) return false;
        _consumed = n;
        return true;
    }
    // what the runtime does when the block did not call consume(): everything (Block.hpp:1636-1651)
    std::size_t consumed() const { return _consumed == static_cast<std::size_t>(-1) ? _s.size() : _consumed; }
};

template <typename T>
class PublishableSpanShim
{
    std::span<T> _s;
    std::size_t _published = static_cast<std::size_t>(-1);

public:
    explicit PublishableSpanShim(std::span<T> s) : _s(s) {}
    std::size_t size() const { return _s.size(); }
    T* data() { return _s.data(); }
    T& operator[](std::size_t i) { return _s[i]; }
    auto begin() { return _s.begin(); }
    auto end() { return _s.end(); }
    void publish(std::size_t n) { _published = n; }
    std::size_t published() const { return _published == static_cast<std::size_t>(-1) ? _s.size() : _published; }
};

// output port stand-in: records the tags a block publishes, with chunk-relative offsets
template <typename T>
struct PortOutShim {
    std::size_t min_samples = 1;
    std::vector<Tag> published_tags;
    void publishTag(const property_map& map, ssize_t offset) { published_tags.push_back(Tag{ offset, map }); }
};
template <typename T>
struct PortInShim {
    std::size_t min_samples = 1;
};


// what gr::Block<Derived> gives a block for input tags (GR/Block.hpp:616-618): the runtime merges the
// tags of the chunk's first sample into _mergedInputTag before calling processBulk
template <typename Derived>
struct BlockShim {
    std::string name = "b200";
    Tag _mergedInputTag{ 0, {} };
    bool input_tags_present() const { return !_mergedInputTag.map.empty(); }
    const Tag& mergedInputTag() const { return _mergedInputTag; }
    // test driver side: what the runtime does before / after a processBulk call
    void offer_input_tag(const property_map& m) { _mergedInputTag = Tag{ 0, m }; }
    void clear_input_tag() { _mergedInputTag.map.clear(); }
};

// gr::Message stand-in: only the `data` member the hot-path blocks read
// (PM/syncword_detection_filter.hpp:137 `headerSpan[0].data.value()`)
struct Message {
    std::optional<property_map> data;
};

struct Async {};

} // namespace gr
