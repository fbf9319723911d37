// Drives the C++ block shell the way the GR4 runtime does (offer a span, honour consume/publish,
// collect published tags) and prints what a VectorSink would have captured, as text lines that
// tests/test_gpu_cpp_shell.py compares with the Python path and the oracle.
//   usage: test_block_shell <capture.cf32> <rrc_taps.f32> <min_bin> <max_bin> <threshold> <chunk>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <fstream>
#include <span>
#include <vector>

#include "../../gr4_packet_modem_b200/blocks/syncword_detection_b200.hpp"

template <typename T>
static std::vector<T> slurp(const char* path)
{
    std::ifstream f(path, std::ios::binary | std::ios::ate);
    if (!f) { std::fprintf(stderr, "cannot open %s\n", path); std::exit(2); }
    const std::streamsize n = f.tellg();
    f.seekg(0);
    std::vector<T> v(static_cast<size_t>(n) / sizeof(T));
    f.read(reinterpret_cast<char*>(v.data()), n);
    return v;
}

int main(int argc, char** argv)
{
    if (argc != 7) return 2;
    using c64 = std::complex<float>;
    const auto x = slurp<c64>(argv[1]);
    gr::packet_modem::SyncwordDetectionB200 blk;
    blk.rrc_taps = slurp<float>(argv[2]);
    static const uint8_t sw[64] = { 0,0,0,0,0,0,1,1,0,1,0,0,0,1,1,1,0,1,1,1,0,1,1,0,1,1,0,0,0,1,1,1,
                                    0,0,1,0,0,1,1,1,0,0,1,0,1,0,0,0,1,0,0,1,0,1,0,1,1,0,1,1,0,0,0,0 };
    blk.syncword.assign(sw, sw + 64);
    blk.constellation = { { 1.0f, 0.0f }, { -1.0f, 0.0f } };
    blk.min_freq_bin = std::atoi(argv[3]);
    blk.max_freq_bin = std::atoi(argv[4]);
    blk.power_threshold = static_cast<float>(std::atof(argv[5]));
    const size_t chunk = static_cast<size_t>(std::atoll(argv[6]));
    // error path first: the reference's exception text must come through
    {
        gr::packet_modem::SyncwordDetectionB200 bad;
        bad.rrc_taps = blk.rrc_taps; bad.syncword = blk.syncword; bad.constellation = blk.constellation;
        bad.min_freq_bin = 3; bad.max_freq_bin = 1;
        try { bad.start(); std::printf("error none\n"); }
        catch (const gr::exception& e) { std::printf("error %s\n", e.what()); }
    }
    blk.start();
    std::vector<c64> out(chunk);
    size_t pos = 0;
    unsigned long long checksum = 0;
    while (x.size() - pos >= blk.fft_size) {
        const size_t n = std::min(chunk, x.size() - pos);
        gr::ConsumableSpanShim<c64> in_span(std::span<const c64>(x.data() + pos, n));
        gr::PublishableSpanShim<c64> out_span(std::span<c64>(out.data(), n));
        blk.out.published_tags.clear();
        const auto st = blk.processBulk(in_span, out_span);
        if (st != gr::work::Status::OK) break;
        const size_t c = in_span.consumed();
        if (c != out_span.published()) { std::printf("mismatch\n"); return 1; }
        for (const auto& t : blk.out.published_tags) {
            std::printf("tag %llu", static_cast<unsigned long long>(pos) + static_cast<unsigned long long>(t.index));
            for (const auto& [k, v] : t.map) {
                std::printf(" %s=", k.c_str());
                std::visit([](auto&& a) {
                    using A = std::decay_t<decltype(a)>;
                    if constexpr (std::is_same_v<A, std::string>) std::printf("%s", a.c_str());
                    else if constexpr (std::is_same_v<A, int>) std::printf("%d", a);
                    else if constexpr (std::is_same_v<A, std::uint64_t>) std::printf("%llu", (unsigned long long)a);
                    else std::printf("%.17g", static_cast<double>(a));
                }, v);
            }
            std::printf("\n");
        }
        for (size_t i = 0; i < c; ++i) {  // order-sensitive checksum of the published samples
            uint32_t w[2];
            std::memcpy(w, &out[i], 8);
            checksum = checksum * 1099511628211ULL + w[0] + 31ULL * w[1];
        }
        pos += c;
        if (c == 0) break;
    }
    std::printf("consumed %zu checksum %llu\n", pos, checksum);
    return 0;
}
