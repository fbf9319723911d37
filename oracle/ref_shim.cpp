// =============================================================================
// ORACLE/_ref — TEST INFRASTRUCTURE.  Thin extern "C" shim around the REFERENCE's
// own headers, included from /root/reference (never copied).  Only the two std-only
// headers of the hot path compile without GNU Radio 4.0's fetched dependencies
// (SURVEY.md §8c):
//   PM/firdes.hpp:30-76       firdes::root_raised_cosine
//   PM/pfb_arb_taps.hpp:13    pfb_arb_taps (1280 remez taps)
// Everything else (#include <gnuradio-4.0/Block.hpp> -> pmtv, fmt, vir-simd, FFTW)
// is unbuildable here, hence the restatement in oracle.hpp.
// =============================================================================
#include <cstddef>
#include <cstring>
#include <sys/types.h>
using ssize_t = ::ssize_t;
#include <gnuradio-4.0/packet-modem/firdes.hpp>
#include <gnuradio-4.0/packet-modem/pfb_arb_taps.hpp>

extern "C" {
int ref_rrc(double gain, double fs, double symrate, double alpha, size_t ntaps, float* out, size_t max_out)
{
    const auto t = gr::packet_modem::firdes::root_raised_cosine(gain, fs, symrate, alpha, ntaps);
    if (t.size() > max_out) return -1;
    std::memcpy(out, t.data(), t.size() * sizeof(float));
    return static_cast<int>(t.size());
}
int ref_pfb_arb_taps(float* out, size_t max_out)
{
    const auto& t = gr::packet_modem::pfb_arb_taps;
    if (t.size() > max_out) return -1;
    std::memcpy(out, t.data(), t.size() * sizeof(float));
    return static_cast<int>(t.size());
}
}
