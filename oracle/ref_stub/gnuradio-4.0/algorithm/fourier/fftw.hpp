// ORACLE — TEST INFRASTRUCTURE.  Stand-in for <gnuradio-4.0/algorithm/fourier/fftw.hpp> (ALG/fourier/fftw.hpp:
// 172-226): the reference's FFT is FFTW 3.3.10, which is not in the tree.  Same interface
// (compute(range, std::vector&&) -> std::vector, power-of-two sizes only), arithmetic = the oracle's
// independent radix-2 FFT, so that the reference block and the oracle's restatement can be compared bit for bit.
#pragma once
#include <complex>
#include <ranges>
#include <vector>

#include "../../../../oracle_fft.hpp"

namespace gr::algorithm {
template <typename TIn, typename TOut>
struct FFTw {
    orc::Fft _f;
    std::vector<TIn> _in;
    template <std::ranges::input_range R>
    std::vector<TOut> compute(const R& in, std::vector<TOut>&& out = {})
    {
        _in.assign(std::ranges::begin(in), std::ranges::end(in));
        if (_f.size() != _in.size()) _f = orc::Fft(_in.size(), orc::FftKind::Radix2);  // throws unless 2^N (:182-184)
        out.resize(_in.size());
        _f.forward(_in.data(), out.data());
        return std::move(out);
    }
};
}  // namespace gr::algorithm
