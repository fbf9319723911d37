// =============================================================================
// ORACLE — TEST INFRASTRUCTURE, NOT PRODUCT CODE (see oracle.hpp).
//
// FFT arithmetic standing in for FFTW 3.3.10 (ALG/fourier/fftw.hpp:172-226 calls
// fftwf_execute; FFTW itself is fetched at configure time by
// gnuradio4/CMakeLists.txt:245-282 and is not in /root/reference).  FFTW's rounding
// depends on the codelets its ESTIMATE planner picks, so no CPU restatement can be
// bit-identical to it: "parity unpinned" at the bit level.  Two arithmetics:
//
//  FftKind::Radix2  — INDEPENDENT of the GPU code.  Iterative radix-2 decimation in
//      time with bit-reversal, the structure of the reference's own std-only FFT
//      (ALG/fourier/fft.hpp:70-83 butterflies, :95-101 bit reversal), with twiddles
//      computed directly in double and rounded once (more accurate than the repeated
//      multiplication of :104-118).  Plain complex arithmetic, every op separately
//      rounded.  This is the arithmetic the parity tests hold the GPU to for detection
//      indices (exact) and estimates (toleranced).
//
//  FftKind::Mirror  — a loop-based restatement of the ARITHMETIC CONTRACT written in
//      gr4_packet_modem_b200/csrc/fft2048.cuh (16x16x8 decomposition, radix-2 DIF
//      small DFTs, fma-based complex multiply, W8 factors folded into the next butterfly,
//      one twiddle table).  With it the oracle
//      reproduces the GPU's zpow and detection records bit for bit, which pins every
//      comparison/threshold/ordering decision of the pipeline, also at low SNR where a
//      1-ulp difference can flip a threshold test.  It shares no code with the product.
// =============================================================================
#pragma once
#include <cmath>
#include <complex>
#include <cstddef>
#include <numbers>
#include <stdexcept>
#include <vector>

namespace orc {

using c64 = std::complex<float>;

enum class FftKind : int { Radix2 = 0, Mirror = 1 };

static inline c64 cmul_plain(c64 a, c64 b)
{
    return c64(a.real() * b.real() - a.imag() * b.imag(), a.real() * b.imag() + a.imag() * b.real());
}
static inline c64 cmul_fma(c64 a, c64 w)
{
    return c64(std::fmaf(a.real(), w.real(), -(a.imag() * w.imag())),
               std::fmaf(a.real(), w.imag(), a.imag() * w.real()));
}

class Fft
{
    size_t _n = 0;
    FftKind _kind = FftKind::Radix2;
    std::vector<c64> _tw;       // Radix2: per-stage twiddles; Mirror: Wt[j], j < 2048
    std::vector<size_t> _rev;   // Radix2 bit reversal

    // ---------------- Radix2 ----------------
    void radix2(const c64* in, c64* out) const
    {
        for (size_t j = 0; j < _n; ++j) out[_rev[j]] = in[j]; // ALG/fourier/fft.hpp:95-101
        size_t tw = 0;
        for (size_t s = 2; s <= _n; s *= 2) { // ALG/fourier/fft.hpp:71-83
            const size_t half = s / 2;
            for (size_t k = 0; k < _n; k += s) {
                for (size_t j = 0; j < half; ++j) {
                    const c64 t = cmul_plain(_tw[tw + j], out[k + j + half]);
                    const c64 u = out[k + j];
                    out[k + j] = c64(u.real() + t.real(), u.imag() + t.imag());
                    out[k + j + half] = c64(u.real() - t.real(), u.imag() - t.imag());
                }
            }
            tw += half;
        }
    }

    // ---------------- Mirror ----------------
    static constexpr float kC = 0.70710678118654752440f;
    static constexpr float kCos = 0.92387953251128675613f;
    static constexpr float kSin = 0.38268343236508977173f;
    static c64 mul_w16(c64 a, int e)
    {
        const float x = a.real(), y = a.imag();
        switch (e) {
        case 0: return a;
        case 1: return cmul_fma(a, c64(kCos, -kSin));
        case 2: return c64(x + y, y - x);          // e1(d): the factor c is applied by the next stage
        case 3: return cmul_fma(a, c64(kSin, -kCos));
        case 4: return c64(y, -x);
        case 5: return cmul_fma(a, c64(-kSin, -kCos));
        case 6: return c64(y - x, (-x) + (-y));    // e3(d): the factor c is applied by the next stage
        default: return cmul_fma(a, c64(-kCos, -kSin));
        }
    }
    // radix-2 DIF, n = 16 or 8; output in bit-reversed order.  Elements that carry W8^1 / W8^3 keep the
    // unscaled term and a "pending c" mark; the next stage pairs two such elements and applies c with fmas
    // (fft2048.cuh: w8pair).
    static void dft_small(c64* v, int n)
    {
        bool pend[16] = {};
        for (int half = n / 2; half >= 1; half /= 2) {
            const int step = 8 / half;
            for (int g = 0; g < n; g += 2 * half) {
                for (int i = 0; i < half; ++i) {
                    const c64 u = v[g + i], w = v[g + i + half];
                    c64 s, d;
                    if (pend[g + i] != pend[g + i + half]) throw std::logic_error("mirror: unpaired W8 factor");
                    if (pend[g + i]) {
                        const c64 a(kC * u.real(), kC * u.imag());
                        s = c64(std::fmaf(kC, w.real(), a.real()), std::fmaf(kC, w.imag(), a.imag()));
                        d = c64(std::fmaf(-kC, w.real(), a.real()), std::fmaf(-kC, w.imag(), a.imag()));
                        pend[g + i] = pend[g + i + half] = false;
                    } else {
                        s = c64(u.real() + w.real(), u.imag() + w.imag());
                        d = c64(u.real() - w.real(), u.imag() - w.imag());
                    }
                    v[g + i] = s;
                    const int e = i * step;
                    v[g + i + half] = mul_w16(d, e);
                    if (e == 2 || e == 6) pend[g + i + half] = true;
                }
            }
        }
        for (int i = 0; i < n; ++i)
            if (pend[i]) throw std::logic_error("mirror: W8 factor left pending");
    }
    static int br4(int k) { return ((k & 1) << 3) | ((k & 2) << 1) | ((k & 4) >> 1) | ((k & 8) >> 3); }
    static int br3(int k) { return ((k & 1) << 2) | (k & 2) | ((k & 4) >> 2); }

    void mirror_a(const c64* x, c64* X) const
    {
        std::vector<c64> b(2048), c(2048);
        c64 v[16];
        for (int n2 = 0; n2 < 16; ++n2)
            for (int n3 = 0; n3 < 8; ++n3) {
                for (int n1 = 0; n1 < 16; ++n1) v[n1] = x[128 * n1 + 8 * n2 + n3];
                dft_small(v, 16);
                for (int k1 = 0; k1 < 16; ++k1) {
                    c64 val = v[br4(k1)];
                    if (k1 != 0) val = cmul_fma(val, _tw[8 * n2 * k1]);
                    b[(k1 * 16 + n2) * 8 + n3] = val;
                }
            }
        for (int k1 = 0; k1 < 16; ++k1)
            for (int n3 = 0; n3 < 8; ++n3) {
                for (int n2 = 0; n2 < 16; ++n2) v[n2] = b[(k1 * 16 + n2) * 8 + n3];
                dft_small(v, 16);
                for (int k2 = 0; k2 < 16; ++k2)
                    c[(k1 + 16 * k2) * 8 + n3] = cmul_fma(v[br4(k2)], _tw[n3 * (k1 + 16 * k2)]);
            }
        for (int p = 0; p < 256; ++p) {
            c64 w[8];
            for (int n3 = 0; n3 < 8; ++n3) w[n3] = c[p * 8 + n3];
            dft_small(w, 8);
            for (int k3 = 0; k3 < 8; ++k3) X[p + 256 * k3] = w[br3(k3)];
        }
    }
    void mirror_b(const c64* Y, c64* C) const
    {
        std::vector<c64> e(2048), f(2048);
        c64 v[16];
        for (int p = 0; p < 256; ++p) {
            c64 w[8];
            for (int f3 = 0; f3 < 8; ++f3) w[f3] = Y[p + 256 * f3];
            dft_small(w, 8);
            for (int m3 = 0; m3 < 8; ++m3) {
                c64 val = w[br3(m3)];
                if (m3 != 0) val = cmul_fma(val, _tw[p * m3]);
                e[p * 8 + m3] = val;
            }
        }
        for (int f1 = 0; f1 < 16; ++f1)
            for (int m3 = 0; m3 < 8; ++m3) {
                for (int f2 = 0; f2 < 16; ++f2) v[f2] = e[(f1 + 16 * f2) * 8 + m3];
                dft_small(v, 16);
                for (int m2 = 0; m2 < 16; ++m2) {
                    c64 val = v[br4(m2)];
                    if (m2 != 0) val = cmul_fma(val, _tw[8 * f1 * m2]);
                    f[(f1 * 16 + m2) * 8 + m3] = val;
                }
            }
        for (int m2 = 0; m2 < 16; ++m2)
            for (int m3 = 0; m3 < 8; ++m3) {
                for (int f1 = 0; f1 < 16; ++f1) v[f1] = f[(f1 * 16 + m2) * 8 + m3];
                dft_small(v, 16);
                for (int m1 = 0; m1 < 16; ++m1) C[128 * m1 + 8 * m2 + m3] = v[br4(m1)];
            }
    }

public:
    Fft() = default;
    Fft(size_t n, FftKind kind) : _n(n), _kind(kind)
    {
        if (n == 0 || (n & (n - 1)) != 0) throw std::runtime_error("FFT size must be 2^N"); // fftw.hpp:182-184
        if (kind == FftKind::Mirror) {
            if (n != 2048) throw std::runtime_error("mirror arithmetic exists for fft_size 2048 only");
            _tw.resize(2048);
            for (size_t j = 0; j < 2048; ++j) {
                const double a = 2.0 * std::numbers::pi * static_cast<double>(j) / 2048.0;
                _tw[j] = c64(static_cast<float>(std::cos(a)), static_cast<float>(-std::sin(a)));
            }
        } else {
            _rev.resize(n);
            size_t bits = 0;
            while ((size_t{ 1 } << bits) < n) ++bits;
            for (size_t j = 0; j < n; ++j) {
                size_t r = 0;
                for (size_t b = 0; b < bits; ++b)
                    if (j & (size_t{ 1 } << b)) r |= size_t{ 1 } << (bits - 1 - b);
                _rev[j] = r;
            }
            for (size_t s = 2; s <= n; s *= 2)
                for (size_t j = 0; j < s / 2; ++j) {
                    const double a = -2.0 * std::numbers::pi * static_cast<double>(j) / static_cast<double>(s);
                    _tw.push_back(c64(static_cast<float>(std::cos(a)), static_cast<float>(std::sin(a))));
                }
        }
    }
    size_t size() const { return _n; }
    FftKind kind() const { return _kind; }

    // forward FFT of the samples / of the template (PM/syncword_detection.hpp:184, 239-241)
    void forward(const c64* in, c64* out) const
    {
        if (_kind == FftKind::Mirror) mirror_a(in, out);
        else radix2(in, out);
    }
    // "IFFT computed as an FFT" of the spectrum product (PM/syncword_detection.hpp:250-251)
    void second(const c64* in, c64* out) const
    {
        if (_kind == FftKind::Mirror) mirror_b(in, out);
        else radix2(in, out);
    }
    // samples_fft[k] * syncword_fft_conj[k] (PM/syncword_detection.hpp:248)
    c64 cmul(c64 x, c64 h) const { return _kind == FftKind::Mirror ? cmul_fma(x, h) : cmul_plain(x, h); }
    // z.real()*z.real() + z.imag()*z.imag() (PM/syncword_detection.hpp:260, 307)
    float norm2(c64 z) const
    {
        return _kind == FftKind::Mirror ? std::fmaf(z.imag(), z.imag(), z.real() * z.real())
                                        : z.real() * z.real() + z.imag() * z.imag();
    }
};

} // namespace orc
