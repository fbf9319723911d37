// =============================================================================
// ORACLE — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement of the receiver-synchronisation hot path of
// daniestevez/gr4-packet-modem.  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may load this.  The product
// (gr4_packet_modem_b200/csrc -> libb200sync.so) never includes, links or calls
// anything in this directory.
//
// Every function cites the reference file:line it follows.  Prefixes:
//   PM/  = blocks/include/gnuradio-4.0/packet-modem/
//   GR/  = gnuradio4/core/include/gnuradio-4.0/
//   ALG/ = gnuradio4/algorithm/include/gnuradio-4.0/algorithm/
//
// PARITY PINNING STATUS
//   * firdes::root_raised_cosine is pinned against the reference's own golden
//     vector (test/qa_firdes.cpp:10-33, 65 taps, tol 1e-7) and against the
//     reference header itself compiled into oracle/_ref (it is std-only).
//   * The FFT arithmetic of the reference is FFTW 3.3.10 (not vendored; fetched
//     by gnuradio4/CMakeLists.txt:245-282).  No golden vectors exist for it, so
//     at the bit level the FFT output is "parity unpinned"; the oracle offers
//     two FFT arithmetics (see oracle_fft.hpp) and is validated against the
//     assertions of test/qa_syncword_detection.cpp on seeded inputs.
//   * All block state machines are line-by-line restatements; they are checked
//     against the assertions of the reference's own qa_*.cpp on seeded inputs,
//     AND bit for bit against the reference's own block headers compiled
//     unmodified against a stand-in runtime (oracle/ref_blocks.cpp ->
//     oracle/_ref/librefblocks.so; tests/test_oracle_vs_reference_blocks.py):
//     pinned against outputs of the reference itself, FFT rounding excepted.
// =============================================================================
#pragma once
#include <algorithm>
#include <bit>
#include <cmath>
#include <complex>
#include <cstddef>
#include <cstdint>
#include <map>
#include <numbers>
#include <numeric>
#include <stdexcept>
#include <string>
#include <vector>

#include "oracle_fft.hpp"

namespace orc {

using c64 = std::complex<float>;

// complex * real and complex + complex with separately rounded float ops, as
// std::complex<float> does when compiled without contraction.
static inline c64 cscale(c64 a, float t) { return c64(a.real() * t, a.imag() * t); }

// ----------------------------------------------------------------------------
// firdes::root_raised_cosine — PM/firdes.hpp:30-76
// ----------------------------------------------------------------------------
inline std::vector<float> root_raised_cosine(double gain, double sampling_freq,
                                             double symbol_rate, double alpha, size_t ntaps)
{
    ntaps |= 1; // PM/firdes.hpp:33
    const double spb = sampling_freq / symbol_rate;
    std::vector<double> taps(ntaps);
    const double pi = std::numbers::pi;
    for (size_t i = 0; i < ntaps; ++i) {
        const double xindx = static_cast<double>(static_cast<std::ptrdiff_t>(i) -
                                                 static_cast<std::ptrdiff_t>(ntaps) / 2);
        const double x1 = pi * xindx / spb;
        double x2 = 4.0 * alpha * xindx / spb;
        double x3 = x2 * x2 - 1.0;
        double num, den;
        if (std::abs(x3) >= 0.000001) { // PM/firdes.hpp:46
            if (i != ntaps / 2) {
                num = std::cos((1.0 + alpha) * x1) +
                      std::sin((1.0 - alpha) * x1) / (4.0 * alpha * xindx / spb);
            } else {
                num = std::cos((1.0 + alpha) * x1) + (1.0 - alpha) * pi / (4.0 * alpha);
            }
            den = x3 * pi;
        } else { // PM/firdes.hpp:55-66
            if (alpha == 1.0) {
                taps[i] = -1.0;
                continue;
            }
            x3 = (1.0 - alpha) * x1;
            x2 = (1.0 + alpha) * x1;
            num = (std::sin(x2) * (1.0 + alpha) * pi -
                   std::cos(x3) * ((1.0 - alpha) * pi * spb) / (4.0 * alpha * xindx) +
                   std::sin(x3) * spb * spb / (4.0 * alpha * xindx * xindx));
            den = -32.0 * pi * alpha * alpha * xindx / spb;
        }
        taps[i] = 4.0 * alpha * num / den;
    }
    const double scale = std::accumulate(taps.cbegin(), taps.cend(), 0.0); // :69
    std::vector<float> out(ntaps);
    for (size_t i = 0; i < ntaps; ++i) out[i] = static_cast<float>(taps[i] * gain / scale);
    return out;
}

// ----------------------------------------------------------------------------
// HistoryBuffer — GR/HistoryBuffer.hpp:58-135.  Double-written circular buffer,
// [0] = newest.  Value-initialised storage (zeros) as std::vector<T>(2*cap).
// ----------------------------------------------------------------------------
template <typename T>
class History
{
    std::vector<T> _buffer;
    size_t _capacity;
    size_t _write_position = 0;
    size_t _size = 0;

public:
    explicit History(size_t capacity = 1) : _buffer(capacity * 2), _capacity(capacity)
    {
        if (capacity == 0) throw std::out_of_range("capacity is zero");
    }
    void push_back(const T& v) // GR/HistoryBuffer.hpp:92-103
    {
        if (_size < _capacity) ++_size;
        if (_write_position == 0) _write_position = _capacity;
        --_write_position;
        _buffer[_write_position] = v;
        _buffer[_write_position + _capacity] = v;
    }
    // GR/HistoryBuffer.hpp:58-72, 127-135 (power-of-two fast path and modulo fallback
    // give the same index)
    T& operator[](size_t i) { return _buffer[(_write_position + i) % _capacity]; }
    const T& operator[](size_t i) const { return _buffer[(_write_position + i) % _capacity]; }
    size_t size() const { return _size; }
    size_t capacity() const { return _capacity; }
    // contiguous view starting at the newest element (what cbegin() gives)
    const T* newest() const { return &_buffer[_write_position]; }
};

// ----------------------------------------------------------------------------
// SyncwordDetection — PM/syncword_detection.hpp
// ----------------------------------------------------------------------------
struct HistoryItem { // PM/syncword_detection.hpp:17-29 (value-initialised => zeros)
    c64 sample{};
    float correlation_power = 0.0f;
    float correlation_power_left = 0.0f;
    float correlation_power_right = 0.0f;
    c64 correlation{};
    int freq_bin = 0;
    float fft_noise_power = 0.0f;
    bool detection = false;
};

// One emitted tag.  `index` is the absolute index in the OUTPUT stream
// (= detected input sample index + 2*time_threshold + 1).  The raw fields the
// estimates were computed from are kept so a GPU implementation that returns raw
// detection records can be compared field by field.
struct SyncwordTag {
    uint64_t index;
    float amplitude;   // "syncword_amplitude"
    float phase;       // "syncword_phase"
    double freq;       // "syncword_freq"
    int32_t freq_bin;  // "syncword_freq_bin"
    float noise_power; // "syncword_noise_power"
    float esn0_db;     // "syncword_esn0_db"
    float time_est;    // "syncword_time_est"
    // raw
    float corr_re, corr_im;
    float pow, pow_left, pow_right, pow_prev, pow_next;
    int32_t _pad;
};

class SyncwordDetection
{
public:
    // settings — PM/syncword_detection.hpp:131-141
    size_t fft_size = 2048;
    size_t samples_per_symbol = 4;
    std::vector<float> rrc_taps;
    std::vector<uint8_t> syncword;
    std::vector<c64> constellation;
    int min_freq_bin = 0;
    int max_freq_bin = 0;
    uint64_t time_threshold = 768;
    float power_threshold = 9.5f;
    // oracle-only knob: which FFT arithmetic stands in for FFTW (see oracle_fft.hpp)
    FftKind fft_kind = FftKind::Radix2;

    // state — PM/syncword_detection.hpp:118-128
    size_t _syncword_samples_size = 0;
    std::vector<std::vector<c64>> _syncword_fft_conj;
    float _syncword_self_corr = 0.0f;
    float _best = 0.0f;
    uint64_t _best_idx = 0;
    uint64_t _items_consumed = 0;
    size_t _history_size = 0;
    History<HistoryItem> _history{ 2 };
    Fft _fft;

    // optional debug taps (oracle only): per-sample zpow / winning bin of everything
    // pushed into the history, in stream order
    bool record_metric = false;
    std::vector<float> metric_pow;
    std::vector<int8_t> metric_bin;

    // PM/syncword_detection.hpp:143-202
    void start()
    {
        if (min_freq_bin > max_freq_bin)
            throw std::runtime_error("min_freq_bin is greater than max_freq_bin");
        _syncword_samples_size = (syncword.size() - 1) * samples_per_symbol + rrc_taps.size();
        if (_syncword_samples_size > fft_size) throw std::runtime_error("fft_size too small");
        _fft = Fft(fft_size, fft_kind);

        std::vector<c64> syncword_samples(_syncword_samples_size);
        for (size_t j = 0; j < syncword.size(); ++j) {
            for (size_t k = 0; k < rrc_taps.size(); ++k) {
                // complex<float> * float, then complex +=   (:157-158)
                const c64 p = cscale(constellation[syncword[j]], rrc_taps[k]);
                c64& d = syncword_samples[j * samples_per_symbol + k];
                d = c64(d.real() + p.real(), d.imag() + p.imag());
            }
        }
        _syncword_self_corr = 0.0f;
        for (auto x : syncword_samples) { // :161-164
            _syncword_self_corr += x.real() * x.real() + x.imag() * x.imag();
        }

        _syncword_fft_conj.clear();
        for (int freq_bin = min_freq_bin; freq_bin <= max_freq_bin; ++freq_bin) {
            double phase = 0.0; // :169-182, including the always-taken else-if branch
            const double phase_incr = static_cast<double>(freq_bin) * std::numbers::pi /
                                      static_cast<double>(_syncword_samples_size);
            std::vector<c64> shifted = syncword_samples;
            for (auto& x : shifted) {
                const c64 e{ static_cast<float>(std::cos(phase)),
                             static_cast<float>(std::sin(phase)) };
                x = cmul_plain(x, e);
                phase += phase_incr;
                if (phase >= std::numbers::pi) {
                    phase -= 2.0 * std::numbers::pi;
                } else if (phase < std::numbers::pi) {
                    phase += 2.0 * std::numbers::pi;
                }
            }
            shifted.resize(fft_size);
            std::vector<c64> spec(fft_size);
            _fft.forward(shifted.data(), spec.data());
            for (auto& z : spec) z = std::conj(z);
            _syncword_fft_conj.push_back(std::move(spec));
        }

        _best = 0.0f;
        _best_idx = 0;
        _items_consumed = 0;
        _history_size = 2 * time_threshold + 1;
        _history = History<HistoryItem>(std::bit_ceil(_history_size + 1)); // :198-199
        metric_pow.clear();
        metric_bin.clear();
    }

    // PM/syncword_detection.hpp:56-115
    SyncwordTag output_tag(const HistoryItem& item, const HistoryItem& previous_item,
                           const HistoryItem& next_item) const
    {
        const double bin_spacing =
            std::numbers::pi / static_cast<double>(_syncword_samples_size);
        double syncword_freq = static_cast<double>(item.freq_bin) * bin_spacing;
        float syncword_phase = std::arg(item.correlation);
        float correlation_power;
        if (item.freq_bin > min_freq_bin && item.freq_bin < max_freq_bin) {
            const double a = static_cast<double>(item.correlation_power_left);
            const double b = static_cast<double>(item.correlation_power);
            const double c = static_cast<double>(item.correlation_power_right);
            const double quad = std::clamp((c - a) / (2.0 * (2.0 * b - (a + c))), -0.5, 0.5);
            const double delta_freq = quad * bin_spacing;
            syncword_freq += delta_freq;
            syncword_phase -= static_cast<float>(
                delta_freq * 0.5 * static_cast<double>(_syncword_samples_size));
            if (syncword_phase >= std::numbers::pi_v<float>) {
                syncword_phase -= 2.0f * std::numbers::pi_v<float>;
            } else if (syncword_phase < -std::numbers::pi_v<float>) {
                syncword_phase += 2.0f * std::numbers::pi_v<float>;
            }
            correlation_power =
                static_cast<float>(b + (c - a) * (c - a) / (16.0 * (b - 0.5 * (a + c))));
        } else {
            correlation_power = item.correlation_power;
        }
        const float syncword_amplitude =
            std::sqrt(correlation_power) / (static_cast<float>(fft_size) * _syncword_self_corr);
        const float syncword_power =
            syncword_amplitude * syncword_amplitude * _syncword_self_corr;
        const float esn0_db =
            10.0f * std::log10((syncword_power * static_cast<float>(samples_per_symbol)) /
                               (item.fft_noise_power *
                                static_cast<float>(_syncword_samples_size)));
        const double a = static_cast<double>(previous_item.correlation_power);
        const double b = static_cast<double>(item.correlation_power);
        const double c = static_cast<double>(next_item.correlation_power);
        const float time_est =
            static_cast<float>(std::clamp((c - a) / (2.0 * (2.0 * b - (a + c))), -0.5, 0.5));
        SyncwordTag t{};
        t.amplitude = syncword_amplitude;
        t.phase = syncword_phase;
        t.freq = syncword_freq;
        t.freq_bin = item.freq_bin;
        t.noise_power = item.fft_noise_power;
        t.esn0_db = esn0_db;
        t.time_est = time_est;
        t.corr_re = item.correlation.real();
        t.corr_im = item.correlation.imag();
        t.pow = item.correlation_power;
        t.pow_left = item.correlation_power_left;
        t.pow_right = item.correlation_power_right;
        t.pow_prev = previous_item.correlation_power;
        t.pow_next = next_item.correlation_power;
        return t;
    }

    // PM/syncword_detection.hpp:204-356.  `in`/`out` are the spans offered by the
    // runtime (same length n).  Returns the number of items consumed == published.
    // Tags are appended with ABSOLUTE output indices.
    size_t processBulk(const c64* in, size_t n, c64* out, std::vector<SyncwordTag>& tags)
    {
        if (n < fft_size) return 0; // :215-227 (INSUFFICIENT_INPUT_ITEMS, consume 0)
        const size_t num_freq_bins = static_cast<size_t>(max_freq_bin - min_freq_bin + 1);
        std::vector<c64> samples_fft(fft_size);
        std::vector<c64> samples_fft_prod(fft_size);
        std::vector<std::vector<c64>> correlation(num_freq_bins, std::vector<c64>(fft_size));
        const size_t _stride = fft_size - _syncword_samples_size + 1; // :236
        size_t j;
        for (j = 0; j + fft_size <= n; j += _stride) {
            _fft.forward(in + j, samples_fft.data()); // :239-241
            for (size_t nfreq = 0; nfreq < num_freq_bins; ++nfreq) { // :246-252
                for (size_t k = 0; k < fft_size; ++k) {
                    samples_fft_prod[k] =
                        _fft.cmul(samples_fft[k], _syncword_fft_conj[nfreq][k]);
                }
                _fft.second(samples_fft_prod.data(), correlation[nfreq].data());
            }
            float fft_noise_power = 0.0f; // :257-265
            for (size_t k = fft_size / 4; k < 3 * fft_size / 4; ++k) {
                const auto z = samples_fft[k];
                fft_noise_power += _fft.norm2(z);
            }
            fft_noise_power /=
                static_cast<float>(fft_size / 2) * static_cast<float>(fft_size);

            for (size_t k = 0; k < _stride; ++k) { // :267-343
                const uint64_t curr_idx = _items_consumed + j + k;
                if (curr_idx - _best_idx > time_threshold) {
                    size_t below_threshold = 0;
                    for (size_t u = 0; u < _history_size; ++u) {
                        if (_history[u].correlation_power < _best / power_threshold) {
                            ++below_threshold;
                        }
                    }
                    if (2 * below_threshold >= _history_size) {
                        const size_t hist_idx = _best_idx + _history_size - curr_idx;
                        _history[_history_size - 1 - hist_idx].detection = true;
                    }
                    _best = 0.0f;
                    _best_idx = curr_idx;
                }
                const size_t z_idx = k == 0 ? 0 : fft_size - k; // :300
                size_t best_freq = 0;
                c64 z{};
                float zpow = -1.0f;
                for (size_t nfreq = 0; nfreq < num_freq_bins; ++nfreq) { // :305-313
                    const c64 zz = correlation[nfreq][z_idx];
                    const float zzpow = _fft.norm2(zz);
                    if (zzpow > zpow) {
                        best_freq = nfreq;
                        z = zz;
                        zpow = zzpow;
                    }
                }
                if (zpow > _best) { // :314-317
                    _best = zpow;
                    _best_idx = curr_idx;
                }
                const auto& pop_history = _history[_history_size - 1]; // :318-325
                out[j + k] = pop_history.sample;
                if (pop_history.detection) {
                    SyncwordTag t = output_tag(pop_history, _history[_history_size],
                                               _history[_history_size - 2]);
                    t.index = _items_consumed + j + k;
                    tags.push_back(t);
                }
                HistoryItem item; // :326-342 (left/right stay unset at the edge bins;
                                  // never read because of the guard at :65 — zero here)
                item.sample = in[j + k];
                item.correlation_power = zpow;
                if (best_freq > 0) {
                    item.correlation_power_left = _fft.norm2(correlation[best_freq - 1][z_idx]);
                }
                if (best_freq < num_freq_bins - 1) {
                    item.correlation_power_right =
                        _fft.norm2(correlation[best_freq + 1][z_idx]);
                }
                item.correlation = z;
                item.freq_bin = min_freq_bin + static_cast<int>(best_freq);
                item.fft_noise_power = fft_noise_power;
                _history.push_back(item);
                if (record_metric) {
                    metric_pow.push_back(zpow);
                    metric_bin.push_back(static_cast<int8_t>(item.freq_bin));
                }
            }
        }
        _items_consumed += j; // :346-350
        return j;
    }
};

// ----------------------------------------------------------------------------
// Rotator — PM/rotator.hpp:44-65
// ----------------------------------------------------------------------------
class Rotator
{
public:
    float phase_incr = 0.0f;
    c64 _exp{ 1.0f, 0.0f };
    c64 _exp_incr{ 1.0f, 0.0f };
    unsigned _counter = 0;
    void settingsChanged() { _exp_incr = { std::cos(phase_incr), std::sin(phase_incr) }; }
    void start()
    {
        _exp = { 1.0f, 0.0f };
        _counter = 0;
    }
    c64 processOne(c64 a)
    {
        const c64 z = cmul_plain(a, _exp);
        _exp = cmul_plain(_exp, _exp_incr);
        if ((++_counter % 512) == 0) {
            // std::abs(complex<float>) is hypotf; complex / float divides both parts
            const float m = std::hypot(_exp.real(), _exp.imag());
            _exp = c64(_exp.real() / m, _exp.imag() / m);
        }
        return z;
    }
};

// ----------------------------------------------------------------------------
// CoarseFrequencyCorrection<float> — PM/coarse_frequency_correction.hpp:40-98
// (sits between SyncwordDetectionFilter and SymbolFilter,
//  PM/packet_receiver.hpp:94-95, 195-202.)  processBulk() takes one chunk whose
// FIRST sample may carry a tag with a "syncword_freq" key (has_freq / freq).
// ----------------------------------------------------------------------------
class CoarseFrequencyCorrection
{
public:
    size_t delay = 0;
    c64 _exp{ 1.0f, 0.0f };
    c64 _exp_incr{ 1.0f, 0.0f };
    unsigned _counter = 0;
    float _next_freq = 0.0f;
    std::ptrdiff_t _next_freq_delay = 0;

    void set_freq(float freq) // :50-59
    {
        _exp = { std::cos(freq * static_cast<float>(delay)), -std::sin(freq * static_cast<float>(delay)) };
        _exp_incr = { std::cos(freq), -std::sin(freq) };
        _counter = 0;
    }
    void processBulk(const c64* in, size_t n, c64* out, bool has_freq, double freq) // :67-98
    {
        if (has_freq) { // pmtv::cast<float>(tag.map.at("syncword_freq")): the tag holds a double
            _next_freq = static_cast<float>(freq);
            _next_freq_delay = static_cast<std::ptrdiff_t>(delay);
        }
        for (size_t j = 0; j < n; ++j) {
            if (_next_freq_delay == 0) set_freq(_next_freq);
            out[j] = cmul_plain(in[j], _exp);
            _exp = cmul_plain(_exp, _exp_incr);
            if ((++_counter % 512) == 0) {
                const float m = std::hypot(_exp.real(), _exp.imag());
                _exp = c64(_exp.real() / m, _exp.imag() / m);
            }
            if (_next_freq_delay >= 0) --_next_freq_delay;
        }
    }
};

// ----------------------------------------------------------------------------
// SyncwordWipeoff<c64, float> — PM/syncword_wipeoff.hpp:38-91.
// processBulk() takes one chunk whose FIRST sample may carry a tag with a
// "syncword_amplitude" key (has_tag).
// ----------------------------------------------------------------------------
class SyncwordWipeoff
{
public:
    std::vector<float> syncword;
    bool _in_syncword = false;
    size_t _position = 0;

    void processBulk(const c64* in, size_t n, c64* out, bool has_tag)
    {
        if (!_in_syncword && has_tag) { // :52-61
            _in_syncword = true;
            _position = 0;
        }
        size_t j = 0;
        if (_in_syncword) { // :65-75
            const size_t m = std::min(n, syncword.size() - _position);
            for (; j < m; ++j) {
                // std::complex<float> * float scales both parts
                out[j] = c64(in[j].real() * syncword[_position], in[j].imag() * syncword[_position]);
                ++_position;
            }
            if (_position == syncword.size()) _in_syncword = false;
        }
        if (!_in_syncword) { // :77-82
            for (; j < n; ++j) out[j] = in[j];
        }
    }
};

// ----------------------------------------------------------------------------
// CostasLoop<float, float> — PM/costas_loop.hpp:56-149.
// processBulk() takes one chunk whose FIRST sample may carry a tag with a
// "syncword_phase" key (has_phase / phase).
//
// Two trig arithmetics (like the two FFT arithmetics of oracle_fft.hpp):
//   Libm   — std::cos / std::sin, what the reference calls (:113-114);
//   Mirror — the GPU's arithmetic contract (csrc/costas.cuh: b200_sincosf), restated here
//            op for op so that the CUDA kernel can be pinned bit-for-bit.
// ----------------------------------------------------------------------------
enum class TrigKind : int { Libm = 0, Mirror = 1 };

inline void mirror_sincosf(float x, float& s, float& c)
{
    const float qb = x * 0.636619772367581343f + 12582912.0f; // round to nearest even by adding 1.5 * 2^23
    const float q = qb - 12582912.0f;
    float r = std::fmaf(q, -1.5703125f, x);
    r = std::fmaf(q, -4.837512969970703125e-4f, r);
    r = std::fmaf(q, -7.54978995489188216e-8f, r);
    const float z = r * r;
    float ps = std::fmaf(-1.9515295891e-4f, z, 8.3321608736e-3f);
    ps = std::fmaf(ps, z, -1.6666654611e-1f);
    ps = ps * z;
    const float sr = std::fmaf(ps, r, r);
    float pc = std::fmaf(2.443315711809948e-5f, z, -1.388731625493765e-3f);
    pc = std::fmaf(pc, z, 4.166664568298827e-2f);
    pc = pc * (z * z);
    const float cr = std::fmaf(-0.5f, z, 1.0f) + pc;
    const int n = static_cast<int>(std::bit_cast<uint32_t>(qb) & 3u);
    const float s0 = (n & 1) ? cr : sr;
    const float c0 = (n & 1) ? sr : cr;
    s = (n & 2) ? -s0 : s0;
    c = ((n + 1) & 2) ? -c0 : c0;
}

class CostasLoop
{
public:
    enum Constellation : int { PILOT = 0, BPSK = 1, QPSK = 2 }; // PM/constellation.hpp
    double loop_bandwidth = 0.01;
    Constellation constellation = BPSK;
    TrigKind trig = TrigKind::Libm;
    float _phase = 0, _freq = 0, _k1 = 0, _k2 = 0;

    void settingsChanged() // :56-90
    {
        double discriminant_gain = 1.0;
        if (constellation == QPSK) discriminant_gain = std::numbers::sqrt2;
        const double loop_bandwidth_2 = loop_bandwidth * loop_bandwidth;
        const double loop_bandwidth_3 = loop_bandwidth_2 * loop_bandwidth;
        const double loop_bandwidth_4 = loop_bandwidth_2 * loop_bandwidth_2;
        const double s = std::cbrt(36.0 * loop_bandwidth_2 +
                                   std::sqrt(3.0) * std::sqrt(432.0 * loop_bandwidth_4 + 848.0 * loop_bandwidth_3 +
                                                              624.0 * loop_bandwidth_2 + 204.0 * loop_bandwidth + 25.0) +
                                   36.0 * loop_bandwidth + 9.0);
        const double z = -(-12.0 * loop_bandwidth - 6.0) / (3.0 * std::cbrt(6.0) * (2.0 * loop_bandwidth + 1.0) * s) +
                         (std::cbrt(2.0) * s) / (std::cbrt(9.0) * (2.0 * loop_bandwidth + 1.0)) - 1.0;
        const double k1 = 1.0 - z * z;
        const double k2 = (1.0 - z) * (1.0 - z);
        _k1 = static_cast<float>(k1 / discriminant_gain);
        _k2 = static_cast<float>(k2 / discriminant_gain);
    }

    void processBulk(const c64* in, size_t n, c64* out, bool has_phase, float phase) // :94-149
    {
        if (has_phase) { // set_phase(), :38-45
            _phase = phase;
            _freq = 0;
        }
        for (size_t j = 0; j < n; ++j) {
            float sn, cs;
            if (trig == TrigKind::Libm) {
                cs = std::cos(_phase);
                sn = std::sin(_phase);
            } else {
                mirror_sincosf(_phase, sn, cs);
            }
            const c64 lo{ cs, -sn };
            const c64 z_out = cmul_plain(in[j], lo);
            out[j] = z_out;
            float error = 0;
            switch (constellation) {
            case PILOT: error = z_out.imag(); break;
            case BPSK: error = z_out.real() * z_out.imag(); break;
            case QPSK:
                error = (z_out.real() > 0 ? z_out.imag() : -z_out.imag()) +
                        (z_out.imag() > 0 ? -z_out.real() : z_out.real());
                break;
            }
            _freq += _k2 * error;
            _phase += _k1 * error + _freq;
            if (_phase >= std::numbers::pi_v<float>) {
                _phase -= 2.0f * std::numbers::pi_v<float>;
            } else if (_phase < -std::numbers::pi_v<float>) {
                _phase += 2.0f * std::numbers::pi_v<float>;
            }
        }
    }
};

// ----------------------------------------------------------------------------
// PfbArbResampler<c64, c64, float, TRate> — PM/pfb_arb_resampler.hpp:67-182
// ----------------------------------------------------------------------------
template <typename TRate>
class PfbArbResampler
{
public:
    TRate rate{ 1.0 };
    std::vector<float> taps;
    size_t filter_size = 32;

    std::vector<std::vector<float>> _taps, _diff_taps;
    size_t _arm_size = 0;
    History<c64> _history{ 1 };
    size_t _decim_rate = 0;
    TRate _filt_rate{};
    size_t _last_filter = 0;
    TRate _phase_acc{};

    void settingsChanged() // :67-120
    {
        if (filter_size == 0) throw std::runtime_error("filter_size cannot be 0");
        _arm_size = (taps.size() + filter_size - 1) / filter_size;
        _taps.assign(filter_size, {});
        for (size_t j = 0; j < filter_size; ++j) {
            for (size_t k = j; k < taps.size(); k += filter_size) _taps[j].push_back(taps[k]);
            while (_taps[j].size() < _arm_size) _taps[j].push_back(0.0f);
        }
        _diff_taps.assign(filter_size, {});
        for (size_t j = 0; j < filter_size; ++j) {
            for (size_t k = j; k < taps.size() - 1; k += filter_size)
                _diff_taps[j].push_back(taps[k + 1] - taps[k]);
            while (_diff_taps[j].size() < _arm_size) _diff_taps[j].push_back(0.0f);
        }
        const size_t capacity = std::bit_ceil(_arm_size);
        History<c64> nh(capacity);
        for (size_t i = 0; i < capacity; ++i) nh.push_back(c64{});
        for (std::ptrdiff_t j = static_cast<std::ptrdiff_t>(_history.size()) - 1; j >= 0; --j)
            nh.push_back(_history[static_cast<size_t>(j)]);
        _history = nh;
        const TRate float_rate = static_cast<TRate>(filter_size) / rate;
        _decim_rate = static_cast<size_t>(std::floor(float_rate));
        _filt_rate = float_rate - static_cast<TRate>(_decim_rate);
        _phase_acc = TRate{ 0 };
        _last_filter = (taps.size() / 2) % filter_size;
    }

    // std::inner_product(taps, taps+arm, history.cbegin(), c64{0}):
    // acc = acc + tap * sample   (float * complex<float>, then complex +)
    c64 dot(const std::vector<float>& t) const
    {
        c64 acc{ 0.0f, 0.0f };
        const c64* h = _history.newest();
        for (size_t k = 0; k < _arm_size; ++k) {
            const c64 p = cscale(h[k], t[k]);
            acc = c64(acc.real() + p.real(), acc.imag() + p.imag());
        }
        return acc;
    }

    // :122-182.  Optionally records (input items consumed so far, arm, phase_acc)
    // for every output so closed-form timing can be checked exactly.
    void processBulk(const c64* in, size_t n_in, c64* out, size_t n_out, size_t& consumed,
                     size_t& produced, std::vector<uint32_t>* arms = nullptr,
                     std::vector<uint64_t>* in_counts = nullptr,
                     std::vector<double>* accs = nullptr, uint64_t in_base = 0)
    {
        size_t ii = 0, oo = 0;
        while (ii < n_in && oo < n_out) {
            while (_last_filter >= filter_size && ii < n_in) {
                _history.push_back(in[ii++]);
                _last_filter -= filter_size;
            }
            if (_last_filter >= filter_size) break;
            const c64 filt_out = dot(_taps[_last_filter]);
            const c64 diff_out = dot(_diff_taps[_last_filter]);
            const float pa = static_cast<float>(_phase_acc);
            const c64 diff_scaled = c64(pa * diff_out.real(), pa * diff_out.imag());
            out[oo++] = c64(filt_out.real() + diff_scaled.real(),
                            filt_out.imag() + diff_scaled.imag());
            if (arms) arms->push_back(static_cast<uint32_t>(_last_filter));
            if (in_counts) in_counts->push_back(in_base + ii);
            if (accs) accs->push_back(static_cast<double>(_phase_acc));
            _phase_acc += _filt_rate;
            _last_filter += _decim_rate;
            if (_phase_acc > TRate{ 1 }) {
                _phase_acc -= TRate{ 1 };
                ++_last_filter;
            }
        }
        consumed = ii;
        produced = oo;
    }
};

// ----------------------------------------------------------------------------
// Generic tag: the keys the hot path touches, plus a free-form "other" marker so
// non-syncword keys can be followed through SyncwordDetectionFilter/SymbolFilter.
// ----------------------------------------------------------------------------
struct StreamTag {
    int64_t index = 0;       // meaning depends on context (see users)
    bool has_syncword = false;
    float amplitude = 0.0f;
    float phase = 0.0f;
    double freq = 0.0;
    int32_t freq_bin = 0;
    float noise_power = 0.0f;
    float esn0_db = 0.0f;
    float time_est = 0.0f;
    int32_t other = 0;       // !=0: carries a non-syncword key (value = id)
};

// ----------------------------------------------------------------------------
// SymbolFilter<c64,c64,float> — PM/symbol_filter.hpp:64-252
// The GR4 runtime presents a chunk whose FIRST sample carries the merged input
// tag (GR/Block.hpp:1501-1508).  processBulk() here takes one such chunk and an
// optional tag for its first sample.
// ----------------------------------------------------------------------------
class SymbolFilter
{
public:
    size_t samples_per_symbol = 4;
    std::vector<float> taps;
    size_t num_arms = 32;
    size_t delay = 0;

    std::vector<std::vector<float>> _taps;
    History<c64> _history{ 1 };
    size_t _clock_phase = 0;
    size_t _reset_clock_phase = 0;
    size_t _pfb_arm = 0;
    std::vector<StreamTag> _tags; // .index = countdown, as gr::Tag::index in the reference
    float _scale = 1.0f;

    void settingsChanged() // :64-108
    {
        if (samples_per_symbol == 0) throw std::runtime_error("samples_per_symbol cannot be zero");
        if (num_arms == 0) throw std::runtime_error("num_arms cannot be zero");
        _taps.assign(num_arms, {});
        for (size_t j = 0; j < num_arms; ++j)
            for (size_t k = j; k < taps.size(); k += num_arms) _taps[j].push_back(taps[k]);
        const size_t arm_size = _taps[0].size();
        const size_t capacity = std::bit_ceil(arm_size);
        History<c64> nh(capacity);
        for (size_t i = 0; i < capacity; ++i) nh.push_back(c64{});
        for (std::ptrdiff_t j = static_cast<std::ptrdiff_t>(_history.size()) - 1; j >= 0; --j)
            nh.push_back(_history[static_cast<size_t>(j)]);
        _history = nh;
        _reset_clock_phase =
            (samples_per_symbol - (delay % samples_per_symbol)) % samples_per_symbol;
    }
    void start() { _clock_phase = 0; }

    c64 filt() const // _scale * inner_product(taps[arm], history)   (:164-167, 211-214)
    {
        c64 acc{ 0.0f, 0.0f };
        const c64* h = _history.newest();
        const auto& t = _taps[_pfb_arm];
        for (size_t k = 0; k < t.size(); ++k) {
            const c64 p = cscale(h[k], t[k]);
            acc = c64(acc.real() + p.real(), acc.imag() + p.imag());
        }
        return cscale(acc, _scale);
    }

    // out_tags get .index = output index relative to `out` of this call
    void processBulk(const c64* in, size_t n_in, c64* out, size_t n_out, const StreamTag* tag_in,
                     size_t& consumed, size_t& produced, std::vector<StreamTag>& out_tags)
    {
        size_t oo = 0, ii = 0;
        auto flush_tags = [&]() { // :171-183, 218-228
            while (!_tags.empty() &&
                   _tags[0].index < static_cast<int64_t>(samples_per_symbol / 2)) {
                StreamTag t = _tags[0];
                t.index = static_cast<int64_t>(oo);
                out_tags.push_back(t);
                _tags.erase(_tags.begin());
            }
        };
        if (tag_in) { // :127-206
            StreamTag tag = *tag_in;
            int64_t tag_index_adjust = 0;
            if (tag.has_syncword) {
                size_t new_clock_phase = _reset_clock_phase;
                _scale = 1.0f / tag.amplitude;
                float time_est = tag.time_est;
                if (time_est < 0.0f) { // :146-156
                    new_clock_phase = (new_clock_phase + 1) % samples_per_symbol;
                    time_est += 1.0f;
                    tag.phase = static_cast<float>(static_cast<double>(tag.phase) - tag.freq);
                }
                if (_clock_phase == 0 && new_clock_phase == 1) { // :160-189
                    _history.push_back(in[ii++]);
                    out[oo] = filt();
                    flush_tags();
                    ++oo;
                    ++new_clock_phase;
                    for (auto& t : _tags) --t.index;
                    tag_index_adjust = -1;
                } else if (_clock_phase == 1 && new_clock_phase == 0) { // :192-195
                    _history.push_back(in[ii++]);
                    ++new_clock_phase;
                }
                _clock_phase = new_clock_phase;
                _pfb_arm = std::clamp(
                    static_cast<size_t>(std::round(static_cast<float>(num_arms) * time_est)),
                    size_t{ 0 }, num_arms - 1); // :199-202
            }
            tag.index = static_cast<int64_t>(delay) + tag_index_adjust; // :204-205
            _tags.push_back(tag);
        }
        while (oo < n_out && ii < n_in) { // :208-241
            _history.push_back(in[ii++]);
            if (_clock_phase == 0) {
                out[oo] = filt();
                flush_tags();
                ++oo;
            }
            ++_clock_phase;
            if (_clock_phase >= samples_per_symbol) _clock_phase = 0;
            for (auto& t : _tags) --t.index;
        }
        consumed = ii;
        produced = oo;
    }
};

// ----------------------------------------------------------------------------
// SyncwordDetectionFilter — PM/syncword_detection_filter.hpp:54-210
// Messages are reduced to what the block reads: invalid_header flag or
// packet_length on `parsed_header`; presence only on `ignored_syncword`.
// ----------------------------------------------------------------------------
struct HeaderMsg {
    bool invalid_header = false;
    uint64_t packet_length = 0;
};

class SyncwordDetectionFilter
{
public:
    size_t samples_per_symbol = 4;
    size_t syncword_size = 64;
    size_t header_size = 128;
    size_t allowed_margin = 16;
    bool _in_packet = false;
    size_t _position = 0;
    size_t _block_until = 0;
    void start() { _in_packet = false; }

    // One work() call.  `tag_in` (optional) sits on the first input sample.
    // Returns number of stream items consumed == published; *hdr_used / *ign_used tell
    // how many messages (0/1) were consumed.  If a tag is forwarded it is written to
    // *tag_out (published at output offset 0, :105) and *tag_forwarded is set.
    size_t processBulk(const HeaderMsg* headers, size_t n_headers, size_t n_ignored,
                       const c64* in, size_t n_in, c64* out, size_t n_out,
                       const StreamTag* tag_in, size_t& hdr_used, size_t& ign_used,
                       StreamTag* tag_out, bool& tag_forwarded)
    {
        hdr_used = ign_used = 0;
        tag_forwarded = false;
        if (tag_in) { // :76-108
            StreamTag o{};
            bool any = false;
            bool new_in_packet = false;
            if (tag_in->has_syncword) {
                if (!_in_packet) {
                    new_in_packet = true;
                    o = *tag_in;
                    o.other = 0;
                    any = true;
                }
            }
            if (tag_in->other != 0) {
                o.other = tag_in->other;
                if (!any) o.has_syncword = false;
                any = true;
            }
            if (new_in_packet) {
                _in_packet = true;
                _position = 0;
                _block_until = 0;
            }
            if (any) {
                *tag_out = o;
                tag_out->index = 0;
                tag_forwarded = true;
            }
        }
        if (!_in_packet) { // :110-132
            const size_t n = std::min(n_in, n_out);
            std::copy_n(in, n, out);
            return n;
        }
        if (_block_until == 0 && n_headers > 0) { // :136-154
            hdr_used = 1;
            if (headers[0].invalid_header) {
                _block_until = 1;
            } else {
                const uint64_t packet_length = headers[0].packet_length;
                if (packet_length == 0) throw std::runtime_error("received packet_length = 0");
                constexpr size_t crc_size_bytes = 4;
                const size_t payload_symbols = (packet_length + crc_size_bytes) * 4;
                _block_until = samples_per_symbol *
                               (header_size + syncword_size - allowed_margin + payload_symbols);
            }
        }
        if (_block_until == 0 && n_ignored > 0) { // :158-161
            ign_used = 1;
            _block_until = 1;
        }
        size_t consumed = 0;
        const size_t allowed = samples_per_symbol * (syncword_size + header_size + allowed_margin);
        if (_position < allowed) { // :166-172
            const size_t n = std::min({ n_in, n_out, allowed - _position });
            std::copy_n(in, n, out);
            _position += n;
            consumed = n;
        }
        if (_position >= allowed && _block_until != 0) { // :174-185
            const size_t n = std::min(n_in, n_out) - consumed;
            std::copy_n(in + consumed, n, out + consumed);
            _position += n;
            consumed += n;
            if (_position >= _block_until) _in_packet = false;
        }
        return consumed;
    }
};

// ----------------------------------------------------------------------------
// InterpolatingFirFilter<c64,c64,float> — PM/interpolating_fir_filter.hpp:40-99.
// Only used to synthesise stimulus for the oracle-side QA checks.
// ----------------------------------------------------------------------------
inline void interpolating_fir(const std::vector<float>& taps, size_t interpolation, const c64* in,
                              size_t n_in, c64* out)
{
    std::vector<std::vector<float>> poly(interpolation);
    for (size_t j = 0; j < interpolation; ++j)
        for (size_t k = j; k < taps.size(); k += interpolation) poly[j].push_back(taps[k]);
    const size_t capacity = std::bit_ceil((taps.size() + interpolation - 1) / interpolation);
    History<c64> h(capacity);
    for (size_t i = 0; i < capacity; ++i) h.push_back(c64{});
    size_t oo = 0;
    for (size_t i = 0; i < n_in; ++i) {
        h.push_back(in[i]);
        const c64* hist = h.newest();
        for (const auto& branch : poly) {
            c64 acc{ 0.0f, 0.0f };
            for (size_t k = 0; k < branch.size(); ++k) {
                const c64 p = cscale(hist[k], branch[k]);
                acc = c64(acc.real() + p.real(), acc.imag() + p.imag());
            }
            out[oo++] = acc;
        }
    }
}

} // namespace orc
