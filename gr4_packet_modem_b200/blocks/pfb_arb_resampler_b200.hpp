// pfb_arb_resampler_b200.hpp — drop-in shells for gr::packet_modem::PfbArbResampler<c64, c64, float, float>
// (PM/pfb_arb_resampler.hpp) and gr::packet_modem::Rotator<float> (PM/rotator.hpp) running on a B200
// through libb200sync.so, and for the two of them fused into one kernel launch per call
// (RxFrontEndB200 — the SFO/CFO conditioning pair of apps/packet_transceiver.cpp:71-75).
//
// Settings carry the reference names, types and defaults (PM/pfb_arb_resampler.hpp:61-65 `rate`, `taps`,
// `filter_size`; PM/rotator.hpp:42 `phase_incr`); reflection lists as :187-188 and rotator.hpp:70.
// All DSP is behind the C ABI (b200sync_fe_*); the shell only moves spans and counts.
#pragma once
#include <type_traits>

#include "b200_shell_common.hpp"

namespace gr::packet_modem {

namespace b200_detail {
// one b200sync_fe context configured from the reflected settings of either block (or both)
struct FrontEndCtx {
    b200sync_fe* ctx = nullptr;
    bool fp_contract = false;   // opt-in: fused multiply-add per tap (2x faster, not bit-identical to the reference's float order)
    FrontEndCtx() = default;
    FrontEndCtx(const FrontEndCtx&) = delete;
    FrontEndCtx& operator=(const FrontEndCtx&) = delete;
    ~FrontEndCtx() { b200sync_fe_destroy(ctx); }

    // TRate = float or double, as the reference's fourth template argument (PM/pfb_arb_resampler.hpp:14-22)
    template <typename TRate>
    void configure(TRate rate, const std::vector<float>& taps, size_t filter_size, float phase_incr, bool resampler,
                   bool rotator, int device)
    {
        b200sync_fe_destroy(ctx);
        ctx = nullptr;
        b200sync_fe_config cfg{};
        cfg.rate = static_cast<float>(rate);
        cfg.rate_is_f64 = std::is_same_v<TRate, double> ? 1u : 0u;
        cfg.rate_f64 = static_cast<double>(rate);
        cfg.taps = taps.empty() ? nullptr : taps.data();
        cfg.n_taps = static_cast<uint32_t>(taps.size());
        cfg.filter_size = static_cast<uint32_t>(filter_size);
        cfg.phase_incr = phase_incr;
        cfg.enable_resampler = resampler ? 1u : 0u;
        cfg.enable_rotator = rotator ? 1u : 0u;
        cfg.device = device;
        cfg.fp_contract = fp_contract ? 1u : 0u;
        // "filter_size cannot be 0" (PM/pfb_arb_resampler.hpp:70-72) comes back as the error text
        if (b200sync_fe_create(&cfg, &ctx) != 0) throw gr::exception(b200sync_fe_last_error());
    }

    // processBulk body shared by the three shells: PM/pfb_arb_resampler.hpp:122-182 (+ rotator.hpp:56-65)
    template <typename TIn, typename TOut>
    gr::work::Status process(const TIn& inSpan, TOut& outSpan)
    {
        if (!ctx) throw gr::exception("processBulk() before settingsChanged()/start()");
        size_t consumed = 0, produced = 0;
        if (b200sync_fe_process(ctx, reinterpret_cast<const float*>(inSpan.data()), inSpan.size(),
                                reinterpret_cast<float*>(outSpan.data()), outSpan.size(), &consumed, &produced) != 0)
            throw gr::exception(b200sync_fe_last_error());
        if (!inSpan.consume(consumed)) throw gr::exception("consume failed");  // :168-170
        outSpan.publish(produced);
        return gr::work::Status::OK;
    }
};
}  // namespace b200_detail

// TRate: the reference's PfbArbResampler<TIn, TOut, TTaps, TRate> with TIn = TOut = c64, TTaps = float
template <typename TRate = float>
class PfbArbResamplerB200T
#if B200SYNC_HAVE_GR4
    : public gr::Block<PfbArbResamplerB200T<TRate>>
#else
    : public gr::BlockShim<PfbArbResamplerB200T<TRate>>
#endif
{
    b200_detail::FrontEndCtx _fe;

public:
#if B200SYNC_HAVE_GR4
    // both ports Async: the ratio is not a fraction (PM/pfb_arb_resampler.hpp:59-62)
    gr::PortIn<std::complex<float>, gr::Async> in;
    gr::PortOut<std::complex<float>, gr::Async> out;
#else
    gr::PortInShim<std::complex<float>> in;
    gr::PortOutShim<std::complex<float>> out;
#endif
    TRate rate{ 1.0 };
    std::vector<float> taps;
    size_t filter_size = 32;
    int device = 0;  // extra: CUDA device ordinal

    void settingsChanged(const gr::property_map& /* old_settings */, const gr::property_map& /* new_settings */)
    {
        _fe.template configure<TRate>(rate, taps, filter_size, 0.0f, true, false, device);
    }

    template <typename TIn, typename TOut>
    gr::work::Status processBulk(const TIn& inSpan, TOut& outSpan)
    {
        return _fe.process(inSpan, outSpan);
    }
};
using PfbArbResamplerB200 = PfbArbResamplerB200T<float>;
using PfbArbResamplerB200Double = PfbArbResamplerB200T<double>;   // test/qa_pfb_arb_resampler.cpp:45-69

// The reference Rotator is a processOne block; a GPU block works on spans, so the shell exposes
// processBulk (same items out as items in, same tags: default forwarding policy).  The NCO phase is
// evaluated in closed form per item instead of by the float recurrence of PM/rotator.hpp:58-63
// (tolerance stated in tests/test_gpu_frontend.py::test_rotator_tolerance).
class RotatorB200
#if B200SYNC_HAVE_GR4
    : public gr::Block<RotatorB200>
#else
    : public gr::BlockShim<RotatorB200>
#endif
{
    b200_detail::FrontEndCtx _fe;

public:
#if B200SYNC_HAVE_GR4
    gr::PortIn<std::complex<float>> in;
    gr::PortOut<std::complex<float>> out;
#else
    gr::PortInShim<std::complex<float>> in;
    gr::PortOutShim<std::complex<float>> out;
#endif
    float phase_incr = 0;  // rad / sample (PM/rotator.hpp:41-42)
    int device = 0;

    void settingsChanged(const gr::property_map& /* old_settings */, const gr::property_map& /* new_settings */)
    {
        _fe.configure(1.0f, {}, 32, phase_incr, false, true, device);
    }

    // PM/rotator.hpp:50-54: phase back to zero
    void start()
    {
        if (!_fe.ctx) _fe.configure(1.0f, {}, 32, phase_incr, false, true, device);
        else if (b200sync_fe_start(_fe.ctx) != 0) throw gr::exception(b200sync_fe_last_error());
    }

    template <typename TIn, typename TOut>
    gr::work::Status processBulk(const TIn& inSpan, TOut& outSpan)
    {
        return _fe.process(inSpan, outSpan);
    }
};

// PfbArbResampler immediately followed by Rotator as ONE block (one kernel, 16 B/sample instead of 32):
// replaces the pair `resampler -> rotator` of apps/packet_transceiver.cpp:71-75 when both run on the GPU.
class RxFrontEndB200
#if B200SYNC_HAVE_GR4
    : public gr::Block<RxFrontEndB200>
#else
    : public gr::BlockShim<RxFrontEndB200>
#endif
{
    b200_detail::FrontEndCtx _fe;

public:
#if B200SYNC_HAVE_GR4
    gr::PortIn<std::complex<float>, gr::Async> in;
    gr::PortOut<std::complex<float>, gr::Async> out;
#else
    gr::PortInShim<std::complex<float>> in;
    gr::PortOutShim<std::complex<float>> out;
#endif
    float rate{ 1.0 };
    std::vector<float> taps;
    size_t filter_size = 32;
    float phase_incr = 0;
    int device = 0;

    void settingsChanged(const gr::property_map& /* old_settings */, const gr::property_map& /* new_settings */)
    {
        _fe.configure(rate, taps, filter_size, phase_incr, true, true, device);
    }

    template <typename TIn, typename TOut>
    gr::work::Status processBulk(const TIn& inSpan, TOut& outSpan)
    {
        return _fe.process(inSpan, outSpan);
    }
};

}  // namespace gr::packet_modem

#if B200SYNC_HAVE_GR4
ENABLE_REFLECTION_FOR_TEMPLATE(gr::packet_modem::PfbArbResamplerB200T, in, out, rate, taps, filter_size, device);
ENABLE_REFLECTION(gr::packet_modem::RotatorB200, in, out, phase_incr, device);
ENABLE_REFLECTION(gr::packet_modem::RxFrontEndB200, in, out, rate, taps, filter_size, phase_incr, device);
#endif
