// coarse_frequency_correction_b200.hpp — drop-in shell for gr::packet_modem::CoarseFrequencyCorrection<float>
// (PM/coarse_frequency_correction.hpp) running on a B200 through libb200sync.so.
//
// Same setting as the reference (`delay`, :44; reflection list :103-106) and the same default tag
// policy: input tags are forwarded unchanged to the first output sample of the chunk by the runtime
// (GR/Block.hpp:777-790).  A tag with a "syncword_freq" key resets the block's rotator `delay`
// samples later (:73-83); the NCO itself runs behind b200sync_cfc_process.
//
// In the receiver this block feeds SymbolFilter (PM/packet_receiver.hpp:195-202); SymbolFilterB200
// can absorb it (`fused_cfc_delay` setting) so that the samples make one pass through the GPU
// instead of two — then this block is simply left out of the flowgraph.
#pragma once
#include "b200_shell_common.hpp"

namespace gr::packet_modem {

class CoarseFrequencyCorrectionB200
#if B200SYNC_HAVE_GR4
    : public gr::Block<CoarseFrequencyCorrectionB200>
#else
    : public gr::BlockShim<CoarseFrequencyCorrectionB200>
#endif
{
    b200sync_cfc* _ctx = nullptr;

    void configure()
    {
        b200sync_cfc_destroy(_ctx);
        _ctx = nullptr;
        if (b200sync_cfc_create(static_cast<uint32_t>(delay), device, &_ctx) != 0)
            throw gr::exception(b200sync_cfc_last_error());
    }

public:
#if B200SYNC_HAVE_GR4
    gr::PortIn<std::complex<float>> in;
    gr::PortOut<std::complex<float>> out;
#else
    gr::PortInShim<std::complex<float>> in;
    gr::PortOutShim<std::complex<float>> out;
#endif
    size_t delay = 0;
    int device = 0;  // extra: CUDA device ordinal

    CoarseFrequencyCorrectionB200() = default;
    CoarseFrequencyCorrectionB200(const CoarseFrequencyCorrectionB200&) = delete;
    CoarseFrequencyCorrectionB200& operator=(const CoarseFrequencyCorrectionB200&) = delete;
    ~CoarseFrequencyCorrectionB200() { b200sync_cfc_destroy(_ctx); }

    void settingsChanged(const gr::property_map& /* old_settings */, const gr::property_map& /* new_settings */)
    {
        configure();
    }

    void start()
    {
        if (!_ctx) configure();
        else if (b200sync_cfc_start(_ctx) != 0) throw gr::exception(b200sync_cfc_last_error());
    }

    // PM/coarse_frequency_correction.hpp:67-98
    template <typename TIn, typename TOut>
    gr::work::Status processBulk(const TIn& inSpan, TOut& outSpan)
    {
        if (!_ctx) throw gr::exception("processBulk() before settingsChanged()/start()");
        const size_t n = std::min(inSpan.size(), outSpan.size());
        b200sync_stream_tag tin{};
        size_t n_tin = 0;
        if (this->input_tags_present()) {
            const auto& map = this->mergedInputTag().map;
            if (map.contains("syncword_freq")) {  // :75
                tin.index = 0;
                tin.has_syncword = 1;
                tin.sw.syncword_freq = b200sync_shell::pmt_cast<double>(map.at("syncword_freq"));
                n_tin = 1;
            }
#if !B200SYNC_HAVE_GR4
            // what the runtime's default tag policy does (GR/Block.hpp:777-790); real GR4 does it itself
            if (n > 0) out.publishTag(map, 0);
#endif
        }
        if (n > 0 && b200sync_cfc_process(_ctx, reinterpret_cast<const float*>(inSpan.data()), n, n_tin ? &tin : nullptr,
                                          n_tin, reinterpret_cast<float*>(outSpan.data())) != 0)
            throw gr::exception(b200sync_cfc_last_error());
        if (!inSpan.consume(n)) throw gr::exception("consume failed");
        outSpan.publish(n);
        return gr::work::Status::OK;
    }
};

}  // namespace gr::packet_modem

#if B200SYNC_HAVE_GR4
ENABLE_REFLECTION(gr::packet_modem::CoarseFrequencyCorrectionB200, in, out, delay, device);
#endif
