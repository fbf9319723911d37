// syncword_wipeoff_b200.hpp — drop-in shell for gr::packet_modem::SyncwordWipeoff<std::complex<float>, float>
// (PM/syncword_wipeoff.hpp) running on a B200 through libb200sync.so.
//
// Same setting as the reference (`syncword`, :36; reflection list :96) and the same default tag policy:
// input tags are forwarded unchanged to the first output item of the chunk by the runtime
// (GR/Block.hpp:777-790).  A tag with a "syncword_amplitude" key that arrives while no syncword is being
// wiped starts one (:52-61); the multiply itself runs behind b200sync_wo_process.
//
// In the receiver this block feeds CostasLoop through PayloadMetadataInsert (PM/packet_receiver.hpp:203-214), which
// forwards the syncword / header / payload symbols and the syncword tag and drops the inter-packet remainder.
// CostasLoopB200 can absorb this block (`fused_wipeoff_syncword` setting: the wipe-off then happens behind
// PayloadMetadataInsert, on the same 64 symbols), and this block is left out of the flowgraph.
#pragma once
#include "b200_shell_common.hpp"

namespace gr::packet_modem {

class SyncwordWipeoffB200
#if B200SYNC_HAVE_GR4
    : public gr::Block<SyncwordWipeoffB200>
#else
    : public gr::BlockShim<SyncwordWipeoffB200>
#endif
{
    b200sync_wo* _ctx = nullptr;

    void configure()
    {
        b200sync_wo_destroy(_ctx);
        _ctx = nullptr;
        if (b200sync_wo_create(syncword.data(), static_cast<uint32_t>(syncword.size()), device, &_ctx) != 0)
            throw gr::exception(b200sync_cl_last_error());
    }

public:
#if B200SYNC_HAVE_GR4
    gr::PortIn<std::complex<float>> in;
    gr::PortOut<std::complex<float>> out;
#else
    gr::PortInShim<std::complex<float>> in;
    gr::PortOutShim<std::complex<float>> out;
#endif
    std::vector<float> syncword;
    int device = 0;  // extra: CUDA device ordinal

    SyncwordWipeoffB200() = default;
    SyncwordWipeoffB200(const SyncwordWipeoffB200&) = delete;
    SyncwordWipeoffB200& operator=(const SyncwordWipeoffB200&) = delete;
    ~SyncwordWipeoffB200() { b200sync_wo_destroy(_ctx); }

    void settingsChanged(const gr::property_map& /* old_settings */, const gr::property_map& /* new_settings */)
    {
        configure();
    }

    void start()
    {
        if (!_ctx) configure();
        else if (b200sync_wo_start(_ctx) != 0) throw gr::exception(b200sync_cl_last_error());
    }

    // PM/syncword_wipeoff.hpp:38-91
    template <typename TIn, typename TOut>
    gr::work::Status processBulk(const TIn& inSpan, TOut& outSpan)
    {
        if (!_ctx) throw gr::exception("processBulk() before settingsChanged()/start()");
        const size_t n = std::min(inSpan.size(), outSpan.size());
        b200sync_stream_tag tin{};
        size_t n_tin = 0;
        if (this->input_tags_present()) {
            const auto& map = this->mergedInputTag().map;
            if (map.contains("syncword_amplitude")) {  // :54
                tin.index = 0;
                tin.has_syncword = 1;
                n_tin = 1;
            }
#if !B200SYNC_HAVE_GR4
            // what the runtime's default tag policy does (GR/Block.hpp:777-790); real GR4 does it itself
            if (n > 0) out.publishTag(map, 0);
#endif
        }
        if (n > 0 && b200sync_wo_process(_ctx, reinterpret_cast<const float*>(inSpan.data()), n, n_tin ? &tin : nullptr,
                                         n_tin, reinterpret_cast<float*>(outSpan.data())) != 0)
            throw gr::exception(b200sync_cl_last_error());
        if (!inSpan.consume(n)) throw gr::exception("consume failed");  // :84-86
        outSpan.publish(n);
        return gr::work::Status::OK;
    }
};

}  // namespace gr::packet_modem

#if B200SYNC_HAVE_GR4
ENABLE_REFLECTION(gr::packet_modem::SyncwordWipeoffB200, in, out, syncword, device);
#endif
