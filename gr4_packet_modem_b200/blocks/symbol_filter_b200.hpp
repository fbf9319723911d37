// symbol_filter_b200.hpp — drop-in shell for gr::packet_modem::SymbolFilter<c64, c64, float>
// (PM/symbol_filter.hpp) running on a B200 through libb200sync.so.
//
// Same settings as the reference (`samples_per_symbol`, `taps`, `num_arms`, `delay`, :53-59; reflection
// list :257-258), same custom tag policy (:61-62): syncword tags reset the symbol clock, pick the
// polyphase arm from syncword_time_est, set the output scale from syncword_amplitude, and are
// re-published `delay` samples later on the nearest output symbol (:141-237).  That state machine and
// the 44-tap matched filter run behind b200sync_sf_process; the shell translates spans and tags.
#pragma once
#include "b200_shell_common.hpp"

namespace gr::packet_modem {

class SymbolFilterB200
#if B200SYNC_HAVE_GR4
    : public gr::Block<SymbolFilterB200>
#else
    : public gr::BlockShim<SymbolFilterB200>
#endif
{
    b200sync_sf* _ctx = nullptr;
    b200sync_shell::TagStore _tags;
    std::vector<b200sync_stream_tag> _out_tags = std::vector<b200sync_stream_tag>(16);

    void configure()
    {
        b200sync_sf_destroy(_ctx);
        _ctx = nullptr;
        _tags.clear();
        b200sync_sf_config cfg{};
        cfg.samples_per_symbol = static_cast<uint32_t>(samples_per_symbol);
        cfg.taps = taps.empty() ? nullptr : taps.data();
        cfg.n_taps = static_cast<uint32_t>(taps.size());
        cfg.num_arms = static_cast<uint32_t>(num_arms);
        cfg.delay = static_cast<uint32_t>(delay);
        cfg.device = device;
        // "samples_per_symbol cannot be zero" / "num_arms cannot be zero" (:67-72) come back as the text
        if (b200sync_sf_create(&cfg, &_ctx) != 0) throw gr::exception(b200sync_sf_last_error());
        if (fused_cfc_delay >= 0 && b200sync_sf_fuse_cfc(_ctx, 1, static_cast<uint32_t>(fused_cfc_delay)) != 0)
            throw gr::exception(b200sync_sf_last_error());
    }

public:
#if B200SYNC_HAVE_GR4
    gr::PortIn<std::complex<float>> in;
    gr::PortOut<std::complex<float>> out;
    constexpr static gr::TagPropagationPolicy tag_policy = gr::TagPropagationPolicy::TPP_CUSTOM;
#else
    gr::PortInShim<std::complex<float>> in;
    gr::PortOutShim<std::complex<float>> out;
#endif
    size_t samples_per_symbol = 4;
    std::vector<float> taps;
    size_t num_arms = 1;
    size_t delay = 0;
    int device = 0;  // extra: CUDA device ordinal
    // extra: >= 0 absorbs the CoarseFrequencyCorrection{delay = fused_cfc_delay} block that precedes this one
    // in the receiver (PM/packet_receiver.hpp:94-95, 195-202) into the filter's load stage
    int fused_cfc_delay = -1;

    SymbolFilterB200() = default;
    SymbolFilterB200(const SymbolFilterB200&) = delete;
    SymbolFilterB200& operator=(const SymbolFilterB200&) = delete;
    ~SymbolFilterB200() { b200sync_sf_destroy(_ctx); }

    // PM/symbol_filter.hpp:64-108
    void settingsChanged(const gr::property_map& /* old_settings */, const gr::property_map& /* new_settings */)
    {
        configure();
    }

    // PM/symbol_filter.hpp:110
    void start()
    {
        if (!_ctx) configure();
        else if (b200sync_sf_start(_ctx) != 0) throw gr::exception(b200sync_sf_last_error());
        else if (fused_cfc_delay >= 0 && b200sync_sf_fuse_cfc(_ctx, 1, static_cast<uint32_t>(fused_cfc_delay)) != 0)
            throw gr::exception(b200sync_sf_last_error());
        _tags.clear();
    }

    // PM/symbol_filter.hpp:112-252
    template <typename TIn, typename TOut>
    gr::work::Status processBulk(const TIn& inSpan, TOut& outSpan)
    {
        if (!_ctx) throw gr::exception("processBulk() before settingsChanged()/start()");
        // the reference loop stops when either span is exhausted (:207); n items give at most
        // n / sps + 2 symbols (the +2: the extra symbol of the special case :160-186 and a partial period)
        const size_t sps = samples_per_symbol;
        size_t n = inSpan.size();
        if (outSpan.size() < n / sps + 2) n = outSpan.size() < 2 ? 0 : (outSpan.size() - 2) * sps;
        if (n == 0) {
            if (!inSpan.consume(0)) throw gr::exception("consume failed");
            outSpan.publish(0);
            return inSpan.size() == 0 ? gr::work::Status::INSUFFICIENT_INPUT_ITEMS
                                      : gr::work::Status::INSUFFICIENT_OUTPUT_ITEMS;
        }
        b200sync_stream_tag tin{};
        size_t n_tin = 0;
        if (this->input_tags_present()) {
            tin = _tags.to_abi(this->mergedInputTag().map);
            n_tin = 1;
        }
        size_t consumed = 0, produced = 0, n_tout = 0;
        for (;;) {
            const int rc = b200sync_sf_process(_ctx, reinterpret_cast<const float*>(inSpan.data()), n, n_tin ? &tin : nullptr,
                                               n_tin, reinterpret_cast<float*>(outSpan.data()), outSpan.size(), &consumed,
                                               &produced, _out_tags.data(), _out_tags.size(), &n_tout);
            if (rc == B200SYNC_ENOMEM && _out_tags.size() < (1u << 20)) {  // tag buffer (not the span: that is ENOSPC); state unchanged: grow and retry
                _out_tags.resize(_out_tags.size() * 4);
                continue;
            }
            if (rc != 0) throw gr::exception(b200sync_sf_last_error());
            break;
        }
        for (size_t i = 0; i < n_tout; ++i)  // :218-228
            out.publishTag(_tags.from_abi(_out_tags[i]), static_cast<ssize_t>(_out_tags[i].index));
        if (!inSpan.consume(consumed)) throw gr::exception("consume failed");  // :239-241
        outSpan.publish(produced);
        return gr::work::Status::OK;
    }
};

}  // namespace gr::packet_modem

#if B200SYNC_HAVE_GR4
ENABLE_REFLECTION(gr::packet_modem::SymbolFilterB200, in, out, samples_per_symbol, taps, num_arms, delay, device,
                  fused_cfc_delay);
#endif
