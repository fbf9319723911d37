// syncword_detection_filter_b200.hpp — drop-in shell for gr::packet_modem::SyncwordDetectionFilter<c64>
// (PM/syncword_detection_filter.hpp).
//
// The block is control logic: it drops syncword_* tags that arrive while a packet is being decoded and
// stalls the stream until the header parser reports the packet length (:54-210).  That state machine
// lives behind b200sync_sdf_process so that it is the SAME code for host spans (this shell) and for
// device-resident captures (where the copy is a no-op and only the counts matter).  Ports, settings and
// tag policy are the reference's (:40-50; `allowed_margin` is a member but not reflected, :215-222).
#pragma once
#include "b200_shell_common.hpp"

namespace gr::packet_modem {

class SyncwordDetectionFilterB200
#if B200SYNC_HAVE_GR4
    : public gr::Block<SyncwordDetectionFilterB200>
#else
    : public gr::BlockShim<SyncwordDetectionFilterB200>
#endif
{
    b200sync_sdf* _ctx = nullptr;

    void configure()
    {
        b200sync_sdf_destroy(_ctx);
        _ctx = nullptr;
        if (b200sync_sdf_create(static_cast<uint32_t>(samples_per_symbol), static_cast<uint32_t>(syncword_size),
                                static_cast<uint32_t>(header_size), &_ctx) != 0)
            throw gr::exception(b200sync_sf_last_error());
    }

public:
#if B200SYNC_HAVE_GR4
    gr::PortIn<gr::Message, gr::Async> parsed_header;
    gr::PortIn<gr::Message, gr::Async> ignored_syncword;
    gr::PortIn<std::complex<float>> in;
    gr::PortOut<std::complex<float>> out;
    constexpr static gr::TagPropagationPolicy tag_policy = gr::TagPropagationPolicy::TPP_CUSTOM;
#else
    gr::PortInShim<gr::Message> parsed_header;
    gr::PortInShim<gr::Message> ignored_syncword;
    gr::PortInShim<std::complex<float>> in;
    gr::PortOutShim<std::complex<float>> out;
#endif
    size_t samples_per_symbol = 4;
    size_t syncword_size = 64;
    size_t header_size = 128;

    SyncwordDetectionFilterB200() = default;
    SyncwordDetectionFilterB200(const SyncwordDetectionFilterB200&) = delete;
    SyncwordDetectionFilterB200& operator=(const SyncwordDetectionFilterB200&) = delete;
    ~SyncwordDetectionFilterB200() { b200sync_sdf_destroy(_ctx); }

    void settingsChanged(const gr::property_map& /* old_settings */, const gr::property_map& /* new_settings */)
    {
        configure();
    }

    // PM/syncword_detection_filter.hpp:52
    void start()
    {
        if (!_ctx) configure();
        else if (b200sync_sdf_start(_ctx) != 0) throw gr::exception(b200sync_sf_last_error());
    }

    // PM/syncword_detection_filter.hpp:54-210
    template <typename THeader, typename TIgnored, typename TIn, typename TOut>
    gr::work::Status processBulk(const THeader& headerSpan, const TIgnored& ignoredSpan, const TIn& inSpan,
                                 TOut& outSpan)
    {
        if (!_ctx) throw gr::exception("processBulk() before start()");
        // merged input tag -> which classes of keys it carries (:76-92)
        b200sync_stream_tag tin{};
        bool have_tag = false;
        if (this->input_tags_present()) {
            have_tag = true;
            for (const auto& [key, val] : this->mergedInputTag().map) {
                if (b200sync_shell::is_syncword_key(key)) tin.has_syncword = 1;
                else tin.other = 1;
            }
        }
        // first pending parsed_header message (:136-154)
        b200sync_sdf_header hdr{};
        bool have_hdr = false;
        if (headerSpan.size() > 0) {
            const auto& meta = headerSpan[0].data.value();
            have_hdr = true;
            if (meta.contains("invalid_header")) {
                hdr.invalid_header = 1;
            } else {
                hdr.packet_length = b200sync_shell::pmt_cast<uint64_t>(meta.at("packet_length"));
            }
        }
        size_t consumed = 0, header_consumed = 0, ignored_consumed = 0;
        b200sync_stream_tag tout{};
        int forwarded = 0, in_packet = 0;
        if (b200sync_sdf_process(_ctx, reinterpret_cast<const float*>(inSpan.data()), inSpan.size(),
                                 reinterpret_cast<float*>(outSpan.data()), outSpan.size(), have_tag ? &tin : nullptr,
                                 have_hdr ? &hdr : nullptr, ignoredSpan.size(), &consumed, &header_consumed,
                                 &ignored_consumed, &tout, &forwarded, &in_packet) != 0)
            throw gr::exception(b200sync_sf_last_error());
        if (forwarded) {  // :93-107
            gr::property_map output_tags;
            for (const auto& [key, val] : this->mergedInputTag().map) {
                const bool sw = b200sync_shell::is_syncword_key(key);
                if ((sw && tout.has_syncword) || (!sw && tout.other)) output_tags[key] = val;
            }
            if (!output_tags.empty()) out.publishTag(output_tags, 0);
        }
        if (!headerSpan.consume(header_consumed)) throw gr::exception("headerSpan.consume() failed");
        if (!ignoredSpan.consume(ignored_consumed)) throw gr::exception("ignoredSpan.consume() failed");
        if (!inSpan.consume(consumed)) throw gr::exception("inSpan.consume() failed");
        outSpan.publish(consumed);
        // forwardTags() only clears the merged tag when every port moved; the message ports did not
        // (:120-125, :198-202)
        this->_mergedInputTag.map.clear();
        return gr::work::Status::OK;
    }
};

}  // namespace gr::packet_modem

#if B200SYNC_HAVE_GR4
ENABLE_REFLECTION(gr::packet_modem::SyncwordDetectionFilterB200, parsed_header, ignored_syncword, in, out,
                  samples_per_symbol, syncword_size, header_size);
#endif
