// syncword_detection_b200.hpp — drop-in block shell for gr::packet_modem::SyncwordDetection
// running on a B200 through libb200sync.so.
//
// Same public surface as the reference block (PM/syncword_detection.hpp:131-141 settings,
// :143 start(), :204 processBulk(), :106-114 tag keys, :361-372 reflection list), so a flowgraph
// replaces
//     fg.emplaceBlock<gr::packet_modem::SyncwordDetection>({...})
// by
//     fg.emplaceBlock<gr::packet_modem::SyncwordDetectionB200>({...})
// and nothing else.  All DSP happens behind the C ABI (include/b200sync.h); this class only
// translates spans/tags/exceptions.
//
// Build modes
//   * real GNU Radio 4.0 on the include path: derives from gr::Block<SyncwordDetectionB200>, uses
//     gr::PortIn/PortOut and ENABLE_REFLECTION.  (Cannot be compiled in the build container:
//     GR4's fetched dependencies are unavailable — DESIGN.md §7.  Verified by construction only.)
//   * otherwise: gr4_compat.hpp stand-ins with the same member names; this is what
//     tests/cpp/test_block_shell.cpp compiles and runs against the GPU.
#pragma once
#include <complex>
#include <cstdint>
#include <string>
#include <vector>

#include "b200_shell_common.hpp"  // the C ABI + real GR4 headers when present, gr4_compat.hpp otherwise

namespace gr::packet_modem {

class SyncwordDetectionB200
#if B200SYNC_HAVE_GR4
    : public gr::Block<SyncwordDetectionB200>
#endif
{
    using c64 = std::complex<float>;
    b200sync_sd* _ctx = nullptr;
    std::vector<b200sync_sd_tag> _tagbuf = std::vector<b200sync_sd_tag>(256);
    uint64_t _items_consumed = 0;

    static gr::property_map to_map(const b200sync_sd_tag& t)
    {
        // keys and value types of PM/syncword_detection.hpp:106-114
        return {
            { "syncword_amplitude", t.syncword_amplitude },
            { "syncword_phase", t.syncword_phase },
            { "syncword_freq", t.syncword_freq },
            { "syncword_freq_bin", static_cast<int>(t.syncword_freq_bin) },
            { "syncword_noise_power", t.syncword_noise_power },
            { "syncword_esn0_db", t.syncword_esn0_db },
            { "syncword_time_est", t.syncword_time_est },
        };
    }

public:
#if B200SYNC_HAVE_GR4
    gr::PortIn<std::complex<float>> in;
    gr::PortOut<std::complex<float>> out;
#else
    gr::PortInShim<std::complex<float>> in;
    gr::PortOutShim<std::complex<float>> out;
#endif
    // settings — names, types and defaults of PM/syncword_detection.hpp:133-141
    size_t fft_size = 2048;
    size_t samples_per_symbol = 4;
    std::vector<float> rrc_taps;
    std::vector<uint8_t> syncword;
    std::vector<std::complex<float>> constellation;
    int min_freq_bin = 0;
    int max_freq_bin = 0;
    uint64_t time_threshold = 768;
    float power_threshold = 9.5;
    // extra, not in the reference: which CUDA device runs this block instance
    int device = 0;
    // extra: page-lock the pages of the input ring as its spans arrive (b200sync_sd_set_auto_register): the spans then
    // go straight to the copy engine (0.8 instead of 0.55 Gsps at 65536-item spans).  OPT-IN: it is only safe when the
    // memory behind the spans outlives this block — true for a GR4 port's CircularBuffer (shared-owned by its readers,
    // GR/CircularBuffer.hpp), NOT for a caller that hands in temporary buffers: a freed and re-allocated buffer would
    // still be mapped to its old pages for the copy engine.  A failed registration falls back to the staged copy.
    bool register_input_ring = false;

    SyncwordDetectionB200() = default;
    SyncwordDetectionB200(const SyncwordDetectionB200&) = delete;
    SyncwordDetectionB200& operator=(const SyncwordDetectionB200&) = delete;
    ~SyncwordDetectionB200() { b200sync_sd_destroy(_ctx); }

    // PM/syncword_detection.hpp:143-202.  Throws gr::exception with the reference's messages
    // ("min_freq_bin is greater than max_freq_bin", "fft_size too small").
    void start()
    {
        b200sync_sd_destroy(_ctx);
        _ctx = nullptr;
        b200sync_sd_config cfg{};
        cfg.fft_size = static_cast<uint32_t>(fft_size);
        cfg.samples_per_symbol = static_cast<uint32_t>(samples_per_symbol);
        cfg.rrc_taps = rrc_taps.data();
        cfg.n_rrc_taps = static_cast<uint32_t>(rrc_taps.size());
        cfg.syncword = syncword.data();
        cfg.n_syncword = static_cast<uint32_t>(syncword.size());
        cfg.constellation = reinterpret_cast<const float*>(constellation.data());
        cfg.n_constellation = static_cast<uint32_t>(constellation.size());
        cfg.min_freq_bin = min_freq_bin;
        cfg.max_freq_bin = max_freq_bin;
        cfg.time_threshold = time_threshold;
        cfg.power_threshold = power_threshold;
        cfg.device = device;
        if (b200sync_sd_create(&cfg, &_ctx) != 0) throw gr::exception(b200sync_last_error());
        if (register_input_ring) b200sync_sd_set_auto_register(_ctx, 1);
        _items_consumed = 0;
        in.min_samples = fft_size;  // :200-201
        out.min_samples = fft_size;
    }

    void stop()
    {
        b200sync_sd_destroy(_ctx);
        _ctx = nullptr;
    }

    // PM/syncword_detection.hpp:204-356
    template <typename TIn, typename TOut>
    gr::work::Status processBulk(const TIn& inSpan, TOut& outSpan)
    {
        if (!_ctx) throw gr::exception("processBulk() before start()");
        size_t consumed = 0, ntags = 0;
        const int rc = b200sync_sd_process(_ctx, reinterpret_cast<const float*>(inSpan.data()), inSpan.size(),
                                           reinterpret_cast<float*>(outSpan.data()), &consumed, _tagbuf.data(),
                                           _tagbuf.size(), &ntags);
        if (rc < 0) throw gr::exception(b200sync_last_error());
        if (rc == 1) {  // :215-227
            if (!inSpan.consume(0)) throw gr::exception("consume failed");
            outSpan.publish(0);
            return gr::work::Status::INSUFFICIENT_INPUT_ITEMS;
        }
        // more publishable tags than the buffer held (tiny time_threshold or a very long span): they are
        // still queued in the context and belong to THIS chunk, so fetch them before publishing
        for (size_t more = b200sync_sd_tags_ready(_ctx); more > 0; more = b200sync_sd_tags_ready(_ctx)) {
            if (_tagbuf.size() < ntags + more) _tagbuf.resize(ntags + more);
            size_t got = 0;
            if (b200sync_sd_drain_tags(_ctx, _tagbuf.data() + ntags, _tagbuf.size() - ntags, &got) != 0)
                throw gr::exception(b200sync_last_error());
            ntags += got;
        }
        for (size_t i = 0; i < ntags; ++i) {
            // tag.index is an absolute output index; publishTag wants the offset in this chunk (:321-324)
            out.publishTag(to_map(_tagbuf[i]), static_cast<ssize_t>(_tagbuf[i].index - _items_consumed));
        }
        if (!inSpan.consume(consumed)) throw gr::exception("consume failed");  // :346-350
        outSpan.publish(consumed);
        _items_consumed += consumed;
        return gr::work::Status::OK;
    }
};

} // namespace gr::packet_modem

#if B200SYNC_HAVE_GR4
ENABLE_REFLECTION(gr::packet_modem::SyncwordDetectionB200, in, out, fft_size, samples_per_symbol, rrc_taps, syncword,
                  constellation, min_freq_bin, max_freq_bin, time_threshold, power_threshold, device);
#endif
