// costas_loop_b200.hpp — drop-in shell for gr::packet_modem::CostasLoop<float, float> (PM/costas_loop.hpp)
// running on a B200 through libb200sync.so.
//
// Same settings as the reference (`loop_bandwidth` :52, `constellation` :54 as a case-insensitive string;
// reflection list :153-154) and the same default tag policy: input tags are forwarded unchanged to the
// first output item of the chunk by the runtime (GR/Block.hpp:777-790).  A tag with a "syncword_phase"
// key does set_phase() before the chunk's first item (:102-107); the loop itself runs behind
// b200sync_cl_process, its state stays on the device between calls.
//
// Extra setting `fused_wipeoff_syncword`: when not empty the block also does the work of the
// SyncwordWipeoff in front of it (PM/packet_receiver.hpp:203-214) inside the same kernel.
#pragma once
#include <algorithm>
#include <cctype>

#include "b200_shell_common.hpp"

namespace gr::packet_modem {

class CostasLoopB200
#if B200SYNC_HAVE_GR4
    : public gr::Block<CostasLoopB200>
#else
    : public gr::BlockShim<CostasLoopB200>
#endif
{
    b200sync_cl* _ctx = nullptr;

    void configure()
    {
        b200sync_cl_destroy(_ctx);
        _ctx = nullptr;
        std::string key = constellation;
        std::transform(key.begin(), key.end(), key.begin(), [](unsigned char ch) { return std::toupper(ch); });
        b200sync_cl_config cfg{};
        cfg.loop_bandwidth = loop_bandwidth;
        cfg.device = device;
        if (key == "PILOT") cfg.constellation = B200SYNC_CONSTELLATION_PILOT;
        else if (key == "BPSK") cfg.constellation = B200SYNC_CONSTELLATION_BPSK;
        else if (key == "QPSK") cfg.constellation = B200SYNC_CONSTELLATION_QPSK;
        else throw gr::exception("unknown constellation " + constellation);  // enum_cast(...).value() (:63-65)
        if (b200sync_cl_create(&cfg, &_ctx) != 0) throw gr::exception(b200sync_cl_last_error());
        if (!fused_wipeoff_syncword.empty() &&
            b200sync_cl_fuse_wipeoff(_ctx, fused_wipeoff_syncword.data(),
                                     static_cast<uint32_t>(fused_wipeoff_syncword.size())) != 0)
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
    double loop_bandwidth = 0.01;
    std::string constellation = "BPSK";
    std::vector<float> fused_wipeoff_syncword;  // extra: absorb the SyncwordWipeoff in front
    int device = 0;                             // extra: CUDA device ordinal

    CostasLoopB200() = default;
    CostasLoopB200(const CostasLoopB200&) = delete;
    CostasLoopB200& operator=(const CostasLoopB200&) = delete;
    ~CostasLoopB200() { b200sync_cl_destroy(_ctx); }

    void settingsChanged(const gr::property_map& /* old_settings */, const gr::property_map& /* new_settings */)
    {
        configure();
    }

    void start()
    {
        if (!_ctx) configure();
        else if (b200sync_cl_start(_ctx) != 0) throw gr::exception(b200sync_cl_last_error());
    }

    // PM/costas_loop.hpp:94-149
    template <typename TIn, typename TOut>
    gr::work::Status processBulk(const TIn& inSpan, TOut& outSpan)
    {
        if (!_ctx) throw gr::exception("processBulk() before settingsChanged()/start()");
        const size_t n = std::min(inSpan.size(), outSpan.size());
        b200sync_stream_tag tin{};
        size_t n_tin = 0;
        if (this->input_tags_present()) {
            const auto& map = this->mergedInputTag().map;
            // the C ABI's has_syncword stands for the whole syncword_* key set of SyncwordDetection's tag
            // (PM/syncword_detection.hpp:106-114): the loop reads syncword_phase (:104), a fused wipe-off
            // reads the presence of syncword_amplitude (PM/syncword_wipeoff.hpp:54)
            if (map.contains("syncword_phase")) {
                tin.index = 0;
                tin.has_syncword = 1;
                tin.sw.syncword_phase = b200sync_shell::pmt_cast<float>(map.at("syncword_phase"));
                n_tin = 1;
            }
#if !B200SYNC_HAVE_GR4
            if (n > 0) out.publishTag(map, 0);  // default tag policy (GR/Block.hpp:777-790)
#endif
        }
        if (n > 0 && b200sync_cl_process(_ctx, reinterpret_cast<const float*>(inSpan.data()), n, n_tin ? &tin : nullptr,
                                         n_tin, reinterpret_cast<float*>(outSpan.data())) != 0)
            throw gr::exception(b200sync_cl_last_error());
        if (!inSpan.consume(n)) throw gr::exception("consume failed");
        outSpan.publish(n);
        return gr::work::Status::OK;
    }
};

}  // namespace gr::packet_modem

#if B200SYNC_HAVE_GR4
ENABLE_REFLECTION(gr::packet_modem::CostasLoopB200, in, out, loop_bandwidth, constellation, fused_wipeoff_syncword,
                  device);
#endif
