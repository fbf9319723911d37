// multi.cpp — one capture on several GPUs of one box, gathered on the host (north_star: "contiguous
// sample segments with an overlap-save halo ... detections gathered on the host ... no NCCL on the data path").
//
// Host code only.  The capture is cut on the reference's own FFT-block grid (block b starts at sample b*S from
// the stream origin, PM/syncword_detection.hpp:236-238, so every block is the block of the single-GPU run); GPU r
// gets a contiguous range of blocks plus one halo block of input on each side for the +-T metric context.  One host
// thread per GPU drives that GPU's SyncwordDetection context:
//     phase 1   correlator + candidate flags + the shard's (T+1)-entry chain table  -> host
//     ---- std::barrier: every thread now sees all tables (769 x uint16 per GPU, plain host memory) ----
//     compose   entry search offset of shard r = tables[r-1] o ... o tables[0] (0)   
//     phase 2   walk the chain from the true entry offset, refine, records           -> host
// and the calling thread concatenates the shards' records, which are already in index order.  Nothing crosses
// between GPUs; the only thing the host exchanges is those tables.
#include <algorithm>
#include <barrier>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include <sched.h>
#include <sys/stat.h>

#include "b200sync_internal.h"

using namespace b200sync;

struct b200sync_sd_multi {
    std::vector<b200sync_sd*> ctx;
    std::vector<int> dev;
    std::vector<cudaStream_t> st;
    std::vector<cpu_set_t> cpus;     // host cores local to each GPU (empty set: unknown, no binding)
    std::vector<bool> cpus_known;
    uint32_t S = 0, L = 0, F = 2048;
    uint64_t T = 0, delay = 0;
    std::vector<std::vector<b200sync_detection_record>> recs;  // per shard
    std::vector<float> corr_ms, peaks_ms, refine_ms;
};

namespace {

struct Shard {
    uint64_t first_block = 0, n_blocks = 0, total_blocks = 0, first_sample = 0;
    size_t n_samples = 0;
};

// the plan of gr4_packet_modem_b200/sharding.py: plan_shards
std::vector<Shard> plan(uint64_t n, size_t world, uint32_t F, uint32_t S, uint64_t T) {
    std::vector<Shard> out(world);
    const uint64_t tb = n >= F ? (n - F) / S + 1 : 0;
    const uint64_t halo = (T + S) / S;
    for (size_t r = 0; r < world; ++r) {
        const uint64_t fb = r * tb / world, nb = (r + 1) * tb / world - fb;
        const uint64_t cb0 = fb > halo ? fb - halo : 0, cb1 = std::min(tb, fb + nb + halo);
        out[r].first_block = fb;
        out[r].n_blocks = nb;
        out[r].total_blocks = tb;
        out[r].first_sample = cb0 * S;
        out[r].n_samples = cb1 > cb0 ? static_cast<size_t>((cb1 - 1) * S + F - cb0 * S) : 0;
    }
    return out;
}

// host cores on the NUMA node of a GPU: /sys/bus/pci/devices/<bus id>/local_cpulist ("0-31,64-95")
bool gpu_local_cpus(int device, cpu_set_t* set) {
    char bus[32] = {0};
    if (cudaDeviceGetPCIBusId(bus, sizeof(bus), device) != cudaSuccess) return false;
    for (char* c = bus; *c; ++c) *c = static_cast<char>(std::tolower(*c));
    const std::string path = std::string("/sys/bus/pci/devices/") + bus + "/local_cpulist";
    FILE* f = std::fopen(path.c_str(), "r");
    if (!f) return false;
    char line[4096] = {0};
    const bool got = std::fgets(line, sizeof(line), f) != nullptr;
    std::fclose(f);
    if (!got) return false;
    CPU_ZERO(set);
    int count = 0;
    for (char* tok = std::strtok(line, ",\n"); tok; tok = std::strtok(nullptr, ",\n")) {
        int a = 0, b = 0;
        const int k = std::sscanf(tok, "%d-%d", &a, &b);
        if (k == 1) b = a;
        if (k < 1) continue;
        for (int c = a; c <= b && c < CPU_SETSIZE; ++c) {
            CPU_SET(c, set);
            ++count;
        }
    }
    return count > 0;
}

// Source of a shard's samples: device pointers (one per shard), one host capture, or a capture file
struct Source {
    const void* const* d_shards = nullptr;
    const float* host = nullptr;
    const char* file = nullptr;
    uint64_t file_first_item = 0;
};

int run(b200sync_sd_multi* m, const Source& src, uint64_t n, b200sync_detection_record* out, size_t max_recs,
        size_t* n_recs, size_t* n_consumed) {
    const size_t W = m->ctx.size();
    *n_recs = 0;
    *n_consumed = 0;
    if (n < m->F) return 0;
    const std::vector<Shard> shards = plan(n, W, m->F, m->S, m->T);
    const size_t tl = static_cast<size_t>(m->T) + 1;
    std::vector<uint16_t> tables(W * tl);
    std::vector<int> rc(W, 0);
    std::vector<std::string> err(W);
    std::barrier sync(static_cast<std::ptrdiff_t>(W));
    auto worker = [&](size_t r) {
        if (m->cpus_known[r]) sched_setaffinity(0, sizeof(cpu_set_t), &m->cpus[r]);  // best effort
        const Shard& sh = shards[r];
        uint16_t* table = tables.data() + r * tl;
        if (sh.n_blocks == 0) {
            for (size_t j = 0; j < tl; ++j) table[j] = static_cast<uint16_t>(j);  // empty shard: identity map
        } else if (src.d_shards) {
            rc[r] = b200sync_sd_shard_phase1(m->ctx[r], src.d_shards[r], sh.first_sample, sh.n_samples, sh.first_block,
                                             sh.n_blocks, sh.total_blocks, m->st[r], table, tl);
        } else if (src.host) {
            rc[r] = b200sync_sd_shard_phase1_host(m->ctx[r], src.host + 2 * sh.first_sample, sh.first_sample, sh.n_samples,
                                                  sh.first_block, sh.n_blocks, sh.total_blocks, table, tl);
        } else {
            rc[r] = b200sync_sd_shard_phase1_file(m->ctx[r], src.file, src.file_first_item, sh.first_sample, sh.n_samples,
                                                  sh.first_block, sh.n_blocks, sh.total_blocks, table, tl);
        }
        if (rc[r] != 0) err[r] = b200sync_last_error();
        sync.arrive_and_wait();  // all tables are in host memory
        bool all_ok = true;
        for (size_t i = 0; i < W; ++i) all_ok = all_ok && rc[i] == 0;
        m->recs[r].clear();
        if (!all_ok || sh.n_blocks == 0) return;
        uint32_t j = 0;  // search offset 0 at the stream start (:191-193), then through the shards before r
        for (size_t i = 0; i < r; ++i) j = tables[i * tl + j];
        const size_t cap = sh.n_samples / (m->T + 1) + 2;
        m->recs[r].resize(cap);
        size_t got = 0;
        rc[r] = b200sync_sd_shard_phase2(m->ctx[r], j, m->recs[r].data(), cap, &got);
        if (rc[r] != 0) err[r] = b200sync_last_error();
        m->recs[r].resize(rc[r] == 0 ? got : 0);
        b200sync_sd_last_timings(m->ctx[r], &m->corr_ms[r], &m->peaks_ms[r], &m->refine_ms[r]);
    };
    std::vector<std::thread> pool;
    pool.reserve(W);
    for (size_t r = 0; r < W; ++r) pool.emplace_back(worker, r);
    for (auto& t : pool) t.join();
    for (size_t r = 0; r < W; ++r)
        if (rc[r] != 0) return set_last_error(rc[r], "GPU " + std::to_string(m->dev[r]) + ": " + err[r]);
    size_t cnt = 0;
    for (size_t r = 0; r < W; ++r) {
        if (cnt + m->recs[r].size() > max_recs) return set_last_error(B200SYNC_ENOMEM, "record buffer too small");
        std::copy(m->recs[r].begin(), m->recs[r].end(), out + cnt);
        cnt += m->recs[r].size();
    }
    *n_recs = cnt;
    *n_consumed = static_cast<size_t>(shards[0].total_blocks * m->S);
    return 0;
}

}  // namespace

extern "C" {

int b200sync_sd_multi_create(const b200sync_sd_config* cfg, const int* devices, size_t n_devices,
                             b200sync_sd_multi** out) {
    if (!cfg || !out || (!devices && n_devices)) return set_last_error(B200SYNC_EINVAL, "null argument");
    *out = nullptr;
    int have = 0;
    if (cudaGetDeviceCount(&have) != cudaSuccess || have <= 0)
        return set_last_error(B200SYNC_ECUDA, "no CUDA device (there is no CPU fallback)");
    std::vector<int> devs(devices, devices + n_devices);
    if (devs.empty())
        for (int d = 0; d < have; ++d) devs.push_back(d);  // every visible GPU
    auto* m = new (std::nothrow) b200sync_sd_multi();
    if (!m) return set_last_error(B200SYNC_ENOMEM, "out of memory");
    for (int d : devs) {
        b200sync_sd_config c = *cfg;
        c.device = d;
        b200sync_sd* sd = nullptr;
        const int rc = b200sync_sd_create(&c, &sd);
        if (rc != 0) {
            const std::string keep = b200sync_last_error();
            b200sync_sd_multi_destroy(m);
            return set_last_error(rc, keep);
        }
        m->ctx.push_back(sd);
        m->dev.push_back(d);
        cudaStream_t st = nullptr;
        if (cudaSetDevice(d) != cudaSuccess || cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) {
            b200sync_sd_multi_destroy(m);
            return set_last_error(B200SYNC_ECUDA, "stream creation failed");
        }
        m->st.push_back(st);
        cpu_set_t set;
        const bool known = gpu_local_cpus(d, &set);
        if (!known) CPU_ZERO(&set);
        m->cpus.push_back(set);
        m->cpus_known.push_back(known);
    }
    uint32_t L = 0, S = 0;
    uint64_t delay = 0;
    b200sync_sd_info(m->ctx[0], &L, &S, nullptr, nullptr, &delay);
    m->L = L;
    m->S = S;
    m->F = cfg->fft_size ? cfg->fft_size : 2048;
    m->delay = delay;
    m->T = (delay - 1) / 2;
    m->recs.resize(m->ctx.size());
    m->corr_ms.assign(m->ctx.size(), 0.f);
    m->peaks_ms.assign(m->ctx.size(), 0.f);
    m->refine_ms.assign(m->ctx.size(), 0.f);
    *out = m;
    return 0;
}

void b200sync_sd_multi_destroy(b200sync_sd_multi* m) {
    if (!m) return;
    for (size_t i = 0; i < m->st.size(); ++i) {
        cudaSetDevice(m->dev[i]);
        cudaStreamSynchronize(m->st[i]);
        cudaStreamDestroy(m->st[i]);
    }
    for (auto* sd : m->ctx) b200sync_sd_destroy(sd);
    delete m;
}

size_t b200sync_sd_multi_devices(const b200sync_sd_multi* m) { return m ? m->ctx.size() : 0; }

b200sync_sd* b200sync_sd_multi_context(b200sync_sd_multi* m, size_t i) {
    return (m && i < m->ctx.size()) ? m->ctx[i] : nullptr;
}

int b200sync_sd_multi_plan(const b200sync_sd_multi* m, uint64_t n, b200sync_shard* shards) {
    if (!m || !shards) return set_last_error(B200SYNC_EINVAL, "null argument");
    const auto p = plan(n, m->ctx.size(), m->F, m->S, m->T);
    for (size_t r = 0; r < p.size(); ++r) {
        shards[r].device = m->dev[r];
        shards[r].first_block = p[r].first_block;
        shards[r].n_blocks = p[r].n_blocks;
        shards[r].total_blocks = p[r].total_blocks;
        shards[r].first_sample = p[r].first_sample;
        shards[r].n_samples = p[r].n_samples;
    }
    return 0;
}

int b200sync_sd_multi_detect_device(b200sync_sd_multi* m, const void* const* d_shards, uint64_t n,
                                    b200sync_detection_record* recs, size_t max_recs, size_t* n_recs,
                                    size_t* n_consumed) {
    if (!m || !d_shards || !n_recs || !n_consumed || (!recs && max_recs))
        return set_last_error(B200SYNC_EINVAL, "null argument");
    Source s;
    s.d_shards = d_shards;
    return run(m, s, n, recs, max_recs, n_recs, n_consumed);
}

int b200sync_sd_multi_detect_host(b200sync_sd_multi* m, const float* in, uint64_t n, b200sync_detection_record* recs,
                                  size_t max_recs, size_t* n_recs, size_t* n_consumed) {
    if (!m || (!in && n) || !n_recs || !n_consumed || (!recs && max_recs))
        return set_last_error(B200SYNC_EINVAL, "null argument");
    Source s;
    s.host = in;
    return run(m, s, n, recs, max_recs, n_recs, n_consumed);
}

int b200sync_sd_multi_detect_file(b200sync_sd_multi* m, const char* filename, uint64_t first_item, uint64_t max_items,
                                  b200sync_detection_record* recs, size_t max_recs, size_t* n_recs, size_t* n_consumed,
                                  uint64_t* n_items_read) {
    if (!m || !filename || !n_recs || !n_consumed || (!recs && max_recs))
        return set_last_error(B200SYNC_EINVAL, "null argument");
    *n_recs = 0;
    *n_consumed = 0;
    if (n_items_read) *n_items_read = 0;
    struct stat sb {};
    if (stat(filename, &sb) != 0)  // PM/file_source.hpp:31-37
        return set_last_error(B200SYNC_EINVAL, std::string("error opening file: ") + std::strerror(errno));
    if (!S_ISREG(sb.st_mode))
        return set_last_error(B200SYNC_EUNSUPPORTED, "not a seekable file (FIFOs: use b200sync_sd_process)");
    const uint64_t items = static_cast<uint64_t>(sb.st_size) / (2 * sizeof(float));  // whole items only, like fread
    if (first_item > items) first_item = items;
    const uint64_t n = std::min(items - first_item, max_items);
    if (n_items_read) *n_items_read = n;
    Source s;
    s.file = filename;
    s.file_first_item = first_item;
    return run(m, s, n, recs, max_recs, n_recs, n_consumed);
}

int b200sync_sd_multi_last_timings(const b200sync_sd_multi* m, float* correlate_ms, float* peaks_ms, float* refine_ms) {
    if (!m) return set_last_error(B200SYNC_EINVAL, "null context");
    for (size_t r = 0; r < m->ctx.size(); ++r) {
        if (correlate_ms) correlate_ms[r] = m->corr_ms[r];
        if (peaks_ms) peaks_ms[r] = m->peaks_ms[r];
        if (refine_ms) refine_ms[r] = m->refine_ms[r];
    }
    return 0;
}

}  // extern "C"
