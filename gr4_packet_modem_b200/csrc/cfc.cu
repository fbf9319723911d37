// cfc.cu — CoarseFrequencyCorrection behind the C ABI.
//
// Reference semantics: CoarseFrequencyCorrection<float>, PM/coarse_frequency_correction.hpp:40-98;
// wiring between SyncwordDetectionFilter and SymbolFilter with delay = (rrc_taps.size()-1)/2 + sps:
// PM/packet_receiver.hpp:94-95, 195-202.  See cfc.cuh for the closed form.  HBM-bound: 16 B per
// sample (8 in + 8 out), one sincosf per 8 samples.
#include <cmath>
#include <new>
#include <string>

#include "cfc.cuh"

namespace b200sync {

// ---- host planner ----------------------------------------------------------------------------
CfcSeed CfcPlanner::make_segment(long long start, float freq, uint32_t delay) {
    // set_freq(), :50-59 — all in float, like the reference
    const float a = freq * static_cast<float>(delay);
    CfcSeed s{};
    s.start = start;
    s.e_re = std::cos(a);
    s.e_im = -std::sin(a);
    s.i_re = std::cos(freq);
    s.i_im = -std::sin(freq);
    return s;
}

void CfcPlanner::reset(uint32_t delay) {
    delay_ = delay;
    abs_pos_ = 0;
    // _next_freq = 0, _next_freq_delay = 0 (:46-47): set_freq(0) on the very first sample — the identity
    pending_ = false;
    segs_.clear();
    segs_.push_back(make_segment(0, 0.0f, delay));
}

void CfcPlanner::advance(size_t n, const b200sync_stream_tag* tags, size_t n_tags) {
    const long long end = abs_pos_ + static_cast<long long>(n);
    auto fire_until = [&](long long limit_inclusive) {  // the pending reset happens at sample pending_at_
        if (pending_ && pending_at_ <= limit_inclusive) {
            segs_.push_back(make_segment(pending_at_, pending_freq_, delay_));
            pending_ = false;
        }
    };
    for (size_t i = 0; i < n_tags; ++i) {
        if (!tags[i].has_syncword) continue;  // only tags with a syncword_freq key act (:75)
        const long long p = abs_pos_ + static_cast<long long>(tags[i].index);
        // a tag is examined BEFORE the sample loop of its chunk (:73-79): a reset due exactly at p is
        // cancelled by the new tag, one due earlier has happened
        fire_until(p - 1);
        pending_ = true;
        pending_freq_ = static_cast<float>(tags[i].sw.syncword_freq);  // pmtv::cast<float>
        pending_at_ = p + static_cast<long long>(delay_);
    }
    fire_until(end - 1);
    abs_pos_ = end;
}

void CfcPlanner::live_segments(long long from_abs, std::vector<CfcSeed>& out) const {
    out.clear();
    size_t first = 0;
    for (size_t i = 0; i < segs_.size(); ++i)
        if (segs_[i].start <= from_abs) first = i;
    out.assign(segs_.begin() + static_cast<std::ptrdiff_t>(first), segs_.end());
}

void CfcPlanner::prune(long long from_abs) {
    while (segs_.size() > 1 && segs_[1].start <= from_abs) segs_.pop_front();
}

// ---- seeds -> segments ---------------------------------------------------------------------------
__global__ void cfc_expand_kernel(const CfcSeed* __restrict__ seeds, CfcSegment* __restrict__ segs, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const CfcSeed sd = seeds[i];
    CfcSegment s;
    s.start = sd.start;
    s.phase0 = atan2((double)sd.e_im, (double)sd.e_re);
    s.theta = atan2((double)sd.i_im, (double)sd.i_re);
    s.amp0_eps = (float)(hypot((double)sd.e_re, (double)sd.e_im) - 1.0);
    s.amp_eps = (float)(hypot((double)sd.i_re, (double)sd.i_im) - 1.0);
#pragma unroll
    for (int r = 0; r < kCfcGroup; ++r) {
        double sn, cs;
        sincos(r * s.theta, &sn, &cs);
        s.w[r] = make_float2((float)cs, (float)sn);
    }
    segs[i] = s;
}

cudaError_t cfc_upload_segments(const std::vector<CfcSeed>& seeds, CfcSeed** d_seeds, CfcSegment** d_segs, size_t* cap,
                                cudaStream_t st) {
    return cfc_upload_segments(seeds.data(), seeds.size(), d_seeds, d_segs, cap, st);
}

// `seeds` may point into pinned memory: then the copy is truly asynchronous (a pageable source makes
// cudaMemcpyAsync wait for the stream first, which would serialise a pipelined caller)
cudaError_t cfc_upload_segments(const CfcSeed* seeds_p, size_t n_seeds, CfcSeed** d_seeds, CfcSegment** d_segs, size_t* cap,
                                cudaStream_t st) {
    struct View { const CfcSeed* p; size_t n; size_t size() const { return n; } const CfcSeed* data() const { return p; } } seeds{seeds_p, n_seeds};
    if (*cap < seeds.size()) {
        if (*d_seeds) cudaFree(*d_seeds);
        if (*d_segs) cudaFree(*d_segs);
        *d_seeds = nullptr;
        *d_segs = nullptr;
        *cap = 0;
        const size_t want = seeds.size() + seeds.size() / 2 + 64;
        cudaError_t e = cudaMalloc(d_seeds, want * sizeof(CfcSeed));
        if (e != cudaSuccess) return e;
        e = cudaMalloc(d_segs, want * sizeof(CfcSegment));
        if (e != cudaSuccess) return e;
        *cap = want;
    }
    cudaError_t e = cudaMemcpyAsync(*d_seeds, seeds.data(), seeds.size() * sizeof(CfcSeed), cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return e;
    const int n = static_cast<int>(seeds.size());
    cfc_expand_kernel<<<(n + 127) / 128, 128, 0, st>>>(*d_seeds, *d_segs, n);
    count_launch();
    return cudaGetLastError();
}

// ---- stand-alone kernel ------------------------------------------------------------------------
constexpr int kCfcThreads = 256;
constexpr int kCfcPer = 8;  // consecutive samples per thread
constexpr int kCfcTile = kCfcThreads * kCfcPer;

__global__ void __launch_bounds__(kCfcThreads)
cfc_kernel(const float2* __restrict__ in, float2* __restrict__ out, long long n, long long abs0,
           const CfcSegment* __restrict__ segs, int n_segs) {
    __shared__ int sg0;
    const long long t0 = (long long)blockIdx.x * kCfcTile;
    if (threadIdx.x == 0) sg0 = cfc_search(segs, n_segs, abs0 + t0);
    __syncthreads();
    const long long i0 = t0 + (long long)threadIdx.x * kCfcPer;
    if (i0 >= n) return;
    CfcCursor c;
    c.sg = cfc_seek(segs, n_segs, sg0, abs0 + i0);
    c.seg_start = segs[c.sg].start;
    c.seg_next = (c.sg + 1 < n_segs) ? segs[c.sg + 1].start : 0x7fffffffffffffffLL;
    c.amp0_eps = segs[c.sg].amp0_eps;
    c.amp_eps = segs[c.sg].amp_eps;
    if (i0 + kCfcPer <= n && ((reinterpret_cast<uintptr_t>(in + i0) | reinterpret_cast<uintptr_t>(out + i0)) & 15) == 0) {
        float4 v[kCfcPer / 2];
#pragma unroll
        for (int j = 0; j < kCfcPer / 2; ++j) v[j] = __ldcs(reinterpret_cast<const float4*>(in + i0) + j);
#pragma unroll
        for (int j = 0; j < kCfcPer / 2; ++j) {
            const float2 a = cfc_apply(segs, n_segs, c, abs0 + i0 + 2 * j, make_float2(v[j].x, v[j].y));
            const float2 b = cfc_apply(segs, n_segs, c, abs0 + i0 + 2 * j + 1, make_float2(v[j].z, v[j].w));
            __stcs(reinterpret_cast<float4*>(out + i0) + j, make_float4(a.x, a.y, b.x, b.y));
        }
    } else {
        for (int j = 0; j < kCfcPer && i0 + j < n; ++j) out[i0 + j] = cfc_apply(segs, n_segs, c, abs0 + i0 + j, in[i0 + j]);
    }
}

}  // namespace b200sync

using namespace b200sync;

namespace {
thread_local std::string g_cfc_error;
int cfc_fail(int code, const std::string& m) {
    g_cfc_error = m;
    return code;
}
#define CCU(expr)                                                                                     \
    do {                                                                                              \
        cudaError_t _e = (expr);                                                                      \
        if (_e != cudaSuccess) return cfc_fail(B200SYNC_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
    } while (0)
}  // namespace

struct b200sync_cfc {
    uint32_t delay = 0;
    int device = 0;
    CfcPlanner plan;
    CfcSeed* d_seeds = nullptr;
    CfcSegment* d_segs = nullptr;
    size_t segs_cap = 0;
    float2* d_in = nullptr;
    float2* d_out = nullptr;
    size_t buf_cap = 0;
    cudaStream_t stream = nullptr;
    std::vector<CfcSeed> live;
};

namespace {
int cfc_run(b200sync_cfc* c, const float2* d_in, size_t n, const b200sync_stream_tag* tags, size_t n_tags,
            float2* d_out, cudaStream_t st) {
    for (size_t i = 0; i < n_tags; ++i)
        if (tags[i].index >= n || (i > 0 && tags[i].index < tags[i - 1].index))
            return cfc_fail(B200SYNC_EINVAL, "input tags must be sorted by index and inside the span");
    if (n == 0) return 0;
    const long long abs0 = c->plan.abs_pos();
    c->plan.advance(n, tags, n_tags);
    c->plan.live_segments(abs0, c->live);
    CCU(cfc_upload_segments(c->live, &c->d_seeds, &c->d_segs, &c->segs_cap, st));
    const unsigned grid = static_cast<unsigned>((n + kCfcTile - 1) / kCfcTile);
    cfc_kernel<<<grid, kCfcThreads, 0, st>>>(d_in, d_out, static_cast<long long>(n), abs0, c->d_segs,
                                             static_cast<int>(c->live.size()));
    count_launch();
    CCU(cudaGetLastError());
    CCU(cudaStreamSynchronize(st));  // the pageable segment vector must outlive the async copy
    c->plan.prune(c->plan.abs_pos());
    return 0;
}
}  // namespace

extern "C" {

const char* b200sync_cfc_last_error(void) { return g_cfc_error.c_str(); }

int b200sync_cfc_create(uint32_t delay, int32_t device, b200sync_cfc** out) {
    if (!out) return cfc_fail(B200SYNC_EINVAL, "null argument");
    *out = nullptr;
    b200sync_cfc* c = new (std::nothrow) b200sync_cfc();
    if (!c) return cfc_fail(B200SYNC_ENOMEM, "out of memory");
    c->delay = delay;
    c->device = device;
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        delete c;
        return cfc_fail(B200SYNC_ECUDA, std::string("no usable CUDA device: ") + cudaGetErrorString(e));
    }
    c->plan.reset(delay);
    *out = c;
    return 0;
}

void b200sync_cfc_destroy(b200sync_cfc* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) {
        cudaStreamSynchronize(c->stream);
        cudaStreamDestroy(c->stream);
    }
    if (c->d_segs) cudaFree(c->d_segs);
    if (c->d_seeds) cudaFree(c->d_seeds);
    if (c->d_in) cudaFree(c->d_in);
    if (c->d_out) cudaFree(c->d_out);
    delete c;
}

int b200sync_cfc_start(b200sync_cfc* c) {
    if (!c) return cfc_fail(B200SYNC_EINVAL, "null context");
    c->plan.reset(c->delay);
    return 0;
}

int b200sync_cfc_process_device(b200sync_cfc* c, const void* d_in, size_t n, const b200sync_stream_tag* in_tags,
                                size_t n_in_tags, void* d_out, void* cuda_stream) {
    if (!c || (!d_in && n) || (!d_out && n) || (!in_tags && n_in_tags)) return cfc_fail(B200SYNC_EINVAL, "null argument");
    CCU(cudaSetDevice(c->device));
    return cfc_run(c, static_cast<const float2*>(d_in), n, in_tags, n_in_tags, static_cast<float2*>(d_out),
                   static_cast<cudaStream_t>(cuda_stream));
}

int b200sync_cfc_process(b200sync_cfc* c, const float* in, size_t n, const b200sync_stream_tag* in_tags,
                         size_t n_in_tags, float* out) {
    if (!c || (!in && n) || (!out && n) || (!in_tags && n_in_tags)) return cfc_fail(B200SYNC_EINVAL, "null argument");
    CCU(cudaSetDevice(c->device));
    if (c->buf_cap < n) {
        if (c->d_in) cudaFree(c->d_in);
        if (c->d_out) cudaFree(c->d_out);
        c->d_in = c->d_out = nullptr;
        c->buf_cap = 0;
        CCU(cudaMalloc(&c->d_in, n * sizeof(float2)));
        CCU(cudaMalloc(&c->d_out, n * sizeof(float2)));
        c->buf_cap = n;
    }
    CCU(cudaMemcpyAsync(c->d_in, in, n * sizeof(float2), cudaMemcpyHostToDevice, c->stream));
    const int rc = cfc_run(c, c->d_in, n, in_tags, n_in_tags, c->d_out, c->stream);
    if (rc != 0) return rc;
    CCU(cudaMemcpyAsync(out, c->d_out, n * sizeof(float2), cudaMemcpyDeviceToHost, c->stream));
    CCU(cudaStreamSynchronize(c->stream));
    return 0;
}

}  // extern "C"
