// costas.cu — SyncwordWipeoff (PM/syncword_wipeoff.hpp:38-91) and CostasLoop (PM/costas_loop.hpp:92-149)
// for sm_100a, behind b200sync_wo_* / b200sync_cl_* (SURVEY §8(f) rank 2).
//
// CostasLoop is a sequential recurrence per symbol, reset by every "syncword_phase" tag: the host cuts the
// span into the stretches between such tags and the kernel runs ONE THREAD PER STRETCH.  A warp owns 32
// stretches and walks them in tiles of 32 symbols: the warp loads 32 x 256 B coalesced rows (one per
// stretch) into a padded shared-memory tile, every lane then runs the recurrence over its own row in
// place, and the warp stores the rows back coalesced.  HBM traffic is the algorithmic 16 B per symbol;
// the run time is the latency of the longest dependent chain (one packet), so the path scales with the
// number of packets in the span, like the detection stage in front of it.
// SyncwordWipeoff multiplies the 64 symbols after a "syncword_amplitude" tag by the bipolar syncword; fused
// into the loop's load stage it costs nothing (b200sync_cl_fuse_wipeoff).
#include <algorithm>
#include <cmath>
#include <new>
#include <string>
#include <vector>

#include "b200sync_internal.h"
#include "costas.cuh"

namespace b200sync {

constexpr long long kNoWipe = INT64_MIN;

struct ClSegment {
    long long start, end;    // items [start, end) of the span
    long long wipe_start;    // syncword wipe-off interval [wipe_start, wipe_start + n_sync) or kNoWipe
    float phase0;            // set_phase() value (ignored when carry)
    unsigned int carry;      // 1: continue from the device-resident state of the previous call
};

constexpr int kClTile = 32;    // symbols per row of a tile
constexpr int kClWarps = 4;    // warps per CTA
constexpr int kClRow = kClTile + 1;  // float2 row stride: lane l, column j -> banks 2(l + j), half-warps disjoint

template <int CONSTELLATION>
__global__ void __launch_bounds__(kClWarps * 32)
costas_kernel(const float2* in, float2* out, const ClSegment* __restrict__ segs,
              int n_segs, float k1, float k2, const ClState* __restrict__ state_in, ClState* __restrict__ state_out,
              const float* __restrict__ syncword, int n_sync) {
    __shared__ float2 tile[kClWarps][32][kClRow];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int sg = (blockIdx.x * kClWarps + warp) * 32 + lane;
    constexpr unsigned kFull = 0xffffffffu;
    long long pos = 0, end = 0, wipe = kNoWipe;
    ClState st{0.0f, 0.0f};
    if (sg < n_segs) {
        const ClSegment s = segs[sg];
        pos = s.start;
        end = s.end;
        wipe = s.wipe_start;
        if (s.carry) st = *state_in;
        else st.phase = s.phase0;
    }
    float2(*rows)[kClRow] = tile[warp];
    while (__any_sync(kFull, pos < end)) {
        {
            // all 32 row loads of the tile are issued before the first one is used (predicated, no branches:
            // a guarded load inside a branch is not hoisted and every row would expose its own DRAM latency)
            float2 v[32];
#pragma unroll
            for (int s = 0; s < 32; ++s) {
                const long long b = __shfl_sync(kFull, pos, s) + lane;
                const long long e = __shfl_sync(kFull, end, s);
                const float2* src = in + (b < e ? b : 0);
                v[s] = make_float2(0.0f, 0.0f);
                if (b < e) v[s] = *src;
            }
            // SyncwordWipeoff: *out_item++ = *in_item++ * syncword[_position++] (:68-70); x * 1.0f == x elsewhere
#pragma unroll
            for (int s = 0; s < 32; ++s) {
                const long long b = __shfl_sync(kFull, pos, s) + lane;
                const long long w = __shfl_sync(kFull, wipe, s);
                const long long d = b - w;
                float sw = 1.0f;
                if (w != kNoWipe && d >= 0 && d < n_sync) sw = __ldg(syncword + d);
                rows[s][lane] = make_float2(__fmul_rn(v[s].x, sw), __fmul_rn(v[s].y, sw));
            }
        }
        __syncwarp();
        const long long left = end - pos;
        const int cnt = left >= kClTile ? kClTile : (left > 0 ? static_cast<int>(left) : 0);
        {
            // the lane's row lives in registers while the recurrence runs: no shared-memory latency on the
            // dependent chain phase -> sincos -> error -> phase
            float2 r[kClTile];
#pragma unroll
            for (int j = 0; j < kClTile; ++j) r[j] = rows[lane][j];
#pragma unroll
            for (int j = 0; j < kClTile; ++j)
                if (j < cnt) r[j] = costas_step<CONSTELLATION>(r[j], st, k1, k2);
#pragma unroll
            for (int j = 0; j < kClTile; ++j) rows[lane][j] = r[j];
        }
        __syncwarp();
#pragma unroll 8
        for (int s = 0; s < 32; ++s) {
            const long long b = __shfl_sync(kFull, pos, s) + lane;
            const long long e = __shfl_sync(kFull, end, s);
            if (b < e) out[b] = rows[s][lane];
        }
        __syncwarp();
        if (pos < end) pos += kClTile;
    }
    // the last stretch of the span hands its state to the next call — through the OTHER slot of a two-slot
    // buffer, so the CTA of the first stretch can never read a value written by this launch
    if (sg == n_segs - 1) *state_out = st;
}

// SyncwordWipeoff alone: the pass-through copy is a device-to-device copy issued by the caller; this kernel
// rewrites the (few) wiped intervals from the input.  One CTA of 64 threads per interval.
__global__ void wipeoff_kernel(const float2* in, float2* out, long long n,
                               const long long* __restrict__ starts, const float* __restrict__ syncword,
                               int n_sync) {
    const long long w = starts[blockIdx.x];
    for (int d = threadIdx.x; d < n_sync; d += blockDim.x) {
        const long long i = w + d;
        if (i >= 0 && i < n) {
            const float2 x = in[i];
            const float sw = syncword[d];
            out[i] = make_float2(__fmul_rn(x.x, sw), __fmul_rn(x.y, sw));
        }
    }
}

// Which items does SyncwordWipeoff multiply?  Host replay of the block's state machine
// (PM/syncword_wipeoff.hpp:52-75): a tag starts an interval only when none is in progress.
struct WipeoffPlanner {
    long long n_sync = 0;
    bool in_syncword = false;  // _in_syncword
    long long position = 0;    // _position
    void reset() { in_syncword = false; position = 0; }
    // interval starts relative to the span (a carried-in interval starts at -position < 0)
    void plan(size_t n, const b200sync_stream_tag* tags, size_t n_tags, std::vector<long long>& starts) {
        starts.clear();
        if (n_sync == 0) return;
        long long until = 0;  // first item after the interval in progress
        if (in_syncword) {
            starts.push_back(-position);
            until = n_sync - position;
        }
        for (size_t i = 0; i < n_tags; ++i) {
            if (!tags[i].has_syncword) continue;
            const long long p = static_cast<long long>(tags[i].index);
            if (p < until) continue;  // still inside a syncword: the tag is not looked at (:52)
            starts.push_back(p);
            until = p + n_sync;
        }
        const long long nn = static_cast<long long>(n);
        if (!starts.empty() && until > nn) {
            in_syncword = true;
            position = nn - starts.back();
        } else {
            in_syncword = false;
            position = starts.empty() ? position : n_sync;
        }
    }
};

}  // namespace b200sync

using namespace b200sync;

namespace {
thread_local std::string g_cl_error;
int cl_fail(int code, const std::string& m) {
    g_cl_error = m;
    return code;
}
#define LCU(expr)                                                                                     \
    do {                                                                                              \
        cudaError_t _e = (expr);                                                                      \
        if (_e != cudaSuccess) return cl_fail(B200SYNC_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
    } while (0)

int check_tags(size_t n, const b200sync_stream_tag* tags, size_t n_tags) {
    for (size_t i = 0; i < n_tags; ++i)
        if (tags[i].index >= n || (i > 0 && tags[i].index < tags[i - 1].index))
            return cl_fail(B200SYNC_EINVAL, "input tags must be sorted by index and inside the span");
    return 0;
}

struct HostStage {  // device staging of host spans
    float2* d_in = nullptr;
    float2* d_out = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t n) {
        if (cap >= n) return cudaSuccess;
        release();
        cudaError_t e = cudaMalloc(&d_in, n * sizeof(float2));
        if (e == cudaSuccess) e = cudaMalloc(&d_out, n * sizeof(float2));
        if (e == cudaSuccess) cap = n;
        return e;
    }
    void release() {
        if (d_in) cudaFree(d_in);
        if (d_out) cudaFree(d_out);
        d_in = d_out = nullptr;
        cap = 0;
    }
};
}  // namespace

struct b200sync_wo {
    int device = 0;
    WipeoffPlanner plan;
    float* d_syncword = nullptr;
    long long* d_starts = nullptr;
    size_t starts_cap = 0;
    std::vector<long long> starts;
    HostStage stage;
    cudaStream_t stream = nullptr;
};

struct b200sync_cl {
    int device = 0;
    int constellation = B200SYNC_CONSTELLATION_BPSK;
    double loop_bandwidth = 0.01;
    float k1 = 0.0f, k2 = 0.0f;
    bool fresh = true;              // no item processed since start(): the state is (0, 0)
    ClState* d_state = nullptr;     // two slots, ping-pong per call
    int state_cur = 0;              // slot holding the state after the last call
    WipeoffPlanner wipe;            // n_sync == 0: no fused SyncwordWipeoff
    float* d_syncword = nullptr;
    ClSegment* d_segs = nullptr;
    size_t segs_cap = 0;
    std::vector<ClSegment> segs;
    std::vector<long long> starts;
    HostStage stage;
    cudaStream_t stream = nullptr;
};

namespace {

int wo_run(b200sync_wo* w, const float2* d_in, size_t n, const b200sync_stream_tag* tags, size_t n_tags,
           float2* d_out, cudaStream_t st) {
    if (int rc = check_tags(n, tags, n_tags)) return rc;
    if (n == 0) return 0;
    w->plan.plan(n, tags, n_tags, w->starts);
    if (d_in != d_out) LCU(cudaMemcpyAsync(d_out, d_in, n * sizeof(float2), cudaMemcpyDeviceToDevice, st));
    if (w->starts.empty()) return 0;
    if (w->starts_cap < w->starts.size()) {
        if (w->d_starts) cudaFree(w->d_starts);
        w->d_starts = nullptr;
        w->starts_cap = 0;
        const size_t want = w->starts.size() * 2 + 64;
        LCU(cudaMalloc(&w->d_starts, want * sizeof(long long)));
        w->starts_cap = want;
    }
    LCU(cudaMemcpyAsync(w->d_starts, w->starts.data(), w->starts.size() * sizeof(long long), cudaMemcpyHostToDevice, st));
    wipeoff_kernel<<<static_cast<unsigned>(w->starts.size()), 64, 0, st>>>(
        d_in, d_out, static_cast<long long>(n), w->d_starts, w->d_syncword, static_cast<int>(w->plan.n_sync));
    count_launch();
    LCU(cudaGetLastError());
    LCU(cudaStreamSynchronize(st));  // the pageable interval vector must outlive the async copy
    return 0;
}

int cl_run(b200sync_cl* c, const float2* d_in, size_t n, const b200sync_stream_tag* tags, size_t n_tags,
           float2* d_out, cudaStream_t st) {
    if (int rc = check_tags(n, tags, n_tags)) return rc;
    if (n == 0) return 0;
    // stretches between set_phase() tags; the first one continues the previous call (or the initial state)
    auto& segs = c->segs;
    segs.clear();
    ClSegment first{0, 0, kNoWipe, 0.0f, c->fresh ? 0u : 1u};
    segs.push_back(first);
    for (size_t i = 0; i < n_tags; ++i) {
        if (!tags[i].has_syncword) continue;
        const long long p = static_cast<long long>(tags[i].index);
        ClSegment s{p, 0, kNoWipe, tags[i].sw.syncword_phase, 0u};
        if (segs.back().start == p) segs.back() = s;  // set_phase() before the stretch's first item wins
        else segs.push_back(s);
    }
    for (size_t i = 0; i + 1 < segs.size(); ++i) segs[i].end = segs[i + 1].start;
    segs.back().end = static_cast<long long>(n);
    if (c->wipe.n_sync > 0) {
        // every wipe-off interval starts on a stretch boundary (same tags) or is carried in: a stretch
        // sees the latest interval that started at or before it
        c->wipe.plan(n, tags, n_tags, c->starts);
        size_t k = 0;
        for (auto& s : segs) {
            while (k + 1 < c->starts.size() && c->starts[k + 1] <= s.start) ++k;
            if (!c->starts.empty() && c->starts[k] <= s.start && c->starts[k] + c->wipe.n_sync > s.start)
                s.wipe_start = c->starts[k];
        }
    }
    if (c->segs_cap < segs.size()) {
        if (c->d_segs) cudaFree(c->d_segs);
        c->d_segs = nullptr;
        c->segs_cap = 0;
        const size_t want = segs.size() * 2 + 64;
        LCU(cudaMalloc(&c->d_segs, want * sizeof(ClSegment)));
        c->segs_cap = want;
    }
    LCU(cudaMemcpyAsync(c->d_segs, segs.data(), segs.size() * sizeof(ClSegment), cudaMemcpyHostToDevice, st));
    const int n_segs = static_cast<int>(segs.size());
    const unsigned grid = static_cast<unsigned>((n_segs + kClWarps * 32 - 1) / (kClWarps * 32));
    const int n_sync = static_cast<int>(c->wipe.n_sync);
    switch (c->constellation) {
    case B200SYNC_CONSTELLATION_PILOT:
        costas_kernel<kClPilot><<<grid, kClWarps * 32, 0, st>>>(d_in, d_out, c->d_segs, n_segs, c->k1, c->k2,
                                                                c->d_state + c->state_cur, c->d_state + (c->state_cur ^ 1), c->d_syncword, n_sync);
        break;
    case B200SYNC_CONSTELLATION_BPSK:
        costas_kernel<kClBpsk><<<grid, kClWarps * 32, 0, st>>>(d_in, d_out, c->d_segs, n_segs, c->k1, c->k2,
                                                               c->d_state + c->state_cur, c->d_state + (c->state_cur ^ 1), c->d_syncword, n_sync);
        break;
    default:
        costas_kernel<kClQpsk><<<grid, kClWarps * 32, 0, st>>>(d_in, d_out, c->d_segs, n_segs, c->k1, c->k2,
                                                               c->d_state + c->state_cur, c->d_state + (c->state_cur ^ 1), c->d_syncword, n_sync);
        break;
    }
    count_launch();
    LCU(cudaGetLastError());
    LCU(cudaStreamSynchronize(st));  // the pageable segment vector must outlive the async copy
    c->fresh = false;
    c->state_cur ^= 1;
    return 0;
}

// CostasLoop::settingsChanged(), PM/costas_loop.hpp:56-90: closed-form root of the cubic in B_L * T
void loop_coefficients(double bw, int constellation, float* k1f, float* k2f) {
    const double gain = constellation == B200SYNC_CONSTELLATION_QPSK ? 1.41421356237309504880 : 1.0;
    const double b2 = bw * bw, b3 = b2 * bw, b4 = b2 * b2;
    const double s = std::cbrt(36.0 * b2 +
                               std::sqrt(3.0) * std::sqrt(432.0 * b4 + 848.0 * b3 + 624.0 * b2 + 204.0 * bw + 25.0) +
                               36.0 * bw + 9.0);
    const double z = -(-12.0 * bw - 6.0) / (3.0 * std::cbrt(6.0) * (2.0 * bw + 1.0) * s) +
                     (std::cbrt(2.0) * s) / (std::cbrt(9.0) * (2.0 * bw + 1.0)) - 1.0;
    *k1f = static_cast<float>((1.0 - z * z) / gain);
    *k2f = static_cast<float>((1.0 - z) * (1.0 - z) / gain);
}

cudaError_t upload_syncword(float** d, const float* syncword, uint32_t n) {
    if (*d) cudaFree(*d);
    *d = nullptr;
    if (n == 0) return cudaSuccess;
    cudaError_t e = cudaMalloc(d, n * sizeof(float));
    if (e == cudaSuccess) e = cudaMemcpy(*d, syncword, n * sizeof(float), cudaMemcpyHostToDevice);
    return e;
}

template <class Ctx, class Run>
int process_host(Ctx* c, const float* in, size_t n, const b200sync_stream_tag* tags, size_t n_tags, float* out,
                 Run run) {
    LCU(cudaSetDevice(c->device));
    LCU(c->stage.ensure(n));
    if (n) LCU(cudaMemcpyAsync(c->stage.d_in, in, n * sizeof(float2), cudaMemcpyHostToDevice, c->stream));
    const int rc = run(c, c->stage.d_in, n, tags, n_tags, c->stage.d_out, c->stream);
    if (rc != 0) return rc;
    if (n) LCU(cudaMemcpyAsync(out, c->stage.d_out, n * sizeof(float2), cudaMemcpyDeviceToHost, c->stream));
    LCU(cudaStreamSynchronize(c->stream));
    return 0;
}
}  // namespace

extern "C" {

const char* b200sync_cl_last_error(void) { return g_cl_error.c_str(); }

// ---- SyncwordWipeoff ------------------------------------------------------------------------------
int b200sync_wo_create(const float* syncword, uint32_t n_syncword, int32_t device, b200sync_wo** out) {
    if (!out || (!syncword && n_syncword)) return cl_fail(B200SYNC_EINVAL, "null argument");
    *out = nullptr;
    b200sync_wo* w = new (std::nothrow) b200sync_wo();
    if (!w) return cl_fail(B200SYNC_ENOMEM, "out of memory");
    w->device = device;
    w->plan.n_sync = n_syncword;
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&w->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = upload_syncword(&w->d_syncword, syncword, n_syncword);
    if (e != cudaSuccess) {
        b200sync_wo_destroy(w);
        return cl_fail(B200SYNC_ECUDA, std::string("no usable CUDA device: ") + cudaGetErrorString(e));
    }
    *out = w;
    return 0;
}

void b200sync_wo_destroy(b200sync_wo* w) {
    if (!w) return;
    cudaSetDevice(w->device);
    if (w->stream) {
        cudaStreamSynchronize(w->stream);
        cudaStreamDestroy(w->stream);
    }
    if (w->d_syncword) cudaFree(w->d_syncword);
    if (w->d_starts) cudaFree(w->d_starts);
    w->stage.release();
    delete w;
}

int b200sync_wo_start(b200sync_wo* w) {
    if (!w) return cl_fail(B200SYNC_EINVAL, "null context");
    w->plan.reset();
    return 0;
}

int b200sync_wo_process_device(b200sync_wo* w, const void* d_in, size_t n, const b200sync_stream_tag* in_tags,
                               size_t n_in_tags, void* d_out, void* cuda_stream) {
    if (!w || (!d_in && n) || (!d_out && n) || (!in_tags && n_in_tags)) return cl_fail(B200SYNC_EINVAL, "null argument");
    LCU(cudaSetDevice(w->device));
    return wo_run(w, static_cast<const float2*>(d_in), n, in_tags, n_in_tags, static_cast<float2*>(d_out),
                  static_cast<cudaStream_t>(cuda_stream));
}

int b200sync_wo_process(b200sync_wo* w, const float* in, size_t n, const b200sync_stream_tag* in_tags,
                        size_t n_in_tags, float* out) {
    if (!w || (!in && n) || (!out && n) || (!in_tags && n_in_tags)) return cl_fail(B200SYNC_EINVAL, "null argument");
    return process_host(w, in, n, in_tags, n_in_tags, out, wo_run);
}

// ---- CostasLoop -----------------------------------------------------------------------------------
int b200sync_cl_create(const b200sync_cl_config* cfg, b200sync_cl** out) {
    if (!cfg || !out) return cl_fail(B200SYNC_EINVAL, "null argument");
    *out = nullptr;
    if (cfg->constellation > B200SYNC_CONSTELLATION_QPSK)  // magic_enum::enum_cast(...).value() throws (:63-65)
        return cl_fail(B200SYNC_EINVAL, "constellation must be PILOT, BPSK or QPSK");
    b200sync_cl* c = new (std::nothrow) b200sync_cl();
    if (!c) return cl_fail(B200SYNC_ENOMEM, "out of memory");
    c->device = cfg->device;
    c->constellation = static_cast<int>(cfg->constellation);
    c->loop_bandwidth = cfg->loop_bandwidth;
    loop_coefficients(cfg->loop_bandwidth, c->constellation, &c->k1, &c->k2);
    cudaError_t e = cudaSetDevice(cfg->device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMalloc(&c->d_state, 2 * sizeof(ClState));
    if (e == cudaSuccess) e = cudaMemset(c->d_state, 0, 2 * sizeof(ClState));
    if (e != cudaSuccess) {
        b200sync_cl_destroy(c);
        return cl_fail(B200SYNC_ECUDA, std::string("no usable CUDA device: ") + cudaGetErrorString(e));
    }
    *out = c;
    return 0;
}

void b200sync_cl_destroy(b200sync_cl* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) {
        cudaStreamSynchronize(c->stream);
        cudaStreamDestroy(c->stream);
    }
    if (c->d_state) cudaFree(c->d_state);
    if (c->d_syncword) cudaFree(c->d_syncword);
    if (c->d_segs) cudaFree(c->d_segs);
    c->stage.release();
    delete c;
}

int b200sync_cl_start(b200sync_cl* c) {
    if (!c) return cl_fail(B200SYNC_EINVAL, "null context");
    c->fresh = true;
    c->wipe.reset();
    LCU(cudaSetDevice(c->device));
    LCU(cudaMemset(c->d_state, 0, 2 * sizeof(ClState)));
    c->state_cur = 0;
    return 0;
}

int b200sync_cl_info(const b200sync_cl* c, float* k1, float* k2) {
    if (!c) return cl_fail(B200SYNC_EINVAL, "null context");
    if (k1) *k1 = c->k1;
    if (k2) *k2 = c->k2;
    return 0;
}

int b200sync_cl_fuse_wipeoff(b200sync_cl* c, const float* syncword, uint32_t n_syncword) {
    if (!c || (!syncword && n_syncword)) return cl_fail(B200SYNC_EINVAL, "null argument");
    LCU(cudaSetDevice(c->device));
    LCU(upload_syncword(&c->d_syncword, syncword, n_syncword));
    c->wipe.n_sync = n_syncword;
    c->wipe.reset();
    return 0;
}

int b200sync_cl_state(b200sync_cl* c, float* phase, float* freq) {
    if (!c) return cl_fail(B200SYNC_EINVAL, "null context");
    LCU(cudaSetDevice(c->device));
    ClState s;
    LCU(cudaMemcpy(&s, c->d_state + c->state_cur, sizeof(s), cudaMemcpyDeviceToHost));
    if (phase) *phase = s.phase;
    if (freq) *freq = s.freq;
    return 0;
}

int b200sync_cl_process_device(b200sync_cl* c, const void* d_in, size_t n, const b200sync_stream_tag* in_tags,
                               size_t n_in_tags, void* d_out, void* cuda_stream) {
    if (!c || (!d_in && n) || (!d_out && n) || (!in_tags && n_in_tags)) return cl_fail(B200SYNC_EINVAL, "null argument");
    LCU(cudaSetDevice(c->device));
    return cl_run(c, static_cast<const float2*>(d_in), n, in_tags, n_in_tags, static_cast<float2*>(d_out),
                  static_cast<cudaStream_t>(cuda_stream));
}

int b200sync_cl_process(b200sync_cl* c, const float* in, size_t n, const b200sync_stream_tag* in_tags,
                        size_t n_in_tags, float* out) {
    if (!c || (!in && n) || (!out && n) || (!in_tags && n_in_tags)) return cl_fail(B200SYNC_EINVAL, "null argument");
    return process_host(c, in, n, in_tags, n_in_tags, out, cl_run);
}

}  // extern "C"
