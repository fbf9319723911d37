// stimulus.cu — on-GPU synthetic capture generator for sm_100a: the transmit-side
// blocks that feed the receiver in apps/packet_transceiver.cpp:60-80, 140-165 collapsed into ONE pass that
// writes the capture straight into HBM (8 B/sample, nothing read):
//   symbols (syncword + header + payload frames)
//     -> InterpolatingFirFilter<c64, c64, float>{interpolation, taps}  PM/interpolating_fir_filter.hpp:93-99
//     -> Rotator{phase_incr}                                           PM/rotator.hpp:56-65 (closed form)
//     -> Add<c64>(., NoiseSource<c64>{"gaussian", amplitude})          PM/add.hpp, PM/noise_source.hpp:74-78
// Every sample is a pure function of (seed, absolute sample index): symbols and noise come from a 32-bit
// integer hash of the index instead of the reference's sequential generators (GlfsrSource, std::mt19937), so
// any time shard (plus its overlap-save halo) can be generated on its own GPU and agrees bit for bit with
// its neighbours and with a single-GPU run.
//
// ARITHMETIC: the FIR is the reference's std::inner_product over one polyphase branch, newest item first,
// complex<float> x float, every multiply and add separately rounded — bit-exact against the oracle's
// restated InterpolatingFirFilter on the same symbols.  The rotation is exp(i * phase_incr * n) with the
// phase reduced in double (separately rounded ops) and b200_sincosf (costas.cuh), mirrored op for op by the
// tests; against the reference's float recurrence it holds the rotator tolerance (rel-L2 < 1e-5 per 2^15
// samples).  Noise is Box-Muller on two hashed uniforms: Rayleigh amplitude sqrt(-2 ln u1) as in
// random::rayleigh() (PM/random.hpp:194), uniform phase; statistical parity (test/qa_noise_source.cpp:40-44:
// power within 1 %).
#include <new>
#include <string>
#include <vector>

#include "b200sync_internal.h"
#include "costas.cuh"

namespace b200sync {

struct StimParams {
    float2* out;
    long long first;     // absolute index of out[0]
    long long n;
    int interp, arm_len; // samples per symbol, taps per polyphase branch
    int n_sync, n_hdr_payload, frame_len;  // symbols
    double theta;        // Rotator phase increment as the float the block holds
    float sigma;         // NoiseSource _amplitude_complex = amplitude / sqrt(2): std per component
    unsigned int sym_salt, n1_salt, n2_salt;
    int rotate, noise;
};

__host__ __device__ __forceinline__ unsigned int stim_hash(long long idx, unsigned int salt) {
    const unsigned int lo = (unsigned int)((unsigned long long)idx & 0xFFFFFFFFull);
    const unsigned int hi = (unsigned int)(((unsigned long long)idx >> 32) & 0xFFFFFFFFull);
    unsigned int h = lo * 2654435761u + hi * 40503u + salt;
    h = (h ^ (h >> 16)) * 2246822507u;
    h = (h ^ (h >> 13)) * 3266489909u;
    return h ^ (h >> 16);
}

// symbol k of the endless frame sequence: BPSK syncword (real), QPSK header + payload, zeros in the gap
// (f = k mod frame_len is passed in: one 64-bit modulo per thread instead of one per symbol)
__device__ __forceinline__ float2 stim_symbol(const StimParams& P, const float* __restrict__ sync_s, long long k, int f) {
    if (k < 0) return make_float2(0.0f, 0.0f);
    if (f < P.n_sync) return make_float2(sync_s[f], 0.0f);
    if (f >= P.n_sync + P.n_hdr_payload) return make_float2(0.0f, 0.0f);
    const unsigned int h = stim_hash(k, P.sym_salt);
    constexpr float a = 0.70710678118654752440f;
    return make_float2((h & 1u) ? -a : a, (h & 2u) ? -a : a);
}

constexpr int kStimThreads = 256;
constexpr int kStimMaxArm = 32;  // taps per branch held in registers' worth of symbols

// One thread per symbol period: it gathers the arm_len symbols the period's outputs depend on once and
// produces the `interp` outputs of the period (consecutive samples: a warp writes 32 * interp * 8 B in a row).
__global__ void __launch_bounds__(kStimThreads)
stimulus_kernel(const StimParams P, const float* __restrict__ taps_poly /*[interp][arm_len]*/,
                const float* __restrict__ sync_g) {
    extern __shared__ float sm[];
    float* taps_s = sm;                              // [interp][arm_len]
    float* sync_s = sm + P.interp * P.arm_len;       // [n_sync]
    for (int i = threadIdx.x; i < P.interp * P.arm_len; i += blockDim.x) taps_s[i] = taps_poly[i];
    for (int i = threadIdx.x; i < P.n_sync; i += blockDim.x) sync_s[i] = sync_g[i];
    __syncthreads();
    const long long k0 = P.first / P.interp;  // first symbol period touched (first >= 0)
    const long long k = k0 + (long long)blockIdx.x * kStimThreads + threadIdx.x;
    const long long s_lo = k * P.interp;
    if (s_lo >= P.first + P.n) return;
    float2 sym[kStimMaxArm];  // sym[j] = symbol k - j: the history, newest first (GR/HistoryBuffer.hpp)
    int f = (int)(k % P.frame_len);  // position of symbol k in its frame; steps back with wrap-around
#pragma unroll
    for (int j = 0; j < kStimMaxArm; ++j)
        if (j < P.arm_len) {
            sym[j] = stim_symbol(P, sync_s, k - j, f);
            f = (f == 0) ? P.frame_len - 1 : f - 1;
        }
    for (int arm = 0; arm < P.interp; ++arm) {
        const long long n = s_lo + arm;
        if (n < P.first || n >= P.first + P.n) continue;
        // std::inner_product(branch.cbegin(), branch.cend(), _history.cbegin(), TOut{0}) (:96-97)
        float2 acc = make_float2(0.0f, 0.0f);
        const float* t = taps_s + arm * P.arm_len;
#pragma unroll
        for (int j = 0; j < kStimMaxArm; ++j)
            if (j < P.arm_len) {
                acc.x = __fadd_rn(acc.x, __fmul_rn(sym[j].x, t[j]));
                acc.y = __fadd_rn(acc.y, __fmul_rn(sym[j].y, t[j]));
            }
        if (P.rotate) {
            double ph = __dmul_rn((double)n, P.theta);
            ph = __dsub_rn(ph, __dmul_rn(6.283185307179586476925, rint(__dmul_rn(ph, 0.15915494309189533577))));
            float s, c;
            b200_sincosf((float)ph, s, c);
            acc = make_float2(__fsub_rn(__fmul_rn(acc.x, c), __fmul_rn(acc.y, s)),
                              __fadd_rn(__fmul_rn(acc.x, s), __fmul_rn(acc.y, c)));
        }
        if (P.noise) {
            const float u1 = __fmul_rn(__fadd_rn((float)stim_hash(n, P.n1_salt), 1.0f), 2.3283064365386963e-10f);
            const float u2 = __fmul_rn((float)stim_hash(n, P.n2_salt), 1.4629180792671596e-09f);  // 2 pi / 2^32
            const float r = __fmul_rn(sqrtf(__fmul_rn(-2.0f, logf(u1))), P.sigma);
            float s, c;
            b200_sincosf(u2, s, c);
            acc.x = __fadd_rn(acc.x, __fmul_rn(r, c));
            acc.y = __fadd_rn(acc.y, __fmul_rn(r, s));
        }
        P.out[n - P.first] = acc;
    }
}

}  // namespace b200sync

using namespace b200sync;

namespace {
thread_local std::string g_stim_error;
int stim_fail(int code, const std::string& m) {
    g_stim_error = m;
    return code;
}
#define SCU(expr)                                                                                     \
    do {                                                                                              \
        cudaError_t _e = (expr);                                                                      \
        if (_e != cudaSuccess) return stim_fail(B200SYNC_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
    } while (0)
}  // namespace

struct b200sync_stim {
    int device = 0;
    StimParams P{};
    float* d_taps = nullptr;
    float* d_sync = nullptr;
    size_t smem = 0;
};

extern "C" {

const char* b200sync_stim_last_error(void) { return g_stim_error.c_str(); }

int b200sync_stim_create(const b200sync_stim_config* cfg, b200sync_stim** out) {
    if (!cfg || !out || !cfg->taps || (!cfg->syncword_symbols && cfg->n_syncword))
        return stim_fail(B200SYNC_EINVAL, "null argument");
    *out = nullptr;
    if (cfg->interpolation == 0 || cfg->n_taps == 0) return stim_fail(B200SYNC_EINVAL, "interpolation and taps must not be empty");
    const uint32_t arm_len = (cfg->n_taps + cfg->interpolation - 1) / cfg->interpolation;
    if (arm_len > kStimMaxArm)
        return stim_fail(B200SYNC_EUNSUPPORTED, "more than 32 taps per polyphase branch");
    const uint64_t frame = static_cast<uint64_t>(cfg->n_syncword) + cfg->header_symbols + cfg->payload_symbols + cfg->gap_symbols;
    if (frame == 0 || frame > 0x7fffffffull) return stim_fail(B200SYNC_EINVAL, "frame length must be 1 .. 2^31-1 symbols");
    b200sync_stim* s = new (std::nothrow) b200sync_stim();
    if (!s) return stim_fail(B200SYNC_ENOMEM, "out of memory");
    s->device = cfg->device;
    // polyphase split, branch j = taps[j::interpolation], zero padded (PM/interpolating_fir_filter.hpp:52-60)
    std::vector<float> poly(static_cast<size_t>(cfg->interpolation) * arm_len, 0.0f);
    for (uint32_t i = 0; i < cfg->n_taps; ++i) poly[(i % cfg->interpolation) * arm_len + i / cfg->interpolation] = cfg->taps[i];
    cudaError_t e = cudaSetDevice(cfg->device);
    if (e == cudaSuccess) e = cudaMalloc(&s->d_taps, poly.size() * sizeof(float));
    if (e == cudaSuccess) e = cudaMemcpy(s->d_taps, poly.data(), poly.size() * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMalloc(&s->d_sync, std::max<size_t>(cfg->n_syncword, 1) * sizeof(float));
    if (e == cudaSuccess && cfg->n_syncword)
        e = cudaMemcpy(s->d_sync, cfg->syncword_symbols, cfg->n_syncword * sizeof(float), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        b200sync_stim_destroy(s);
        return stim_fail(B200SYNC_ECUDA, std::string("no usable CUDA device: ") + cudaGetErrorString(e));
    }
    StimParams& P = s->P;
    P.interp = static_cast<int>(cfg->interpolation);
    P.arm_len = static_cast<int>(arm_len);
    P.n_sync = static_cast<int>(cfg->n_syncword);
    P.n_hdr_payload = static_cast<int>(cfg->header_symbols + cfg->payload_symbols);
    P.frame_len = static_cast<int>(frame);
    P.theta = static_cast<double>(cfg->phase_incr);
    P.rotate = cfg->phase_incr != 0.0f;
    P.sigma = cfg->noise_amplitude / 1.41421356237309504880f;  // _amplitude_complex (PM/noise_source.hpp:47)
    P.noise = cfg->noise_amplitude != 0.0f;
    const unsigned int sd = static_cast<unsigned int>(cfg->seed) ^ static_cast<unsigned int>(cfg->seed >> 32);
    P.sym_salt = sd * 7919u + 17u;
    P.n1_salt = sd * 104729u + 1u;
    P.n2_salt = sd * 1299709u + 2u;
    s->smem = (poly.size() + cfg->n_syncword) * sizeof(float);
    if (s->smem > 48 * 1024) {
        b200sync_stim_destroy(s);
        return stim_fail(B200SYNC_EUNSUPPORTED, "taps + syncword exceed 48 KiB of shared memory");
    }
    *out = s;
    return 0;
}

void b200sync_stim_destroy(b200sync_stim* s) {
    if (!s) return;
    cudaSetDevice(s->device);
    if (s->d_taps) cudaFree(s->d_taps);
    if (s->d_sync) cudaFree(s->d_sync);
    delete s;
}

int b200sync_stim_generate_device(b200sync_stim* s, uint64_t first_sample, size_t n, void* d_out, void* cuda_stream) {
    if (!s || (!d_out && n)) return stim_fail(B200SYNC_EINVAL, "null argument");
    if (n == 0) return 0;
    SCU(cudaSetDevice(s->device));
    StimParams P = s->P;
    P.out = static_cast<float2*>(d_out);
    P.first = static_cast<long long>(first_sample);
    P.n = static_cast<long long>(n);
    const long long k0 = P.first / P.interp, k1 = (P.first + P.n - 1) / P.interp;
    const long long periods = k1 - k0 + 1;
    const unsigned grid = static_cast<unsigned>((periods + kStimThreads - 1) / kStimThreads);
    stimulus_kernel<<<grid, kStimThreads, s->smem, static_cast<cudaStream_t>(cuda_stream)>>>(P, s->d_taps, s->d_sync);
    count_launch();
    SCU(cudaGetLastError());
    return 0;
}

}  // extern "C"
