// b200sync_internal.h — declarations shared by the .cu translation units of libb200sync.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include <string>

#include "../../include/b200sync.h"

namespace b200sync {

constexpr int kFft = 2048;         // the fft_size correlator.cu is hand-scheduled for (reference default); every other
                                   // power of two in [64, 8192] takes correlator_generic.cu
constexpr int kGroupThreads = 128; // threads cooperating on one FFT block
#ifndef B200_CORR_THREADS
#define B200_CORR_THREADS 768
#endif
constexpr int kCorrThreads = B200_CORR_THREADS;  // 6 FFT groups per CTA, 1 persistent CTA per SM
constexpr int kGenericFlag = 1 << 30;  // or-ed into the `fft` argument of the launchers: fft_size 2048 on the generic path
constexpr int kMaxHyp = 601;       // max frequency hypotheses: min/max_freq_bin = -/+300 spans the whole band for the default
                                   // syncword (bin spacing pi / 297 rad/sample); sizes a shared-memory array of refine_kernel
constexpr int kMaxTimeThreshold = 1023;  // parallel chain kernels keep one bitmap word per lane; the time-sharded and
                                         // batched-channel entry points need them
constexpr int kMaxTimeThresholdSeq = 4095;  // beyond 1023 the peak walk runs in order (one warp): any capture, slower

using DetectionRecord = b200sync_detection_record;

// device-resident scalar state of the peak detector (a6): where the sequential
// search of PM/syncword_detection.hpp:267-298 resumes, and the detection list size
struct PeakState {
    unsigned long long r_abs;  // next search start (absolute sample index)
    unsigned int det_count;
    unsigned int _pad;
};

// Streaming: the in-order peak walk fused in front of the refine stage (correlator.cu refine_kernel).  cand / pass:
// the bitmaps of [lo, hi) written by the flags kernel; r_abs_in: where the search resumes (the host knows it from
// the previous call); CTA 0 writes the new PeakState to state_out and to header (the slot in front of the records,
// so that one D2H copy brings state and records).
struct StreamWalk {
    const uint32_t* cand = nullptr;
    const uint32_t* pass = nullptr;
    long long range = 0, lo = 0, hi = 0;
    int T = 0;
    unsigned long long r_abs_in = 0;
    PeakState* state_out = nullptr;
    PeakState* header = nullptr;
    // Completion flag for records that land in MAPPED HOST memory (no D2H copy, no stream synchronisation): every CTA
    // counts itself on *done after its record is written and fenced system-wide; the last one stores `seq` into
    // header->_pad, which the host polls.  done == nullptr: the header is written by CTA 0 as soon as it is known.
    unsigned int* done = nullptr;
    unsigned int seq = 0;
};

// Programmatic dependent launch (streaming path: correlator -> flags -> walk + refine back to back in one stream): a
// kernel launched with launch_pdl() may become resident while its predecessor still runs — its launch latency and
// its prologue (tables into shared memory) hide behind the predecessor — and calls pdl_wait() before it touches
// anything the predecessor wrote; the predecessor calls pdl_launch_dependents() once all of its CTAs are resident
// work (at its start).  Both are no-ops in a kernel launched the ordinary way.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// api.cu: every kernel launch site calls this (b200sync_launch_count of the C ABI)
void count_launch(int n = 1);
// api.cu: set the calling thread's b200sync_last_error() text and return `code`
int set_last_error(int code, const std::string& msg);

// correlator.cu
// `fft`: fft_size of the context (| kGenericFlag, below); anything but plain 2048 dispatches to correlator_generic.cu,
// where d_hperm holds natural-order conj spectra and d_tw the per-stage radix-2 factors (fft - 1 entries)
cudaError_t launch_template_spectra(const float2* d_td, float2* d_hperm, int K, const float2* d_tw,
                                    cudaStream_t st, int fft);
cudaError_t launch_correlate(const float2* d_in, long long in_base, float* d_zpow, long long z_base,
                             const float2* d_hperm, int K, int S, int fft, long long b0, long long nb,
                             const float2* d_tw, float2* d_out_delayed, long long out_base, long long out_lo,
                             long long out_hi, int delay, int num_sms, cudaStream_t st, long long nb_chan = 0,
                             long long in_chan_stride = 0, long long z_chan_stride = 0, float2* d_gm = nullptr,
                             long long gm_b0 = 0, long long gm_chan_stride = 0, const float2* in_host = nullptr,
                             long long in_host_base = 0);
// in_host (streaming, zero copy): device-visible address of the span in pinned host memory, in_host[0] = absolute sample
// in_host_base; the kernel reads the blocks from there and fills the device window d_in itself.  Only when
// correlate_takes_host_input() says so for the span.
bool correlate_takes_host_input(int K, int fft, long long nb, int num_sms);
// Group extrema of the metric (d_gm): per FFT block, group 0 = lag 0 alone, group q >= 1 = lags 32q-31 .. 32q (clipped to
// the stride S) — the 32 consecutive samples a warp of the correlator holds.  (max, min) per group, written by the
// correlator, read by the peak stage instead of the samples themselves.
// Row of a group: 8 floats = [max, -, -, -, min of samples 0..7, 8..15, 16..23, 24..31 of the group].
constexpr int kGmF2PerGroup = 4;   // float2 per group
inline int gm_groups_per_block(int S) { return 1 + (S - 1 + 31) / 32; }
// MEASURED DEAD END, off unless built with -DB200_GM_FLAGS (round 2, 2^30 samples, one B200): the peak stage drops
// from 3.56 to 2.57 ms at K = 9 (flags kernel 2.9 -> 1.9 ms), but the correlator's epilogue that produces the extrema
// (REDUX + 3 shuffles per 32 samples, 5 scattered stores per group) costs 1.07 ms of its own at K = 9 and 1.5 ms at
// K = 1, where it matters most: no net gain, so the tile-based flags kernel stays the default.  A first version with
// one minimum per group decided only half of the threshold tests from the extrema (noise maxima sit where ~50 % of
// the groups hold a sample below the threshold) and was 2x slower than the tile kernel.
#ifdef B200_GM_FLAGS
inline bool gm_supported(int S, int T) { return S >= 64 && S <= 2016 && T >= 32 && T <= 1023; }
#else
inline bool gm_supported(int, int) { return false; }
#endif
cudaError_t launch_refine(const float2* d_in, long long in_base, const float* d_zpow, long long z_base,
                          const float2* d_hperm, int K, int S, int fft, int min_freq_bin, const float2* d_tw,
                          const unsigned long long* d_det_idx, const unsigned int* d_det_count,
                          unsigned int det_cap, DetectionRecord* d_recs, int num_sms, cudaStream_t st, int nch = 1,
                          long long in_chan_stride = 0, long long z_chan_stride = 0, long long det_chan_stride = 0,
                          const StreamWalk* walk = nullptr);
// flags only (streaming): fills *walk's cand / pass / range / lo / hi / T for the fused walk of launch_refine
cudaError_t launch_peak_flags_stream(const float* d_zpow, long long z_base, long long z_end, long long lo, long long hi,
                                     int T, float power_threshold, void* d_ws, size_t ws_bytes, int num_sms,
                                     cudaStream_t st, StreamWalk* walk);

// correlator_generic.cu: any power-of-two fft_size in [64, 8192], radix-2 arithmetic (= the oracle's independent FFT).
// `fft` carries kGenericFlag when the context runs fft_size 2048 on the generic path (B200SYNC_FORCE_GENERIC, tests).
bool generic_fft_supported(unsigned fft);
cudaError_t launch_template_spectra_generic(int fft, const float2* d_td, float2* d_hc, int K, const float2* d_tw,
                                            cudaStream_t st);
cudaError_t launch_correlate_generic(int fft, const float2* d_in, long long in_base, float* d_zpow, long long z_base,
                                     const float2* d_hc, int K, int S, long long b0, long long nb, const float2* d_tw,
                                     float2* d_out_delayed, long long out_base, long long out_lo, long long out_hi,
                                     int delay, int num_sms, cudaStream_t st, long long nb_chan,
                                     long long in_chan_stride, long long z_chan_stride);
cudaError_t launch_refine_generic(int fft, const float2* d_in, long long in_base, const float* d_zpow, long long z_base,
                                  const float2* d_hc, int K, int S, int min_freq_bin, const float2* d_tw,
                                  const unsigned long long* d_det_idx, const unsigned int* d_det_count,
                                  unsigned int det_cap, DetectionRecord* d_recs, int num_sms, cudaStream_t st, int nch,
                                  long long in_chan_stride, long long z_chan_stride, long long det_chan_stride);

// peaks.cu
struct PeakWorkspace;  // opaque: bitmaps, tables, per-segment entry states
size_t peak_workspace_bytes_sms(long long max_range, int T, int num_sms);
// Decide peaks for p in [lo, hi) given zpow known on [.., hi+T] (indices < 0 read as 0).
//   phase 1: candidate / threshold bitmaps + per-segment chain tables + range table
//   phase 2: walk the chain from entry state j_in, append detections, update state
cudaError_t launch_peak_phase1(const float* d_zpow, long long z_base, long long z_end, long long lo,
                               long long hi, int T,
                               float power_threshold, void* d_ws, size_t ws_bytes,
                               uint16_t* d_range_table /*[T+1] or nullptr*/, int num_sms,
                               cudaStream_t st, int nch = 1, long long z_stride = 0, size_t ws_stride = 0,
                               const float2* d_gm = nullptr, long long gm_b0 = 0, long long gm_blocks = 0, int S = 0,
                               long long gm_chan_stride = 0);
size_t peak_plan_bytes(long long range, int T, int num_sms);  // workspace of one channel for exactly this range
cudaError_t launch_peak_stream(const float* d_zpow, long long z_base, long long z_end, long long lo, long long hi,
                               int T, float power_threshold, void* d_ws, size_t ws_bytes, PeakState* d_state,
                               unsigned long long* d_det_idx, unsigned int det_cap, int num_sms, cudaStream_t st);
cudaError_t launch_peak_phase2(long long lo, long long hi, int T, void* d_ws, size_t ws_bytes,
                               int j_in /*-1: derive from state->r_abs*/, PeakState* d_state,
                               unsigned long long* d_det_idx, unsigned int det_cap, int num_sms,
                               cudaStream_t st, int nch = 1, size_t ws_stride = 0, size_t det_stride = 0);

}  // namespace b200sync
