/* b200sync.h — C ABI of libb200sync.so: the B200 (sm_100a) receiver-synchronisation
 * hot path of gr4-packet-modem.
 *
 * The reference has no C ABI for DSP: its blocks are header-only C++ classes called
 * by the GNU Radio 4.0 runtime (gr::Block<T>::workInternal -> T::processBulk).  The
 * entry points below are what a cgo/JNI/ctypes/C++ binding of that block contract
 * needs; each one names the reference interface it stands in for.  Paths:
 *   PM/ = blocks/include/gnuradio-4.0/packet-modem/   (reference repository)
 *
 * Conventions
 *   - complex samples are interleaved float32 (re, im), i.e. std::complex<float>;
 *   - every function returns 0 on success or a negative B200SYNC_E* code; the
 *     message of the last failure on the calling thread is b200sync_last_error();
 *   - a context is NOT thread-safe (like a GR4 block instance: one worker thread per
 *     block, GR/Scheduler.hpp:387-398); distinct contexts may be used concurrently;
 *   - there is no CPU fallback: if no sm_100-class device is usable, create fails.
 */
#ifndef B200SYNC_H
#define B200SYNC_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200SYNC_OK 0
#define B200SYNC_EINVAL (-1)      /* bad argument / setting (gr::exception in the reference) */
#define B200SYNC_ECUDA (-2)       /* CUDA runtime failure                                     */
#define B200SYNC_ENOMEM (-3)      /* caller buffer too small / allocation failed             */
#define B200SYNC_EUNSUPPORTED (-4)/* valid in the reference, not implemented on the GPU path */
#define B200SYNC_ENOSPC (-5)      /* output span too small for the items the call would produce */

const char* b200sync_last_error(void);
/* ABI version of this header (bumped on incompatible change). */
int b200sync_abi_version(void);
/* Number of kernel launches issued by this process so far (bench.py's gpu_launches). */
uint64_t b200sync_launch_count(void);
/* Page-lock a long-lived host buffer (GR4's port ring buffers, GR/CircularBuffer.hpp) so that the streaming calls
 * (b200sync_*_process with host spans) hand its spans straight to the copy engine instead of staging them through
 * the context's own pinned buffer: a 65536-item span then costs 79 instead of 112 us (profiles/r2_stream_throughput.txt).
 * Thin wrappers over cudaHostRegister / cudaHostUnregister so that the host program need not link the CUDA runtime;
 * unregister before the buffer is freed.  Returns 0, B200SYNC_EINVAL or B200SYNC_ECUDA. */
int b200sync_host_register(const void* ptr, size_t bytes);
int b200sync_host_unregister(const void* ptr);
/* The same without knowing the rings: when on (or with B200SYNC_AUTO_REGISTER=1 in the environment), a SyncwordDetection
 * context page-locks the pages of every pageable input span it is handed by b200sync_sd_process — a flowgraph's ring is
 * covered after one pass over it — and releases them in b200sync_sd_destroy.  Off by default: the buffers MUST outlive the
 * context (GR4's do: an edge's buffer lives as long as the graph; a buffer freed and re-allocated at the same address would
 * still be mapped to its old pages for the copy engine), and page-locked memory counts against RLIMIT_MEMLOCK. */
int b200sync_sd_set_auto_register(struct b200sync_sd* sd, int on);

/* ------------------------------------------------------------------------------
 * SyncwordDetection                                   PM/syncword_detection.hpp
 * ------------------------------------------------------------------------------ */

/* The reflected settings of the block, PM/syncword_detection.hpp:131-141, 361-372. */
typedef struct b200sync_sd_config {
    uint32_t fft_size;            /* default 2048; any power of two in [64, 8192] (2048: the hand-scheduled kernel;
                                   * others: the generic radix-2 path, correlator_generic.cu)          */
    uint32_t samples_per_symbol;  /* default 4                                          */
    const float* rrc_taps;        /* rrc_taps                                           */
    uint32_t n_rrc_taps;
    const uint8_t* syncword;      /* syncword (symbol indices into constellation)       */
    uint32_t n_syncword;
    const float* constellation;   /* constellation, interleaved (re, im)                */
    uint32_t n_constellation;
    int32_t min_freq_bin;         /* default 0                                          */
    int32_t max_freq_bin;         /* default 0                                          */
    uint64_t time_threshold;      /* default 768                                        */
    float power_threshold;        /* default 9.5                                        */
    int32_t device;               /* CUDA device ordinal                                */
} b200sync_sd_config;

/* Raw per-detection fields = the HistoryItem of PM/syncword_detection.hpp:17-29 for the
 * detected sample plus its neighbours' powers (:321-323), computed on the GPU. */
typedef struct b200sync_detection_record {
    uint64_t index;        /* absolute INPUT sample index where the syncword starts     */
    float corr_re, corr_im;/* HistoryItem::correlation                                  */
    float pow;             /* HistoryItem::correlation_power                            */
    float pow_left;        /* correlation_power_left  (0 at the lowest bin)             */
    float pow_right;       /* correlation_power_right (0 at the highest bin)            */
    float pow_prev;        /* previous_item.correlation_power                           */
    float pow_next;        /* next_item.correlation_power                               */
    float noise_power;     /* HistoryItem::fft_noise_power                              */
    int32_t freq_bin;      /* HistoryItem::freq_bin                                     */
    int32_t _pad;
} b200sync_detection_record;

/* The tag SyncwordDetection::output_tag() publishes, PM/syncword_detection.hpp:106-114. */
typedef struct b200sync_sd_tag {
    uint64_t index;          /* absolute OUTPUT stream index (= record.index + 2*time_threshold + 1) */
    double syncword_freq;
    float syncword_amplitude;
    float syncword_phase;
    int32_t syncword_freq_bin;
    float syncword_noise_power;
    float syncword_esn0_db;
    float syncword_time_est;
} b200sync_sd_tag;

typedef struct b200sync_sd b200sync_sd;

/* emplaceBlock<SyncwordDetection>({settings}) + start(): validates the settings
 * (PM/syncword_detection.hpp:145-152), builds the modulated syncword and its K conjugated
 * spectra (:154-189, spectra by the GPU FFT) and resets the streaming state (:191-201). */
int b200sync_sd_create(const b200sync_sd_config* cfg, b200sync_sd** out);
void b200sync_sd_destroy(b200sync_sd* sd);
/* start() again: reset streaming state, keep settings. */
int b200sync_sd_start(b200sync_sd* sd);

/* Derived quantities (read-only): _syncword_samples_size (:148), stride (:236),
 * _syncword_self_corr (:161-164), number of hypotheses, output delay 2T+1. */
int b200sync_sd_info(const b200sync_sd* sd, uint32_t* syncword_samples, uint32_t* stride,
                     float* self_corr, uint32_t* num_hypotheses, uint64_t* delay);

/* processBulk(inSpan, outSpan), PM/syncword_detection.hpp:204-356, HOST spans.
 *   in, n_in      the ConsumableSpan offered by the runtime
 *   out           the PublishableSpan (same length); may be NULL to skip the delayed copy
 *   *n_consumed   items consumed == items published (a multiple of the stride; 0 and
 *                 return value 1 == INSUFFICIENT_INPUT_ITEMS when n_in < fft_size, :215-227)
 *   tags          tags to publish; tag.index - (items consumed before this call) is the
 *                 offset to pass to out.publishTag(); at most max_tags, count in *n_tags.
 * Host<->device copies happen inside this call.  Once input has been consumed the call cannot fail:
 * tags beyond max_tags stay queued inside the context, b200sync_sd_tags_ready() counts those whose output
 * index has already been published and b200sync_sd_drain_tags() hands them out (same index convention);
 * a caller publishes them in the same chunk (the shell does, PM/syncword_detection.hpp:320-325). */
int b200sync_sd_process(b200sync_sd* sd, const float* in, size_t n_in, float* out, size_t* n_consumed,
                        b200sync_sd_tag* tags, size_t max_tags, size_t* n_tags);
size_t b200sync_sd_tags_ready(const b200sync_sd* sd);
int b200sync_sd_drain_tags(b200sync_sd* sd, b200sync_sd_tag* tags, size_t max_tags, size_t* n_tags);

/* Offline bulk entry point over a DEVICE-resident capture (no reference counterpart: a
 * 65536-item GR ring chunk is far too small to fill a B200).  Equivalent to start()
 * followed by one processBulk over the whole capture: runs blocks while j + fft_size <= n.
 *   d_in          device pointer, n complex samples
 *   d_out_delayed optional device pointer (>= n items): receives out[i] = in[i - delay]
 *   cuda_stream   cudaStream_t (NULL = default stream); work is enqueued and completed
 *                 before return (records come back to host memory)
 *   recs          host array; receives the records of every tag the reference block would
 *                 have published (index + delay < items consumed), sorted by index. */
int b200sync_sd_detect_device(b200sync_sd* sd, const void* d_in, size_t n, void* d_out_delayed,
                              void* cuda_stream, b200sync_detection_record* recs, size_t max_recs,
                              size_t* n_recs, size_t* n_consumed);

/* Same over a HOST capture: chunks are staged through pinned double buffers so H2D
 * copies overlap compute.  `in` may be pageable or pinned. */
int b200sync_sd_detect_host(b200sync_sd* sd, const float* in, size_t n, b200sync_detection_record* recs,
                            size_t max_recs, size_t* n_recs, size_t* n_consumed);
/* The same with the block's output span: out (host, >= items consumed) receives out[i] = in[i - delay], zeros
 * first (PM/syncword_detection.hpp:318-319).  For host spans the delay line is a host copy (the samples never
 * needed the GPU); it runs on a few host threads while the GPU works. */
int b200sync_sd_detect_host_out(b200sync_sd* sd, const float* in, size_t n, float* out,
                                b200sync_detection_record* recs, size_t max_recs, size_t* n_recs, size_t* n_consumed);

/* Raw capture ingestion: the same over a capture FILE in the format
 * FileSource<std::complex<float>> reads (PM/file_source.hpp:47-53; apps/packet_receiver_file.cpp:29-31;
 * apps/README.md:15-19): interleaved little-endian float32 I/Q, no header.  Items [first_item,
 * first_item + max_items) of the file (clipped to its length; a trailing partial item is ignored, like
 * fread) are read through pinned staging buffers so disk reads, H2D copies and kernels overlap; the
 * capture stays resident on the device.  Record indices are relative to first_item.  *n_items_read
 * (optional) receives the number of items taken from the file.  A missing / unreadable file fails with
 * the reference's message ("error opening file: ...", :33-36); a FIFO is B200SYNC_EUNSUPPORTED (stream
 * it through b200sync_sd_process instead). */
int b200sync_sd_detect_file(b200sync_sd* sd, const char* filename, uint64_t first_item, uint64_t max_items,
                            b200sync_detection_record* recs, size_t max_recs, size_t* n_recs,
                            size_t* n_consumed, uint64_t* n_items_read);

/* Batched channel mode (BASELINE config 5, SURVEY §8e "channel mode"): n_channels independent
 * streams of n samples each, channel c at d_in + c * channel_stride complex samples.  Equivalent to
 * n_channels SyncwordDetection block instances (one per channel of the reference flowgraph), each
 * running start() + one processBulk over its own stream.  Channels are pipelined over a few CUDA
 * streams that start after the work already enqueued on cuda_stream; the call returns when all
 * records are in host memory.
 *   recs      host array [n_channels * max_recs_per_channel]; channel c's records start at
 *             recs + c * max_recs_per_channel, their number is n_recs[c]
 *   n_recs    host array [n_channels];  *n_consumed: items consumed per channel. */
int b200sync_sd_detect_channels_device(b200sync_sd* sd, const void* d_in, size_t channel_stride,
                                       size_t n_channels, size_t n, void* cuda_stream,
                                       b200sync_detection_record* recs, size_t max_recs_per_channel,
                                       size_t* n_recs, size_t* n_consumed);

/* Time-sharded operation (one context per GPU).  The shard owns FFT
 * blocks [first_block, first_block + n_blocks) of a longer stream; d_in points at absolute
 * sample first_sample_abs and must cover every block the shard computes (one extra block
 * of halo on each side that exists in the stream).
 *   phase 1 computes zpow and the shard's chain table: entry search offset j (0..T) at the
 *           shard's first decided sample -> exit offset, written to table[T+1];
 *   phase 2 takes the true entry offset (composition of the previous shards' tables, done by
 *           the caller) and returns the shard's detection records. */
/* Optional, before a phase-1 call: that call also writes the shard's slice of the delayed output stream
 * (out[i] = in[i - delay], :318-319).  Shard r owns output items [first_block*S, (first_block+n_blocks)*S) of the
 * capture; d_out_delayed (device) holds absolute output items [out_first_abs, out_first_abs + out_len), items
 * outside it are not stored.  The setting is consumed by the next phase-1 call. */
int b200sync_sd_shard_output(b200sync_sd* sd, void* d_out_delayed, uint64_t out_first_abs, size_t out_len);
/* the same into HOST memory, for b200sync_sd_shard_phase1_host (a host copy on a few threads while the GPU works) */
int b200sync_sd_shard_output_host(b200sync_sd* sd, float* out_delayed, uint64_t out_first_abs, size_t out_len);
int b200sync_sd_shard_phase1(b200sync_sd* sd, const void* d_in, uint64_t first_sample_abs, size_t n_in,
                             uint64_t first_block, uint64_t n_blocks, uint64_t total_blocks,
                             void* cuda_stream, uint16_t* table, size_t table_len);
/* phase 1 with the shard's samples in HOST memory (pageable or pinned): copied in pieces on a copy stream,
 * the correlator chases the copies (as b200sync_sd_detect_host does for a whole capture). */
int b200sync_sd_shard_phase1_host(b200sync_sd* sd, const float* in, uint64_t first_sample_abs, size_t n_in,
                                  uint64_t first_block, uint64_t n_blocks, uint64_t total_blocks, uint16_t* table,
                                  size_t table_len);
/* phase 1 with the shard's samples in a capture FILE (format of FileSource<c64>, see b200sync_sd_detect_file): items
 * [capture_first_item + first_sample_abs, + n_in) are read through the pinned staging ring, copies and correlator
 * overlap the reads. */
int b200sync_sd_shard_phase1_file(b200sync_sd* sd, const char* filename, uint64_t capture_first_item,
                                  uint64_t first_sample_abs, size_t n_in, uint64_t first_block, uint64_t n_blocks,
                                  uint64_t total_blocks, uint16_t* table, size_t table_len);
int b200sync_sd_shard_phase2(b200sync_sd* sd, uint32_t entry_offset, b200sync_detection_record* recs,
                             size_t max_recs, size_t* n_recs);

/* One capture on several GPUs of one box, detections gathered on the host (BASELINE north_star).
 * Equivalent to ONE SyncwordDetection block (start() + one processBulk over the whole capture,
 * PM/syncword_detection.hpp:204-356): the records are those of the single-GPU call, in index order.  The
 * context owns one b200sync_sd per GPU; every call runs one host thread per GPU (bound to that GPU's local
 * cores when /sys tells which they are), the shards' (T+1)-entry chain tables are composed in host memory,
 * and no data moves between GPUs: there is no collective and no NCCL anywhere on this path.
 *   devices / n_devices   CUDA ordinals to use; n_devices = 0: every visible device.  cfg->device is ignored. */
typedef struct b200sync_sd_multi b200sync_sd_multi;
typedef struct b200sync_shard {
    int32_t device;         /* CUDA ordinal that owns the shard                                         */
    int32_t _pad;
    uint64_t first_block;   /* first FFT block the shard decides; blocks are on the stream's own grid   */
    uint64_t n_blocks;
    uint64_t total_blocks;  /* blocks of the whole capture, (n - fft_size) / stride + 1 (:238)          */
    uint64_t first_sample;  /* first input sample the shard needs (one halo block before first_block)   */
    uint64_t n_samples;     /* input samples the shard needs (halo on both sides included)              */
} b200sync_shard;
int b200sync_sd_multi_create(const b200sync_sd_config* cfg, const int* devices, size_t n_devices,
                             b200sync_sd_multi** out);
void b200sync_sd_multi_destroy(b200sync_sd_multi* m);
size_t b200sync_sd_multi_devices(const b200sync_sd_multi* m);
/* the per-GPU context (e.g. for b200sync_sd_records_to_tags, which needs any one of them) */
b200sync_sd* b200sync_sd_multi_context(b200sync_sd_multi* m, size_t i);
/* how an n-sample capture is cut: shards[n_devices] */
int b200sync_sd_multi_plan(const b200sync_sd_multi* m, uint64_t n, b200sync_shard* shards);
/* capture already resident: d_shards[r] is a pointer ON shards[r].device to that shard's samples
 * (absolute samples [first_sample, first_sample + n_samples) of the n-sample capture) */
int b200sync_sd_multi_detect_device(b200sync_sd_multi* m, const void* const* d_shards, uint64_t n,
                                    b200sync_detection_record* recs, size_t max_recs, size_t* n_recs,
                                    size_t* n_consumed);
/* capture in host memory (pageable or pinned): each GPU pulls its shard, compute chases the copies */
int b200sync_sd_multi_detect_host(b200sync_sd_multi* m, const float* in, uint64_t n, b200sync_detection_record* recs,
                                  size_t max_recs, size_t* n_recs, size_t* n_consumed);
/* capture file (see b200sync_sd_detect_file): each GPU's host thread reads its own shard of the file */
int b200sync_sd_multi_detect_file(b200sync_sd_multi* m, const char* filename, uint64_t first_item, uint64_t max_items,
                                  b200sync_detection_record* recs, size_t max_recs, size_t* n_recs, size_t* n_consumed,
                                  uint64_t* n_items_read);
/* per-GPU stage times of the last call: arrays of n_devices floats (any may be NULL) */
int b200sync_sd_multi_last_timings(const b200sync_sd_multi* m, float* correlate_ms, float* peaks_ms, float* refine_ms);

/* output_tag(), PM/syncword_detection.hpp:56-115, evaluated on the host in the reference's
 * own float/double mix from the raw records. */
int b200sync_sd_records_to_tags(const b200sync_sd* sd, const b200sync_detection_record* recs, size_t n,
                                b200sync_sd_tag* tags);

/* Device time (CUDA events recorded on the caller's stream) of the stages of the last
 * detect_device / detect_host call: correlator launches, peak stage, refine + record copy.
 * With detect_host the correlator figure includes waiting for the H2D copies it chases. */
int b200sync_sd_last_timings(const b200sync_sd* sd, float* correlate_ms, float* peaks_ms, float* refine_ms);

/* Debug/verification tap: copy the per-sample winning correlation power of the last
 * detect_device/detect_host call (zpow[0..n)) to host memory. */
int b200sync_sd_copy_metric(const b200sync_sd* sd, float* zpow, size_t n);

/* ------------------------------------------------------------------------------
 * Front end: PfbArbResampler + Rotator fused     PM/pfb_arb_resampler.hpp, PM/rotator.hpp
 * (the SFO / CFO conditioning stage in front of SyncwordDetection,
 *  apps/packet_transceiver.cpp:71-75).  Either stage can be disabled, which gives the two
 *  reference blocks individually.
 * ------------------------------------------------------------------------------ */
typedef struct b200sync_fe_config {
    float rate;                /* PfbArbResampler::rate, TRate = float (PM/pfb_arb_resampler.hpp:63)  */
    const float* taps;         /* PfbArbResampler::taps (prototype filter)            (:64)           */
    uint32_t n_taps;
    uint32_t filter_size;      /* PfbArbResampler::filter_size, default 32            (:65)           */
    float phase_incr;          /* Rotator::phase_incr in rad/sample                   (PM/rotator.hpp:42) */
    uint32_t enable_resampler; /* 0: bypass the resampler (Rotator block alone)                       */
    uint32_t enable_rotator;   /* 0: bypass the rotator (PfbArbResampler block alone)                 */
    int32_t device;
    uint32_t rate_is_f64;      /* 1: TRate = double — the rate is rate_f64, `rate` is ignored; the timing recurrence */
    double rate_f64;           /*    then runs in double like PfbArbResampler<.., double> (test/qa_pfb_arb_resampler.cpp) */
    uint32_t fp_contract;      /* 0 (default): every tap a separately rounded multiply and add, the order of std::inner_product
                                *    (:147-160) — the resampled stream is BIT-EXACT.  1: fused multiply-add per tap (what
                                *    GCC's -ffp-contract=fast makes of the reference on an FMA machine): one rounding per tap,
                                *    relative L2 difference ~1e-7, the kernel 2x faster.                                   */
    uint32_t _reserved;
} b200sync_fe_config;

typedef struct b200sync_fe b200sync_fe;

/* settingsChanged() of both blocks + start(): polyphase split, derivative filter, decim/filt
 * rates (PM/pfb_arb_resampler.hpp:67-120), _exp_incr (PM/rotator.hpp:44-48); state reset. */
int b200sync_fe_create(const b200sync_fe_config* cfg, b200sync_fe** out);
void b200sync_fe_destroy(b200sync_fe* fe);
int b200sync_fe_start(b200sync_fe* fe);
const char* b200sync_fe_last_error(void);
/* Upper bound on the outputs n_in further inputs can produce (buffer sizing for Async ports). */
size_t b200sync_fe_max_output(const b200sync_fe* fe, size_t n_in);

/* processBulk(inSpan, outSpan) of the resampler (PM/pfb_arb_resampler.hpp:122-182) followed by
 * Rotator::processOne on every produced item; streaming state (filter history, output phase)
 * lives behind the handle.  Consumes all n_in items unless the output span fills up first.
 * Host spans: */
int b200sync_fe_process(b200sync_fe* fe, const float* in, size_t n_in, float* out, size_t max_out,
                        size_t* n_consumed, size_t* n_produced);
/* Device spans, asynchronous on cuda_stream (the counts are computed on the host up front). */
int b200sync_fe_process_device(b200sync_fe* fe, const void* d_in, size_t n_in, void* d_out, size_t max_out,
                               void* cuda_stream, size_t* n_consumed, size_t* n_produced);

/* ------------------------------------------------------------------------------
 * Stream tags crossing SyncwordDetectionFilter and SymbolFilter.  A GR4 tag is a property_map;
 * the hot path only reads the syncword_* keys (PM/syncword_detection.hpp:106-114), every other
 * key travels as an opaque id.
 * ------------------------------------------------------------------------------ */
typedef struct b200sync_stream_tag {
    uint64_t index;         /* item index relative to the span of the call it is passed to / returned from */
    uint32_t has_syncword;  /* the tag carries the syncword_* keys in `sw`                                */
    uint32_t other;         /* != 0: opaque id of non-syncword keys carried by the same tag               */
    b200sync_sd_tag sw;     /* syncword fields (sw.index unused)                                          */
} b200sync_stream_tag;

/* ------------------------------------------------------------------------------
 * SymbolFilter<c64, c64, float>                               PM/symbol_filter.hpp
 * ------------------------------------------------------------------------------ */
typedef struct b200sync_sf_config {
    uint32_t samples_per_symbol;  /* PM/symbol_filter.hpp:55 */
    const float* taps;            /* :56 prototype filter, taps.size() == num_arms * arm length */
    uint32_t n_taps;
    uint32_t num_arms;            /* :58 */
    uint32_t delay;               /* :59 */
    int32_t device;
} b200sync_sf_config;

typedef struct b200sync_sf b200sync_sf;

/* settingsChanged() + start() (PM/symbol_filter.hpp:64-110). */
int b200sync_sf_create(const b200sync_sf_config* cfg, b200sync_sf** out);
void b200sync_sf_destroy(b200sync_sf* sf);
int b200sync_sf_start(b200sync_sf* sf);
const char* b200sync_sf_last_error(void);

/* processBulk (PM/symbol_filter.hpp:112-252) over a span that may carry any number of tags
 * (sorted by index; the GR4 shell passes at most one, at index 0, because the runtime cuts chunks at
 * tags).  Consumes all n_in items; produces one item per symbol clock tick; returns the re-indexed
 * tags (delayed by `delay`, placed on the nearest output symbol, syncword_phase adjusted when
 * time_est < 0).  B200SYNC_ENOSPC if max_out is too small, B200SYNC_ENOMEM if
 * max_out_tags is (state unchanged in both cases; only the latter is worth a retry with a larger buffer).
 * A device-span call with >= 4096 tags whose output span holds n_in / sps + n_in_tags + 2 items (no tag pattern can
 * produce more) and whose tag buffer holds n_in_tags + pending tags overlaps the host replay of the tag state machine
 * with the kernel: the span is run as a few consecutive sub-spans without synchronising in between. */
int b200sync_sf_process(b200sync_sf* sf, const float* in, size_t n_in, const b200sync_stream_tag* in_tags,
                        size_t n_in_tags, float* out, size_t max_out, size_t* n_consumed, size_t* n_produced,
                        b200sync_stream_tag* out_tags, size_t max_out_tags, size_t* n_out_tags);
int b200sync_sf_process_device(b200sync_sf* sf, const void* d_in, size_t n_in, const b200sync_stream_tag* in_tags,
                               size_t n_in_tags, void* d_out, size_t max_out, void* cuda_stream, size_t* n_consumed,
                               size_t* n_produced, b200sync_stream_tag* out_tags, size_t max_out_tags,
                               size_t* n_out_tags);

/* Fuses a CoarseFrequencyCorrection block (below) into the filter's load stage: the span handed to
 * b200sync_sf_process* is then the INPUT of CoarseFrequencyCorrection{delay = cfc_delay} and the
 * output is that of the SymbolFilter behind it (PM/packet_receiver.hpp:94-115, 195-202) — one pass
 * over the samples instead of two, bit-identical to running the two contexts back to back.  Call
 * after create()/start() and before the first process call. */
int b200sync_sf_fuse_cfc(b200sync_sf* sf, int enable, uint32_t cfc_delay);

/* ------------------------------------------------------------------------------
 * CoarseFrequencyCorrection<float>            PM/coarse_frequency_correction.hpp
 * (sits between SyncwordDetectionFilter and SymbolFilter)
 * ------------------------------------------------------------------------------ */
typedef struct b200sync_cfc b200sync_cfc;
/* emplaceBlock<CoarseFrequencyCorrection<>>({{"delay", delay}}) (:44, 100-103). */
int b200sync_cfc_create(uint32_t delay, int32_t device, b200sync_cfc** out);
void b200sync_cfc_destroy(b200sync_cfc* c);
/* back to the initial state: unit rotator, no pending frequency (:40-47) */
int b200sync_cfc_start(b200sync_cfc* c);
const char* b200sync_cfc_last_error(void);
/* processBulk (:67-98) over a span carrying any number of tags (sorted by index; only tags with
 * has_syncword, i.e. a "syncword_freq" key, act: `delay` samples after such a tag the rotator is
 * reset to phase -freq*delay and frequency -freq; a newer tag arriving first replaces it).  Tags are
 * forwarded unchanged by the block's default tag policy, so none are returned.  in == out is
 * allowed for device spans. */
int b200sync_cfc_process(b200sync_cfc* c, const float* in, size_t n, const b200sync_stream_tag* in_tags,
                         size_t n_in_tags, float* out);
int b200sync_cfc_process_device(b200sync_cfc* c, const void* d_in, size_t n, const b200sync_stream_tag* in_tags,
                                size_t n_in_tags, void* d_out, void* cuda_stream);

/* ------------------------------------------------------------------------------
 * SyncwordWipeoff<c64, float>                               PM/syncword_wipeoff.hpp
 * CostasLoop<float, float>                                  PM/costas_loop.hpp
 * (the consumers of the syncword_amplitude / syncword_phase tags behind the
 *  SymbolFilter, PM/packet_receiver.hpp:117-125, 203-218.)
 * ------------------------------------------------------------------------------ */
typedef struct b200sync_wo b200sync_wo;
/* emplaceBlock<SyncwordWipeoff<>>({{"syncword", syncword}}) (:34-36). */
int b200sync_wo_create(const float* syncword, uint32_t n_syncword, int32_t device, b200sync_wo** out);
void b200sync_wo_destroy(b200sync_wo* w);
/* back to the initial state: not inside a syncword (:27-28) */
int b200sync_wo_start(b200sync_wo* w);
/* processBulk (:38-91) over a span carrying any number of tags (sorted by index): a tag with
 * has_syncword (a "syncword_amplitude" key) that arrives while no syncword is being wiped starts one:
 * the next syncword.size() items are multiplied by the syncword, everything else is copied.  The state
 * (_in_syncword, _position) carries over to the next call.  Tags are forwarded unchanged by the block's
 * default tag policy, so none are returned.  in == out is allowed for device spans. */
int b200sync_wo_process(b200sync_wo* w, const float* in, size_t n, const b200sync_stream_tag* in_tags,
                        size_t n_in_tags, float* out);
int b200sync_wo_process_device(b200sync_wo* w, const void* d_in, size_t n, const b200sync_stream_tag* in_tags,
                               size_t n_in_tags, void* d_out, void* cuda_stream);

#define B200SYNC_CONSTELLATION_PILOT 0
#define B200SYNC_CONSTELLATION_BPSK 1
#define B200SYNC_CONSTELLATION_QPSK 2
typedef struct b200sync_cl_config {
    double loop_bandwidth;   /* CostasLoop::loop_bandwidth, default 0.01 (PM/costas_loop.hpp:52) */
    uint32_t constellation;  /* CostasLoop::constellation, default BPSK  (:53-54)                */
    int32_t device;
} b200sync_cl_config;
typedef struct b200sync_cl b200sync_cl;
/* emplaceBlock<CostasLoop<>>({settings}) + settingsChanged(): loop coefficients K1, K2 from the cubic
 * in loop_bandwidth (:56-90). */
int b200sync_cl_create(const b200sync_cl_config* cfg, b200sync_cl** out);
void b200sync_cl_destroy(b200sync_cl* c);
/* back to the initial state: _phase = _freq = 0 (:22-23) */
int b200sync_cl_start(b200sync_cl* c);
const char* b200sync_cl_last_error(void);   /* shared by the b200sync_wo_* and b200sync_cl_* calls */
/* the float coefficients _k1, _k2 (:88-89) */
int b200sync_cl_info(const b200sync_cl* c, float* k1, float* k2);
/* processBulk (:94-149) over a span carrying any number of tags (sorted by index): a tag with
 * has_syncword (a "syncword_phase" key) does set_phase(sw.syncword_phase) before its item (:102-107).
 * The stretches between such tags are independent and run one GPU thread each; the loop state at the end
 * of the span stays on the device for the next call.  in == out is allowed for device spans.
 * Cost model: the run time of a call is the recurrence of its LONGEST stretch on one thread (about 0.3 us per symbol:
 * a chain of dependent sin/cos, multiply and loop-filter operations), whatever the number of stretches — the throughput
 * comes from many packets per span (2^28 symbols, one tag per 6208 symbols: 1.9 ms).  A span without any tag is a
 * single stretch: 65536 untagged symbols take about 20 ms.  In the receiver the loop only ever sees packets
 * (PayloadMetadataInsert drops what lies between them, PM/packet_receiver.hpp:203-214), each opened by a tag. */
int b200sync_cl_process(b200sync_cl* c, const float* in, size_t n, const b200sync_stream_tag* in_tags,
                        size_t n_in_tags, float* out);
int b200sync_cl_process_device(b200sync_cl* c, const void* d_in, size_t n, const b200sync_stream_tag* in_tags,
                               size_t n_in_tags, void* d_out, void* cuda_stream);
/* Fuses a SyncwordWipeoff{syncword} block into the loop's load stage: the span handed to
 * b200sync_cl_process* is then the INPUT of SyncwordWipeoff and the output that of the CostasLoop behind
 * it (PM/packet_receiver.hpp:203-218) — one
 * pass over the symbols instead of two, bit-identical to running the two contexts back to back.
 * n_syncword = 0 removes the fusion.  Call after create()/start() and before the first process call.
 * In packet_receiver.hpp a PayloadMetadataInsert sits between the two blocks; it forwards the syncword,
 * header and payload symbols with the syncword tag and drops the inter-packet remainder
 * (PM/payload_metadata_insert.hpp:18-30).  The wipe-off needs only that tag and the 64 symbols behind it,
 * which arrive unchanged, so the fused block takes the CostasLoop's place BEHIND PayloadMetadataInsert. */
int b200sync_cl_fuse_wipeoff(b200sync_cl* c, const float* syncword, uint32_t n_syncword);
/* the loop state after the last processed item (copied from the device; synchronises) */
int b200sync_cl_state(b200sync_cl* c, float* phase, float* freq);

/* ------------------------------------------------------------------------------
 * On-GPU stimulus: frames -> InterpolatingFirFilter -> Rotator -> + NoiseSource,
 * PM/interpolating_fir_filter.hpp:93-99, PM/rotator.hpp:56-65, PM/noise_source.hpp:74-78, PM/add.hpp
 * as wired in apps/packet_transceiver.cpp:60-80, 140-165 — one pass that writes a synthetic capture into
 * device memory.  Every sample is a pure function of (seed, absolute sample index), so time shards can be
 * generated independently on different GPUs.
 * ------------------------------------------------------------------------------ */
typedef struct b200sync_stim_config {
    const float* taps;              /* InterpolatingFirFilter::taps (TX RRC, PM/packet_transmitter_rrc_taps.hpp) */
    uint32_t n_taps;
    uint32_t interpolation;         /* InterpolatingFirFilter::interpolation = samples per symbol               */
    const float* syncword_symbols;  /* the syncword as real BPSK symbols (+1 / -1), sent first in every frame   */
    uint32_t n_syncword;
    uint32_t header_symbols;        /* QPSK symbols after the syncword (128)                                    */
    uint32_t payload_symbols;       /* QPSK symbols of the payload ((bytes + 4) * 4)                            */
    uint32_t gap_symbols;           /* zero symbols between frames (0: back-to-back stream mode)                */
    float phase_incr;               /* Rotator::phase_incr, rad/sample (0: no rotator)                          */
    float noise_amplitude;          /* NoiseSource::amplitude, "gaussian": E|n|^2 = amplitude^2 (0: no noise)   */
    uint64_t seed;
    int32_t device;
} b200sync_stim_config;
typedef struct b200sync_stim b200sync_stim;
int b200sync_stim_create(const b200sync_stim_config* cfg, b200sync_stim** out);
void b200sync_stim_destroy(b200sync_stim* s);
const char* b200sync_stim_last_error(void);
/* Writes samples [first_sample, first_sample + n) of the endless stream to d_out (asynchronous on cuda_stream). */
int b200sync_stim_generate_device(b200sync_stim* s, uint64_t first_sample, size_t n, void* d_out, void* cuda_stream);

/* ------------------------------------------------------------------------------
 * SyncwordDetectionFilter<c64>                    PM/syncword_detection_filter.hpp
 * Control logic only (which syncword tags survive while inside a packet) plus the pass-through
 * copy of host spans; with device-resident data pass in = out = NULL and only the counts matter.
 * ------------------------------------------------------------------------------ */
typedef struct b200sync_sdf_header {   /* one message of the `parsed_header` port (:136-154) */
    uint32_t invalid_header;           /* meta.contains("invalid_header")                    */
    uint64_t packet_length;            /* meta.at("packet_length")                           */
} b200sync_sdf_header;

typedef struct b200sync_sdf b200sync_sdf;
int b200sync_sdf_create(uint32_t samples_per_symbol, uint32_t syncword_size, uint32_t header_size,
                        b200sync_sdf** out);
void b200sync_sdf_destroy(b200sync_sdf* f);
int b200sync_sdf_start(b200sync_sdf* f);
/* One processBulk(headerSpan, ignoredSpan, inSpan, outSpan) call (:54-210).
 *   tag_in      merged input tag on the first item, or NULL
 *   header      first pending parsed_header message, or NULL; n_ignored pending ignored_syncword messages
 *   tag_out     receives the tag to publish at output offset 0 when *tag_forwarded is set (:82-107) */
int b200sync_sdf_process(b200sync_sdf* f, const float* in, size_t n_in, float* out, size_t n_out,
                         const b200sync_stream_tag* tag_in, const b200sync_sdf_header* header, size_t n_ignored,
                         size_t* n_consumed, size_t* header_consumed, size_t* ignored_consumed,
                         b200sync_stream_tag* tag_out, int* tag_forwarded, int* in_packet);

#ifdef __cplusplus
}
#endif
#endif /* B200SYNC_H */
