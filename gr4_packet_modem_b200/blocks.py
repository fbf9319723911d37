"""Host-side mirror of the reference's block interface for the hot path, on top of the
C ABI (include/b200sync.h).  Same setting names, defaults, lifecycle and error behaviour as
the reference blocks so the parity tests read like the reference's own QA:

    reference (C++, GR4)                                   here
    ----------------------------------------------------   -------------------------------
    fg.emplaceBlock<SyncwordDetection>({{"rrc_taps",..}})   SyncwordDetection(rrc_taps=..)
    start()                       PM/syncword_detection.hpp:143    .start()
    processBulk(inSpan, outSpan)  PM/syncword_detection.hpp:204    .process_bulk(in_span)
    out.publishTag(map, offset)   PM/syncword_detection.hpp:321    returned tag list
    throw gr::exception(...)                                        raises B200SyncError

All arithmetic happens in libb200sync.so on the GPU; nothing here computes.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native
from ._native import B200SyncError, DetectionRecord, SdConfig, SyncwordTag, check

# numpy views of the C structs of include/b200sync.h (zero-copy, no per-record Python objects)
RECORD_DTYPE = np.dtype([("index", "<u8"), ("corr_re", "<f4"), ("corr_im", "<f4"), ("pow", "<f4"),
                         ("pow_left", "<f4"), ("pow_right", "<f4"), ("pow_prev", "<f4"), ("pow_next", "<f4"),
                         ("noise_power", "<f4"), ("freq_bin", "<i4"), ("_pad", "<i4")])
TAG_DTYPE = np.dtype([("index", "<u8"), ("syncword_freq", "<f8"), ("syncword_amplitude", "<f4"),
                      ("syncword_phase", "<f4"), ("syncword_freq_bin", "<i4"), ("syncword_noise_power", "<f4"),
                      ("syncword_esn0_db", "<f4"), ("syncword_time_est", "<f4")])
assert RECORD_DTYPE.itemsize == C.sizeof(DetectionRecord) and TAG_DTYPE.itemsize == C.sizeof(SyncwordTag)


_RAW = np.dtype((np.void, RECORD_DTYPE.itemsize))


def host_register(buf: np.ndarray) -> None:
    """Page-lock a long-lived host buffer (b200sync_host_register): spans inside it go straight to the copy engine."""
    check(_native.lib().b200sync_host_register(buf.ctypes.data, buf.nbytes))


def host_unregister(buf: np.ndarray) -> None:
    check(_native.lib().b200sync_host_unregister(buf.ctypes.data))


def _copy_records(buf: np.ndarray, n: int) -> np.ndarray:
    """Copy of the first n records as one memcpy (numpy copies structured arrays field by field)."""
    return buf.view(_RAW)[:n].copy().view(RECORD_DTYPE)


def _tag_dict(t: SyncwordTag) -> dict:
    """The property_map published by output_tag (PM/syncword_detection.hpp:106-114)."""
    return {
        "syncword_amplitude": t.syncword_amplitude,
        "syncword_phase": t.syncword_phase,
        "syncword_freq": t.syncword_freq,
        "syncword_freq_bin": t.syncword_freq_bin,
        "syncword_noise_power": t.syncword_noise_power,
        "syncword_esn0_db": t.syncword_esn0_db,
        "syncword_time_est": t.syncword_time_est,
    }


class SyncwordDetection:
    """gr::packet_modem::SyncwordDetection on the GPU (PM/syncword_detection.hpp).

    Settings are the reflected members of the reference block (:131-141, :361-372)."""

    def __init__(self, rrc_taps, syncword, constellation, min_freq_bin: int = 0, max_freq_bin: int = 0,
                 time_threshold: int = 768, power_threshold: float = 9.5, fft_size: int = 2048,
                 samples_per_symbol: int = 4, device: int = 0):
        self.fft_size = int(fft_size)
        self.samples_per_symbol = int(samples_per_symbol)
        self.rrc_taps = np.ascontiguousarray(rrc_taps, dtype=np.float32)
        self.syncword = np.ascontiguousarray(syncword, dtype=np.uint8)
        self.constellation = np.ascontiguousarray(constellation, dtype=np.complex64)
        self.min_freq_bin = int(min_freq_bin)
        self.max_freq_bin = int(max_freq_bin)
        self.time_threshold = int(time_threshold)
        self.power_threshold = float(power_threshold)
        self.device = int(device)
        self._h = C.c_void_p()
        self._items_consumed = 0
        self.start()

    # -- lifecycle ---------------------------------------------------------------------
    def start(self) -> None:
        """Validate settings, build the syncword spectra, reset streaming state
        (PM/syncword_detection.hpp:143-202).  Raises where the reference throws."""
        L = _native.lib()
        self._destroy()
        cfg = SdConfig(self.fft_size, self.samples_per_symbol, self.rrc_taps.ctypes.data, self.rrc_taps.size,
                       self.syncword.ctypes.data, self.syncword.size, self.constellation.ctypes.data,
                       self.constellation.size, self.min_freq_bin, self.max_freq_bin, self.time_threshold,
                       self.power_threshold, self.device)
        h = C.c_void_p()
        check(L.b200sync_sd_create(C.byref(cfg), C.byref(h)))
        self._h = h
        self._items_consumed = 0
        ss, st, dl, nh = C.c_uint32(), C.c_uint32(), C.c_uint64(), C.c_uint32()
        sc = C.c_float()
        check(L.b200sync_sd_info(self._h, C.byref(ss), C.byref(st), C.byref(sc), C.byref(nh), C.byref(dl)))
        self._syncword_samples_size = ss.value
        self.stride = st.value
        self._syncword_self_corr = sc.value
        self.num_hypotheses = nh.value
        self.delay = dl.value

    def _destroy(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            _native.lib().b200sync_sd_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self._destroy()
        except Exception:
            pass

    # -- processBulk -------------------------------------------------------------------
    def set_auto_register(self, on: bool = True) -> None:
        """Page-lock the pages of every pageable input span handed to process_bulk (b200sync_sd_set_auto_register)."""
        check(_native.lib().b200sync_sd_set_auto_register(self._h, int(bool(on))))

    def process_bulk(self, in_span, want_output: bool = True, max_tags: int = 4096):
        """One processBulk(inSpan, outSpan) call with host spans.

        Returns (status, consumed, out, tags): status "OK" or "INSUFFICIENT_INPUT_ITEMS"
        (:215-227); `out` = the published output items (input delayed by 2*time_threshold+1,
        :318-319); tags = [(offset_in_this_chunk, absolute_output_index, property_map)]."""
        L = _native.lib()
        x = np.ascontiguousarray(in_span, dtype=np.complex64)
        out = np.empty(x.size, np.complex64) if want_output else None
        tags = (SyncwordTag * max_tags)()
        nc, nt = C.c_size_t(0), C.c_size_t(0)
        rc = check(L.b200sync_sd_process(self._h, x.ctypes.data, x.size, out.ctypes.data if want_output else None,
                                         C.byref(nc), tags, max_tags, C.byref(nt)))
        base = self._items_consumed
        self._items_consumed += nc.value
        res = [(int(tags[i].index - base), int(tags[i].index), _tag_dict(tags[i])) for i in range(nt.value)]
        # tags beyond max_tags stay queued in the context and belong to this chunk: fetch them too
        while L.b200sync_sd_tags_ready(self._h) > 0:
            check(L.b200sync_sd_drain_tags(self._h, tags, max_tags, C.byref(nt)))
            res += [(int(tags[i].index - base), int(tags[i].index), _tag_dict(tags[i])) for i in range(nt.value)]
        status = "INSUFFICIENT_INPUT_ITEMS" if rc == 1 else "OK"
        return status, nc.value, (out[:nc.value] if want_output else None), res

    def run(self, x, chunk: int = 65536, want_output: bool = False):
        """Drive process_bulk the way the GR4 runtime does: offer a span, advance by what was
        consumed, re-offer the remainder (GR/Block.hpp:1537-1651)."""
        x = np.ascontiguousarray(x, dtype=np.complex64)
        pos, outs, tags = 0, [], []
        while x.size - pos >= self.fft_size:
            status, c, o, t = self.process_bulk(x[pos:pos + chunk], want_output)
            if c == 0:
                break
            if want_output:
                outs.append(o)
            tags.extend(t)
            pos += c
        out = np.concatenate(outs) if outs else np.zeros(0, np.complex64)
        return pos, out, tags

    # -- offline bulk entry points -----------------------------------------------------
    def _rec_buffer(self, max_recs: int) -> np.ndarray:
        if getattr(self, "_recbuf", None) is None or self._recbuf.size < max_recs:
            self._recbuf = np.empty(max_recs, RECORD_DTYPE)
        return self._recbuf

    def records_to_tags(self, recs: np.ndarray, reuse: bool = False) -> np.ndarray:
        """output_tag() on raw records (b200sync_sd_records_to_tags).  reuse: the tags land in a buffer the context
        keeps (a view, valid until the next call) instead of a fresh array."""
        recs = np.ascontiguousarray(recs, dtype=RECORD_DTYPE)
        if reuse:
            if getattr(self, "_tagbuf", None) is None or self._tagbuf.size < recs.size:
                self._tagbuf = np.empty(max(recs.size, 1024), TAG_DTYPE)
            tags = self._tagbuf[:recs.size]
        else:
            tags = np.empty(recs.size, TAG_DTYPE)
        check(_native.lib().b200sync_sd_records_to_tags(self._h, recs.ctypes.data, recs.size, tags.ctypes.data))
        return tags

    def detect_device(self, d_in_ptr: int, n: int, stream_ptr: int = 0, d_out_ptr: int = 0, max_recs: int = 0,
                      copy: bool = True):
        """Whole device-resident capture (b200sync_sd_detect_device).  d_in_ptr: device address of
        n complex64 samples.  Returns (consumed, records, tags).  copy=False: records and tags are views into
        buffers the context keeps — like the spans of a processBulk call they are valid until the next call (a
        fresh 2 MB array per call costs more page faults than the conversion itself)."""
        L = _native.lib()
        if max_recs <= 0:
            max_recs = n // (self.time_threshold + 1) + 2
        recs = self._rec_buffer(max_recs)
        nr, nc = C.c_size_t(0), C.c_size_t(0)
        check(L.b200sync_sd_detect_device(self._h, C.c_void_p(d_in_ptr), n, C.c_void_p(d_out_ptr or None),
                                          C.c_void_p(stream_ptr or None), recs.ctypes.data, max_recs, C.byref(nr),
                                          C.byref(nc)))
        r = _copy_records(recs, nr.value) if copy else recs[:nr.value]
        return nc.value, r, self.records_to_tags(r, reuse=not copy)

    def detect_host(self, x, max_recs: int = 0, out=None):
        """Whole host capture, H2D pipelined with compute (b200sync_sd_detect_host).  out: optional host
        destination of the block's output span (complex64 array or address of >= n items): the input delayed by
        2*time_threshold+1 (b200sync_sd_detect_host_out)."""
        L = _native.lib()
        if isinstance(x, np.ndarray):
            x = np.ascontiguousarray(x, dtype=np.complex64)
            ptr, n = x.ctypes.data, x.size
        else:  # (address, n) of e.g. a pinned torch tensor
            ptr, n = x
        if max_recs <= 0:
            max_recs = n // (self.time_threshold + 1) + 2
        recs = self._rec_buffer(max_recs)
        nr, nc = C.c_size_t(0), C.c_size_t(0)
        if out is None:
            check(L.b200sync_sd_detect_host(self._h, C.c_void_p(ptr), n, recs.ctypes.data, max_recs, C.byref(nr),
                                            C.byref(nc)))
        else:
            optr = out.ctypes.data if isinstance(out, np.ndarray) else int(out)
            check(L.b200sync_sd_detect_host_out(self._h, C.c_void_p(ptr), n, C.c_void_p(optr), recs.ctypes.data,
                                                max_recs, C.byref(nr), C.byref(nc)))
        r = _copy_records(recs, nr.value)
        return nc.value, r, self.records_to_tags(r)

    def shard_output(self, d_out_ptr: int, out_first_abs: int, out_len: int) -> None:
        """The next shard_phase1* call also writes its slice of the delayed output stream (b200sync_sd_shard_output)."""
        check(_native.lib().b200sync_sd_shard_output(self._h, C.c_void_p(d_out_ptr or None), int(out_first_abs),
                                                     int(out_len)))

    def detect_file(self, filename, first_item: int = 0, max_items: int | None = None, max_recs: int = 0):
        """Whole raw cf32 capture file — the format FileSource<c64> reads (PM/file_source.hpp) — staged
        through pinned buffers (b200sync_sd_detect_file).  Returns (consumed, records, tags, items_read)."""
        import os

        L = _native.lib()
        path = os.fsencode(filename)
        limit = (1 << 64) - 1 if max_items is None else int(max_items)
        if max_recs <= 0:
            try:
                items = max(0, os.path.getsize(filename) // 8 - first_item)
            except OSError:
                items = 0
            max_recs = min(items, limit) // (self.time_threshold + 1) + 2
        recs = self._rec_buffer(max_recs)
        nr, nc, ni = C.c_size_t(0), C.c_size_t(0), C.c_uint64(0)
        check(L.b200sync_sd_detect_file(self._h, path, int(first_item), limit, recs.ctypes.data, max_recs,
                                        C.byref(nr), C.byref(nc), C.byref(ni)))
        r = _copy_records(recs, nr.value)
        return nc.value, r, self.records_to_tags(r), ni.value

    def detect_channels_device(self, d_in_ptr: int, n_channels: int, n: int, channel_stride: int = 0,
                               stream_ptr: int = 0, max_recs: int = 0):
        """Batched channel mode (b200sync_sd_detect_channels_device): n_channels independent streams of n
        samples, channel c at d_in_ptr + 8 * c * channel_stride.  Returns (consumed_per_channel,
        [records of channel 0, records of channel 1, ...])."""
        L = _native.lib()
        stride = channel_stride or n
        if max_recs <= 0:
            max_recs = n // (self.time_threshold + 1) + 2
        recs = self._rec_buffer(max_recs * n_channels)
        counts = np.zeros(n_channels, np.uintp)
        nc = C.c_size_t(0)
        check(L.b200sync_sd_detect_channels_device(self._h, C.c_void_p(d_in_ptr), stride, n_channels, n,
                                                   C.c_void_p(stream_ptr or None), recs.ctypes.data, max_recs,
                                                   counts.ctypes.data, C.byref(nc)))
        raw = recs.view(_RAW)
        out = [raw[c * max_recs:c * max_recs + int(counts[c])].copy().view(RECORD_DTYPE) for c in range(n_channels)]
        return nc.value, out

    def shard_output_host(self, out_ptr: int, out_first_abs: int, out_len: int) -> None:
        """... into host memory, for shard_phase1_host (b200sync_sd_shard_output_host)."""
        check(_native.lib().b200sync_sd_shard_output_host(self._h, C.c_void_p(out_ptr or None), int(out_first_abs),
                                                          int(out_len)))

    def shard_phase1(self, d_in_ptr: int, first_sample_abs: int, n_in: int, first_block: int, n_blocks: int,
                     total_blocks: int, stream_ptr: int = 0) -> np.ndarray:
        L = _native.lib()
        table = np.zeros(self.time_threshold + 1, np.uint16)
        check(L.b200sync_sd_shard_phase1(self._h, C.c_void_p(d_in_ptr), first_sample_abs, n_in, first_block,
                                         n_blocks, total_blocks, C.c_void_p(stream_ptr or None), table.ctypes.data,
                                         table.size))
        return table

    def shard_phase1_host(self, x, first_sample_abs: int, first_block: int, n_blocks: int,
                          total_blocks: int) -> np.ndarray:
        """Phase 1 with the shard's samples in host memory (b200sync_sd_shard_phase1_host); x is a complex64
        array or (address, n) of e.g. a pinned torch tensor."""
        L = _native.lib()
        if isinstance(x, np.ndarray):
            x = np.ascontiguousarray(x, dtype=np.complex64)
            ptr, n = x.ctypes.data, x.size
        else:
            ptr, n = x
        table = np.zeros(self.time_threshold + 1, np.uint16)
        check(L.b200sync_sd_shard_phase1_host(self._h, C.c_void_p(ptr), first_sample_abs, n, first_block, n_blocks,
                                              total_blocks, table.ctypes.data, table.size))
        return table

    def shard_phase2(self, entry_offset: int, max_recs: int, copy: bool = True):
        """copy=False: records and tags are views into the context's buffers, valid until the next call (detect_device)."""
        L = _native.lib()
        recs = self._rec_buffer(max(max_recs, 1))
        nr = C.c_size_t(0)
        check(L.b200sync_sd_shard_phase2(self._h, entry_offset, recs.ctypes.data, max_recs, C.byref(nr)))
        r = _copy_records(recs, nr.value) if copy else recs[:nr.value]
        return r, self.records_to_tags(r, reuse=not copy)

    def last_timings(self) -> dict:
        """Device milliseconds of the stages of the last offline call (CUDA events)."""
        a, b, c = C.c_float(), C.c_float(), C.c_float()
        check(_native.lib().b200sync_sd_last_timings(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return {"correlate_ms": a.value, "peaks_ms": b.value, "refine_ms": c.value}

    def metric(self, n: int) -> np.ndarray:
        """Per-sample winning correlation power of the last offline call (verification tap)."""
        z = np.zeros(n, np.float32)
        check(_native.lib().b200sync_sd_copy_metric(self._h, z.ctypes.data, n))
        return z


class SyncwordDetectionMulti:
    """ONE SyncwordDetection block over several GPUs of one box (b200sync_sd_multi_*): the capture is cut into
    contiguous time shards on the reference's own FFT-block grid, one host thread drives each GPU, the shards'
    chain tables are composed on the host and the detections are gathered on the host — the records equal those
    of the single-GPU call.  No collective, no NCCL.  devices=None: every visible GPU."""

    def __init__(self, rrc_taps, syncword, constellation, min_freq_bin: int = 0, max_freq_bin: int = 0,
                 time_threshold: int = 768, power_threshold: float = 9.5, fft_size: int = 2048,
                 samples_per_symbol: int = 4, devices=None):
        L = _native.lib()
        self.rrc_taps = np.ascontiguousarray(rrc_taps, dtype=np.float32)
        self.syncword = np.ascontiguousarray(syncword, dtype=np.uint8)
        self.constellation = np.ascontiguousarray(constellation, dtype=np.complex64)
        self.time_threshold = int(time_threshold)
        cfg = SdConfig(int(fft_size), int(samples_per_symbol), self.rrc_taps.ctypes.data, self.rrc_taps.size,
                       self.syncword.ctypes.data, self.syncword.size, self.constellation.ctypes.data,
                       self.constellation.size, int(min_freq_bin), int(max_freq_bin), self.time_threshold,
                       float(power_threshold), 0)
        devs = np.ascontiguousarray(devices if devices is not None else [], dtype=np.int32)
        h = C.c_void_p()
        check(L.b200sync_sd_multi_create(C.byref(cfg), devs.ctypes.data if devs.size else None, devs.size, C.byref(h)))
        self._h = h
        self.n_devices = int(L.b200sync_sd_multi_devices(self._h))
        self.delay = 2 * self.time_threshold + 1

    def __del__(self):
        try:
            if getattr(self, "_h", None) is not None and self._h.value:
                _native.lib().b200sync_sd_multi_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    def plan(self, n: int):
        """How an n-sample capture is cut: list of dicts (device, first_block, n_blocks, total_blocks,
        first_sample, n_samples), one per GPU."""
        sh = (_native.Shard * self.n_devices)()
        check(_native.lib().b200sync_sd_multi_plan(self._h, int(n), sh))
        return [{k: int(getattr(s, k)) for k in ("device", "first_block", "n_blocks", "total_blocks", "first_sample",
                                                  "n_samples")} for s in sh]

    def _tags(self, r):
        L = _native.lib()
        tags = np.zeros(len(r), TAG_DTYPE)
        if len(r):
            ctx = C.c_void_p(L.b200sync_sd_multi_context(self._h, 0))
            rr = np.ascontiguousarray(r)
            check(L.b200sync_sd_records_to_tags(ctx, rr.ctypes.data, len(rr), tags.ctypes.data))
        return tags

    def _finish(self, recs, nr, nc):
        r = _copy_records(recs, nr.value)
        return nc.value, r, self._tags(r)

    def detect_host(self, x, max_recs: int = 0):
        """Capture in host memory: a complex64 array, or (address, n) of e.g. a pinned torch tensor."""
        if isinstance(x, np.ndarray):
            x = np.ascontiguousarray(x, dtype=np.complex64)
            ptr, n = x.ctypes.data, x.size
        else:
            ptr, n = x
        if max_recs <= 0:
            max_recs = n // (self.time_threshold + 1) + 2 * self.n_devices + 2
        recs = np.zeros(max_recs, RECORD_DTYPE)
        nr, nc = C.c_size_t(0), C.c_size_t(0)
        check(_native.lib().b200sync_sd_multi_detect_host(self._h, C.c_void_p(ptr), n, recs.ctypes.data, max_recs,
                                                          C.byref(nr), C.byref(nc)))
        return self._finish(recs, nr, nc)

    def detect_device(self, shard_ptrs, n: int, max_recs: int = 0):
        """Capture already resident: shard_ptrs[r] = device address (on plan(n)[r]["device"]) of that shard's
        samples."""
        ptrs = (C.c_void_p * self.n_devices)(*[C.c_void_p(int(p)) for p in shard_ptrs])
        if max_recs <= 0:
            max_recs = n // (self.time_threshold + 1) + 2 * self.n_devices + 2
        recs = np.zeros(max_recs, RECORD_DTYPE)
        nr, nc = C.c_size_t(0), C.c_size_t(0)
        check(_native.lib().b200sync_sd_multi_detect_device(self._h, ptrs, int(n), recs.ctypes.data, max_recs,
                                                            C.byref(nr), C.byref(nc)))
        return self._finish(recs, nr, nc)

    def detect_file(self, filename, first_item: int = 0, max_items: int | None = None, max_recs: int = 0):
        """Raw cf32 capture file (PM/file_source.hpp format); every GPU's host thread reads its own shard."""
        import os

        path = os.fsencode(filename)
        limit = (1 << 64) - 1 if max_items is None else int(max_items)
        if max_recs <= 0:
            try:
                items = max(0, os.path.getsize(filename) // 8 - first_item)
            except OSError:
                items = 0
            max_recs = min(items, limit) // (self.time_threshold + 1) + 2 * self.n_devices + 2
        recs = np.zeros(max_recs, RECORD_DTYPE)
        nr, nc, ni = C.c_size_t(0), C.c_size_t(0), C.c_uint64(0)
        check(_native.lib().b200sync_sd_multi_detect_file(self._h, path, int(first_item), limit, recs.ctypes.data,
                                                          max_recs, C.byref(nr), C.byref(nc), C.byref(ni)))
        return self._finish(recs, nr, nc) + (ni.value,)

    def last_timings(self) -> list:
        a = (C.c_float * self.n_devices)()
        b = (C.c_float * self.n_devices)()
        c = (C.c_float * self.n_devices)()
        check(_native.lib().b200sync_sd_multi_last_timings(self._h, a, b, c))
        return [{"correlate_ms": a[i], "peaks_ms": b[i], "refine_ms": c[i]} for i in range(self.n_devices)]


class FrontEnd:
    """PfbArbResampler<c64,c64,float,float> followed by Rotator<float>, fused in one kernel
    (PM/pfb_arb_resampler.hpp, PM/rotator.hpp; wiring apps/packet_transceiver.cpp:71-75).
    Settings carry the reference names: rate, taps, filter_size, phase_incr."""

    def __init__(self, rate: float = 1.0, taps=None, filter_size: int = 32, phase_incr: float = 0.0,
                 enable_resampler: bool = True, enable_rotator: bool = True, device: int = 0,
                 rate_dtype=np.float32, fp_contract: bool = False):
        """rate_dtype: np.float32 = PfbArbResampler<.., TRate = float> (the default template argument),
        np.float64 = TRate = double (what test/qa_pfb_arb_resampler.cpp instantiates).
        fp_contract: fused multiply-add per filter tap — 2x faster, one rounding per tap instead of the reference's
        two (relative L2 difference ~1e-7); off by default, where the resampled stream is bit-exact."""
        self.fp_contract = bool(fp_contract)
        self.rate_is_f64 = np.dtype(rate_dtype) == np.float64
        self.rate = float(rate) if self.rate_is_f64 else float(np.float32(rate))
        self.taps = np.ascontiguousarray(taps if taps is not None else np.zeros(0), dtype=np.float32)
        self.filter_size = int(filter_size)
        self.phase_incr = float(np.float32(phase_incr))
        self.enable_resampler, self.enable_rotator = bool(enable_resampler), bool(enable_rotator)
        self.device = int(device)
        self._h = C.c_void_p()
        self.start()

    def start(self) -> None:
        """settingsChanged() + start() of both blocks; raises where the reference throws
        ("filter_size cannot be 0", PM/pfb_arb_resampler.hpp:70-72)."""
        from ._native import FeConfig, check_fe

        L = _native.lib()
        self._destroy()
        cfg = FeConfig(self.rate, self.taps.ctypes.data if self.taps.size else None, self.taps.size,
                       self.filter_size, self.phase_incr, int(self.enable_resampler), int(self.enable_rotator),
                       self.device, int(self.rate_is_f64), float(self.rate), int(self.fp_contract), 0)
        h = C.c_void_p()
        check_fe(L.b200sync_fe_create(C.byref(cfg), C.byref(h)))
        self._h = h

    def _destroy(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            _native.lib().b200sync_fe_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self._destroy()
        except Exception:
            pass

    def restart(self) -> None:
        """start() on the live context: streaming state reset, device allocations kept (b200sync_fe_start)."""
        from ._native import check_fe

        check_fe(_native.lib().b200sync_fe_start(self._h))

    def max_output(self, n_in: int) -> int:
        return int(_native.lib().b200sync_fe_max_output(self._h, n_in))

    def process_bulk(self, in_span, max_out: int | None = None):
        """processBulk(inSpan, outSpan) with host spans -> (consumed, produced items)."""
        from ._native import check_fe

        x = np.ascontiguousarray(in_span, dtype=np.complex64)
        if max_out is None:
            max_out = self.max_output(x.size)
        out = np.empty(max_out, np.complex64)
        nc, npd = C.c_size_t(0), C.c_size_t(0)
        check_fe(_native.lib().b200sync_fe_process(self._h, x.ctypes.data, x.size, out.ctypes.data, max_out,
                                                   C.byref(nc), C.byref(npd)))
        return nc.value, out[:npd.value]

    def process_device(self, d_in_ptr: int, n_in: int, d_out_ptr: int, max_out: int, stream_ptr: int = 0):
        from ._native import check_fe

        nc, npd = C.c_size_t(0), C.c_size_t(0)
        check_fe(_native.lib().b200sync_fe_process_device(self._h, C.c_void_p(d_in_ptr), n_in, C.c_void_p(d_out_ptr),
                                                          max_out, C.c_void_p(stream_ptr or None), C.byref(nc),
                                                          C.byref(npd)))
        return nc.value, npd.value


STREAM_TAG_DTYPE = np.dtype([("index", "<u8"), ("has_syncword", "<u4"), ("other", "<u4"), ("sw", TAG_DTYPE)])
assert STREAM_TAG_DTYPE.itemsize == C.sizeof(_native.StreamTag)


def stream_tags_from_detection(tags: np.ndarray) -> np.ndarray:
    """SyncwordDetection output tags (absolute output indices) as a stream-tag array."""
    st = np.zeros(tags.size, STREAM_TAG_DTYPE)
    st["index"] = tags["index"]
    st["has_syncword"] = 1
    st["sw"] = tags
    return st


class SymbolFilter:
    """gr::packet_modem::SymbolFilter<c64, c64, float> on the GPU (PM/symbol_filter.hpp).
    Settings: samples_per_symbol, taps, num_arms, delay (:53-59)."""

    def __init__(self, taps, num_arms: int, samples_per_symbol: int = 4, delay: int = 0, device: int = 0,
                 fused_cfc_delay: int | None = None):
        """fused_cfc_delay: when not None, a CoarseFrequencyCorrection{delay = fused_cfc_delay} runs as the
        filter's load stage (b200sync_sf_fuse_cfc): the input span is then the CFC block's input."""
        self.taps = np.ascontiguousarray(taps, dtype=np.float32)
        self.num_arms, self.samples_per_symbol, self.delay = int(num_arms), int(samples_per_symbol), int(delay)
        self.device = int(device)
        self.fused_cfc_delay = fused_cfc_delay
        self._h = C.c_void_p()
        self.start()

    def start(self) -> None:
        from ._native import SfConfig, check_sf

        self._destroy()
        cfg = SfConfig(self.samples_per_symbol, self.taps.ctypes.data if self.taps.size else None, self.taps.size,
                       self.num_arms, self.delay, self.device)
        h = C.c_void_p()
        check_sf(_native.lib().b200sync_sf_create(C.byref(cfg), C.byref(h)))
        self._h = h
        if self.fused_cfc_delay is not None:
            check_sf(_native.lib().b200sync_sf_fuse_cfc(self._h, 1, int(self.fused_cfc_delay)))

    def _destroy(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            _native.lib().b200sync_sf_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self._destroy()
        except Exception:
            pass

    def process_bulk(self, in_span, in_tags: np.ndarray | None = None):
        """processBulk over a host span carrying `in_tags` (STREAM_TAG_DTYPE, indices relative to the
        span).  Returns (consumed, symbols, out_tags) with out_tags indices relative to the symbols."""
        from ._native import check_sf

        x = np.ascontiguousarray(in_span, dtype=np.complex64)
        it = np.ascontiguousarray(in_tags if in_tags is not None else np.zeros(0, STREAM_TAG_DTYPE), STREAM_TAG_DTYPE)
        max_out = x.size // max(self.samples_per_symbol, 1) + it.size + 2
        out = np.empty(max_out, np.complex64)
        ot = np.zeros(it.size + 64, STREAM_TAG_DTYPE)
        nc, npd, nt = C.c_size_t(0), C.c_size_t(0), C.c_size_t(0)
        check_sf(_native.lib().b200sync_sf_process(self._h, x.ctypes.data, x.size, it.ctypes.data if it.size else None,
                                                   it.size, out.ctypes.data, max_out, C.byref(nc), C.byref(npd),
                                                   ot.ctypes.data, ot.size, C.byref(nt)))
        return nc.value, out[:npd.value], ot[:nt.value].copy()


    def restart(self) -> None:
        """start() on the live context: state reset, device allocations kept (b200sync_sf_start)."""
        from ._native import check_sf

        check_sf(_native.lib().b200sync_sf_start(self._h))

    def process_device(self, d_in_ptr: int, n_in: int, d_out_ptr: int, max_out: int, in_tags: np.ndarray | None = None,
                       stream_ptr: int = 0):
        """processBulk with device spans (b200sync_sf_process_device) -> (consumed, produced, out_tags)."""
        from ._native import check_sf

        it = np.ascontiguousarray(in_tags if in_tags is not None else np.zeros(0, STREAM_TAG_DTYPE), STREAM_TAG_DTYPE)
        ot = np.zeros(it.size + 64, STREAM_TAG_DTYPE)
        nc, npd, nt = C.c_size_t(0), C.c_size_t(0), C.c_size_t(0)
        check_sf(_native.lib().b200sync_sf_process_device(self._h, C.c_void_p(d_in_ptr), n_in,
                                                          it.ctypes.data if it.size else None, it.size,
                                                          C.c_void_p(d_out_ptr), max_out, C.c_void_p(stream_ptr or None),
                                                          C.byref(nc), C.byref(npd), ot.ctypes.data, ot.size,
                                                          C.byref(nt)))
        return nc.value, npd.value, ot[:nt.value]


class CoarseFrequencyCorrection:
    """gr::packet_modem::CoarseFrequencyCorrection<float> on the GPU (PM/coarse_frequency_correction.hpp).
    Setting: delay (:44).  Tags are forwarded unchanged (default tag policy), so only samples come back."""

    def __init__(self, delay: int = 0, device: int = 0):
        from ._native import check_cfc

        self.delay, self.device = int(delay), int(device)
        self._h = C.c_void_p()
        h = C.c_void_p()
        check_cfc(_native.lib().b200sync_cfc_create(self.delay, self.device, C.byref(h)))
        self._h = h

    def __del__(self):
        try:
            if getattr(self, "_h", None) is not None and self._h.value:
                _native.lib().b200sync_cfc_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    def start(self) -> None:
        from ._native import check_cfc

        check_cfc(_native.lib().b200sync_cfc_start(self._h))

    restart = start

    def process_bulk(self, in_span, in_tags: np.ndarray | None = None) -> np.ndarray:
        """processBulk over a host span carrying `in_tags` (STREAM_TAG_DTYPE, indices relative to the span)."""
        from ._native import check_cfc

        x = np.ascontiguousarray(in_span, dtype=np.complex64)
        it = np.ascontiguousarray(in_tags if in_tags is not None else np.zeros(0, STREAM_TAG_DTYPE), STREAM_TAG_DTYPE)
        out = np.empty_like(x)
        check_cfc(_native.lib().b200sync_cfc_process(self._h, x.ctypes.data if x.size else None, x.size,
                                                     it.ctypes.data if it.size else None, it.size,
                                                     out.ctypes.data if x.size else None))
        return out

    def process_device(self, d_in_ptr: int, n: int, d_out_ptr: int, in_tags: np.ndarray | None = None,
                       stream_ptr: int = 0) -> None:
        from ._native import check_cfc

        it = np.ascontiguousarray(in_tags if in_tags is not None else np.zeros(0, STREAM_TAG_DTYPE), STREAM_TAG_DTYPE)
        check_cfc(_native.lib().b200sync_cfc_process_device(self._h, C.c_void_p(d_in_ptr), n,
                                                            it.ctypes.data if it.size else None, it.size,
                                                            C.c_void_p(d_out_ptr), C.c_void_p(stream_ptr or None)))


def _tags_arg(in_tags):
    it = np.ascontiguousarray(in_tags if in_tags is not None else np.zeros(0, STREAM_TAG_DTYPE), STREAM_TAG_DTYPE)
    return it, (it.ctypes.data if it.size else None), it.size


class SyncwordWipeoff:
    """gr::packet_modem::SyncwordWipeoff<c64, float> on the GPU (PM/syncword_wipeoff.hpp).  Setting: syncword
    (:36).  Tags are forwarded unchanged (default tag policy), so only items come back."""

    def __init__(self, syncword, device: int = 0):
        from ._native import check_cl

        self.syncword = np.ascontiguousarray(syncword, dtype=np.float32)
        h = C.c_void_p()
        self._h = C.c_void_p()
        check_cl(_native.lib().b200sync_wo_create(self.syncword.ctypes.data if self.syncword.size else None,
                                                  self.syncword.size, int(device), C.byref(h)))
        self._h = h

    def __del__(self):
        try:
            if getattr(self, "_h", None) is not None and self._h.value:
                _native.lib().b200sync_wo_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    def start(self) -> None:
        from ._native import check_cl

        check_cl(_native.lib().b200sync_wo_start(self._h))

    def process_bulk(self, in_span, in_tags: np.ndarray | None = None) -> np.ndarray:
        """processBulk over a host span carrying `in_tags` (STREAM_TAG_DTYPE, indices relative to the span)."""
        from ._native import check_cl

        x = np.ascontiguousarray(in_span, dtype=np.complex64)
        _it, tp, nt = _tags_arg(in_tags)
        out = np.empty_like(x)
        check_cl(_native.lib().b200sync_wo_process(self._h, x.ctypes.data if x.size else None, x.size, tp, nt,
                                                   out.ctypes.data if x.size else None))
        return out

    def process_device(self, d_in_ptr: int, n: int, d_out_ptr: int, in_tags: np.ndarray | None = None,
                       stream_ptr: int = 0) -> None:
        from ._native import check_cl

        _it, tp, nt = _tags_arg(in_tags)
        check_cl(_native.lib().b200sync_wo_process_device(self._h, C.c_void_p(d_in_ptr), n, tp, nt,
                                                          C.c_void_p(d_out_ptr), C.c_void_p(stream_ptr or None)))


class CostasLoop:
    """gr::packet_modem::CostasLoop<float, float> on the GPU (PM/costas_loop.hpp).  Settings: loop_bandwidth
    (:52), constellation "PILOT" | "BPSK" | "QPSK", case-insensitive (:53-54, 63-65).  Tags are forwarded
    unchanged (default tag policy), so only items come back."""

    CONSTELLATIONS = {"PILOT": 0, "BPSK": 1, "QPSK": 2}

    def __init__(self, loop_bandwidth: float = 0.01, constellation: str = "BPSK", device: int = 0):
        from ._native import ClConfig, B200SyncError, check_cl

        key = str(constellation).upper()
        if key not in self.CONSTELLATIONS:  # magic_enum::enum_cast(...).value() throws in the reference
            raise B200SyncError(f"unknown constellation {constellation!r}")
        self.loop_bandwidth, self.constellation = float(loop_bandwidth), key
        cfg = ClConfig(self.loop_bandwidth, self.CONSTELLATIONS[key], int(device))
        h = C.c_void_p()
        self._h = C.c_void_p()
        check_cl(_native.lib().b200sync_cl_create(C.byref(cfg), C.byref(h)))
        self._h = h

    def __del__(self):
        try:
            if getattr(self, "_h", None) is not None and self._h.value:
                _native.lib().b200sync_cl_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    def start(self) -> None:
        from ._native import check_cl

        check_cl(_native.lib().b200sync_cl_start(self._h))

    def fuse_wipeoff(self, syncword) -> None:
        """Put a SyncwordWipeoff{syncword} in front of the loop, inside the same kernel."""
        from ._native import check_cl

        sw = np.ascontiguousarray(syncword if syncword is not None else [], dtype=np.float32)
        check_cl(_native.lib().b200sync_cl_fuse_wipeoff(self._h, sw.ctypes.data if sw.size else None, sw.size))

    @property
    def coefficients(self):
        """(_k1, _k2)"""
        from ._native import check_cl

        k1, k2 = C.c_float(), C.c_float()
        check_cl(_native.lib().b200sync_cl_info(self._h, C.byref(k1), C.byref(k2)))
        return k1.value, k2.value

    @property
    def state(self):
        """(_phase, _freq) after the last processed item"""
        from ._native import check_cl

        ph, fr = C.c_float(), C.c_float()
        check_cl(_native.lib().b200sync_cl_state(self._h, C.byref(ph), C.byref(fr)))
        return ph.value, fr.value

    def process_bulk(self, in_span, in_tags: np.ndarray | None = None) -> np.ndarray:
        """processBulk over a host span carrying `in_tags` (STREAM_TAG_DTYPE, indices relative to the span)."""
        from ._native import check_cl

        x = np.ascontiguousarray(in_span, dtype=np.complex64)
        _it, tp, nt = _tags_arg(in_tags)
        out = np.empty_like(x)
        check_cl(_native.lib().b200sync_cl_process(self._h, x.ctypes.data if x.size else None, x.size, tp, nt,
                                                   out.ctypes.data if x.size else None))
        return out

    def process_device(self, d_in_ptr: int, n: int, d_out_ptr: int, in_tags: np.ndarray | None = None,
                       stream_ptr: int = 0) -> None:
        from ._native import check_cl

        _it, tp, nt = _tags_arg(in_tags)
        check_cl(_native.lib().b200sync_cl_process_device(self._h, C.c_void_p(d_in_ptr), n, tp, nt,
                                                          C.c_void_p(d_out_ptr), C.c_void_p(stream_ptr or None)))


class SyncwordDetectionFilter:
    """gr::packet_modem::SyncwordDetectionFilter (PM/syncword_detection_filter.hpp): host control logic
    behind the C ABI."""

    def __init__(self, samples_per_symbol: int = 4, syncword_size: int = 64, header_size: int = 128):
        h = C.c_void_p()
        check(_native.lib().b200sync_sdf_create(samples_per_symbol, syncword_size, header_size, C.byref(h)))
        self._h = h
        check(_native.lib().b200sync_sdf_start(self._h))

    def __del__(self):
        try:
            if self._h.value:
                _native.lib().b200sync_sdf_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    def process_bulk(self, in_span, n_out: int | None = None, tag: np.void | None = None, header=None,
                     n_ignored: int = 0):
        """header: None | ("parsed", packet_length) | ("invalid",).  Returns
        (consumed, out, forwarded_tag_or_None, header_consumed, ignored_consumed, in_packet)."""
        from ._native import SdfHeader, StreamTag, check_sf

        x = np.ascontiguousarray(in_span, dtype=np.complex64)
        n_out = x.size if n_out is None else n_out
        out = np.zeros(n_out, np.complex64)
        hdr = None
        if header is not None:
            hdr = SdfHeader(1, 0) if header[0] == "invalid" else SdfHeader(0, int(header[1]))
        tin = None
        if tag is not None:
            tin = StreamTag.from_buffer_copy(np.asarray(tag, STREAM_TAG_DTYPE).tobytes())
        tout = StreamTag()
        nc, hu, iu = C.c_size_t(0), C.c_size_t(0), C.c_size_t(0)
        fwd, inpkt = C.c_int(0), C.c_int(0)
        check_sf(_native.lib().b200sync_sdf_process(self._h, x.ctypes.data, x.size, out.ctypes.data, n_out,
                                                    C.byref(tin) if tin is not None else None,
                                                    C.byref(hdr) if hdr is not None else None, n_ignored,
                                                    C.byref(nc), C.byref(hu), C.byref(iu), C.byref(tout),
                                                    C.byref(fwd), C.byref(inpkt)))
        t = np.frombuffer(bytes(tout), STREAM_TAG_DTYPE)[0] if fwd.value else None
        return nc.value, out[:nc.value], t, hu.value, iu.value, bool(inpkt.value)


class PfbArbResampler(FrontEnd):
    """gr::packet_modem::PfbArbResampler<c64, c64, float, TRate> alone (PM/pfb_arb_resampler.hpp); TRate = float
    unless rate_dtype=np.float64."""

    def __init__(self, rate: float, taps, filter_size: int = 32, device: int = 0, rate_dtype=np.float32,
                 fp_contract: bool = False):
        super().__init__(rate=rate, taps=taps, filter_size=filter_size, enable_rotator=False, device=device,
                         rate_dtype=rate_dtype, fp_contract=fp_contract)


class Rotator(FrontEnd):
    """gr::packet_modem::Rotator<float> alone (PM/rotator.hpp)."""

    def __init__(self, phase_incr: float, device: int = 0):
        super().__init__(phase_incr=phase_incr, enable_resampler=False, device=device)


__all__ = ["SyncwordDetection", "DetectionRecord", "SyncwordTag", "B200SyncError", "RECORD_DTYPE", "TAG_DTYPE",
           "FrontEnd", "PfbArbResampler", "Rotator", "SymbolFilter", "SyncwordDetectionFilter",
           "STREAM_TAG_DTYPE", "stream_tags_from_detection"]
