"""ORACLE — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes front end of oracle/liboracle.so (the CPU restatement of the reference's
receiver-synchronisation blocks, see oracle/oracle.hpp) and, when present, of
oracle/_ref/libref.so (the reference's own std-only headers compiled in place).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_REF_PATH = os.path.join(_HERE, "_ref", "libref.so")

FFT_RADIX2 = 0  # independent arithmetic (structure of ALG/fourier/fft.hpp)
FFT_MIRROR = 1  # bit-exact mirror of the GPU arithmetic contract (fft2048.cuh)


def build(ref: bool = True) -> None:
    """Compile the oracle (and the reference shim when /root/reference exists)."""
    subprocess.run(["make", "-s", "-C", _HERE, "all"], check=True)
    if ref and os.path.isdir("/root/reference/blocks/include"):
        subprocess.run(["make", "-s", "-C", _HERE, "ref"], check=True)


class SyncwordTag(C.Structure):
    # oracle.hpp: struct SyncwordTag
    _fields_ = [
        ("index", C.c_uint64),
        ("amplitude", C.c_float),
        ("phase", C.c_float),
        ("freq", C.c_double),
        ("freq_bin", C.c_int32),
        ("noise_power", C.c_float),
        ("esn0_db", C.c_float),
        ("time_est", C.c_float),
        ("corr_re", C.c_float),
        ("corr_im", C.c_float),
        ("pow", C.c_float),
        ("pow_left", C.c_float),
        ("pow_right", C.c_float),
        ("pow_prev", C.c_float),
        ("pow_next", C.c_float),
        ("_pad", C.c_int32),
    ]


class StreamTag(C.Structure):
    # oracle.hpp: struct StreamTag
    _fields_ = [
        ("index", C.c_int64),
        ("has_syncword", C.c_bool),
        ("amplitude", C.c_float),
        ("phase", C.c_float),
        ("freq", C.c_double),
        ("freq_bin", C.c_int32),
        ("noise_power", C.c_float),
        ("esn0_db", C.c_float),
        ("time_est", C.c_float),
        ("other", C.c_int32),
    ]


_lib = None
_ref = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build(ref=False)
        L = C.CDLL(_LIB_PATH)
        L.orc_rrc.restype = C.c_int
        L.orc_rrc.argtypes = [C.c_double] * 4 + [C.c_size_t, C.c_void_p, C.c_size_t]
        L.orc_fft.restype = C.c_int
        L.orc_fft.argtypes = [C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_void_p]
        L.orc_sd_create.restype = C.c_void_p
        L.orc_sd_create.argtypes = [C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                                    C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_uint64, C.c_float, C.c_int,
                                    C.c_int]
        L.orc_sd_destroy.argtypes = [C.c_void_p]
        L.orc_sd_info.argtypes = [C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_float)]
        L.orc_sd_process.restype = C.c_longlong
        L.orc_sd_process.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t,
                                     C.POINTER(C.c_size_t)]
        L.orc_sd_metric.restype = C.c_size_t
        L.orc_sd_metric.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        L.orc_sd_template.restype = C.c_int
        L.orc_sd_template.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        L.orc_rotator.argtypes = [C.c_float, C.c_void_p, C.c_size_t, C.c_void_p]
        L.orc_cfc_create.restype = C.c_void_p
        L.orc_cfc_create.argtypes = [C.c_size_t]
        L.orc_cfc_destroy.argtypes = [C.c_void_p]
        L.orc_cfc_process.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_int, C.c_double]
        L.orc_wo_create.restype = C.c_void_p
        L.orc_wo_create.argtypes = [C.c_void_p, C.c_size_t]
        L.orc_wo_destroy.argtypes = [C.c_void_p]
        L.orc_wo_process.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_int]
        L.orc_cl_create.restype = C.c_void_p
        L.orc_cl_create.argtypes = [C.c_double, C.c_int, C.c_int]
        L.orc_cl_destroy.argtypes = [C.c_void_p]
        L.orc_cl_process.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_int, C.c_float]
        pf = C.POINTER(C.c_float)
        L.orc_cl_state.argtypes = [C.c_void_p, pf, pf, pf, pf]
        L.orc_mirror_sincosf.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
        L.orc_resampler_create.restype = C.c_void_p
        L.orc_resampler_create.argtypes = [C.c_double, C.c_int, C.c_void_p, C.c_size_t, C.c_size_t]
        L.orc_resampler_destroy.argtypes = [C.c_void_p]
        L.orc_resampler_process.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                                            C.POINTER(C.c_size_t), C.POINTER(C.c_size_t), C.c_void_p,
                                            C.c_void_p, C.c_void_p]
        L.orc_symfilt_create.restype = C.c_void_p
        L.orc_symfilt_create.argtypes = [C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t]
        L.orc_symfilt_destroy.argtypes = [C.c_void_p]
        L.orc_symfilt_process.restype = C.c_int
        L.orc_symfilt_process.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p,
                                          C.POINTER(C.c_size_t), C.POINTER(C.c_size_t), C.c_void_p, C.c_size_t]
        L.orc_sdf_create.restype = C.c_void_p
        L.orc_sdf_create.argtypes = [C.c_size_t, C.c_size_t, C.c_size_t]
        L.orc_sdf_destroy.argtypes = [C.c_void_p]
        L.orc_sdf_process.restype = C.c_longlong
        L.orc_sdf_process.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.c_size_t, C.c_void_p, C.c_size_t,
                                      C.c_void_p, C.c_size_t, C.c_void_p, C.POINTER(C.c_size_t),
                                      C.POINTER(C.c_size_t), C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.orc_interp_fir.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p]
        L.orc_sizeof_syncword_tag.restype = C.c_size_t
        L.orc_sizeof_stream_tag.restype = C.c_size_t
        assert L.orc_sizeof_syncword_tag() == C.sizeof(SyncwordTag)
        assert L.orc_sizeof_stream_tag() == C.sizeof(StreamTag)
        _lib = L
    return _lib


def ref_lib():
    """The reference's own compiled headers, or None when oracle/_ref was not built."""
    global _ref
    if _ref is None and os.path.exists(_REF_PATH):
        R = C.CDLL(_REF_PATH)
        R.ref_rrc.restype = C.c_int
        R.ref_rrc.argtypes = [C.c_double] * 4 + [C.c_size_t, C.c_void_p, C.c_size_t]
        R.ref_pfb_arb_taps.restype = C.c_int
        R.ref_pfb_arb_taps.argtypes = [C.c_void_p, C.c_size_t]
        _ref = R
    return _ref


def _c64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.complex64)


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def root_raised_cosine(gain, fs, symrate, alpha, ntaps) -> np.ndarray:
    out = np.zeros(ntaps + 2, np.float32)
    n = lib().orc_rrc(gain, fs, symrate, alpha, ntaps, out.ctypes.data, out.size)
    assert n > 0
    return out[:n].copy()


def ref_root_raised_cosine(gain, fs, symrate, alpha, ntaps) -> np.ndarray:
    out = np.zeros(ntaps + 2, np.float32)
    n = ref_lib().ref_rrc(gain, fs, symrate, alpha, ntaps, out.ctypes.data, out.size)
    assert n > 0
    return out[:n].copy()


def ref_pfb_arb_taps() -> np.ndarray:
    out = np.zeros(4096, np.float32)
    n = ref_lib().ref_pfb_arb_taps(out.ctypes.data, out.size)
    assert n > 0
    return out[:n].copy()


def fft(x, kind=FFT_RADIX2, which=0) -> np.ndarray:
    x = _c64(x)
    out = np.empty_like(x)
    rc = lib().orc_fft(kind, which, x.size, x.ctypes.data, out.ctypes.data)
    assert rc == 0
    return out


class SyncwordDetection:
    """Restated gr::packet_modem::SyncwordDetection (PM/syncword_detection.hpp)."""

    def __init__(self, rrc_taps, syncword, constellation, min_freq_bin=0, max_freq_bin=0, time_threshold=768,
                 power_threshold=9.5, fft_size=2048, samples_per_symbol=4, fft_kind=FFT_RADIX2,
                 record_metric=False):
        rrc = _f32(rrc_taps)
        sw = np.ascontiguousarray(syncword, dtype=np.uint8)
        cst = _c64(constellation)
        self.time_threshold = int(time_threshold)
        self.fft_size = int(fft_size)
        self._h = lib().orc_sd_create(fft_size, samples_per_symbol, rrc.ctypes.data, rrc.size, sw.ctypes.data,
                                      sw.size, cst.ctypes.data, cst.size, min_freq_bin, max_freq_bin,
                                      time_threshold, power_threshold, fft_kind, int(record_metric))
        if not self._h:
            raise ValueError("oracle SyncwordDetection rejected the settings")
        L = C.c_uint32()
        sc = C.c_float()
        lib().orc_sd_info(self._h, C.byref(L), C.byref(sc))
        self.syncword_samples_size = L.value
        self.self_corr = sc.value
        self.stride = self.fft_size - L.value + 1

    def __del__(self):
        if getattr(self, "_h", None):
            try:
                lib().orc_sd_destroy(self._h)
            except Exception:  # interpreter shutdown
                pass
            self._h = None

    def process_bulk(self, x, want_output=True, max_tags=1 << 16):
        """One processBulk call over span x.  Returns (consumed, out[:consumed] or None, tags)."""
        x = _c64(x)
        out = np.zeros(x.size, np.complex64) if want_output else None
        tags = (SyncwordTag * max_tags)()
        nt = C.c_size_t(0)
        c = lib().orc_sd_process(self._h, x.ctypes.data, x.size, out.ctypes.data if want_output else None, tags,
                                 max_tags, C.byref(nt))
        if c < 0:
            raise RuntimeError("oracle tag buffer overflow")
        return int(c), (out[:c] if want_output else None), [tags[i] for i in range(nt.value)]

    def run(self, x, chunk=65536, want_output=False):
        """Feed x the way the GR4 runtime does: offer a span, consume what the block took,
        re-offer the rest (GR/Block.hpp:1537-1651).  Returns (consumed_total, out, tags)."""
        x = _c64(x)
        pos = 0
        outs, tags = [], []
        while x.size - pos >= self.fft_size:
            c, o, t = self.process_bulk(x[pos:pos + chunk], want_output)
            if c == 0:
                break
            if want_output:
                outs.append(o)
            tags.extend(t)
            pos += c
        out = np.concatenate(outs) if outs else np.zeros(0, np.complex64)
        return pos, out, tags

    def metric(self, n):
        p = np.zeros(n, np.float32)
        b = np.zeros(n, np.int8)
        m = lib().orc_sd_metric(self._h, p.ctypes.data, b.ctypes.data, n)
        return p[:m], b[:m]

    def template(self, k):
        out = np.zeros(self.fft_size, np.complex64)
        assert lib().orc_sd_template(self._h, k, out.ctypes.data) == 0
        return out


def rotator(x, phase_incr) -> np.ndarray:
    x = _c64(x)
    out = np.empty_like(x)
    lib().orc_rotator(phase_incr, x.ctypes.data, x.size, out.ctypes.data)
    return out


class CoarseFrequencyCorrection:
    """Restated CoarseFrequencyCorrection<float> (PM/coarse_frequency_correction.hpp)."""

    def __init__(self, delay=0):
        self._h = lib().orc_cfc_create(delay)

    def __del__(self):
        if getattr(self, "_h", None):
            try:
                lib().orc_cfc_destroy(self._h)
            except Exception:  # interpreter shutdown
                pass
            self._h = None

    def process_bulk(self, x, freq=None):
        """One chunk; `freq` is the syncword_freq of the tag on its first sample (None: no tag)."""
        x = _c64(x)
        out = np.empty_like(x)
        lib().orc_cfc_process(self._h, x.ctypes.data, x.size, out.ctypes.data, int(freq is not None),
                              float(freq if freq is not None else 0.0))
        return out

    def run(self, x, tags):
        """Whole stream with (index, syncword_freq) tags: the runtime cuts chunks so that every tag sits
        on the first sample of a chunk (GR/Block.hpp:1501-1508)."""
        x = _c64(x)
        out = np.empty_like(x)
        cuts = [0] + [int(i) for i, _ in tags] + [x.size]
        freqs = [None] + [f for _, f in tags]
        for a, b, f in zip(cuts[:-1], cuts[1:], freqs):
            if b > a or f is not None:
                out[a:b] = self.process_bulk(x[a:b], f)
        return out


def _chunked(x, tags, fn):
    """Whole stream with (index, value) tags: the runtime cuts chunks so that every tag sits on the first
    sample of a chunk (GR/Block.hpp:1501-1508); fn(chunk, value_or_None) -> out chunk."""
    x = _c64(x)
    out = np.empty_like(x)
    cuts = [0] + [int(i) for i, _ in tags] + [x.size]
    vals = [None] + [v for _, v in tags]
    for a, b, v in zip(cuts[:-1], cuts[1:], vals):
        if b > a:
            out[a:b] = fn(x[a:b], v)
    return out


class SyncwordWipeoff:
    """Restated SyncwordWipeoff<c64, float> (PM/syncword_wipeoff.hpp)."""

    def __init__(self, syncword):
        sw = np.ascontiguousarray(syncword, dtype=np.float32)
        self._h = lib().orc_wo_create(sw.ctypes.data, sw.size)

    def __del__(self):
        if getattr(self, "_h", None):
            try:
                lib().orc_wo_destroy(self._h)
            except Exception:  # interpreter shutdown
                pass
            self._h = None

    def process_bulk(self, x, has_tag=False):
        x = _c64(x)
        out = np.empty_like(x)
        lib().orc_wo_process(self._h, x.ctypes.data, x.size, out.ctypes.data, int(bool(has_tag)))
        return out

    def run(self, x, tag_indices):
        """Whole stream; a syncword_amplitude tag at every index of `tag_indices` (sorted, distinct)."""
        return _chunked(x, [(i, True) for i in tag_indices], lambda c, v: self.process_bulk(c, v is not None))


TRIG_LIBM = 0    # std::cos / std::sin, as the reference
TRIG_MIRROR = 1  # bit-exact mirror of the GPU's b200_sincosf (csrc/costas.cuh)
PILOT, BPSK_LOOP, QPSK_LOOP = 0, 1, 2


class CostasLoop:
    """Restated CostasLoop<float, float> (PM/costas_loop.hpp)."""

    def __init__(self, loop_bandwidth=0.01, constellation=BPSK_LOOP, trig=TRIG_LIBM):
        self._h = lib().orc_cl_create(float(loop_bandwidth), int(constellation), int(trig))

    def __del__(self):
        if getattr(self, "_h", None):
            try:
                lib().orc_cl_destroy(self._h)
            except Exception:  # interpreter shutdown
                pass
            self._h = None

    def process_bulk(self, x, phase=None):
        """One chunk; `phase` is the syncword_phase of the tag on its first sample (None: no tag)."""
        x = _c64(x)
        out = np.empty_like(x)
        lib().orc_cl_process(self._h, x.ctypes.data, x.size, out.ctypes.data, int(phase is not None),
                             float(phase if phase is not None else 0.0))
        return out

    def run(self, x, tags):
        """Whole stream with (index, syncword_phase) tags (sorted, distinct indices)."""
        return _chunked(x, list(tags), self.process_bulk)

    def state(self):
        """(_phase, _freq, _k1, _k2)"""
        v = [C.c_float() for _ in range(4)]
        lib().orc_cl_state(self._h, *[C.byref(a) for a in v])
        return tuple(a.value for a in v)


def mirror_sincosf(x):
    x = np.ascontiguousarray(x, dtype=np.float32)
    s, c = np.empty_like(x), np.empty_like(x)
    lib().orc_mirror_sincosf(x.ctypes.data, x.size, s.ctypes.data, c.ctypes.data)
    return s, c


class PfbArbResampler:
    """Restated PfbArbResampler<c64,c64,float,TRate> (PM/pfb_arb_resampler.hpp)."""

    def __init__(self, rate, taps, filter_size=32, use_double=False):
        t = _f32(taps)
        self._h = lib().orc_resampler_create(float(rate), int(use_double), t.ctypes.data, t.size, filter_size)
        if not self._h:
            raise ValueError("oracle PfbArbResampler rejected the settings")

    def __del__(self):
        if getattr(self, "_h", None):
            try:
                lib().orc_resampler_destroy(self._h)
            except Exception:  # interpreter shutdown
                pass
            self._h = None

    def process_bulk(self, x, n_out, timing=False):
        x = _c64(x)
        out = np.zeros(n_out, np.complex64)
        cons, prod = C.c_size_t(0), C.c_size_t(0)
        arms = np.zeros(n_out, np.uint32) if timing else None
        cnts = np.zeros(n_out, np.uint64) if timing else None
        accs = np.zeros(n_out, np.float64) if timing else None
        lib().orc_resampler_process(self._h, x.ctypes.data, x.size, out.ctypes.data, n_out, C.byref(cons),
                                    C.byref(prod), arms.ctypes.data if timing else None,
                                    cnts.ctypes.data if timing else None, accs.ctypes.data if timing else None)
        p = prod.value
        if timing:
            return cons.value, out[:p], arms[:p], cnts[:p], accs[:p]
        return cons.value, out[:p]


class SymbolFilter:
    """Restated SymbolFilter<c64,c64,float> (PM/symbol_filter.hpp)."""

    def __init__(self, taps, num_arms, samples_per_symbol=4, delay=0):
        t = _f32(taps)
        self._h = lib().orc_symfilt_create(samples_per_symbol, t.ctypes.data, t.size, num_arms, delay)
        if not self._h:
            raise ValueError("oracle SymbolFilter rejected the settings")

    def __del__(self):
        if getattr(self, "_h", None):
            try:
                lib().orc_symfilt_destroy(self._h)
            except Exception:  # interpreter shutdown
                pass
            self._h = None

    def process_bulk(self, x, n_out, tag: StreamTag | None = None, max_tags=64):
        x = _c64(x)
        out = np.zeros(n_out, np.complex64)
        cons, prod = C.c_size_t(0), C.c_size_t(0)
        ot = (StreamTag * max_tags)()
        n = lib().orc_symfilt_process(self._h, x.ctypes.data, x.size, out.ctypes.data, n_out,
                                      C.byref(tag) if tag is not None else None, C.byref(cons), C.byref(prod), ot,
                                      max_tags)
        assert n >= 0
        return cons.value, out[:prod.value], [ot[i] for i in range(n)]


class SyncwordDetectionFilter:
    """Restated SyncwordDetectionFilter (PM/syncword_detection_filter.hpp)."""

    def __init__(self, samples_per_symbol=4, syncword_size=64, header_size=128):
        self._h = lib().orc_sdf_create(samples_per_symbol, syncword_size, header_size)

    def __del__(self):
        if getattr(self, "_h", None):
            try:
                lib().orc_sdf_destroy(self._h)
            except Exception:  # interpreter shutdown
                pass
            self._h = None

    def process_bulk(self, x, n_out=None, tag: StreamTag | None = None, header=None, n_ignored=0):
        """header: None | ("parsed", packet_length) | ("invalid",)"""
        x = _c64(x)
        n_out = x.size if n_out is None else n_out
        out = np.zeros(n_out, np.complex64)
        hk, plen = 0, 0
        if header is not None:
            hk, plen = (2, 0) if header[0] == "invalid" else (1, int(header[1]))
        hu, iu = C.c_size_t(0), C.c_size_t(0)
        to = StreamTag()
        fwd, inpkt = C.c_int(0), C.c_int(0)
        c = lib().orc_sdf_process(self._h, hk, plen, n_ignored, x.ctypes.data, x.size, out.ctypes.data, n_out,
                                  C.byref(tag) if tag is not None else None, C.byref(hu), C.byref(iu),
                                  C.byref(to), C.byref(fwd), C.byref(inpkt))
        if c < 0:
            raise RuntimeError("oracle SyncwordDetectionFilter threw")
        return int(c), out[:c], (to if fwd.value else None), hu.value, iu.value, bool(inpkt.value)


def interpolating_fir(x, taps, interpolation) -> np.ndarray:
    x = _c64(x)
    t = _f32(taps)
    out = np.zeros(x.size * interpolation, np.complex64)
    lib().orc_interp_fir(t.ctypes.data, t.size, interpolation, x.ctypes.data, x.size, out.ctypes.data)
    return out
