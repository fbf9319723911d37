"""ORACLE — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes front end of oracle/_ref/librefblocks.so: the REFERENCE's own RX-synchronisation blocks
(blocks/include/gnuradio-4.0/packet-modem/*.hpp, compiled unmodified against the stand-in GR4 runtime of
oracle/ref_stub/, see oracle/ref_blocks.cpp).  Exists only where /root/reference does (the build container):
tests that use it skip elsewhere, and tests/golden/ref_blocks_golden.npz carries its outputs to the GPU box.

The drivers below play the GR4 scheduler the way oracle/pyoracle.py does for the restated blocks: chunks are
cut so that every tag sits on the first item of a chunk (GR/Block.hpp:1501-1508), consume()/publish() are
honoured, the remainder is offered again.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "_ref", "librefblocks.so")
_lib = None


class RefTag(C.Structure):
    _fields_ = [("index", C.c_int64), ("freq", C.c_double), ("amplitude", C.c_float), ("phase", C.c_float),
                ("noise_power", C.c_float), ("esn0_db", C.c_float), ("time_est", C.c_float), ("freq_bin", C.c_int32),
                ("no_syncword", C.c_int32), ("other", C.c_int32)]


def available() -> bool:
    return os.path.exists(_PATH)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(_PATH)
        vp, sz, psz = C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)
        L.refblk_sizeof_tag.restype = sz
        assert L.refblk_sizeof_tag() == C.sizeof(RefTag)
        L.refblk_sd_create.restype = vp
        L.refblk_sd_create.argtypes = [vp, sz, vp, sz, vp, sz, C.c_int, C.c_int, sz, C.c_float, sz]
        L.refblk_sd_destroy.argtypes = [vp]
        L.refblk_sd_process.argtypes = [vp, vp, sz, vp, psz, vp, sz, psz]
        L.refblk_rotator.argtypes = [C.c_float, vp, sz, vp]
        for n in ("cfc", "wo", "cl"):
            getattr(L, f"refblk_{n}_destroy").argtypes = [vp]
            getattr(L, f"refblk_{n}_process").argtypes = [vp, vp, sz, vp, vp]
        L.refblk_cfc_create.restype = vp
        L.refblk_cfc_create.argtypes = [sz]
        L.refblk_wo_create.restype = vp
        L.refblk_wo_create.argtypes = [vp, sz]
        L.refblk_cl_create.restype = vp
        L.refblk_cl_create.argtypes = [C.c_double, C.c_char_p]
        pf = C.POINTER(C.c_float)
        L.refblk_cl_state.argtypes = [vp, pf, pf, pf, pf]
        L.refblk_sf_create.restype = vp
        L.refblk_sf_create.argtypes = [vp, sz, sz, sz, sz]
        L.refblk_sf_destroy.argtypes = [vp]
        L.refblk_sf_process.argtypes = [vp, vp, sz, vp, sz, vp, psz, psz, vp, sz, psz]
        L.refblk_rs_create.restype = vp
        L.refblk_rs_create.argtypes = [C.c_float, vp, sz, sz]
        L.refblk_rs_destroy.argtypes = [vp]
        L.refblk_rs_process.argtypes = [vp, vp, sz, vp, sz, psz, psz]
        L.refblk_interp_fir.argtypes = [vp, sz, sz, vp, sz, vp]
        L.refblk_sdf_create.restype = vp
        L.refblk_sdf_create.argtypes = [sz, sz, sz]
        L.refblk_sdf_destroy.argtypes = [vp]
        L.refblk_sdf_process.restype = C.c_longlong
        pi = C.POINTER(C.c_int)
        L.refblk_sdf_process.argtypes = [vp, C.c_int, C.c_uint64, sz, vp, sz, vp, sz, vp, psz, psz, vp, pi, pi]
        _lib = L
    return _lib


def _c64(a):
    return np.ascontiguousarray(a, dtype=np.complex64)


def _tag(**kw) -> RefTag:
    t = RefTag()
    for k, v in kw.items():
        setattr(t, k, v)
    return t


class _Handle:
    _destroy = None

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                getattr(lib(), self._destroy)(self._h)
                self._h = None
        except Exception:
            pass


class SyncwordDetection(_Handle):
    """gr::packet_modem::SyncwordDetection itself (PM/syncword_detection.hpp), FFT = oracle radix-2."""
    _destroy = "refblk_sd_destroy"

    def __init__(self, rrc_taps, syncword, constellation, min_freq_bin=0, max_freq_bin=0, time_threshold=768,
                 power_threshold=9.5, fft_size=2048):
        rrc = np.ascontiguousarray(rrc_taps, np.float32)
        sw = np.ascontiguousarray(syncword, np.uint8)
        cst = _c64(constellation)
        self._h = lib().refblk_sd_create(rrc.ctypes.data, rrc.size, sw.ctypes.data, sw.size, cst.ctypes.data, cst.size,
                                         min_freq_bin, max_freq_bin, time_threshold, power_threshold, int(fft_size))
        self.fft_size = int(fft_size)
        if not self._h:
            raise ValueError("reference SyncwordDetection::start() threw")
        self.published = 0

    def run(self, x, chunk=65536):
        """-> (items consumed, delayed output, [RefTag with absolute output index])"""
        x = _c64(x)
        outs, tags, pos = [], [], 0
        tb = (RefTag * 4096)()
        while x.size - pos >= self.fft_size:
            seg = x[pos:pos + chunk]
            out = np.empty(seg.size, np.complex64)
            c, nt = C.c_size_t(0), C.c_size_t(0)
            st = lib().refblk_sd_process(self._h, seg.ctypes.data, seg.size, out.ctypes.data, C.byref(c), tb, 4096,
                                         C.byref(nt))
            if st != 0 or c.value == 0:
                break
            for i in range(nt.value):
                t = RefTag.from_buffer_copy(tb[i])
                t.index += self.published
                tags.append(t)
            outs.append(out[:c.value])
            pos += c.value
            self.published += c.value
        return pos, (np.concatenate(outs) if outs else np.zeros(0, np.complex64)), tags


def rotator(x, phase_incr):
    x = _c64(x)
    out = np.empty_like(x)
    lib().refblk_rotator(phase_incr, x.ctypes.data, x.size, out.ctypes.data)
    return out


def _run_tagged(fn, h, x, tags):
    """x through a one-in-one-out block; tags = [(index, RefTag)] sorted, distinct; chunks cut at tags."""
    x = _c64(x)
    out = np.empty_like(x)
    cuts = [0] + [int(i) for i, _ in tags] + [x.size]
    vals = [None] + [t for _, t in tags]
    for a, b, t in zip(cuts[:-1], cuts[1:], vals):
        if b > a:
            seg, o = x[a:b], np.empty(b - a, np.complex64)
            st = fn(h, seg.ctypes.data, seg.size, o.ctypes.data, C.byref(t) if t is not None else None)
            assert st == 0, st
            out[a:b] = o
    return out


class CoarseFrequencyCorrection(_Handle):
    _destroy = "refblk_cfc_destroy"

    def __init__(self, delay=0):
        self._h = lib().refblk_cfc_create(delay)

    def run(self, x, tags):
        """tags = [(index, syncword_freq)]"""
        return _run_tagged(lib().refblk_cfc_process, self._h, x, [(i, _tag(freq=f)) for i, f in tags])


class SyncwordWipeoff(_Handle):
    _destroy = "refblk_wo_destroy"

    def __init__(self, syncword):
        sw = np.ascontiguousarray(syncword, np.float32)
        self._h = lib().refblk_wo_create(sw.ctypes.data, sw.size)

    def run(self, x, tag_indices):
        return _run_tagged(lib().refblk_wo_process, self._h, x, [(i, _tag(amplitude=1.0)) for i in tag_indices])


class CostasLoop(_Handle):
    _destroy = "refblk_cl_destroy"

    def __init__(self, loop_bandwidth=0.01, constellation="BPSK"):
        self._h = lib().refblk_cl_create(loop_bandwidth, constellation.encode())
        if not self._h:
            raise ValueError("reference CostasLoop::settingsChanged() threw")

    def run(self, x, tags):
        """tags = [(index, syncword_phase)]"""
        return _run_tagged(lib().refblk_cl_process, self._h, x, [(i, _tag(phase=p)) for i, p in tags])

    def state(self):
        v = [C.c_float() for _ in range(4)]
        lib().refblk_cl_state(self._h, *[C.byref(a) for a in v])
        return tuple(a.value for a in v)


class SymbolFilter(_Handle):
    _destroy = "refblk_sf_destroy"

    def __init__(self, taps, num_arms, samples_per_symbol=4, delay=0):
        t = np.ascontiguousarray(taps, np.float32)
        self._h = lib().refblk_sf_create(t.ctypes.data, t.size, num_arms, samples_per_symbol, delay)
        if not self._h:
            raise ValueError("reference SymbolFilter::settingsChanged() threw")

    def run(self, x, tags, chunk=50000):
        """tags = [(index, RefTag)]; -> (symbols, [(output index, RefTag)])"""
        x = _c64(x)
        by_index = dict(tags)
        cuts = sorted(by_index)
        pos, ys, otags, nout = 0, [], [], 0
        tb = (RefTag * 64)()
        while pos < x.size:
            end = min([c for c in cuts if c > pos] + [x.size, pos + chunk])
            seg = x[pos:end]
            out = np.empty(seg.size + 2, np.complex64)
            c, p, nt = C.c_size_t(0), C.c_size_t(0), C.c_size_t(0)
            t = by_index.get(pos)
            st = lib().refblk_sf_process(self._h, seg.ctypes.data, seg.size, out.ctypes.data, out.size,
                                         C.byref(t) if t is not None else None, C.byref(c), C.byref(p), tb, 64,
                                         C.byref(nt))
            assert st == 0 and c.value == seg.size, (st, c.value, seg.size)
            for i in range(nt.value):
                q = RefTag.from_buffer_copy(tb[i])
                otags.append((nout + q.index, q))
            ys.append(out[:p.value].copy())
            nout += p.value
            pos += c.value
        return np.concatenate(ys), otags


class PfbArbResampler(_Handle):
    _destroy = "refblk_rs_destroy"

    def __init__(self, rate, taps, filter_size=32):
        t = np.ascontiguousarray(taps, np.float32)
        self._h = lib().refblk_rs_create(rate, t.ctypes.data, t.size, filter_size)
        if not self._h:
            raise ValueError("reference PfbArbResampler::settingsChanged() threw")

    def run(self, x, chunk=40000, out_chunk=40000):
        x = _c64(x)
        pos, ys = 0, []
        while pos < x.size:
            seg = x[pos:pos + chunk]
            out = np.empty(out_chunk, np.complex64)
            c, p = C.c_size_t(0), C.c_size_t(0)
            st = lib().refblk_rs_process(self._h, seg.ctypes.data, seg.size, out.ctypes.data, out.size, C.byref(c),
                                         C.byref(p))
            assert st == 0
            ys.append(out[:p.value].copy())
            pos += c.value
            if c.value == 0 and p.value == 0:
                break
        return pos, np.concatenate(ys)


def interpolating_fir(x, taps, interpolation):
    x = _c64(x)
    t = np.ascontiguousarray(taps, np.float32)
    out = np.empty(x.size * interpolation, np.complex64)
    st = lib().refblk_interp_fir(t.ctypes.data, t.size, interpolation, x.ctypes.data, x.size, out.ctypes.data)
    assert st == 0, st
    return out


class SyncwordDetectionFilter(_Handle):
    """gr::packet_modem::SyncwordDetectionFilter itself; one processBulk call per process_bulk()."""
    _destroy = "refblk_sdf_destroy"

    def __init__(self, samples_per_symbol=4, syncword_size=64, header_size=128):
        self._h = lib().refblk_sdf_create(samples_per_symbol, syncword_size, header_size)

    def process_bulk(self, x, n_out=None, tag: RefTag | None = None, header=None, n_ignored=0):
        """header: None | ("parsed", packet_length) | ("invalid",).  Returns
        (consumed, out, forwarded RefTag or None, header_consumed, ignored_consumed, in_packet)."""
        x = _c64(x)
        n_out = x.size if n_out is None else n_out
        out = np.zeros(n_out, np.complex64)
        hk, plen = 0, 0
        if header is not None:
            hk, plen = (2, 0) if header[0] == "invalid" else (1, int(header[1]))
        hu, iu = C.c_size_t(0), C.c_size_t(0)
        to = RefTag()
        fwd, inpkt = C.c_int(0), C.c_int(0)
        c = lib().refblk_sdf_process(self._h, hk, plen, n_ignored, x.ctypes.data, x.size, out.ctypes.data, n_out,
                                     C.byref(tag) if tag is not None else None, C.byref(hu), C.byref(iu), C.byref(to),
                                     C.byref(fwd), C.byref(inpkt))
        if c < 0:
            raise RuntimeError("reference SyncwordDetectionFilter threw")
        return int(c), out[:c], (to if fwd.value else None), hu.value, iu.value, bool(inpkt.value)
