"""Seeded synthetic captures for tests and bench.py (not on the product data path).

Signal model of BASELINE.json configs / SURVEY.md §8(d): back-to-back frames of
  64-symbol BPSK syncword + 128 QPSK header symbols + QPSK payload ((bytes+4)*4 symbols),
x4 RRC pulse shaping with the transmitter's taps (PM/packet_transmitter_rrc_taps.hpp:8-28),
carrier frequency offset (what Rotator applies, apps/packet_transceiver.cpp:74-75) and complex
AWGN with N0 = 0.32 * sps * 10^(-EsN0/10) (apps/packet_transceiver.cpp:48-52).

Two generators with the same structure: numpy (host, any size that fits RAM) and torch
(device, used by bench.py for the 2^30-sample capture).  They are NOT bit-identical to each
other; parity tests always feed the same array to the GPU path and to the oracle.
"""
from __future__ import annotations

import math

import numpy as np

from .firdes import SYNCWORD, root_raised_cosine

TX_POWER = 0.32  # apps/packet_transceiver.cpp:48


def tx_rrc_taps(sps: int = 4) -> np.ndarray:
    """PM/packet_transmitter_rrc_taps.hpp:8-28."""
    rrc = root_raised_cosine(1.0, float(sps), 1.0, 0.35, sps * 11)
    m = max(float(np.sum(np.abs(rrc[j::sps]))) for j in range(sps))
    return (rrc * np.float32(0.9 / m)).astype(np.float32)


def frame_symbols(rng: np.random.Generator, payload_bytes: int) -> np.ndarray:
    nq = 128 + (payload_bytes + 4) * 4
    q = rng.integers(0, 4, nq)
    qpsk = ((1 - 2 * (q & 1)) + 1j * (1 - 2 * (q >> 1))) / math.sqrt(2.0)
    sw = 1.0 - 2.0 * SYNCWORD.astype(np.float64)
    return np.concatenate([sw.astype(np.complex128), qpsk])


def packet_capture(n_samples: int, seed: int = 1, esn0_db: float = 20.0, cfo: float = 0.005,
                   payload_bytes: int = 1500, sps: int = 4, gap_symbols: int = 0, noise_seed: int = 2,
                   signal: bool = True) -> tuple[np.ndarray, np.ndarray]:
    """Returns (capture complex64[n_samples], syncword_start_sample_indices)."""
    rng = np.random.default_rng(seed)
    taps = tx_rrc_taps(sps).astype(np.float64)
    nsym = n_samples // sps + 2
    syms = np.zeros(nsym, np.complex128)
    starts = []
    pos = 0
    if signal:
        while pos < nsym:
            f = frame_symbols(rng, payload_bytes)
            m = min(f.size, nsym - pos)
            syms[pos:pos + m] = f[:m]
            if m >= 64:
                starts.append(pos * sps)
            pos += f.size + gap_symbols
    up = np.zeros(nsym * sps, np.complex128)
    up[::sps] = syms
    x = np.convolve(up, taps)[:n_samples]
    if cfo != 0.0:
        x = x * np.exp(1j * cfo * np.arange(n_samples, dtype=np.float64))
    n0 = TX_POWER * sps * 10.0 ** (-0.1 * esn0_db)
    nrng = np.random.default_rng(noise_seed)
    noise = (nrng.standard_normal(n_samples) + 1j * nrng.standard_normal(n_samples)) * math.sqrt(n0 / 2.0)
    x = (x + noise).astype(np.complex64)
    return x, np.array([s for s in starts if s + 297 <= n_samples], dtype=np.int64)


def _hash32_torch(idx, salt: int):
    """32-bit mixing hash of an int64 index tensor (pure function of index and salt)."""
    lo = idx & 0xFFFFFFFF
    hi = (idx >> 32) & 0xFFFFFFFF
    h = (lo * 2654435761 + hi * 40503 + salt) & 0xFFFFFFFF
    h = ((h ^ (h >> 16)) * 2246822507) & 0xFFFFFFFF
    h = ((h ^ (h >> 13)) * 3266489909) & 0xFFFFFFFF
    return h ^ (h >> 16)


def packet_capture_torch(n_samples: int, device, seed: int = 1, esn0_db: float = 20.0, cfo: float = 0.005,
                         payload_bytes: int = 1500, sps: int = 4, chunk: int = 1 << 24, start: int = 0):
    """Device-side generator for large captures: samples [start, start + n_samples) of an endless
    stream, as a torch.complex64 tensor on `device`.  Every sample is a pure function of
    (seed, absolute index) — symbols and noise come from integer hashes — so any segment
    (e.g. one GPU's time shard plus halo) can be generated independently and is consistent
    with its neighbours.  Same frame structure and statistics as packet_capture()."""
    import torch

    taps = torch.tensor(tx_rrc_taps(sps), dtype=torch.float32, device=device)
    ntaps = taps.numel()
    frame_len = 64 + 128 + (payload_bytes + 4) * 4
    sw = torch.tensor(1.0 - 2.0 * SYNCWORD.astype(np.float32), device=device)
    out = torch.empty(n_samples, dtype=torch.complex64, device=device)
    n0 = TX_POWER * sps * 10.0 ** (-0.1 * esn0_db)
    hist = (ntaps + sps - 1) // sps  # symbols of filter memory
    kernel = taps.flip(0).view(1, 1, -1)
    for c0 in range(0, n_samples, chunk):
        m = min(chunk, n_samples - c0)
        s0 = start + c0
        sym0 = s0 // sps - hist
        nsym = m // sps + hist + 2
        idx = torch.arange(sym0, sym0 + nsym, device=device, dtype=torch.int64)
        h = _hash32_torch(idx.clamp(min=0), seed * 7919 + 17)
        re = 1.0 - 2.0 * (h & 1).to(torch.float32)
        im = 1.0 - 2.0 * ((h >> 1) & 1).to(torch.float32)
        inframe = torch.remainder(idx.clamp(min=0), frame_len)
        is_sw = inframe < 64
        re = torch.where(is_sw, sw[inframe.clamp(max=63)], re * (1.0 / math.sqrt(2.0)))
        im = torch.where(is_sw, torch.zeros_like(im), im * (1.0 / math.sqrt(2.0)))
        valid = (idx >= 0).to(torch.float32)
        re, im = re * valid, im * valid
        up = torch.zeros(2, 1, nsym * sps, dtype=torch.float32, device=device)
        up[0, 0, ::sps] = re
        up[1, 0, ::sps] = im
        del re, im, h, inframe, is_sw, valid, idx
        up = torch.nn.functional.pad(up, (ntaps - 1, 0))
        y = torch.nn.functional.conv1d(up, kernel)  # causal FIR: y[n] = sum_k taps[k] up[n-k]
        del up
        off = s0 - sym0 * sps
        sig = torch.complex(y[0, 0, off:off + m], y[1, 0, off:off + m])
        del y
        n = torch.arange(s0, s0 + m, device=device, dtype=torch.int64)
        ph = torch.remainder(n.to(torch.float64) * cfo, 2.0 * math.pi).to(torch.float32)
        sig = sig * torch.complex(torch.cos(ph), torch.sin(ph))
        del ph
        u1 = (_hash32_torch(n, seed * 104729 + 1).to(torch.float32) + 1.0) * (1.0 / 4294967296.0)
        u2 = _hash32_torch(n, seed * 1299709 + 2).to(torch.float32) * (2.0 * math.pi / 4294967296.0)
        del n
        r = torch.sqrt(-2.0 * torch.log(u1.clamp(min=1e-12))) * math.sqrt(n0 / 2.0)
        out[c0:c0 + m] = sig + torch.complex(r * torch.cos(u2), r * torch.sin(u2))
        del sig, u1, u2, r
    return out


class DeviceStimulus:
    """The native generator (csrc/stimulus.cu, b200sync_stim_*): frames -> InterpolatingFirFilter ->
    Rotator -> + gaussian NoiseSource in one kernel that writes the capture into HBM; every sample a pure
    function of (seed, absolute index).  Same signal model and parameters as packet_capture()."""

    def __init__(self, seed: int = 1, esn0_db: float | None = 20.0, cfo: float = 0.005, payload_bytes: int = 1500,
                 sps: int = 4, gap_symbols: int = 0, taps=None, device: int = 0):
        import ctypes as C

        from . import _native

        self.taps = np.ascontiguousarray(tx_rrc_taps(sps) if taps is None else taps, dtype=np.float32)
        self.sync = np.ascontiguousarray(1.0 - 2.0 * SYNCWORD.astype(np.float32))
        self.sps, self.seed, self.cfo = int(sps), int(seed), float(np.float32(cfo))
        self.header_symbols, self.payload_symbols, self.gap_symbols = 128, (payload_bytes + 4) * 4, int(gap_symbols)
        self.frame_len = self.sync.size + self.header_symbols + self.payload_symbols + self.gap_symbols
        # NoiseSource amplitude = sqrt(N0), N0 = 0.32 * sps * 10^(-EsN0/10) (apps/packet_transceiver.cpp:48-52)
        self.noise_amplitude = 0.0 if esn0_db is None else float(np.float32(math.sqrt(TX_POWER * sps * 10.0 ** (-0.1 * esn0_db))))
        cfg = _native.StimConfig(self.taps.ctypes.data, self.taps.size, self.sps, self.sync.ctypes.data, self.sync.size,
                                 self.header_symbols, self.payload_symbols, self.gap_symbols, self.cfo,
                                 self.noise_amplitude, self.seed, int(device))
        h = C.c_void_p()
        self._h = C.c_void_p()
        _native.check_stim(_native.lib().b200sync_stim_create(C.byref(cfg), C.byref(h)))
        self._h = h

    def __del__(self):
        try:
            from . import _native

            if getattr(self, "_h", None) is not None and self._h.value:
                _native.lib().b200sync_stim_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def generate_device(self, d_out_ptr: int, n: int, first_sample: int = 0, stream_ptr: int = 0) -> None:
        import ctypes as C

        from . import _native

        _native.check_stim(_native.lib().b200sync_stim_generate_device(self._h, int(first_sample), int(n),
                                                                       C.c_void_p(d_out_ptr),
                                                                       C.c_void_p(stream_ptr or None)))

    def generate(self, n: int, device, first_sample: int = 0):
        """torch.complex64 tensor on `device` holding samples [first_sample, first_sample + n)."""
        import torch

        out = torch.empty(n, dtype=torch.complex64, device=device)
        self.generate_device(out.data_ptr(), n, first_sample, torch.cuda.current_stream().cuda_stream)
        return out

    def symbols(self, k0: int, k1: int) -> np.ndarray:
        """Host restatement of the symbol sequence [k0, k1) (integer hash of the symbol index)."""
        k = np.arange(k0, k1, dtype=np.int64)
        f = np.remainder(np.maximum(k, 0), self.frame_len)
        sd = (self.seed ^ (self.seed >> 32)) & 0xFFFFFFFF
        h = hash32(np.maximum(k, 0), (sd * 7919 + 17) & 0xFFFFFFFF)
        a = np.float32(0.70710678118654752440)
        q = np.where(h & 1, -a, a).astype(np.float32) + 1j * np.where(h & 2, -a, a).astype(np.float32)
        s = np.where(f < self.sync.size, self.sync[np.minimum(f, self.sync.size - 1)].astype(np.complex64),
                     q.astype(np.complex64))
        s = np.where((f >= self.sync.size + self.header_symbols + self.payload_symbols) | (k < 0), 0, s)
        return s.astype(np.complex64)


def hash32(idx: np.ndarray, salt: int) -> np.ndarray:
    """csrc/stimulus.cu: stim_hash (same mixing function as _hash32_torch)."""
    idx = np.asarray(idx, dtype=np.int64).astype(np.uint64)
    lo, hi = idx & np.uint64(0xFFFFFFFF), (idx >> np.uint64(32)) & np.uint64(0xFFFFFFFF)
    m = np.uint64(0xFFFFFFFF)
    h = (lo * np.uint64(2654435761) + hi * np.uint64(40503) + np.uint64(salt)) & m
    h = ((h ^ (h >> np.uint64(16))) * np.uint64(2246822507)) & m
    h = ((h ^ (h >> np.uint64(13))) * np.uint64(3266489909)) & m
    return (h ^ (h >> np.uint64(16))).astype(np.uint64)
