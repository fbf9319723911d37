"""Host-side tap design helper: the GR3-equivalent root-raised-cosine design the reference
flowgraphs use to make `rrc_taps` (PM/firdes.hpp:30-76; call sites
PM/packet_receiver.hpp:60-74, benchmarks/benchmark_syncword_detection.cpp:48-62).
Settings preparation only — not on the data path."""
from __future__ import annotations

import math

import numpy as np


def root_raised_cosine(gain: float, sampling_freq: float, symbol_rate: float, alpha: float, ntaps: int) -> np.ndarray:
    ntaps |= 1
    spb = sampling_freq / symbol_rate
    taps = np.zeros(ntaps, np.float64)
    for i in range(ntaps):
        xindx = float(i - ntaps // 2)
        x1 = math.pi * xindx / spb
        x2 = 4.0 * alpha * xindx / spb
        x3 = x2 * x2 - 1.0
        if abs(x3) >= 0.000001:
            if i != ntaps // 2:
                num = math.cos((1.0 + alpha) * x1) + math.sin((1.0 - alpha) * x1) / (4.0 * alpha * xindx / spb)
            else:
                num = math.cos((1.0 + alpha) * x1) + (1.0 - alpha) * math.pi / (4.0 * alpha)
            den = x3 * math.pi
        else:
            if alpha == 1.0:
                taps[i] = -1.0
                continue
            x3 = (1.0 - alpha) * x1
            x2 = (1.0 + alpha) * x1
            num = (math.sin(x2) * (1.0 + alpha) * math.pi
                   - math.cos(x3) * ((1.0 - alpha) * math.pi * spb) / (4.0 * alpha * xindx)
                   + math.sin(x3) * spb * spb / (4.0 * alpha * xindx * xindx))
            den = -32.0 * math.pi * alpha * alpha * xindx / spb
        taps[i] = 4.0 * alpha * num / den
    scale = 0.0
    for t in taps:  # std::accumulate order
        scale += t
    return (taps * gain / scale).astype(np.float32)


def unit_energy_rrc(samples_per_symbol: int = 4, span_symbols: int = 11, alpha: float = 0.35) -> np.ndarray:
    """The receiver's matched-filter taps: RRC normalised to unit energy in float32
    (PM/packet_receiver.hpp:60-74)."""
    rrc = root_raised_cosine(1.0, float(samples_per_symbol), 1.0, alpha, samples_per_symbol * span_symbols)
    norm = np.float32(0.0)
    for x in rrc:
        norm = np.float32(norm + x * x)
    norm = np.float32(math.sqrt(norm))
    return (rrc / norm).astype(np.float32)


def pfb_matched_filter_taps(samples_per_symbol: int = 4, num_arms: int = 32, span_symbols: int = 11,
                            alpha: float = 0.35) -> np.ndarray:
    """SymbolFilter's polyphase RRC prototype: num_arms x (sps*span) taps with gain num_arms / ||rrc||
    (PM/packet_receiver.hpp:96-110; 1408 taps = 32 arms x 44 at the defaults)."""
    rrc = root_raised_cosine(1.0, float(samples_per_symbol), 1.0, alpha, samples_per_symbol * span_symbols)
    norm = math.sqrt(float(np.sum(rrc.astype(np.float64) ** 2)))
    t = root_raised_cosine(num_arms / norm, float(num_arms * samples_per_symbol), 1.0, alpha,
                           num_arms * samples_per_symbol * span_symbols)
    return t[:-1]


def lowpass_prototype_taps(num_arms: int = 32, taps_per_arm: int = 40) -> np.ndarray:
    """A num_arms*taps_per_arm-tap low-pass prototype (Blackman-windowed sinc, cutoff 0.5/num_arms, DC gain
    num_arms) for PfbArbResampler when the caller has no design of its own.  The reference ships a
    hard-coded 1280-tap remez design (PM/pfb_arb_taps.hpp:13); any prototype of the same shape costs the
    same on the GPU."""
    n = num_arms * taps_per_arm
    k = np.arange(n, dtype=np.float64) - (n - 1) / 2.0
    h = np.sinc(k / num_arms) * np.blackman(n)
    return (h * (num_arms / h.sum())).astype(np.float32)


# CCSDS 64-bit attached sync marker as used by the reference (PM/packet_receiver.hpp:45-59)
SYNCWORD = np.array([0, 0, 0, 0, 0, 0, 1, 1, 0, 1, 0, 0, 0, 1, 1, 1, 0, 1, 1, 1, 0, 1, 1, 0, 1, 1, 0, 0, 0, 1, 1, 1,
                     0, 0, 1, 0, 0, 1, 1, 1, 0, 0, 1, 0, 1, 0, 0, 0, 1, 0, 0, 1, 0, 1, 0, 1, 1, 0, 1, 1, 0, 0, 0, 0],
                    dtype=np.uint8)
BPSK = np.array([1.0 + 0.0j, -1.0 + 0.0j], dtype=np.complex64)
