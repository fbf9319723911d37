"""CPU tests (build container only): the oracle's restated blocks against the REFERENCE's own block code,
compiled unmodified from /root/reference against a stand-in GR4 runtime (oracle/ref_blocks.cpp,
oracle/ref_stub/) into oracle/_ref/librefblocks.so.  The only piece that is not the reference's is the FFT
(FFTW is not in the tree): the stand-in uses the oracle's radix-2 arithmetic, so everything must agree BIT
FOR BIT — state machines, estimator, delay line, tag placement, filters, loops.  Skipped where the reference
tree is absent (the GPU box); tests/golden/ref_blocks_golden.npz (test_reference_blocks_golden below and
tests/test_gpu_reference_golden.py) carries the same reference outputs there."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def ref(oracle):
    from oracle import refblocks

    if not refblocks.available():
        pytest.skip("oracle/_ref/librefblocks.so not built (no /root/reference here)")
    return refblocks


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def _noise(n, seed, scale=1.0):
    rng = np.random.default_rng(seed)
    return ((rng.standard_normal(n) + 1j * rng.standard_normal(n)) * scale).astype(np.complex64)


@pytest.mark.parametrize("bins,esn0,thr,T", [(4, 20.0, 9.5, 768), (0, 20.0, 9.5, 768), (2, 3.0, 7.0, 100),
                                             (8, 0.0, 9.5, 768)])
def test_syncword_detection_is_the_reference(oracle, ref, rx_params, bins, esn0, thr, T):
    from gr4_packet_modem_b200.stimulus import packet_capture

    x, _ = packet_capture(1 << 18, seed=3 + bins, esn0_db=esn0, cfo=0.004 * bins, payload_bytes=120)
    r = ref.SyncwordDetection(**rx_params, min_freq_bin=-bins, max_freq_bin=bins, time_threshold=T, power_threshold=thr)
    rc, rout, rtags = r.run(x, chunk=65536)
    o = oracle.SyncwordDetection(**rx_params, min_freq_bin=-bins, max_freq_bin=bins, time_threshold=T,
                                 power_threshold=thr, fft_kind=oracle.FFT_RADIX2)
    oc, oout, otags = o.run(x, chunk=65536, want_output=True)
    assert rc == oc and np.array_equal(_bits(rout), _bits(oout))
    assert [t.index for t in rtags] == [t.index for t in otags] and len(rtags) >= 10
    for a, b in zip(rtags, otags):
        assert a.freq == b.freq and a.freq_bin == b.freq_bin
        for k in ("amplitude", "phase", "noise_power", "esn0_db", "time_est"):
            assert np.float32(getattr(a, k)).tobytes() == np.float32(getattr(b, k)).tobytes(), k
    # other chunkings of the same stream (the runtime may offer any span >= fft_size)
    r2 = ref.SyncwordDetection(**rx_params, min_freq_bin=-bins, max_freq_bin=bins, time_threshold=T, power_threshold=thr)
    rc2, rout2, rtags2 = r2.run(x, chunk=7001)
    assert rc2 <= rc and np.array_equal(_bits(rout2), _bits(rout[:rc2]))
    assert [t.index for t in rtags2] == [t.index for t in rtags if t.index < rc2]


@pytest.mark.parametrize("fft_size", [512, 1024, 4096])
def test_syncword_detection_other_fft_sizes_is_the_reference(oracle, ref, rx_params, fft_size):
    """fft_size is a setting of the block (PM/syncword_detection.hpp:133): the restatement follows the reference's class
    bit for bit at other sizes too."""
    from gr4_packet_modem_b200.stimulus import packet_capture

    x, _ = packet_capture(1 << 17, seed=21, esn0_db=8.0, cfo=0.006, payload_bytes=60)
    kw = dict(min_freq_bin=-2, max_freq_bin=2, time_threshold=300, power_threshold=8.0, fft_size=fft_size)
    rc, rout, rtags = ref.SyncwordDetection(**rx_params, **kw).run(x, chunk=30000)
    oc, oout, otags = oracle.SyncwordDetection(**rx_params, **kw, fft_kind=oracle.FFT_RADIX2).run(x, chunk=30000,
                                                                                                 want_output=True)
    assert rc == oc and np.array_equal(_bits(rout), _bits(oout))
    assert [t.index for t in rtags] == [t.index for t in otags] and len(rtags) >= 10
    for a, b in zip(rtags, otags):
        assert a.freq == b.freq and a.freq_bin == b.freq_bin
        for k in ("amplitude", "phase", "noise_power", "esn0_db", "time_est"):
            assert np.float32(getattr(a, k)).tobytes() == np.float32(getattr(b, k)).tobytes(), k


def test_reference_settings_errors(ref, rx_params):
    """start() throws for min_freq_bin > max_freq_bin and for a syncword longer than the FFT (:145-152)."""
    with pytest.raises(ValueError):
        ref.SyncwordDetection(**rx_params, min_freq_bin=1, max_freq_bin=0)


def test_filters_and_loops_are_the_reference(oracle, ref, rx_params):
    from gr4_packet_modem_b200.firdes import SYNCWORD, lowpass_prototype_taps, pfb_matched_filter_taps
    from gr4_packet_modem_b200.stimulus import packet_capture, tx_rrc_taps

    n = 150000
    x = _noise(n, 5)
    # Rotator
    assert np.array_equal(_bits(ref.rotator(x, 0.005)), _bits(oracle.rotator(x, 0.005)))
    # PfbArbResampler<float rate>
    rate = float(np.float32(1.0) + np.float32(1e-6) * np.float32(1.2))
    taps = lowpass_prototype_taps(32, 40)
    rc, ry = ref.PfbArbResampler(rate, taps, 32).run(x)
    oc, oy = oracle.PfbArbResampler(rate, taps, 32, use_double=False).process_bulk(x, n + 1000)
    assert rc == oc == n and np.array_equal(_bits(ry), _bits(oy))
    for rate2 in (0.9999988, 1.1234568, 0.75):
        rc, ry = ref.PfbArbResampler(rate2, taps, 32).run(x[:40000])
        oc, oy = oracle.PfbArbResampler(rate2, taps, 32, use_double=False).process_bulk(x[:40000], 60000)
        assert rc == oc and np.array_equal(_bits(ry), _bits(oy))
    # InterpolatingFirFilter
    syms = _noise(5000, 6)
    assert np.array_equal(_bits(ref.interpolating_fir(syms, tx_rrc_taps(4), 4)),
                          _bits(oracle.interpolating_fir(syms, tx_rrc_taps(4), 4)))
    # CoarseFrequencyCorrection
    ftags = [(0, 0.003), (5000, -0.0213), (5010, 0.05), (30001, 0.0101), (70000, -0.15), (n - 1, 0.01)]
    for delay in (0, 26):
        assert np.array_equal(_bits(ref.CoarseFrequencyCorrection(delay).run(x, ftags)),
                              _bits(oracle.CoarseFrequencyCorrection(delay).run(x, ftags)))
    # SyncwordWipeoff
    sw = np.where(np.asarray(SYNCWORD) != 0, -1.0, 1.0).astype(np.float32)
    idx = [0, 5, 63, 64, 200, 230, 264, 4000, 4064, n - 10]
    assert np.array_equal(_bits(ref.SyncwordWipeoff(sw).run(x, idx)), _bits(oracle.SyncwordWipeoff(sw).run(x, idx)))
    # CostasLoop, the three constellations
    ptags = [(0, 0.1), (6208, -1.3), (12416, 3.0), (30000, 0.0)]
    for name, code in (("PILOT", 0), ("BPSK", 1), ("qpsk", 2)):
        rl, ol = ref.CostasLoop(0.01, name), oracle.CostasLoop(0.01, code, oracle.TRIG_LIBM)
        assert np.array_equal(_bits(rl.run(x[:50000] * np.float32(0.5), ptags)),
                              _bits(ol.run(x[:50000] * np.float32(0.5), ptags)))
        assert rl.state() == ol.state()
    with pytest.raises(ValueError):
        ref.CostasLoop(0.01, "8PSK")


def test_symbol_filter_is_the_reference(oracle, ref, rx_params):
    """SymbolFilter driven by real detection tags (negative and positive time estimates, both special cases
    of :160-195 occur over 60 packets): symbols and re-indexed tags."""
    from gr4_packet_modem_b200.firdes import pfb_matched_filter_taps
    from gr4_packet_modem_b200.stimulus import packet_capture

    x, _ = packet_capture(1 << 18, seed=12, esn0_db=10.0, cfo=0.003, payload_bytes=60)
    o = oracle.SyncwordDetection(**rx_params, min_freq_bin=-4, max_freq_bin=4, fft_kind=oracle.FFT_RADIX2)
    oc, delayed, tags = o.run(x, chunk=65536, want_output=True)
    sf_taps = pfb_matched_filter_taps()
    rtags = [(t.index, ref.RefTag(index=0, freq=t.freq, amplitude=t.amplitude, phase=t.phase, noise_power=t.noise_power,
                                  esn0_db=t.esn0_db, time_est=t.time_est, freq_bin=t.freq_bin)) for t in tags]
    rsym, rot = ref.SymbolFilter(sf_taps, 32, 4, delay=44).run(delayed, rtags)
    osf = oracle.SymbolFilter(sf_taps, 32, 4, delay=44)
    by_index = {t.index: t for t in tags}
    cuts = sorted(by_index)
    pos, ys, oot, nout = 0, [], [], 0
    while pos < delayed.size:
        end = min([c for c in cuts if c > pos] + [delayed.size, pos + 50000])
        tag = None
        if pos in by_index:
            t = by_index[pos]
            tag = oracle.StreamTag()
            tag.has_syncword = True
            tag.amplitude, tag.time_est, tag.phase, tag.freq, tag.other = t.amplitude, t.time_est, t.phase, t.freq, 0
        c, ysym, ot = osf.process_bulk(delayed[pos:end], end - pos + 2, tag)
        oot += [(nout + q.index, q.phase) for q in ot]
        ys.append(ysym)
        nout += ysym.size
        pos += c
    osym = np.concatenate(ys)
    assert len(tags) > 50 and rsym.size == osym.size and np.array_equal(_bits(rsym), _bits(osym))
    assert [i for i, _ in rot] == [i for i, _ in oot]
    assert [np.float32(q.phase).tobytes() for _, q in rot] == [np.float32(p).tobytes() for _, p in oot]
    assert any(t.time_est < 0 for t in tags) and any(t.time_est > 0 for t in tags)


def test_detection_filter_state_machine_is_the_reference(oracle, ref):
    """SyncwordDetectionFilter driven in lock step through 4000 random processBulk calls: chunk sizes, output
    spans shorter than the input, syncword tags inside and outside packets, tags with other keys, parsed /
    invalid header messages and ignored_syncword messages arriving early, late or never."""
    rng = np.random.default_rng(99)
    r, o = ref.SyncwordDetectionFilter(4, 64, 128), oracle.SyncwordDetectionFilter(4, 64, 128)
    x = _noise(1 << 16, 1)
    in_packet_calls = forwarded = dropped = 0
    for step in range(4000):
        n = int(rng.integers(1, 3000))
        a = int(rng.integers(0, x.size - n))
        n_out = n if rng.random() < 0.7 else int(rng.integers(1, n + 1))
        kind = rng.random()
        rt = ot = None
        if kind < 0.35:
            rt, ot = ref.RefTag(), oracle.StreamTag()
            has_sw = rng.random() < 0.8
            other = int(rng.integers(1, 100)) if rng.random() < 0.3 or not has_sw else 0
            rt.no_syncword, rt.other = 0 if has_sw else 1, other
            ot.has_syncword, ot.other = has_sw, other
            for k, v in (("amplitude", 1.5), ("phase", -0.3), ("freq", 0.0123), ("time_est", 0.25)):
                setattr(rt, k, v)
                setattr(ot, k, v)
        hdr = None
        h = rng.random()
        if h < 0.25:
            hdr = ("parsed", int(rng.integers(1, 400)))
        elif h < 0.32:
            hdr = ("invalid",)
        n_ign = int(rng.random() < 0.1)
        rr = r.process_bulk(x[a:a + n], n_out, rt, hdr, n_ign)
        oo = o.process_bulk(x[a:a + n], n_out, ot, hdr, n_ign)
        assert rr[0] == oo[0] and rr[3:] == oo[3:], step
        assert np.array_equal(_bits(rr[1]), _bits(oo[1])), step
        assert (rr[2] is None) == (oo[2] is None), step
        if rr[2] is not None:
            assert (rr[2].no_syncword == 0) == bool(oo[2].has_syncword) and rr[2].other == oo[2].other, step
            forwarded += 1
        elif rt is not None:
            dropped += 1
        in_packet_calls += rr[5]
    assert in_packet_calls > 500 and forwarded > 300 and dropped > 100   # every branch was exercised
