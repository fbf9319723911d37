"""GPU parity against the REFERENCE's own block code, live: oracle/_ref/librefblocks.so (PM/*.hpp compiled
unmodified in the build container against the stand-in runtime, FFT = radix-2 stand-in for FFTW) travels to
the GPU box as a built library, so the CUDA path and the reference's classes can be fed the same seeded
captures here.  north_star's bars: detection indices and counts exact; |df| < 1e-5 rad/sample, |dphi| < 1e-3
rad; filter outputs bit for bit where the arithmetic is exact (delay line, SymbolFilter, PfbArbResampler,
SyncwordWipeoff), relative L2 < 1e-5 where an NCO or sin/cos is involved.  Skipped when the library is absent."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ref():
    from oracle import refblocks

    if not refblocks.available():
        pytest.skip("oracle/_ref/librefblocks.so not built")
    return refblocks


def _rel(a, b):
    return float(np.linalg.norm(a.astype(np.complex128) - b.astype(np.complex128)) / np.linalg.norm(b.astype(np.complex128)))


@pytest.mark.parametrize("bins,esn0,thr,T,cfo", [(4, 20.0, 9.5, 768, 0.005), (0, 15.0, 9.5, 768, 0.001),
                                                 (8, 1.0, 8.0, 768, -0.06), (16, 0.0, 7.0, 768, 0.14),
                                                 (2, 6.0, 7.0, 100, 0.01)])
def test_detection_vs_reference_block(ref, rx_params, bins, esn0, thr, T, cfo):
    from gr4_packet_modem_b200 import SyncwordDetection
    from gr4_packet_modem_b200.stimulus import packet_capture

    x, _ = packet_capture(1 << 19, seed=900 + bins, esn0_db=esn0, cfo=cfo, payload_bytes=100)
    rc, rout, rtags = ref.SyncwordDetection(**rx_params, min_freq_bin=-bins, max_freq_bin=bins, time_threshold=T,
                                            power_threshold=thr).run(x, chunk=65536)
    sd = SyncwordDetection(**rx_params, min_freq_bin=-bins, max_freq_bin=bins, time_threshold=T, power_threshold=thr)
    consumed, recs, tags = sd.detect_host(x)
    assert consumed == rc and len(rtags) >= 20
    assert tags["index"].tolist() == [t.index for t in rtags]
    assert tags["syncword_freq_bin"].tolist() == [t.freq_bin for t in rtags]
    assert np.max(np.abs(tags["syncword_freq"] - np.array([t.freq for t in rtags]))) < 1e-5
    dphi = np.angle(np.exp(1j * (tags["syncword_phase"].astype(np.float64) - np.array([t.phase for t in rtags]))))
    assert np.max(np.abs(dphi)) < 1e-3
    assert np.max(np.abs(tags["syncword_time_est"] - np.array([t.time_est for t in rtags]))) < 1e-3
    assert np.allclose(tags["syncword_amplitude"], [t.amplitude for t in rtags], rtol=1e-4)
    assert np.max(np.abs(tags["syncword_esn0_db"] - np.array([t.esn0_db for t in rtags]))) < 2e-2
    # the block contract through processBulk-sized chunks: delayed output bit for bit, same tags
    sd2 = SyncwordDetection(**rx_params, min_freq_bin=-bins, max_freq_bin=bins, time_threshold=T, power_threshold=thr)
    pos, out, t2 = sd2.run(x, chunk=65536, want_output=True)
    assert pos == rc and np.array_equal(out.view(np.uint32), rout.view(np.uint32))
    assert [t[1] for t in t2] == [t.index for t in rtags]


@pytest.mark.parametrize("fft_size,bins", [(1024, 2), (4096, 4)])
def test_other_fft_sizes_vs_reference_block_bit_for_bit(ref, rx_params, fft_size, bins):
    """fft_size != 2048: the GPU runs the radix-2 arithmetic that stands in for FFTW under the reference's class here,
    so every tag value and the delayed output equal the reference block's BIT FOR BIT (no tolerance)."""
    from gr4_packet_modem_b200 import SyncwordDetection
    from gr4_packet_modem_b200.stimulus import packet_capture

    x, _ = packet_capture(1 << 19, seed=77, esn0_db=3.0, cfo=0.007, payload_bytes=100)
    kw = dict(min_freq_bin=-bins, max_freq_bin=bins, time_threshold=768, power_threshold=8.0, fft_size=fft_size)
    rc, rout, rtags = ref.SyncwordDetection(**rx_params, **kw).run(x, chunk=65536)
    pos, out, tags = SyncwordDetection(**rx_params, **kw).run(x, chunk=65536, want_output=True)
    assert pos == rc and len(rtags) >= 20
    assert np.array_equal(out.view(np.uint32), rout.view(np.uint32))
    assert [t[1] for t in tags] == [t.index for t in rtags]
    for (_, _, m), t in zip(tags, rtags):
        assert m["syncword_freq"] == t.freq and m["syncword_freq_bin"] == t.freq_bin
        for k in ("amplitude", "phase", "noise_power", "esn0_db", "time_est"):
            assert np.float32(m["syncword_" + k]).tobytes() == np.float32(getattr(t, k)).tobytes(), k


def test_filters_vs_reference_blocks(ref, rx_params):
    from gr4_packet_modem_b200 import (CoarseFrequencyCorrection, CostasLoop, FrontEnd, PfbArbResampler, SymbolFilter,
                                       SyncwordWipeoff)
    from gr4_packet_modem_b200.blocks import STREAM_TAG_DTYPE
    from gr4_packet_modem_b200.firdes import SYNCWORD, lowpass_prototype_taps, pfb_matched_filter_taps
    from gr4_packet_modem_b200.stimulus import packet_capture

    raw, _ = packet_capture(1 << 18, seed=66, esn0_db=10.0, cfo=0.0, payload_bytes=80)
    rate = float(np.float32(1.0) + np.float32(1e-6) * np.float32(1.2))
    fe_taps = np.asarray(lowpass_prototype_taps(32, 40), np.float32)
    rc, ry = ref.PfbArbResampler(rate, fe_taps, 32).run(raw)
    c, y = PfbArbResampler(rate, fe_taps, 32).process_bulk(raw)
    assert c == rc and np.array_equal(y.view(np.uint32), ry.view(np.uint32))                 # exact arithmetic
    rz = ref.rotator(ry, 0.005)
    _, z = FrontEnd(rate=rate, taps=fe_taps, phase_incr=0.005).process_bulk(raw)
    for lo in range(0, z.size - (1 << 15), 1 << 15):                                          # closed-form NCO
        w = slice(lo, lo + (1 << 15))
        assert _rel(z[w], rz[w]) < 1e-5 * (1 + lo / (1 << 15))
    # detection on the reference's rotated stream, then the filters behind it, each fed the reference's data
    rcons, rdel, rtags = ref.SyncwordDetection(**rx_params, min_freq_bin=-4, max_freq_bin=4).run(rz, chunk=65536)
    assert len(rtags) > 40
    it = np.zeros(len(rtags), STREAM_TAG_DTYPE)
    it["index"] = [t.index for t in rtags]
    it["has_syncword"] = 1
    for k, a in (("syncword_freq", "freq"), ("syncword_amplitude", "amplitude"), ("syncword_phase", "phase"),
                 ("syncword_time_est", "time_est")):
        it["sw"][k] = [getattr(t, a) for t in rtags]
    rcor = ref.CoarseFrequencyCorrection(26).run(rdel, [(t.index, t.freq) for t in rtags])
    gcor = CoarseFrequencyCorrection(26).process_bulk(rdel, it)
    assert _rel(gcor, rcor) < 1e-5
    sf_taps = pfb_matched_filter_taps()
    rsym, rot = ref.SymbolFilter(sf_taps, 32, 4, delay=44).run(rcor, [(t.index, t) for t in rtags], chunk=1 << 30)
    c, gsym, got = SymbolFilter(sf_taps, 32, 4, delay=44).process_bulk(rcor, it)
    assert c == rcor.size and np.array_equal(gsym.view(np.uint32), rsym.view(np.uint32))      # exact arithmetic
    assert got["index"].tolist() == [i for i, _ in rot]
    assert [np.float32(p).tobytes() for p in got["sw"]["syncword_phase"]] == [np.float32(q.phase).tobytes() for _, q in rot]
    sw = np.where(np.asarray(SYNCWORD) != 0, -1.0, 1.0).astype(np.float32)
    rwo = ref.SyncwordWipeoff(sw).run(rsym, [i for i, _ in rot])
    gwo = SyncwordWipeoff(sw).process_bulk(rsym, got)
    assert np.array_equal(gwo.view(np.uint32), rwo.view(np.uint32))                           # exact arithmetic
    for name in ("BPSK", "QPSK", "PILOT"):
        rl = ref.CostasLoop(0.01, name).run(rwo, [(i, q.phase) for i, q in rot])
        gl = CostasLoop(0.01, name).process_bulk(rwo, got)
        assert _rel(gl, rl) < 1e-5, name
