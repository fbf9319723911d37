"""GPU parity tests for the on-GPU stimulus (csrc/stimulus.cu): frames ->
InterpolatingFirFilter (PM/interpolating_fir_filter.hpp:93-99) -> Rotator (PM/rotator.hpp:56-65) ->
+ gaussian NoiseSource (PM/noise_source.hpp:74-78).

* pulse shaping: BIT-EXACT against the oracle's restated InterpolatingFirFilter on the same symbols;
* rotation: bit-exact against a host restatement of the kernel's closed form (phase reduced in double,
  oracle.mirror_sincosf), and within the rotator tolerance of the reference's float recurrence;
* noise: the reference's generator is sequential (std::mt19937), the kernel's is a hash of the sample
  index, so parity is the reference's own QA property (test/qa_noise_source.cpp:40-44: power within 1 %)
  plus zero mean, circularity and whiteness;
* index purity: any window generated on its own == the same window of a longer run, bit for bit
  (what lets every GPU generate its own time shard)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _gen(stim, n, first=0):
    import torch

    return stim.generate(n, torch.device("cuda", 0), first).cpu().numpy()


def test_pulse_shaping_bit_exact(oracle):
    from gr4_packet_modem_b200.stimulus import DeviceStimulus

    stim = DeviceStimulus(seed=5, esn0_db=None, cfo=0.0, payload_bytes=100)
    n = 200000
    got = _gen(stim, n)
    syms = stim.symbols(0, n // 4)
    want = oracle.interpolating_fir(syms, stim.taps, 4)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    # frame structure: syncword symbols first in every frame, QPSK elsewhere, nothing before symbol 0
    assert stim.frame_len == 64 + 128 + 104 * 4
    assert np.array_equal(syms[:64].real, stim.sync) and np.all(syms[:64].imag == 0)
    assert np.array_equal(syms[stim.frame_len:stim.frame_len + 64].real, stim.sync)
    assert np.allclose(np.abs(syms[64:stim.frame_len]), 1.0, atol=1e-6)
    # burst mode: gap symbols are zeros (apps/packet_transceiver.cpp burst mode: idle between packets)
    burst = DeviceStimulus(seed=5, esn0_db=None, cfo=0.0, payload_bytes=100, gap_symbols=300)
    sb = burst.symbols(0, 3000)
    assert np.all(sb[608:908] == 0) and np.all(np.abs(sb[908 + 64:908 + 608]) > 0.9)
    gb = _gen(burst, 12000)
    assert np.array_equal(gb.view(np.uint32), oracle.interpolating_fir(sb, burst.taps, 4).view(np.uint32))
    # other interpolation factors / tap counts (ragged last polyphase branch)
    from gr4_packet_modem_b200.firdes import root_raised_cosine

    taps5 = root_raised_cosine(1.0, 5.0, 1.0, 0.35, 43)
    s5 = DeviceStimulus(seed=9, esn0_db=None, cfo=0.0, payload_bytes=10, sps=5, taps=taps5)
    g5 = _gen(s5, 5 * 4000)
    assert np.array_equal(g5.view(np.uint32), oracle.interpolating_fir(s5.symbols(0, 4000), taps5, 5).view(np.uint32))


def test_rotation_closed_form_and_reference_tolerance(oracle):
    from gr4_packet_modem_b200.stimulus import DeviceStimulus

    n = 1 << 17
    cfo = 0.005
    base = _gen(DeviceStimulus(seed=3, esn0_db=None, cfo=0.0), n)
    got = _gen(DeviceStimulus(seed=3, esn0_db=None, cfo=cfo), n)
    theta = float(np.float32(cfo))
    ph = np.arange(n, dtype=np.float64) * theta
    ph = ph - 6.283185307179586476925 * np.rint(ph * 0.15915494309189533577)
    s, c = oracle.mirror_sincosf(ph.astype(np.float32))
    want = np.empty(n, np.complex64)
    want.real = base.real * c - base.imag * s
    want.imag = base.real * s + base.imag * c
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    ref = oracle.rotator(base, cfo)   # the reference's float recurrence
    for lo in range(0, n, 1 << 15):
        w = slice(lo, lo + (1 << 15))
        assert np.linalg.norm(got[w] - ref[w]) / np.linalg.norm(ref[w]) < 1e-5 * (1 + lo / (1 << 15))


def test_noise_statistics_reference_qa():
    """test/qa_noise_source.cpp:14-46: power of 'gaussian' noise within 1 % of amplitude^2."""
    from gr4_packet_modem_b200.stimulus import DeviceStimulus, TX_POWER

    n = 1 << 22
    for esn0 in (0.0, 20.0):
        sig = _gen(DeviceStimulus(seed=11, esn0_db=None, cfo=0.005), n)
        stim = DeviceStimulus(seed=11, esn0_db=esn0, cfo=0.005)
        noise = (_gen(stim, n).astype(np.complex128) - sig.astype(np.complex128))
        a2 = stim.noise_amplitude ** 2
        assert abs(a2 - TX_POWER * 4 * 10 ** (-0.1 * esn0)) < 1e-6 * a2 + 1e-9
        power = np.mean(np.abs(noise) ** 2)
        assert abs(power - a2) < 1e-2 * a2
        assert abs(np.mean(noise)) < 4 * np.sqrt(a2 / n)
        assert abs(np.mean(noise.real ** 2) - np.mean(noise.imag ** 2)) < 1e-2 * a2     # circular
        assert abs(np.mean(noise.real * noise.imag)) < 4 * (a2 / 2) / np.sqrt(n)
        for lag in (1, 2, 4, 7):                                                        # white
            assert abs(np.mean(noise[lag:] * np.conj(noise[:-lag]))) < 5 * a2 / np.sqrt(n)
        kurt = np.mean(noise.real ** 4) / np.mean(noise.real ** 2) ** 2                 # gaussian: 3
        assert abs(kurt - 3.0) < 0.05
    # different seeds: different symbols and noise
    a = _gen(DeviceStimulus(seed=1, esn0_db=10.0), 4096)
    b = _gen(DeviceStimulus(seed=2, esn0_db=10.0), 4096)
    assert np.mean(np.abs(a - b)) > 0.1


def test_index_purity_windows():
    from gr4_packet_modem_b200.stimulus import DeviceStimulus

    stim = DeviceStimulus(seed=7, esn0_db=10.0, cfo=0.005)
    n = 1 << 20
    whole = _gen(stim, n)
    for first, cnt in ((0, 1), (1, 7), (3, 4096), (24832 - 5, 300), (n - 1001, 1001), (123457, 65537)):
        w = _gen(stim, cnt, first)
        assert np.array_equal(w.view(np.uint32), whole[first:first + cnt].view(np.uint32)), (first, cnt)
    # far into the stream (beyond 2^32 samples): windows agree with each other
    far = (1 << 33) + 12345
    a = _gen(stim, 50000, far)
    b = _gen(stim, 20000, far + 30000)
    assert np.array_equal(a[30000:].view(np.uint32), b.view(np.uint32))


def test_generated_capture_is_detected(oracle, rx_params):
    """The receiver's view: one syncword per frame at the expected sample, GPU == mirror oracle."""
    from gr4_packet_modem_b200 import SyncwordDetection
    from gr4_packet_modem_b200.stimulus import DeviceStimulus

    stim = DeviceStimulus(seed=1, esn0_db=20.0, cfo=0.005, payload_bytes=1500)
    n = 1 << 20
    x = _gen(stim, n)
    sd = SyncwordDetection(**rx_params, min_freq_bin=-4, max_freq_bin=4)
    consumed, recs, tags = sd.detect_host(x)
    o = oracle.SyncwordDetection(**rx_params, min_freq_bin=-4, max_freq_bin=4, fft_kind=oracle.FFT_MIRROR)
    oc, _, otags = o.run(x, chunk=1 << 16)
    assert oc == consumed and recs["index"].tolist() == [t.index - 1537 for t in otags]
    frame = stim.frame_len * 4
    expect = [k * frame for k in range(n // frame + 1) if k * frame + 1537 + 297 < consumed]
    assert recs["index"].tolist() == expect
    assert np.all(np.abs(tags["syncword_freq"] - 0.005) < 1e-3)   # quadratic bin interpolation, bin spacing 0.0106
