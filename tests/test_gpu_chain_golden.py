"""The GPU receiver chain against the committed fixture tests/golden/sync_chain_golden.npz (oracle outputs for a
seeded capture; generator tests/golden/make_chain_golden.py): SyncwordDetection -> [SyncwordDetectionFilter gate]
-> CoarseFrequencyCorrection -> SymbolFilter -> SyncwordWipeoff -> CostasLoop through the Python mirrors of the
C ABI, unfused and fused.  Bars (north_star): detection indices and counts exact; |df| < 1e-5 rad/sample,
|dphi| < 1e-3 rad; filter outputs relative L2 < 1e-5 (the two closed-form NCOs against the reference's float
recurrences; 1e-4 behind the Costas loop, whose feedback sees those differences)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _rel(a, b):
    return float(np.linalg.norm(a.astype(np.complex128) - b.astype(np.complex128)) / np.linalg.norm(b.astype(np.complex128)))


@pytest.mark.parametrize("fused", [False, True])
def test_gpu_chain_matches_golden(rx_params, fused):
    from gr4_packet_modem_b200 import (CoarseFrequencyCorrection, CostasLoop, SymbolFilter, SyncwordDetection,
                                       SyncwordWipeoff)
    from gr4_packet_modem_b200.blocks import STREAM_TAG_DTYPE
    from gr4_packet_modem_b200.firdes import SYNCWORD, pfb_matched_filter_taps

    g = np.load(os.path.join(GOLDEN, "sync_chain_golden.npz"))
    x, payload = g["capture"], int(g["payload_bytes"])
    rrc = rx_params["rrc_taps"]
    sd = SyncwordDetection(**rx_params, min_freq_bin=-4, max_freq_bin=4)
    consumed, delayed, tags = sd.run(x, chunk=1 << 16, want_output=True)
    assert consumed == int(g["consumed"])
    assert [t[1] for t in tags] == g["tag_index"].tolist()
    kv = [t[2] for t in tags]
    assert [d["syncword_freq_bin"] for d in kv] == g["tag_freq_bin"].tolist()
    assert np.max(np.abs(np.array([d["syncword_freq"] for d in kv]) - g["tag_freq"])) < 1e-5
    dphi = np.angle(np.exp(1j * (np.array([d["syncword_phase"] for d in kv], np.float64) - g["tag_phase"])))
    assert np.max(np.abs(dphi)) < 1e-3
    assert np.max(np.abs(np.array([d["syncword_time_est"] for d in kv]) - g["tag_time_est"])) < 1e-3
    assert np.allclose([d["syncword_amplitude"] for d in kv], g["tag_amplitude"], rtol=1e-4)
    # the gate of SyncwordDetectionFilter (PM/syncword_detection_filter.hpp:141-152)
    block = 4 * (128 + 64 - 16 + 4 * (payload + 4))
    kept, until = [], -1
    for _, idx, d in tags:
        if idx >= until:
            kept.append((idx, d))
            until = idx + block
    assert [i for i, _ in kept] == g["kept_index"].tolist()
    it = np.zeros(len(kept), STREAM_TAG_DTYPE)
    it["index"] = [i for i, _ in kept]
    it["has_syncword"] = 1
    for k in ("syncword_freq", "syncword_amplitude", "syncword_phase", "syncword_time_est"):
        it["sw"][k] = [d[k] for _, d in kept]
    cfc_delay = (rrc.size - 1) // 2 + 4
    sf_taps = pfb_matched_filter_taps()
    if fused:
        c, sym, ot = SymbolFilter(sf_taps, 32, 4, delay=rrc.size - 1, fused_cfc_delay=cfc_delay).process_bulk(delayed, it)
    else:
        corrected = CoarseFrequencyCorrection(cfc_delay).process_bulk(delayed, it)
        c, sym, ot = SymbolFilter(sf_taps, 32, 4, delay=rrc.size - 1).process_bulk(corrected, it)
    assert c == delayed.size and sym.size == g["symbols"].size
    assert ot["index"].tolist() == g["symbol_tag_index"].tolist()
    assert np.max(np.abs(np.angle(np.exp(1j * (ot["sw"]["syncword_phase"].astype(np.float64) - g["symbol_tag_phase"]))))) < 1e-3
    assert _rel(sym, g["symbols"]) < 1e-5
    sw = np.where(np.asarray(SYNCWORD) != 0, -1.0, 1.0).astype(np.float32)
    if fused:
        cl = CostasLoop(0.01, "BPSK")
        cl.fuse_wipeoff(sw)
        locked = cl.process_bulk(sym, ot)
    else:
        locked = CostasLoop(0.01, "BPSK").process_bulk(SyncwordWipeoff(sw).process_bulk(sym, ot), ot)
    assert _rel(locked, g["locked"]) < 1e-4


def test_gpu_front_end_matches_reference_golden():
    """tests/golden/frontend_golden.npz holds the outputs of the REFERENCE's PfbArbResampler and Rotator blocks:
    the GPU resampler reproduces them bit for bit, the fused front end (closed-form NCO) within the rotator
    tolerance."""
    from gr4_packet_modem_b200 import FrontEnd, PfbArbResampler

    g = np.load(os.path.join(GOLDEN, "frontend_golden.npz"))
    raw, taps, rate = g["raw"], g["taps"], float(g["rate"])
    c, y = PfbArbResampler(rate, taps, 32).process_bulk(raw)
    assert c == raw.size and np.array_equal(y.view(np.uint32), g["resampled"].view(np.uint32))
    c, z = FrontEnd(rate=rate, taps=taps, phase_incr=float(g["phase_incr"])).process_bulk(raw)
    assert z.size == g["rotated"].size and _rel(z, g["rotated"]) < 1e-5


def test_gpu_low_snr_matches_reference_golden(rx_params):
    """tests/golden/lowsnr_golden.npz: the tags of the REFERENCE's SyncwordDetection block at Es/N0 1 dB, CFO
    -0.07 rad/sample, K = 17, threshold 8.  GPU: same indices and bins exactly (streaming and bulk entry
    points), estimates within north_star's tolerances."""
    from gr4_packet_modem_b200 import SyncwordDetection

    g = np.load(os.path.join(GOLDEN, "lowsnr_golden.npz"))
    b = int(g["bins"])
    sd = SyncwordDetection(**rx_params, min_freq_bin=-b, max_freq_bin=b, power_threshold=float(g["power_threshold"]))
    consumed, recs, tags = sd.detect_host(g["capture"])
    assert consumed == int(g["consumed"]) and len(g["tag_index"]) > 10
    assert tags["index"].tolist() == g["tag_index"].tolist()
    assert tags["syncword_freq_bin"].tolist() == g["tag_freq_bin"].tolist()
    assert np.max(np.abs(tags["syncword_freq"] - g["tag_freq"])) < 1e-5
    assert np.max(np.abs(np.angle(np.exp(1j * (tags["syncword_phase"].astype(np.float64) - g["tag_phase"]))))) < 1e-3
    assert np.max(np.abs(tags["syncword_time_est"] - g["tag_time_est"])) < 1e-3
    assert np.max(np.abs(tags["syncword_esn0_db"] - g["tag_esn0_db"])) < 1e-2
    sd2 = SyncwordDetection(**rx_params, min_freq_bin=-b, max_freq_bin=b, power_threshold=float(g["power_threshold"]))
    pos, _, t2 = sd2.run(g["capture"], chunk=5000)
    assert [t[1] for t in t2] == [i for i in g["tag_index"].tolist() if i < pos]
