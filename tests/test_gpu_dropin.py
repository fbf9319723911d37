"""The literal drop-in test: oracle/_ref/dropin_test (tests/cpp/test_dropin_vs_reference.cpp, built where the
reference tree exists by `make -C oracle ref`) runs ONE templated driver over the REFERENCE's block classes —
compiled unmodified from /root/reference against a stand-in GR4 runtime — and over this repository's B200
shells of the same blocks, configured by the same code, fed the same spans and tags.  Bars (north_star):
detection indices / counts / bins exact, |df| < 1e-5 rad/sample, |dphi| < 1e-3 rad, exact-arithmetic blocks
(delay line, SymbolFilter, SyncwordWipeoff, PfbArbResampler) bit for bit, NCO-based blocks rel-L2 < 1e-5."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "oracle", "_ref", "dropin_test")


def test_reference_classes_and_b200_shells_side_by_side(rx_params, tmp_path):
    if not os.path.exists(EXE):
        pytest.skip("oracle/_ref/dropin_test not built (it needs the reference tree at build time)")
    from gr4_packet_modem_b200.firdes import lowpass_prototype_taps, pfb_matched_filter_taps
    from gr4_packet_modem_b200.stimulus import packet_capture

    x, _ = packet_capture(400000, seed=71, esn0_db=12.0, cfo=0.004, payload_bytes=150)
    x.tofile(tmp_path / "x.cf32")
    rx_params["rrc_taps"].tofile(tmp_path / "rrc.f32")
    np.asarray(pfb_matched_filter_taps(), np.float32).tofile(tmp_path / "sf.f32")
    np.asarray(lowpass_prototype_taps(32, 40), np.float32).tofile(tmp_path / "fe.f32")
    r = subprocess.run([EXE, str(tmp_path / "x.cf32"), str(tmp_path / "rrc.f32"), str(tmp_path / "sf.f32"),
                        str(tmp_path / "fe.f32")], capture_output=True, text=True, check=True)
    res = {}
    for line in r.stdout.strip().splitlines():
        name, *kv = line.split()
        res[name] = {k: float(v) for k, v in (p.split("=") for p in kv)}
    sd = res["syncword_detection"]
    assert sd["tags"] > 40 and sd["consumed_equal"] == 1 and sd["delayed_bits_equal"] == 1
    assert sd["indices_equal"] == 1 and sd["bins_equal"] == 1 and sd["keys"] == 7
    assert sd["dfreq"] < 1e-5 and sd["dphase"] < 1e-3 and sd["dtime"] < 1e-3 and sd["damp"] < 1e-4
    assert res["coarse_frequency_correction"]["size_equal"] == 1 and res["coarse_frequency_correction"]["rel_l2"] < 1e-5
    sf = res["symbol_filter"]
    assert sf["symbols"] > 90000 and sf["bits_equal"] == 1 and sf["tags"] > 40 and sf["tags_equal"] == 1
    assert res["syncword_wipeoff"]["bits_equal"] == 1
    assert res["costas_loop"]["size_equal"] == 1 and res["costas_loop"]["rel_l2"] < 1e-5
    rs = res["pfb_arb_resampler"]
    assert rs["outputs"] > 399000 and rs["consumed_equal"] == 1 and rs["bits_equal"] == 1
