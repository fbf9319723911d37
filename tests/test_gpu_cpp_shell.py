"""GPU test of the C++ block shell (gr4_packet_modem_b200/blocks/syncword_detection_b200.hpp): the
GR4-shaped processBulk loop in C++ gives the same tags, at the same indices, with the same values,
and the same delayed output stream as the oracle's restated block."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_block_shell_matches_oracle(oracle, rx_params, tmp_path):
    from gr4_packet_modem_b200.stimulus import packet_capture

    exe = tmp_path / "test_block_shell"
    libdir = os.path.join(ROOT, "gr4_packet_modem_b200")
    subprocess.run(["g++", "-std=c++20", "-O2", "-o", str(exe), os.path.join(ROOT, "tests/cpp/test_block_shell.cpp"),
                    f"-L{libdir}", "-lb200sync", f"-Wl,-rpath,{libdir}"], check=True)
    x, _ = packet_capture(500000, seed=21, esn0_db=10.0, cfo=0.004, payload_bytes=150)
    x.tofile(tmp_path / "cap.cf32")
    rx_params["rrc_taps"].tofile(tmp_path / "rrc.f32")
    chunk = 65536
    r = subprocess.run([str(exe), str(tmp_path / "cap.cf32"), str(tmp_path / "rrc.f32"), "-4", "4", "9.5", str(chunk)],
                       capture_output=True, text=True, check=True)
    lines = r.stdout.strip().splitlines()
    assert lines[0] == "error min_freq_bin is greater than max_freq_bin"  # PM/syncword_detection.hpp:146
    tags = [l.split() for l in lines if l.startswith("tag ")]
    consumed = int(lines[-1].split()[1])
    checksum = int(lines[-1].split()[3])

    o = oracle.SyncwordDetection(**rx_params, min_freq_bin=-4, max_freq_bin=4, power_threshold=9.5,
                                 fft_kind=oracle.FFT_MIRROR)
    oc, oout, otags = o.run(x, chunk=chunk, want_output=True)
    assert consumed == oc
    assert [int(t[1]) for t in tags] == [t.index for t in otags] and len(tags) > 5
    for t, ot in zip(tags, otags):
        kv = dict(p.split("=") for p in t[2:])
        assert float(kv["syncword_freq"]) == ot.freq
        assert np.float32(float(kv["syncword_amplitude"])) == np.float32(ot.amplitude)
        assert np.float32(float(kv["syncword_phase"])) == np.float32(ot.phase)
        assert int(kv["syncword_freq_bin"]) == ot.freq_bin
        assert np.float32(float(kv["syncword_time_est"])) == np.float32(ot.time_est)
        assert np.float32(float(kv["syncword_esn0_db"])) == np.float32(ot.esn0_db)
        assert np.float32(float(kv["syncword_noise_power"])) == np.float32(ot.noise_power)
    cs = 0
    w = oout.view(np.uint32).reshape(-1, 2).astype(np.uint64)
    for a, b in w:
        cs = (cs * 1099511628211 + int(a) + 31 * int(b)) & 0xFFFFFFFFFFFFFFFF
    assert cs == checksum
