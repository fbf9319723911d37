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


def _build(tmp_path, src, name):
    exe = tmp_path / name
    libdir = os.path.join(ROOT, "gr4_packet_modem_b200")
    subprocess.run(["g++", "-std=c++20", "-O2", "-o", str(exe), os.path.join(ROOT, "tests/cpp", src),
                    f"-L{libdir}", "-lb200sync", f"-Wl,-rpath,{libdir}"], check=True)
    return exe


def test_cpp_rx_chain_shells(oracle, rx_params, tmp_path):
    """The six block shells chained the way packet_receiver.hpp wires them (resampler -> rotator ->
    SyncwordDetection -> SyncwordDetectionFilter -> CoarseFrequencyCorrection -> SymbolFilter), driven from
    C++ with GR4-style chunks:
    every stage's stream and tags against the oracle's restated blocks (bit-exact except behind the
    rotator, whose parity is a tolerance) and against the Python mirrors (bit-exact everywhere)."""
    from gr4_packet_modem_b200 import CoarseFrequencyCorrection, FrontEnd, SyncwordDetection
    from gr4_packet_modem_b200.blocks import STREAM_TAG_DTYPE
    from gr4_packet_modem_b200.firdes import lowpass_prototype_taps, pfb_matched_filter_taps
    from gr4_packet_modem_b200.stimulus import packet_capture

    exe = _build(tmp_path, "test_rx_chain_shells.cpp", "test_rx_chain_shells")
    payload = 150
    raw, _ = packet_capture(400000, seed=33, esn0_db=15.0, cfo=0.0, payload_bytes=payload)
    rate = float(np.float32(1.0) + np.float32(1e-6) * np.float32(1.2))
    fe_taps, sf_taps, rrc = lowpass_prototype_taps(32, 40), pfb_matched_filter_taps(), rx_params["rrc_taps"]
    raw.tofile(tmp_path / "raw.cf32")
    rrc.tofile(tmp_path / "rrc.f32")
    np.asarray(fe_taps, np.float32).tofile(tmp_path / "fe.f32")
    np.asarray(sf_taps, np.float32).tofile(tmp_path / "sf.f32")
    prefix = str(tmp_path / "out")
    r = subprocess.run([str(exe), str(tmp_path / "raw.cf32"), str(tmp_path / "rrc.f32"), str(tmp_path / "fe.f32"),
                        str(tmp_path / "sf.f32"), repr(rate), "0.005", str(payload), "50000", prefix],
                       capture_output=True, text=True, check=True)
    lines = r.stdout.strip().splitlines()
    assert lines[0] == "error filter_size cannot be 0"  # PM/pfb_arb_resampler.hpp:70-72
    assert "fused_equals_pair 1" in lines
    assert "fused_cfc_equals_pair 1" in lines
    load = lambda s: np.fromfile(prefix + f"_{s}.cf32", np.complex64)  # noqa: E731
    y, z, d, f, g, sym = (load(s) for s in ("resampled", "rotated", "delayed", "filtered", "corrected", "symbols"))

    def tags_of(stage):
        out = []
        for l in lines:
            t = l.split()
            if t[:2] == ["tag", stage]:
                out.append((int(t[2]), dict(p.split("=") for p in t[3:])))
        return out

    # resampler: bit-exact vs the oracle block; rotator: the Python mirror exactly, the oracle within tolerance
    oc, oy = oracle.PfbArbResampler(rate, fe_taps, 32, use_double=False).process_bulk(raw, raw.size + 1000)
    assert np.array_equal(y.view(np.uint32), oy.view(np.uint32))
    _, zp = FrontEnd(rate=rate, taps=fe_taps, phase_incr=0.005).process_bulk(raw)
    assert np.array_equal(z.view(np.uint32), zp.view(np.uint32))
    oz = oracle.rotator(oy, 0.005)
    for lo in range(0, z.size - (1 << 15), 1 << 15):
        w = slice(lo, lo + (1 << 15))
        assert np.linalg.norm(z[w] - oz[w]) / np.linalg.norm(oz[w]) < 1e-5 * (1 + lo / (1 << 15))

    # SyncwordDetection on the C++ rotator output: oracle (mirror arithmetic) bit for bit
    o = oracle.SyncwordDetection(**rx_params, min_freq_bin=-4, max_freq_bin=4, fft_kind=oracle.FFT_MIRROR)
    oc, od, otags = o.run(z, chunk=1 << 16, want_output=True)
    assert d.size == oc and np.array_equal(d.view(np.uint32), od.view(np.uint32))
    t_sd = tags_of("syncword_detection")
    assert [i for i, _ in t_sd] == [t.index for t in otags] and len(t_sd) > 20
    for (_, kv), ot in zip(t_sd, otags):
        assert float(kv["syncword_freq"]) == ot.freq
        assert np.float32(float(kv["syncword_phase"])) == np.float32(ot.phase)
        assert np.float32(float(kv["syncword_time_est"])) == np.float32(ot.time_est)
        assert np.float32(float(kv["syncword_amplitude"])) == np.float32(ot.amplitude)

    # SyncwordDetectionFilter: pass-through stream; tags dropped exactly while a packet is open
    assert np.array_equal(f.view(np.uint32), d[:f.size].view(np.uint32)) and f.size == d.size
    t_sdf = tags_of("syncword_detection_filter")
    block = 4 * (128 + 64 - 16 + 4 * (payload + 4))  # PM/syncword_detection_filter.hpp:141-152
    expect, until = [], -1
    for i, kv in t_sd:
        if i >= until:
            expect.append(i)
            until = i + block
    assert [i for i, _ in t_sdf] == expect and 5 < len(expect) <= len(t_sd)
    assert all(len(kv) == 7 for _, kv in t_sdf)

    # CoarseFrequencyCorrection (delay 26): tags forwarded unchanged; the Python mirror exactly (the C++ shell
    # sees one tag per chunk, the mirror all of them in one span), the oracle's recurrence within tolerance
    t_cfc = tags_of("coarse_frequency_correction")
    assert t_cfc == t_sdf and g.size == f.size
    cfc_delay = (rrc.size - 1) // 2 + 4
    it = np.zeros(len(t_sdf), STREAM_TAG_DTYPE)
    it["index"] = [i for i, _ in t_sdf]
    it["has_syncword"] = 1
    it["sw"]["syncword_freq"] = [float(kv["syncword_freq"]) for _, kv in t_sdf]
    gp = CoarseFrequencyCorrection(cfc_delay).process_bulk(f, it)
    assert np.array_equal(g.view(np.uint32), gp.view(np.uint32))
    og = oracle.CoarseFrequencyCorrection(cfc_delay).run(f, [(i, float(kv["syncword_freq"])) for i, kv in t_sdf])
    assert np.linalg.norm(g - og) / np.linalg.norm(og) < 1e-5
    assert not np.array_equal(g, f)
    f = g  # what SymbolFilter was fed

    # SymbolFilter: the oracle block driven by the same tags, chunks cut at tags (GR/Block.hpp:1501-1508)
    osf = oracle.SymbolFilter(sf_taps, 32, 4, delay=rrc.size - 1)
    by_index = dict(t_sdf)
    pos, ys, ot_all, nout = 0, [], [], 0
    cuts = sorted(by_index)
    while pos < f.size:
        nxt = min([c for c in cuts if c > pos] + [f.size])
        end = min(nxt, pos + 50000)
        tag = None
        if pos in by_index:
            kv = by_index[pos]
            tag = oracle.StreamTag()
            tag.has_syncword = True
            tag.amplitude, tag.time_est = float(kv["syncword_amplitude"]), float(kv["syncword_time_est"])
            tag.phase, tag.freq, tag.other = float(kv["syncword_phase"]), float(kv["syncword_freq"]), 0
        c, ysym, ot = osf.process_bulk(f[pos:end], end - pos + 2, tag)
        assert c == end - pos
        ot_all += [(nout + t.index, t.phase) for t in ot]
        ys.append(ysym)
        nout += ysym.size
        pos += c
    osym = np.concatenate(ys)
    assert sym.size == osym.size and np.array_equal(sym.view(np.uint32), osym.view(np.uint32))
    t_sf = tags_of("symbol_filter")
    assert [i for i, _ in t_sf] == [i for i, _ in ot_all] and len(t_sf) >= len(t_sdf) - 1
    for (_, kv), (_, ph) in zip(t_sf, ot_all):
        assert np.float32(float(kv["syncword_phase"])) == np.float32(ph) and len(kv) == 7

    # SyncwordWipeoff + CostasLoop behind the SymbolFilter (PM/packet_receiver.hpp:117-125, 203-214): the
    # oracle blocks driven by the SymbolFilter's tags, mirror trig bit for bit, libm within tolerance; in a
    # locked packet the wiped syncword comes out as 64 symbols of one sign on the real axis
    from gr4_packet_modem_b200.firdes import SYNCWORD

    assert "fused_wipeoff_equals_pair 1" in lines
    assert any(l.startswith("costas_error unknown constellation") for l in lines)
    wiped, locked = load("wiped"), load("locked")
    sw = np.where(np.asarray(SYNCWORD) != 0, -1.0, 1.0).astype(np.float32)
    idx = [i for i, _ in t_sf]
    ow = oracle.SyncwordWipeoff(sw).run(sym, idx)
    assert np.array_equal(wiped.view(np.uint32), ow.view(np.uint32))
    ptags = [(i, float(kv["syncword_phase"])) for i, kv in t_sf]
    ol = oracle.CostasLoop(0.01, 1, oracle.TRIG_MIRROR).run(ow, ptags)
    assert np.array_equal(locked.view(np.uint32), ol.view(np.uint32))
    ol_ref = oracle.CostasLoop(0.01, 1, oracle.TRIG_LIBM).run(ow, ptags)
    assert np.linalg.norm(locked - ol_ref) / np.linalg.norm(ol_ref) < 1e-5
    good = 0
    for i in idx:
        s = locked[i + 8:i + 64]
        if s.size == 56 and np.all(s.real > 0.3) and np.max(np.abs(s.imag)) < 0.6:
            good += 1
    assert good >= len(idx) - 2, (good, len(idx))
