"""Regenerates tests/golden/sync_chain_golden.npz: a seeded 2^17-sample capture and what the REFERENCE's own
blocks (oracle/_ref/librefblocks.so, see oracle/ref_blocks.cpp) — and, bit for bit the same, the oracle's
restated blocks — make of it along the receiver chain (PM/packet_receiver.hpp:76-125, 195-218):
SyncwordDetection (independent radix-2 FFT arithmetic) -> CoarseFrequencyCorrection -> SymbolFilter ->
SyncwordWipeoff -> CostasLoop (libm trig = the reference's arithmetic).

The reference's runtime and FFTW cannot be built here; its block sources can, against a stand-in
runtime with the oracle's radix-2 FFT.  These vectors carry their outputs to the GPU box and pin the oracle —
and through it the GPU path — across rounds: tests/test_oracle_golden.py::test_chain_golden_pins_oracle (CPU) and
tests/test_gpu_chain_golden.py (GPU).  Run in the build container:  python tests/golden/make_chain_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def oracle_chain(po, x, sf_taps, rrc, syncword_bits, bpsk, payload_bytes):
    """-> dict of the chain's observable outputs for capture x."""
    o = po.SyncwordDetection(rrc, syncword_bits, bpsk, -4, 4, 768, 9.5, fft_kind=po.FFT_RADIX2)
    consumed, delayed, tags = o.run(x, chunk=1 << 16, want_output=True)
    # SyncwordDetectionFilter: the syncword tag that opens a packet closes the gate for the packet's length
    block = 4 * (128 + 64 - 16 + 4 * (payload_bytes + 4))   # PM/syncword_detection_filter.hpp:141-152
    kept, until = [], -1
    for t in tags:
        if t.index >= until:
            kept.append(t)
            until = t.index + block
    cfc_delay = (rrc.size - 1) // 2 + 4                       # PM/packet_receiver.hpp:94-95
    corrected = po.CoarseFrequencyCorrection(cfc_delay).run(delayed, [(t.index, t.freq) for t in kept])
    sf = po.SymbolFilter(sf_taps, 32, 4, delay=rrc.size - 1)
    by_index = {t.index: t for t in kept}
    cuts = sorted(by_index)
    pos, ys, otags, nout = 0, [], [], 0
    while pos < corrected.size:
        end = min([c for c in cuts if c > pos] + [corrected.size])
        tag = None
        if pos in by_index:
            t = by_index[pos]
            tag = po.StreamTag()
            tag.has_syncword = True
            tag.amplitude, tag.time_est, tag.phase, tag.freq, tag.other = t.amplitude, t.time_est, t.phase, t.freq, 0
        c, ysym, ot = sf.process_bulk(corrected[pos:end], end - pos + 2, tag)
        otags += [(nout + q.index, q.phase) for q in ot]
        ys.append(ysym)
        nout += ysym.size
        pos += c
    sym = np.concatenate(ys)
    sw = np.where(np.asarray(syncword_bits) != 0, -1.0, 1.0).astype(np.float32)
    wiped = po.SyncwordWipeoff(sw).run(sym, [i for i, _ in otags])
    locked = po.CostasLoop(0.01, 1, po.TRIG_LIBM).run(wiped, otags)
    return dict(consumed=np.int64(consumed),
                tag_index=np.array([t.index for t in tags], np.int64),
                tag_freq=np.array([t.freq for t in tags], np.float64),
                tag_phase=np.array([t.phase for t in tags], np.float32),
                tag_time_est=np.array([t.time_est for t in tags], np.float32),
                tag_amplitude=np.array([t.amplitude for t in tags], np.float32),
                tag_freq_bin=np.array([t.freq_bin for t in tags], np.int32),
                kept_index=np.array([t.index for t in kept], np.int64),
                symbol_tag_index=np.array([i for i, _ in otags], np.int64),
                symbol_tag_phase=np.array([p for _, p in otags], np.float32),
                symbols=sym, locked=locked)


def reference_chain(rb, x, sf_taps, rrc, syncword_bits, bpsk, payload_bytes):
    """The same chain through the REFERENCE's own blocks (oracle/_ref/librefblocks.so: the reference headers
    compiled unmodified against a stand-in runtime, FFT = the oracle's radix-2)."""
    consumed, delayed, tags = rb.SyncwordDetection(rrc, syncword_bits, bpsk, -4, 4, 768, 9.5).run(x, chunk=1 << 16)
    block = 4 * (128 + 64 - 16 + 4 * (payload_bytes + 4))
    kept, until = [], -1
    for t in tags:
        if t.index >= until:
            kept.append(t)
            until = t.index + block
    cfc_delay = (rrc.size - 1) // 2 + 4
    corrected = rb.CoarseFrequencyCorrection(cfc_delay).run(delayed, [(t.index, t.freq) for t in kept])
    sym, otags = rb.SymbolFilter(sf_taps, 32, 4, delay=rrc.size - 1).run(corrected, [(t.index, t) for t in kept], chunk=1 << 30)
    sw = np.where(np.asarray(syncword_bits) != 0, -1.0, 1.0).astype(np.float32)
    wiped = rb.SyncwordWipeoff(sw).run(sym, [i for i, _ in otags])
    locked = rb.CostasLoop(0.01, "BPSK").run(wiped, [(i, q.phase) for i, q in otags])
    return dict(consumed=np.int64(consumed),
                tag_index=np.array([t.index for t in tags], np.int64),
                tag_freq=np.array([t.freq for t in tags], np.float64),
                tag_phase=np.array([t.phase for t in tags], np.float32),
                tag_time_est=np.array([t.time_est for t in tags], np.float32),
                tag_amplitude=np.array([t.amplitude for t in tags], np.float32),
                tag_freq_bin=np.array([t.freq_bin for t in tags], np.int32),
                kept_index=np.array([t.index for t in kept], np.int64),
                symbol_tag_index=np.array([i for i, _ in otags], np.int64),
                symbol_tag_phase=np.array([q.phase for _, q in otags], np.float32),
                symbols=sym, locked=locked)


def settings():
    from gr4_packet_modem_b200.firdes import BPSK, SYNCWORD, pfb_matched_filter_taps, unit_energy_rrc

    return unit_energy_rrc(), SYNCWORD, BPSK, np.asarray(pfb_matched_filter_taps(), np.float32)


def main():
    from gr4_packet_modem_b200.stimulus import packet_capture
    from oracle import pyoracle as po

    po.build(ref=False)
    payload = 200
    x, starts = packet_capture(1 << 17, seed=2026, esn0_db=8.0, cfo=0.012, payload_bytes=payload, noise_seed=17)
    rrc, sw, bpsk, sf_taps = settings()
    out = oracle_chain(po, x, sf_taps, rrc, sw, bpsk, payload)
    from oracle import refblocks as rb

    source = "oracle restatement (reference tree absent)"
    if rb.available():
        # the fixture is what the REFERENCE's own block code produces; the oracle must reproduce it bit for bit
        ref = reference_chain(rb, x, sf_taps, rrc, sw, bpsk, payload)
        for k, v in ref.items():
            assert np.atleast_1d(v).tobytes() == np.atleast_1d(out[k]).tobytes(), k
        out = ref
        source = "reference blocks (oracle/_ref/librefblocks.so: PM/*.hpp unmodified, stand-in runtime, radix-2 FFT)"
    np.savez_compressed(os.path.join(HERE, "sync_chain_golden.npz"), capture=x, true_starts=starts,
                        payload_bytes=np.int64(payload), source=np.array(source), **out)
    print("source:", source)
    print("tags", out["tag_index"].size, "kept", out["kept_index"].size, "symbols", out["symbols"].size)


if __name__ == "__main__":
    main()


def make_frontend_golden():
    """tests/golden/frontend_golden.npz: a raw 2^15-sample stream through the REFERENCE's PfbArbResampler
    (rate 1 + 1.2 ppm, PM/pfb_arb_taps.hpp taps) and Rotator (0.005 rad/sample)."""
    from gr4_packet_modem_b200.firdes import lowpass_prototype_taps
    from gr4_packet_modem_b200.stimulus import packet_capture
    from oracle import pyoracle as po
    from oracle import refblocks as rb

    po.build(ref=True)
    assert rb.available(), "needs /root/reference (oracle/_ref/librefblocks.so)"
    raw, _ = packet_capture(1 << 15, seed=77, esn0_db=12.0, cfo=0.0, payload_bytes=100, noise_seed=5)
    rate = float(np.float32(1.0) + np.float32(1e-6) * np.float32(1.2))
    taps = np.asarray(lowpass_prototype_taps(32, 40), np.float32)
    consumed, resampled = rb.PfbArbResampler(rate, taps, 32).run(raw)
    rotated = rb.rotator(resampled, 0.005)
    oc, oy = po.PfbArbResampler(rate, taps, 32, use_double=False).process_bulk(raw, raw.size + 100)
    assert oc == consumed and oy.tobytes() == resampled.tobytes()
    assert po.rotator(resampled, 0.005).tobytes() == rotated.tobytes()
    np.savez_compressed(os.path.join(HERE, "frontend_golden.npz"), raw=raw, rate=np.float32(rate), taps=taps,
                        phase_incr=np.float32(0.005), resampled=resampled, rotated=rotated,
                        source=np.array("reference blocks (oracle/_ref/librefblocks.so)"))
    print("frontend golden:", resampled.size, "outputs")


if __name__ == "__main__":
    make_frontend_golden()


def make_lowsnr_golden():
    """tests/golden/lowsnr_golden.npz (BASELINE configs[3] in miniature): Es/N0 1 dB, CFO -0.07 rad/sample,
    K = 17 hypotheses, power_threshold 8 — what the REFERENCE's SyncwordDetection block tags."""
    from gr4_packet_modem_b200.stimulus import packet_capture
    from oracle import pyoracle as po
    from oracle import refblocks as rb

    po.build(ref=True)
    assert rb.available(), "needs /root/reference (oracle/_ref/librefblocks.so)"
    rrc, sw, bpsk, _ = settings()
    x, starts = packet_capture(1 << 16, seed=404, esn0_db=1.0, cfo=-0.07, payload_bytes=40, noise_seed=9)
    consumed, delayed, tags = rb.SyncwordDetection(rrc, sw, bpsk, -8, 8, 768, 8.0).run(x, chunk=1 << 16)
    oc, od, ot = po.SyncwordDetection(rrc, sw, bpsk, -8, 8, 768, 8.0, fft_kind=po.FFT_RADIX2).run(x, chunk=1 << 16,
                                                                                               want_output=True)
    assert oc == consumed and od.tobytes() == delayed.tobytes() and [t.index for t in ot] == [t.index for t in tags]
    np.savez_compressed(os.path.join(HERE, "lowsnr_golden.npz"), capture=x, true_starts=starts,
                        consumed=np.int64(consumed), bins=np.int64(8), power_threshold=np.float32(8.0),
                        tag_index=np.array([t.index for t in tags], np.int64),
                        tag_freq=np.array([t.freq for t in tags], np.float64),
                        tag_phase=np.array([t.phase for t in tags], np.float32),
                        tag_time_est=np.array([t.time_est for t in tags], np.float32),
                        tag_freq_bin=np.array([t.freq_bin for t in tags], np.int32),
                        tag_esn0_db=np.array([t.esn0_db for t in tags], np.float32),
                        source=np.array("reference blocks (oracle/_ref/librefblocks.so)"))
    print("low-SNR golden:", len(tags), "tags of", len(starts), "frames")


if __name__ == "__main__":
    make_lowsnr_golden()
