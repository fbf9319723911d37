"""GPU parity tests for SyncwordDetection: CUDA path (through the C ABI) vs the CPU oracle.

Bars (BASELINE.json north_star):
  * detected sample indices and detection counts: bit-exact;
  * vs the oracle's MIRROR arithmetic: every float of the metric and of the records bit-exact;
  * vs the oracle's independent RADIX-2 arithmetic: |df| < 1e-5 rad/sample, |dphi| < 1e-3 rad,
    amplitude / time estimate within 1e-4.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL_FREQ = 1e-5   # rad/sample
TOL_PHASE = 1e-3  # rad
TOL_AMP = 1e-4
TOL_TIME = 1e-3


def _gpu(rx_params, **kw):
    from gr4_packet_modem_b200 import SyncwordDetection

    return SyncwordDetection(**rx_params, **kw)


def _qa_stimulus(oracle, rx_params, freq_error, nsym=200000, seed=1234):
    """The stimulus of test/qa_syncword_detection.cpp:24-76 with a fixed seed."""
    from gr4_packet_modem_b200.firdes import SYNCWORD

    rng = np.random.default_rng(seed)
    sym = rng.integers(0, 2, nsym).astype(np.uint8)
    locs = [100, 1000, 1250, 10000, 13721, 43124, 58000, 127018]
    for l in locs:
        sym[l:l + 64] = SYNCWORD
    x = oracle.interpolating_fir(rx_params["constellation"][sym], rx_params["rrc_taps"], 4)
    return oracle.rotator(x, freq_error), locs


def _wrap(d):
    return (d + np.pi) % (2 * np.pi) - np.pi


@pytest.mark.parametrize("freq_error", [0.0, 0.005, 0.015, -0.005, -0.015])
def test_reference_qa_assertions_streaming(oracle, rx_params, freq_error):
    """test/qa_syncword_detection.cpp:99-146 against the GPU block, driven in 65536-item chunks."""
    x, locs = _qa_stimulus(oracle, rx_params, freq_error)
    sd = _gpu(rx_params, min_freq_bin=-4, max_freq_bin=4, power_threshold=20.0)
    consumed, out, tags = sd.run(x, want_output=True)
    delay = 2 * 768 + 1
    assert consumed <= x.size and consumed + 2048 > x.size
    assert np.all(out[:delay] == 0)
    assert np.array_equal(out[delay:], x[:consumed - delay])
    assert len(tags) == len(locs)
    for (off, idx, m), loc in zip(tags, locs):
        assert idx == delay + 4 * loc
        assert 0.95 < m["syncword_amplitude"] < 1.01
        assert m["syncword_esn0_db"] >= 30.0
        assert abs(m["syncword_freq"] - freq_error) < 5e-4
        assert m["syncword_freq_bin"] == round(freq_error / (np.pi / 297))
        assert m["syncword_noise_power"] < 5e-4
        if freq_error == 0.0:
            assert abs(m["syncword_phase"]) < 1e-6
        assert abs(m["syncword_time_est"]) < 0.05


@pytest.mark.parametrize("esn0_db,thr,bins", [(20.0, 9.5, 4), (0.0, 9.5, 4), (3.0, 6.0, 1), (20.0, 9.5, 0),
                                              # BASELINE configs[3]: Es/N0 0 dB, K = 17 / 33, threshold sweep ends
                                              (0.0, 6.0, 8), (0.0, 20.0, 8), (0.0, 12.0, 16)])
def test_offline_bit_exact_vs_mirror_oracle(oracle, rx_params, esn0_db, thr, bins):
    """Whole pipeline bit-for-bit: metric, detection set, raw records, estimates."""
    from gr4_packet_modem_b200.stimulus import packet_capture

    n = 1 << 20
    x, _ = packet_capture(n, seed=3, esn0_db=esn0_db, cfo=0.005, payload_bytes=200)
    sd = _gpu(rx_params, min_freq_bin=-bins, max_freq_bin=bins, power_threshold=thr)
    consumed, recs, tags = sd.detect_host(x)
    o = oracle.SyncwordDetection(**rx_params, min_freq_bin=-bins, max_freq_bin=bins, power_threshold=thr,
                                 fft_kind=oracle.FFT_MIRROR, record_metric=True)
    oc, _, otags = o.run(x, chunk=1 << 20)
    assert consumed == oc
    zp, _ = o.metric(oc)
    assert np.array_equal(sd.metric(consumed).view(np.uint32), zp.view(np.uint32)), "metric not bit-exact"
    assert (recs["index"] + sd.delay).tolist() == [t.index for t in otags]
    assert len(recs) > 0
    for r, t, ot in zip(recs, tags, otags):
        for a, b in [(r["corr_re"], ot.corr_re), (r["corr_im"], ot.corr_im), (r["pow"], ot.pow),
                     (r["pow_prev"], ot.pow_prev), (r["pow_next"], ot.pow_next), (r["noise_power"], ot.noise_power)]:
            assert np.float32(a).view(np.uint32) == np.float32(b).view(np.uint32)
        assert r["freq_bin"] == ot.freq_bin
        if -bins < r["freq_bin"] < bins:
            assert np.float32(r["pow_left"]) == np.float32(ot.pow_left)
            assert np.float32(r["pow_right"]) == np.float32(ot.pow_right)
        assert t["syncword_freq"] == ot.freq
        assert np.float32(t["syncword_amplitude"]) == np.float32(ot.amplitude)
        assert np.float32(t["syncword_phase"]) == np.float32(ot.phase)
        assert np.float32(t["syncword_time_est"]) == np.float32(ot.time_est)
        assert np.float32(t["syncword_esn0_db"]) == np.float32(ot.esn0_db)


def test_offline_vs_independent_oracle(oracle, rx_params):
    """Config 1/2 signal model (20 dB, CFO 0.005): indices exact, estimates within tolerance, against
    the oracle arithmetic that shares nothing with the GPU FFT."""
    from gr4_packet_modem_b200.stimulus import packet_capture

    n = 1 << 21
    x, starts = packet_capture(n, seed=1, esn0_db=20.0, cfo=0.005, payload_bytes=1500)
    sd = _gpu(rx_params, min_freq_bin=-4, max_freq_bin=4, power_threshold=9.5)
    consumed, recs, tags = sd.detect_host(x)
    o = oracle.SyncwordDetection(**rx_params, min_freq_bin=-4, max_freq_bin=4, power_threshold=9.5,
                                 fft_kind=oracle.FFT_RADIX2, record_metric=True)
    oc, _, otags = o.run(x, chunk=1 << 20)
    assert consumed == oc
    assert tags["index"].tolist() == [t.index for t in otags]
    # every true syncword far enough from the end is found, at exactly its start sample
    found = set(recs["index"].tolist())
    for s in starts:
        if s + sd.delay < consumed:
            assert int(s) in found
    zp, _ = o.metric(oc)
    z = sd.metric(consumed)
    assert np.linalg.norm(z - zp) / np.linalg.norm(zp) < 1e-5
    for t, ot in zip(tags, otags):
        assert abs(t["syncword_freq"] - ot.freq) < TOL_FREQ
        assert abs(_wrap(t["syncword_phase"] - ot.phase)) < TOL_PHASE
        assert abs(t["syncword_amplitude"] - ot.amplitude) < TOL_AMP
        assert abs(t["syncword_time_est"] - ot.time_est) < TOL_TIME
        assert t["syncword_freq_bin"] == ot.freq_bin
        assert abs(t["syncword_esn0_db"] - ot.esn0_db) < 1e-2


@pytest.mark.parametrize("chunk", [2048, 5000, 65536, 300000])
def test_streaming_equals_offline_and_oracle(oracle, rx_params, chunk):
    """processBulk in arbitrary chunkings gives the same tags at the same indices (state carried across
    calls: _items_consumed, history, pending detections), same delayed output."""
    from gr4_packet_modem_b200.stimulus import packet_capture

    n = 600000
    x, _ = packet_capture(n, seed=5, esn0_db=6.0, cfo=-0.012, payload_bytes=100)
    sd = _gpu(rx_params, min_freq_bin=-2, max_freq_bin=2, power_threshold=9.5)
    consumed, out, tags = sd.run(x, chunk=chunk, want_output=True)
    o = oracle.SyncwordDetection(**rx_params, min_freq_bin=-2, max_freq_bin=2, power_threshold=9.5,
                                 fft_kind=oracle.FFT_MIRROR)
    oc, oout, otags = o.run(x, chunk=chunk, want_output=True)
    assert consumed == oc
    assert np.array_equal(out, oout)
    assert [i for _, i, _ in tags] == [t.index for t in otags]
    assert len(tags) > 3
    for (_, _, m), ot in zip(tags, otags):
        assert m["syncword_freq"] == ot.freq
        assert np.float32(m["syncword_phase"]) == np.float32(ot.phase)


def test_edge_inputs(oracle, rx_params):
    """Empty / short / all-zero / exactly-one-block inputs (reference guards :215-227; the
    benchmark's NullSource input, benchmarks/README.md:50-53)."""
    sd = _gpu(rx_params, min_freq_bin=-4, max_freq_bin=4)
    status, c, out, tags = sd.process_bulk(np.zeros(100, np.complex64))
    assert status == "INSUFFICIENT_INPUT_ITEMS" and c == 0 and tags == []
    status, c, out, tags = sd.process_bulk(np.zeros(2048, np.complex64))
    assert status == "OK" and c == sd.stride == 1752 and tags == []
    sd.start()
    z = np.zeros(1 << 18, np.complex64)
    consumed, recs, tags = sd.detect_host(z)
    assert consumed == ((z.size - 2048) // 1752 + 1) * 1752 and len(recs) == 0
    assert np.all(sd.metric(consumed) == 0.0)
    # a constant (DC) input: heavy ties in the metric; must agree with the oracle exactly
    dc = np.full(1 << 16, 0.25 + 0.5j, np.complex64)
    consumed, recs, tags = sd.detect_host(dc)
    o = oracle.SyncwordDetection(**rx_params, min_freq_bin=-4, max_freq_bin=4, fft_kind=oracle.FFT_MIRROR)
    oc, _, otags = o.run(dc, chunk=1 << 16)
    assert consumed == oc and (recs["index"] + sd.delay).tolist() == [t.index for t in otags]


def test_settings_errors(rx_params):
    """start() throws like the reference (PM/syncword_detection.hpp:145-152)."""
    from gr4_packet_modem_b200.blocks import B200SyncError

    with pytest.raises(B200SyncError, match="min_freq_bin is greater than max_freq_bin"):
        _gpu(rx_params, min_freq_bin=2, max_freq_bin=1)
    with pytest.raises(B200SyncError, match="fft_size too small"):
        _gpu(dict(rx_params, syncword=np.zeros(600, np.uint8)))


def test_settings_the_gpu_path_refuses(rx_params):
    """Settings the reference accepts and this build does not implement fail LOUDLY at start() with
    B200SYNC_EUNSUPPORTED — never a silent fallback (the reference: any power-of-two fft_size,
    PM/syncword_detection.hpp:133 / ALG/fourier/fftw.hpp:182-184 — here 64 ... 8192; any time_threshold, :140; any bin range)."""
    import torch
    from gr4_packet_modem_b200.blocks import B200SyncError

    with pytest.raises(B200SyncError, match="Input data must have 2\\^N samples"):      # the reference throws too (fftw.hpp:182-184)
        _gpu(rx_params, fft_size=3000)
    with pytest.raises(B200SyncError, match="fft_size outside \\[64, 8192\\]"):
        _gpu(rx_params, fft_size=16384)
    with pytest.raises(B200SyncError, match="too many frequency hypotheses"):
        _gpu(rx_params, min_freq_bin=-301, max_freq_bin=301)
    with pytest.raises(B200SyncError, match="time_threshold > 4095"):
        _gpu(rx_params, time_threshold=4096)
    # time_threshold in (1023, 4095] works for one stream, but not for the entry points built on the chain tables
    sd = _gpu(rx_params, time_threshold=2000)
    x = torch.zeros(1 << 16, dtype=torch.complex64, device="cuda:0")
    with pytest.raises(B200SyncError, match="time-sharded operation needs time_threshold <= 1023"):
        sd.shard_phase1(x.data_ptr(), 0, 1 << 16, 0, 10, 36)
    with pytest.raises(B200SyncError, match="batched channel mode needs time_threshold <= 1023"):
        sd.detect_channels_device(x.data_ptr(), 2, 1 << 15, 1 << 15)


@pytest.mark.parametrize("T", [0, 1, 5, 31, 32, 33, 100, 500, 1023, 1024, 2047, 4095])
def test_time_threshold_sweep_bit_exact(oracle, rx_params, T):
    """Both flags kernels (generic T < 32, group-scan T >= 32) and the chain kernels for every
    piece length, against the oracle's sequential state machine, on a noisy capture with many
    marginal threshold decisions."""
    from gr4_packet_modem_b200.stimulus import packet_capture

    x, _ = packet_capture(1 << 18, seed=9, esn0_db=2.0, cfo=0.02, payload_bytes=60)
    sd = _gpu(rx_params, min_freq_bin=-1, max_freq_bin=1, power_threshold=4.0, time_threshold=T)
    consumed, recs, tags = sd.detect_host(x)
    o = oracle.SyncwordDetection(**rx_params, min_freq_bin=-1, max_freq_bin=1, power_threshold=4.0,
                                 time_threshold=T, fft_kind=oracle.FFT_MIRROR)
    oc, _, otags = o.run(x, chunk=1 << 18, )
    assert consumed == oc
    assert tags["index"].tolist() == [t.index for t in otags]
    if T >= 31:  # tiny windows of a 4x-oversampled correlation never pass the threshold test
        assert len(tags) > (10 if T <= 1023 else 3)
    if T > 1023:  # beyond the parallel chain kernels: also the streaming path, in odd chunks
        sd2 = _gpu(rx_params, min_freq_bin=-1, max_freq_bin=1, power_threshold=4.0, time_threshold=T)
        c2, _, t2 = sd2.run(x, chunk=50000)
        assert [i for _, i, _ in t2] == [t.index for t in otags if t.index < c2]
    # and streaming in odd chunks
    sd.start()
    c2, _, t2 = sd.run(x, chunk=7001)
    o2 = oracle.SyncwordDetection(**rx_params, min_freq_bin=-1, max_freq_bin=1, power_threshold=4.0,
                                  time_threshold=T, fft_kind=oracle.FFT_MIRROR)
    oc2, _, ot2 = o2.run(x, chunk=7001)
    assert c2 == oc2 and [i for _, i, _ in t2] == [t.index for t in ot2]


@pytest.mark.parametrize("world", [2, 3, 5])
def test_time_sharded_equals_single_run(rx_params, world):
    """SURVEY §8e on one GPU: the capture cut into `world` time shards (own context each, halo blocks,
    chain tables composed on the host) gives exactly the records of the single-context run."""
    import torch

    from gr4_packet_modem_b200.sharding import entry_offsets, plan_shards
    from gr4_packet_modem_b200.stimulus import packet_capture

    n = 1 << 21
    x, _ = packet_capture(n, seed=31, esn0_db=4.0, cfo=0.01, payload_bytes=80)
    kw = dict(min_freq_bin=-2, max_freq_bin=2, power_threshold=7.0)
    ref_c, ref_recs, _ = _gpu(rx_params, **kw).detect_host(x)
    shards = plan_shards(n, world, 2048, 1752, 768)
    ctxs, tables, bufs = [], [], []
    for s in shards:
        sd = _gpu(rx_params, **kw)
        seg = torch.from_numpy(x[s.first_sample:s.first_sample + s.n_samples].copy()).cuda()
        tables.append(sd.shard_phase1(seg.data_ptr(), s.first_sample, s.n_samples, s.first_block, s.n_blocks,
                                      s.total_blocks))
        ctxs.append(sd)
        bufs.append(seg)
    got = []
    for sd, j in zip(ctxs, entry_offsets(tables)):
        recs, _ = sd.shard_phase2(j, n // 769 + 2)
        got.append(recs)
    got = np.concatenate(got)
    assert len(got) == len(ref_recs) > 20
    assert np.array_equal(got.view(np.uint8), ref_recs.view(np.uint8))
    # the same with the shards' samples in HOST memory (b200sync_sd_shard_phase1_host: the correlator
    # chases the H2D copies); contexts are reused, which also covers start-over after a finished shard run
    tables_h = [sd.shard_phase1_host(x[s.first_sample:s.first_sample + s.n_samples], s.first_sample, s.first_block,
                                     s.n_blocks, s.total_blocks) for sd, s in zip(ctxs, shards)]
    assert all(np.array_equal(a, b) for a, b in zip(tables, tables_h))
    got_h = np.concatenate([sd.shard_phase2(j, n // 769 + 2)[0] for sd, j in zip(ctxs, entry_offsets(tables_h))])
    assert np.array_equal(got_h.view(np.uint8), ref_recs.view(np.uint8))


@pytest.mark.parametrize("n_channels,stride_pad", [(1, 0), (5, 0), (7, 1000)])
def test_batched_channels_equal_single_runs(oracle, rx_params, n_channels, stride_pad):
    """BASELINE config 5, channel mode: n independent channels in one call give, per channel, exactly the
    records of a single-stream run of that channel — and the first channel matches the oracle."""
    import torch

    from gr4_packet_modem_b200.stimulus import packet_capture

    n = 300000
    stride = n + stride_pad
    chans = [packet_capture(n, seed=40 + c, noise_seed=90 + c, esn0_db=(3.0 if c % 2 else 15.0), cfo=0.004 * (c - 2),
                            payload_bytes=60 + 40 * c, signal=(c != 3))[0] for c in range(n_channels)]
    buf = np.zeros(n_channels * stride, np.complex64)
    for c, x in enumerate(chans):
        buf[c * stride:c * stride + n] = x
    d = torch.from_numpy(buf).cuda()
    kw = dict(min_freq_bin=-4, max_freq_bin=4)
    sd = _gpu(rx_params, **kw)
    consumed, per = sd.detect_channels_device(d.data_ptr(), n_channels, n, stride,
                                              torch.cuda.current_stream().cuda_stream)
    single = _gpu(rx_params, **kw)
    for c, x in enumerate(chans):
        c1, r1, _ = single.detect_host(x)
        assert c1 == consumed
        assert np.array_equal(per[c].view(np.uint8), r1.view(np.uint8)), c
    o = oracle.SyncwordDetection(**rx_params, **kw, fft_kind=oracle.FFT_RADIX2)
    oc, _, otags = o.run(chans[0], chunk=1 << 18)
    assert oc == consumed and (per[0]["index"] + sd.delay).tolist() == [t.index for t in otags]
    assert sum(len(p) for p in per) > 5
    # a second call on the same context (buffers reused) gives the same answer
    _, per2 = sd.detect_channels_device(d.data_ptr(), n_channels, n, stride)
    assert all(np.array_equal(a.view(np.uint8), b.view(np.uint8)) for a, b in zip(per, per2))


def test_detect_file_equals_detect_host(oracle, rx_params, tmp_path):
    """Raw capture ingestion: a cf32 file in FileSource<c64>'s format
    (PM/file_source.hpp:47-53), larger than the staging ring (3 x 32 MiB) so that slots are reused, gives
    the records of the same samples handed over in host memory; offsets, a trailing partial item, short
    and missing files behave like fread / the reference's error."""
    from gr4_packet_modem_b200 import SyncwordDetection
    from gr4_packet_modem_b200.blocks import B200SyncError
    from gr4_packet_modem_b200.stimulus import packet_capture

    n = (1 << 24) + 12345   # 128 MiB + a ragged tail: 5 pieces through 3 slots
    base, _ = packet_capture(1 << 22, seed=21, esn0_db=12.0, cfo=0.004, payload_bytes=300)
    x = np.tile(base, n // base.size + 1)[:n]
    path = tmp_path / "capture.cf32"
    with open(path, "wb") as f:
        x.tofile(f)
        f.write(b"\x01\x02\x03")   # partial trailing item: ignored like fread's item count
    sd = SyncwordDetection(**rx_params, min_freq_bin=-4, max_freq_bin=4)
    c0, r0, t0 = sd.detect_host(x)
    c1, r1, t1, items = sd.detect_file(path)
    assert items == n and c1 == c0 and len(r1) == len(r0) > 1000
    assert r1.tobytes() == r0.tobytes() and t1.tobytes() == t0.tobytes()
    # window of the file: indices relative to first_item
    first, cnt = 1000003, 3000000
    c2, r2, _ = sd.detect_host(x[first:first + cnt])
    c3, r3, _, items = sd.detect_file(path, first_item=first, max_items=cnt)
    assert items == cnt and c3 == c2 and r3.tobytes() == r2.tobytes()
    # the mirror oracle on a prefix of the file
    m = 1 << 20
    o = oracle.SyncwordDetection(**rx_params, min_freq_bin=-4, max_freq_bin=4, fft_kind=oracle.FFT_MIRROR)
    oc, _, otags = o.run(x[:m], chunk=1 << 16)
    c4, r4, t4, _ = sd.detect_file(path, max_items=m)
    assert c4 == oc and t4["index"].tolist() == [t.index for t in otags]
    # fewer items than one FFT block: nothing consumed (:215-227); beyond the end: empty
    assert sd.detect_file(path, max_items=2047)[0] == 0
    assert sd.detect_file(path, first_item=n + 10)[3] == 0
    with pytest.raises(B200SyncError, match="error opening file"):
        sd.detect_file(tmp_path / "missing.cf32")


def test_full_size_capture_round_trip(rx_params):
    """BASELINE configs[1] at its full size (2^30 samples, 8 GiB resident) through size-independent
    properties: generator -> detector round trip (every frame of the synthetic stream is found exactly at
    its first sample, nothing else is), records sorted and unique, and the same capture cut into 3 time
    shards (own contexts, halo blocks, composed chain tables) gives byte-identical records."""
    import torch

    from gr4_packet_modem_b200 import SyncwordDetection
    from gr4_packet_modem_b200.sharding import entry_offsets, plan_shards
    from gr4_packet_modem_b200.stimulus import DeviceStimulus

    free, _ = torch.cuda.mem_get_info()
    if free < 40 << 30:
        pytest.skip("needs 40 GiB of free device memory")
    n = 1 << 30
    dev = torch.device("cuda", 0)
    stim = DeviceStimulus(seed=1, esn0_db=20.0, cfo=0.005)
    x = stim.generate(n, dev)
    st = torch.cuda.current_stream().cuda_stream
    sd = SyncwordDetection(**rx_params, min_freq_bin=-4, max_freq_bin=4)
    consumed, recs, tags = sd.detect_device(x.data_ptr(), n, st)
    assert consumed == ((n - 2048) // 1752 + 1) * 1752
    idx = recs["index"].astype(np.int64)
    frame = stim.frame_len * 4
    expect = np.arange(0, consumed, frame)
    expect = expect[expect + sd.delay < consumed]
    assert np.array_equal(idx, expect)                      # sorted, unique, exactly the frame starts
    assert np.all(np.abs(tags["syncword_freq"] - 0.005) < 1e-3)
    # (the peak on sample 0 has no predecessor: its interpolation sees the zero-initialised history, :321-323)
    assert np.all(np.abs(tags["syncword_time_est"][1:]) < 0.1)
    del sd
    shards = plan_shards(n, 3, 2048, 1752, 768)
    ctxs, tables = [], []
    for s in shards:
        c = SyncwordDetection(**rx_params, min_freq_bin=-4, max_freq_bin=4)
        tables.append(c.shard_phase1(x.data_ptr() + 8 * s.first_sample, s.first_sample, s.n_samples, s.first_block,
                                     s.n_blocks, s.total_blocks, st))
        ctxs.append(c)
    got = np.concatenate([c.shard_phase2(j, n // 769 + 2)[0] for c, j in zip(ctxs, entry_offsets(tables))])
    assert np.array_equal(got.view(np.uint8), recs.view(np.uint8))


def test_contexts_are_independent_across_threads(rx_params):
    """SURVEY §8(b) Tier 2: a context is not thread-safe, but distinct contexts (one CUDA stream each) may be
    driven concurrently from different threads — what GR4 does with one block instance per worker
    (GR/Scheduler.hpp:387-398).  Four threads stream four different captures through processBulk-sized
    chunks and bulk calls at the same time; every result equals the single-threaded run."""
    import threading

    from gr4_packet_modem_b200 import CostasLoop, SyncwordDetection
    from gr4_packet_modem_b200.stimulus import packet_capture

    caps = [packet_capture(1 << 19, seed=50 + i, esn0_db=6.0, cfo=0.002 * i, payload_bytes=100)[0] for i in range(4)]

    def work(i, out):
        sd = SyncwordDetection(**rx_params, min_freq_bin=-4, max_freq_bin=4)
        pos, delayed, tags = sd.run(caps[i], chunk=65536, want_output=True)
        c, recs, _ = sd.detect_host(caps[i])
        cl = CostasLoop(0.01, "QPSK")
        locked = np.concatenate([cl.process_bulk(delayed[a:a + 50000]) for a in range(0, delayed.size, 50000)])
        out[i] = (pos, delayed.copy(), [t[1] for t in tags], c, recs.copy(), locked)

    seq, par = {}, {}
    for i in range(4):
        work(i, seq)
    ts = [threading.Thread(target=work, args=(i, par)) for i in range(4)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    for i in range(4):
        a, b = seq[i], par[i]
        assert a[0] == b[0] and a[2] == b[2] and a[3] == b[3] and len(a[2]) > 20
        assert np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32))
        assert a[4].tobytes() == b[4].tobytes()
        assert np.array_equal(a[5].view(np.uint32), b[5].view(np.uint32))


@pytest.mark.parametrize("n", [1 << 20, (1 << 20) + 777, 3 * 1752 + 2048])
def test_device_delayed_output_stays_inside_n_items(oracle, rx_params, n):
    """The fused delayed output of detect_device (PM/syncword_detection.hpp:318-319) is published up to the
    consumed count only: an exactly n-item device buffer followed by a guard area must come back with the
    guard untouched, and equal to the host-span path's output.  (Round-1 advisor finding: the last block used
    to store up to `delay` items past the publish limit.)"""
    import torch
    from gr4_packet_modem_b200.stimulus import packet_capture

    x, _ = packet_capture(n, seed=5, esn0_db=20.0, cfo=0.005, payload_bytes=200)
    sd = _gpu(rx_params, min_freq_bin=-4, max_freq_bin=4, power_threshold=9.5)
    dev = torch.device("cuda:0")
    d_in = torch.from_numpy(x.view(np.float32)).to(dev)
    guard = 4096
    d_out = torch.full((2 * (n + guard),), 777.0, dtype=torch.float32, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    consumed, recs, tags = sd.detect_device(d_in.data_ptr(), n, st, d_out.data_ptr())
    torch.cuda.synchronize()
    out = d_out.cpu().numpy().view(np.complex64)
    assert np.all(out[consumed:].view(np.float32) == 777.0), "stored past the publish limit"
    sd2 = _gpu(rx_params, min_freq_bin=-4, max_freq_bin=4, power_threshold=9.5)
    c2, out2, tags2 = sd2.run(x, chunk=1 << 18, want_output=True)
    assert c2 == consumed
    assert np.array_equal(out[:consumed].view(np.uint32), out2.view(np.uint32))


@pytest.mark.parametrize("T", [33, 100, 769, 1000])
def test_streaming_time_threshold_not_multiple_of_32(oracle, rx_params, T):
    """Streaming keeps exactly 2T+2 samples of metric history; the fast flags kernel rounds its window origin
    down to a group boundary, which used to reach below the buffer for T % 32 != 0 (round-1 advisor finding).
    Decisions must equal the offline run and the oracle's sequential loop for such T."""
    from gr4_packet_modem_b200.stimulus import packet_capture

    n = 1 << 19
    x, _ = packet_capture(n, seed=9, esn0_db=6.0, cfo=0.004, payload_bytes=100)
    kw = dict(min_freq_bin=-2, max_freq_bin=2, power_threshold=8.0, time_threshold=T)
    sd = _gpu(rx_params, **kw)
    c_s, _, tags_s = sd.run(x, chunk=40000)
    sd2 = _gpu(rx_params, **kw)
    c_o, recs_o, _ = sd2.detect_host(x)
    o = oracle.SyncwordDetection(**rx_params, **kw, fft_kind=oracle.FFT_MIRROR)
    oc, _, otags = o.run(x, chunk=1 << 19)
    want = [t.index for t in otags]
    assert len(want) > 0
    assert c_o == oc
    assert (recs_o["index"] + sd2.delay).tolist() == want
    assert [idx for _, idx, _ in tags_s] == [i for i in want if i < c_s]


def test_process_bulk_more_tags_than_the_buffer_holds(oracle, rx_params):
    """b200sync_sd_process cannot fail after it has consumed input: tags beyond max_tags stay queued and are
    drained in the same chunk (round-1 advisor finding: they used to be lost with ENOMEM)."""
    rng = np.random.default_rng(11)
    n = 1 << 17
    x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    kw = dict(min_freq_bin=0, max_freq_bin=0, power_threshold=1.0001, time_threshold=3)
    sd = _gpu(rx_params, **kw)
    status, c, _, tags = sd.process_bulk(x, want_output=False, max_tags=7)
    assert status == "OK" and c > 0
    sd2 = _gpu(rx_params, **kw)
    _, c2, _, tags2 = sd2.process_bulk(x, want_output=False, max_tags=1 << 16)
    assert c2 == c and len(tags2) > 7
    assert [t[1] for t in tags] == [t[1] for t in tags2]
    assert all(0 <= off < c for off, _, _ in tags)


def _assert_bit_exact_vs_oracle(sd, recs, tags, o, oc, otags, bins):
    zp, _ = o.metric(oc)
    assert np.array_equal(sd.metric(oc).view(np.uint32), zp.view(np.uint32)), "metric not bit-exact"
    assert (recs["index"] + sd.delay).tolist() == [t.index for t in otags]
    for r, t, ot in zip(recs, tags, otags):
        for a, b in [(r["corr_re"], ot.corr_re), (r["corr_im"], ot.corr_im), (r["pow"], ot.pow),
                     (r["pow_prev"], ot.pow_prev), (r["pow_next"], ot.pow_next), (r["noise_power"], ot.noise_power)]:
            assert np.float32(a).view(np.uint32) == np.float32(b).view(np.uint32)
        assert r["freq_bin"] == ot.freq_bin
        if -bins < r["freq_bin"] < bins:
            assert np.float32(r["pow_left"]) == np.float32(ot.pow_left)
            assert np.float32(r["pow_right"]) == np.float32(ot.pow_right)
        assert t["syncword_freq"] == ot.freq
        assert np.float32(t["syncword_amplitude"]) == np.float32(ot.amplitude)
        assert np.float32(t["syncword_phase"]) == np.float32(ot.phase)
        assert np.float32(t["syncword_time_est"]) == np.float32(ot.time_est)
        assert np.float32(t["syncword_esn0_db"]) == np.float32(ot.esn0_db)


@pytest.mark.parametrize("fft_size,esn0_db,bins,T", [(512, 20.0, 4, 768), (1024, 3.0, 2, 768), (4096, 20.0, 4, 768),
                                                     (4096, 0.0, 8, 300), (8192, 6.0, 1, 768), (1024, 20.0, 0, 40)])
def test_other_fft_sizes_bit_exact_vs_independent_oracle(oracle, rx_params, fft_size, esn0_db, bins, T):
    """fft_size != 2048 (PM/syncword_detection.hpp:133; the reference takes any 2^N): correlator_generic.cu runs the
    oracle's INDEPENDENT radix-2 arithmetic op for op, so metric, detections, records and estimates are bit for bit
    those of the CPU restatement (and of the reference's block code over the same FFT) — offline and streaming."""
    from gr4_packet_modem_b200.stimulus import packet_capture

    n = (1 << 19) + 777
    x, _ = packet_capture(n, seed=5, esn0_db=esn0_db, cfo=0.004, payload_bytes=150)
    kw = dict(min_freq_bin=-bins, max_freq_bin=bins, fft_size=fft_size, time_threshold=T)
    sd = _gpu(rx_params, **kw)
    consumed, recs, tags = sd.detect_host(x)
    o = oracle.SyncwordDetection(**rx_params, **kw, fft_kind=oracle.FFT_RADIX2, record_metric=True)
    oc, oout, otags = o.run(x, chunk=1 << 19)
    assert consumed == oc and len(recs) > 3
    _assert_bit_exact_vs_oracle(sd, recs, tags, o, oc, otags, bins)
    # streaming in ring-sized chunks, delayed output included
    sd2 = _gpu(rx_params, **kw)
    c2, out2, tags2 = sd2.run(x, chunk=65536, want_output=True)
    o2 = oracle.SyncwordDetection(**rx_params, **kw, fft_kind=oracle.FFT_RADIX2)
    oc2, oout2, otags2 = o2.run(x, chunk=65536, want_output=True)
    assert c2 == oc2 and np.array_equal(out2.view(np.uint32), oout2[:oc2].view(np.uint32))
    assert [i for _, i, _ in tags2] == [t.index for t in otags2]
    for (_, _, m), ot in zip(tags2, otags2):
        assert m["syncword_freq"] == ot.freq and m["syncword_freq_bin"] == ot.freq_bin
        assert np.float32(m["syncword_phase"]) == np.float32(ot.phase)
        assert np.float32(m["syncword_amplitude"]) == np.float32(ot.amplitude)


def test_other_fft_size_device_output_shards_and_channels(oracle, rx_params):
    """The remaining entry points with fft_size = 4096: delayed output written on the device, time shards composed
    on the host, batched channels — all equal to the single offline run."""
    import torch

    from gr4_packet_modem_b200.sharding import entry_offsets, plan_shards
    from gr4_packet_modem_b200.stimulus import packet_capture

    F, T = 4096, 768
    n = 1 << 20
    x, _ = packet_capture(n, seed=11, esn0_db=8.0, cfo=-0.006, payload_bytes=100)
    kw = dict(min_freq_bin=-3, max_freq_bin=3, fft_size=F)
    sd = _gpu(rx_params, **kw)
    S = sd.stride
    assert S == F - 297 + 1
    ref_c, ref_recs, _ = sd.detect_host(x)
    assert len(ref_recs) > 20
    # device spans with the delayed output
    d = torch.from_numpy(x).cuda()
    out = torch.full((n,), 7 + 7j, dtype=torch.complex64, device="cuda:0")
    c, recs, _ = sd.detect_device(d.data_ptr(), n, torch.cuda.current_stream().cuda_stream, d_out_ptr=out.data_ptr())
    torch.cuda.synchronize()
    assert c == ref_c and np.array_equal(recs.view(np.uint8), ref_recs.view(np.uint8))
    o = out.cpu().numpy()
    delay = 2 * T + 1
    assert np.array_equal(o[delay:c], x[:c - delay]) and np.all(o[c:] == 7 + 7j)
    # time shards
    shards = plan_shards(n, 3, F, S, T)
    ctxs, tables, keep = [], [], []
    for s in shards:
        c_ = _gpu(rx_params, **kw)
        seg = torch.from_numpy(x[s.first_sample:s.first_sample + s.n_samples].copy()).cuda()
        tables.append(c_.shard_phase1(seg.data_ptr(), s.first_sample, s.n_samples, s.first_block, s.n_blocks,
                                      s.total_blocks))
        ctxs.append(c_)
        keep.append(seg)
    got = np.concatenate([c_.shard_phase2(j, n // (T + 1) + 2)[0] for c_, j in zip(ctxs, entry_offsets(tables))])
    assert np.array_equal(got.view(np.uint8), ref_recs.view(np.uint8))
    # batched channels
    nch, nc = 3, 300000
    buf = np.concatenate([x[i * 100000:i * 100000 + nc] for i in range(nch)])
    db = torch.from_numpy(buf).cuda()
    consumed, per = sd.detect_channels_device(db.data_ptr(), nch, nc, nc)
    single = _gpu(rx_params, **kw)
    for i in range(nch):
        c1, r1, _ = single.detect_host(buf[i * nc:(i + 1) * nc])
        assert c1 == consumed and np.array_equal(per[i].view(np.uint8), r1.view(np.uint8))


def test_generic_path_at_2048_bit_exact_vs_independent_oracle(oracle, rx_params, monkeypatch):
    """B200SYNC_FORCE_GENERIC=1 puts the default fft_size on the generic path: there the GPU is bit-identical to the
    independent arithmetic, and its detections equal those of the hand-scheduled 2048 kernel on the same capture."""
    from gr4_packet_modem_b200.stimulus import packet_capture

    n = 1 << 20
    x, _ = packet_capture(n, seed=3, esn0_db=0.0, cfo=0.005, payload_bytes=200)
    kw = dict(min_freq_bin=-4, max_freq_bin=4)
    c_t, recs_t, _ = _gpu(rx_params, **kw).detect_host(x)
    monkeypatch.setenv("B200SYNC_FORCE_GENERIC", "1")
    sd = _gpu(rx_params, **kw)
    monkeypatch.delenv("B200SYNC_FORCE_GENERIC")
    consumed, recs, tags = sd.detect_host(x)
    o = oracle.SyncwordDetection(**rx_params, **kw, fft_kind=oracle.FFT_RADIX2, record_metric=True)
    oc, _, otags = o.run(x, chunk=1 << 20)
    assert consumed == oc == c_t
    _assert_bit_exact_vs_oracle(sd, recs, tags, o, oc, otags, 4)
    assert recs["index"].tolist() == recs_t["index"].tolist()


def test_many_hypotheses_bit_exact_vs_mirror_oracle(oracle, rx_params):
    """K = 257 (min/max_freq_bin = -/+128, beyond round 1's limit of 129 hypotheses; the reference takes any range):
    metric, detections and records bit for bit, winners far from bin 0 included (CFO 1.2 rad/sample = bin 113)."""
    from gr4_packet_modem_b200.stimulus import packet_capture

    n = 1 << 18
    x, _ = packet_capture(n, seed=8, esn0_db=10.0, cfo=1.2, payload_bytes=60)
    kw = dict(min_freq_bin=-128, max_freq_bin=128)
    sd = _gpu(rx_params, **kw)
    consumed, recs, tags = sd.detect_host(x)
    o = oracle.SyncwordDetection(**rx_params, **kw, fft_kind=oracle.FFT_MIRROR, record_metric=True)
    oc, _, otags = o.run(x, chunk=1 << 18)
    assert consumed == oc and len(recs) > 10 and abs(int(np.median(recs["freq_bin"])) - 113) <= 1
    _assert_bit_exact_vs_oracle(sd, recs, tags, o, oc, otags, 128)


def test_registered_host_ring_streams_like_a_pageable_one(rx_params):
    """b200sync_host_register: a long-lived pageable ring (GR4's port buffer) is page-locked once; spans inside it then
    take the pinned path (no staging copy).  Same consumed counts, output and tags as the pageable run."""
    from gr4_packet_modem_b200 import host_register, host_unregister
    from gr4_packet_modem_b200.stimulus import packet_capture

    chunk = 1 << 16
    x, _ = packet_capture(12 * chunk, seed=17, esn0_db=10.0, cfo=0.003, payload_bytes=100)
    kw = dict(min_freq_bin=-4, max_freq_bin=4)

    def drive(ring, out, auto=False):
        sd = _gpu(rx_params, **kw)
        if auto:
            sd.set_auto_register(True)
        pos, outs, tags = 0, [], []
        while x.size - pos >= chunk:
            ring[:] = x[pos:pos + chunk]
            status, c, o, t = sd.process_bulk(ring, True)
            outs.append(o.copy())
            tags.extend(t)
            pos += c
        return pos, np.concatenate(outs), [(i, m["syncword_freq"], m["syncword_phase"]) for _, i, m in tags]

    plain = drive(np.empty(chunk, np.complex64), None)
    ring = np.empty(chunk, np.complex64)
    host_register(ring)
    try:
        import torch
        assert torch.cuda.is_available()
        reg = drive(ring, None)
    finally:
        host_unregister(ring)
    assert plain[0] == reg[0] and np.array_equal(plain[1], reg[1]) and plain[2] == reg[2] and len(reg[2]) > 5
    # b200sync_sd_set_auto_register: the context page-locks the ring itself, span by span (a view that starts inside a
    # page and spans of changing length exercise the bookkeeping of overlapping ranges), and releases it at destroy
    import torch

    big = np.empty(chunk + 3000, np.complex64)
    ring2 = big[777:777 + chunk]
    auto = drive(ring2, None, auto=True)
    assert auto[0] == plain[0] and np.array_equal(auto[1], plain[1]) and auto[2] == plain[2]
    assert not torch.from_numpy(big).is_pinned()   # released again when the context went away
    # a capture walked front to back (every span starts inside the previous registration and runs past its end: the
    # registration has to grow, a copy may not straddle two of them)
    sd = _gpu(rx_params, **kw)
    sd.set_auto_register(True)
    xs = x.copy()
    c_w, out_w, tags_w = sd.run(xs, chunk=50000, want_output=True)
    c_p, out_p, tags_p = _gpu(rx_params, **kw).run(x, chunk=50000, want_output=True)
    assert c_w == c_p and np.array_equal(out_w, out_p) and [t[1] for t in tags_w] == [t[1] for t in tags_p]
    del sd
