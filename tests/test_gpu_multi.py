"""GPU tests of the host-gathered multi-GPU entry points (b200sync_sd_multi_*): ONE capture cut into time shards,
one host thread per shard, chain tables composed on the host, records gathered on the host.  The records must
equal the single-context run bit for bit.  On a one-GPU box the shards are several contexts on device 0 — the
host logic (plan, threads, barrier, table composition, record gathering) is the same; bench.py --gpus N and the
world-2 gloo test cover the process-per-GPU form."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _devices(k):
    import torch

    n = torch.cuda.device_count()
    return [i % n for i in range(k)]


@pytest.mark.parametrize("shards,n,T", [(2, 1 << 21, 768), (3, (1 << 21) + 12345, 768), (5, 1 << 20, 100),
                                        (8, 1 << 22, 768), (4, 5000, 768)])
def test_multi_host_equals_single(rx_params, shards, n, T):
    from gr4_packet_modem_b200 import SyncwordDetection, SyncwordDetectionMulti
    from gr4_packet_modem_b200.stimulus import packet_capture

    x, _ = packet_capture(n, seed=21, esn0_db=8.0, cfo=0.004, payload_bytes=150)
    kw = dict(min_freq_bin=-2, max_freq_bin=2, power_threshold=9.5, time_threshold=T)
    c1, r1, t1 = SyncwordDetection(**rx_params, **kw).detect_host(x)
    m = SyncwordDetectionMulti(**rx_params, **kw, devices=_devices(shards))
    assert m.n_devices == shards
    c2, r2, t2 = m.detect_host(x)
    assert c2 == c1
    assert len(r1) > 0 or n < 10000
    assert np.array_equal(r1.view(np.uint8), r2.view(np.uint8)), "records differ from the single-context run"
    assert np.array_equal(t1.view(np.uint8), t2.view(np.uint8))
    # twice on the same context: state fully reset per call
    c3, r3, _ = m.detect_host(x)
    assert c3 == c1 and np.array_equal(r1.view(np.uint8), r3.view(np.uint8))


def test_multi_device_resident_and_file(rx_params, tmp_path):
    import torch
    from gr4_packet_modem_b200 import SyncwordDetection, SyncwordDetectionMulti
    from gr4_packet_modem_b200.stimulus import packet_capture

    n = (1 << 21) + 999
    x, _ = packet_capture(n, seed=22, esn0_db=20.0, cfo=0.005, payload_bytes=300)
    kw = dict(min_freq_bin=-4, max_freq_bin=4, power_threshold=9.5)
    c1, r1, _ = SyncwordDetection(**rx_params, **kw).detect_host(x)
    m = SyncwordDetectionMulti(**rx_params, **kw, devices=_devices(3))
    plan = m.plan(n)
    assert [p["first_block"] for p in plan] == sorted(p["first_block"] for p in plan)
    assert sum(p["n_blocks"] for p in plan) == plan[0]["total_blocks"] == (n - 2048) // 1752 + 1
    bufs = []
    for p in plan:
        seg = x[p["first_sample"]:p["first_sample"] + p["n_samples"]]
        assert seg.size == p["n_samples"]
        bufs.append(torch.from_numpy(seg.view(np.float32).copy()).to(torch.device("cuda", p["device"])))
    torch.cuda.synchronize()
    c2, r2, _ = m.detect_device([b.data_ptr() for b in bufs], n)
    assert c2 == c1 and np.array_equal(r1.view(np.uint8), r2.view(np.uint8))
    # the same capture from a raw cf32 file, with an item offset
    path = os.path.join(tmp_path, "cap.cf32")
    pre = 1000
    np.concatenate([np.zeros(pre, np.complex64), x]).tofile(path)
    c3, r3, _, items = m.detect_file(path, first_item=pre)
    assert items == n and c3 == c1 and np.array_equal(r1.view(np.uint8), r3.view(np.uint8))
    with pytest.raises(Exception, match="error opening file"):
        m.detect_file(os.path.join(tmp_path, "missing.cf32"))


def test_shard_delayed_output_and_host_output(rx_params):
    """The block's output span (input delayed by 2T+1, PM/syncword_detection.hpp:318-319) from the sharded path
    (every shard writes the slice it owns) and from the host-span bulk call equals the streaming block's."""
    import torch
    from gr4_packet_modem_b200 import SyncwordDetection
    from gr4_packet_modem_b200.sharding import entry_offsets, plan_shards
    from gr4_packet_modem_b200.stimulus import packet_capture

    n = (1 << 20) + 4321
    x, _ = packet_capture(n, seed=23, esn0_db=20.0, cfo=0.005, payload_bytes=300)
    kw = dict(min_freq_bin=-1, max_freq_bin=1, power_threshold=9.5)
    ref = SyncwordDetection(**rx_params, **kw)
    c0, out0, tags0 = ref.run(x, chunk=1 << 18, want_output=True)
    # host-span bulk call with an output span
    sd = SyncwordDetection(**rx_params, **kw)
    out_h = np.full(n, 7 + 7j, np.complex64)
    c1, r1, _ = sd.detect_host(x, out=out_h)
    assert c1 == c0 and np.array_equal(out_h[:c1].view(np.uint32), out0.view(np.uint32))
    assert np.all(out_h[c1:] == 7 + 7j)
    # three shards on one device, each writing its slice into ONE n-item device buffer
    world = 3
    shards = plan_shards(n, world, 2048, 1752, 768)
    dev = torch.device("cuda:0")
    d_out = torch.full((2 * (c0 + 512),), 5.0, dtype=torch.float32, device=dev)
    sds = [SyncwordDetection(**rx_params, **kw) for _ in range(world)]
    bufs, tables = [], []
    for sh, s in zip(shards, sds):
        seg = torch.from_numpy(x[sh.first_sample:sh.first_sample + sh.n_samples].view(np.float32).copy()).to(dev)
        bufs.append(seg)
        s.shard_output(d_out.data_ptr(), 0, c0)
        tables.append(s.shard_phase1(seg.data_ptr(), sh.first_sample, sh.n_samples, sh.first_block, sh.n_blocks,
                                     sh.total_blocks, torch.cuda.current_stream().cuda_stream))
    idx = []
    for s, j in zip(sds, entry_offsets(tables)):
        r, _ = s.shard_phase2(j, n // 769 + 2)
        idx += (r["index"] + s.delay).tolist()
    torch.cuda.synchronize()
    got = d_out.cpu().numpy().view(np.complex64)
    assert np.array_equal(got[:c0].view(np.uint32), out0.view(np.uint32))
    assert np.all(got[c0:].view(np.float32) == 5.0)
    assert idx == [t[1] for t in tags0]


@pytest.mark.parametrize("mode", ["d2h", "memcpy"])
def test_pinned_host_output_comes_back_over_pcie(rx_params, mode):
    """A PINNED host output span takes the other path: the correlator writes the delayed stream on the device and
    finished pieces are copied back on a D2H stream (bulk host call and sharded host call).  Same bits as the
    streaming block's output span."""
    import subprocess
    import sys

    if os.environ.get("B200SYNC_HOST_OUTPUT") != mode:
        # the library reads the override once per process: run this very test in a child with it set
        env = dict(os.environ, B200SYNC_HOST_OUTPUT=mode)
        r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-m", "gpu", f"{__file__}::test_pinned_host_output_comes_back_over_pcie[{mode}]"],
                           env=env, capture_output=True, text=True, cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        return
    import torch
    from gr4_packet_modem_b200 import SyncwordDetection
    from gr4_packet_modem_b200.sharding import entry_offsets, plan_shards
    from gr4_packet_modem_b200.stimulus import packet_capture

    n = (1 << 21) + 77
    x, _ = packet_capture(n, seed=31, esn0_db=20.0, cfo=0.005, payload_bytes=300)
    kw = dict(min_freq_bin=-1, max_freq_bin=1, power_threshold=9.5)
    c0, out0, tags0 = SyncwordDetection(**rx_params, **kw).run(x, chunk=1 << 18, want_output=True)
    hx = torch.empty(n, dtype=torch.complex64, pin_memory=True)
    hx.copy_(torch.from_numpy(x))
    hout = torch.full((n,), 9.0, dtype=torch.complex64, pin_memory=True)
    sd = SyncwordDetection(**rx_params, **kw)
    c1, r1, _ = sd.detect_host((hx.data_ptr(), n), out=hout.data_ptr())
    got = hout.numpy()
    assert c1 == c0 and np.array_equal(got[:c1].view(np.uint32), out0.view(np.uint32))
    assert np.all(got[c1:] == 9.0)
    assert (r1["index"] + sd.delay).tolist() == [t[1] for t in tags0]
    # sharded, every shard's slice into ONE pinned host buffer
    world = 3
    shards = plan_shards(n, world, 2048, 1752, 768)
    hout.fill_(4.0)
    sds = [SyncwordDetection(**rx_params, **kw) for _ in range(world)]
    tables = []
    for sh, s in zip(shards, sds):
        s.shard_output_host(hout.data_ptr(), 0, c0)
        tables.append(s.shard_phase1_host((hx.data_ptr() + 8 * sh.first_sample, sh.n_samples), sh.first_sample,
                                          sh.first_block, sh.n_blocks, sh.total_blocks))
    idx = []
    for s, j in zip(sds, entry_offsets(tables)):
        r, _ = s.shard_phase2(j, n // 769 + 2)
        idx += (r["index"] + s.delay).tolist()
    got = hout.numpy()
    assert np.array_equal(got[:c0].view(np.uint32), out0.view(np.uint32))
    assert np.all(got[c0:] == 4.0)
    assert idx == [t[1] for t in tags0]
