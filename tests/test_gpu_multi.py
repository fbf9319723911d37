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
