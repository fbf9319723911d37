"""BASELINE configs[3] / SURVEY §8(d) config 4: low-SNR wide-CFO search.  Es/N0 0 dB captures at several carrier
offsets inside and outside the search range, K in {9, 17, 33} hypotheses, power_threshold sweep: detection
probability P_d (true syncword starts found, +/-1 sample), false alarms per Msample, and — on a prefix —
index agreement with the CPU oracle (mirror arithmetic, must be exact).
Usage (GPU box): python scripts/threshold_sweep.py [log2n] > gpurun_out/threshold_sweep.json"""
import json
import sys

sys.path.insert(0, ".")
import numpy as np
import torch

from gr4_packet_modem_b200 import SyncwordDetection
from gr4_packet_modem_b200.firdes import BPSK, SYNCWORD, unit_energy_rrc
from gr4_packet_modem_b200.stimulus import DeviceStimulus
from oracle import pyoracle as po

log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 26
n = 1 << log2n
dev = torch.device("cuda", 0)
st = torch.cuda.current_stream().cuda_stream
rrc = unit_energy_rrc()
payload = 200
rows = []
po.build(ref=False)
for cfo in (0.005, 0.04, -0.08, 0.15):
    stim = DeviceStimulus(seed=7, esn0_db=0.0, cfo=cfo, payload_bytes=payload)
    x = stim.generate(n, dev)
    frame = stim.frame_len * 4
    truth = np.arange(0, n - 4096, frame)
    xp = x[:1 << 21].cpu().numpy()
    for bins in (4, 8, 16):
        covered = abs(cfo) <= (bins + 0.5) * np.pi / 297
        for thr in (6.0, 8.0, 9.5, 12.0, 16.0, 20.0):
            sd = SyncwordDetection(rrc, SYNCWORD, BPSK, -bins, bins, 768, thr)
            consumed, recs, tags = sd.detect_device(x.data_ptr(), n, st)
            idx = recs["index"].astype(np.int64)
            t = truth[truth + 1537 + 297 < consumed]
            near = np.abs(idx[:, None] - t[np.clip(np.searchsorted(t, idx), 1, len(t) - 1) - 1][:, None]).ravel() if len(idx) and len(t) > 1 else np.array([])
            pos = np.searchsorted(idx, t)
            hit = np.zeros(len(t), bool)
            for d in (-1, 0):
                j = np.clip(pos + d, 0, max(len(idx) - 1, 0))
                if len(idx):
                    hit |= np.abs(idx[j] - t) <= 1
            is_true = np.zeros(len(idx), bool)
            if len(idx) and len(t):
                k = np.clip(np.searchsorted(t, idx), 0, len(t) - 1)
                is_true = (np.abs(t[k] - idx) <= 1) | (np.abs(t[np.maximum(k - 1, 0)] - idx) <= 1)
            row = {"cfo": cfo, "K": 2 * bins + 1, "covered": bool(covered), "power_threshold": thr,
                   "frames": int(len(t)), "detections": int(len(idx)), "P_d": float(hit.mean()) if len(t) else None,
                   "false_alarms_per_Msample": float((~is_true).sum() / (consumed / 1e6))}
            if thr in (9.5,):
                o = po.SyncwordDetection(rrc, SYNCWORD, BPSK, -bins, bins, 768, thr, fft_kind=po.FFT_MIRROR)
                oc, _, otags = o.run(xp, chunk=1 << 16)
                c2, r2, _ = sd.detect_host(xp)
                row["oracle_prefix_samples"] = int(oc)
                row["oracle_indices_equal"] = bool(c2 == oc and (r2["index"] + 1537).tolist() == [tg.index for tg in otags])
            rows.append(row)
print(json.dumps({"log2n": log2n, "esn0_db": 0.0, "payload_bytes": payload, "rows": rows}, indent=1))
