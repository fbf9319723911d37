"""How sensitive are the DECISIONS of the path to the FFT arithmetic?  (SURVEY hard part 2: "quantify, don't assume".)

On a 2^log2n-sample Es/N0 0 dB capture, for K in {9, 17, 33} hypotheses:
  * the GPU's per-sample metric (packed-FP32 16x16x8 FFT, fft2048.cuh) against the metric of an arithmetic that shares
    nothing with it — the oracle's radix-2 FFT, which is bit-identical to the reference's own block code compiled
    against that FFT (tests/test_oracle_vs_reference_blocks.py) — relative differences, and how many per-sample
    candidate flags (forward-isolated maxima, PM/syncword_detection.hpp:267-317) differ;
  * for power_threshold in 6 .. 20: the detection sets of the two arithmetics — symmetric difference (expected 0) —
    and the smallest decision margins seen: how far the median count of an examined peak was from flipping
    (2*count vs 2T+1, :279) and how close the closest runner-up came to a candidate within its window.
Also: relative L2 error of the GPU's closed-form rotator against the reference's float recurrence (PM/rotator.hpp:56-65),
per 2^15-sample window over 2^22 samples (the bar of the GPU-vs-reference test grows with the window index).
Usage (GPU box): python scripts/lowsnr_sensitivity.py [log2n] > gpurun_out/r2_lowsnr_sensitivity.json"""
import json
import sys
import threading

sys.path.insert(0, ".")
import numpy as np
import torch
from scipy.ndimage import maximum_filter1d

from gr4_packet_modem_b200 import FrontEnd, SyncwordDetection
from gr4_packet_modem_b200.firdes import BPSK, SYNCWORD, unit_energy_rrc
from gr4_packet_modem_b200.stimulus import DeviceStimulus
from oracle import pyoracle as po

T = 768
log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 26
n = 1 << log2n
dev = torch.device("cuda", 0)
st = torch.cuda.current_stream().cuda_stream
rrc = unit_energy_rrc()
po.build(ref=False)
x = DeviceStimulus(seed=7, esn0_db=0.0, cfo=0.04, payload_bytes=200).generate(n, dev)
xh = x.cpu().numpy()
THRS = [6.0, 8.0, 9.5, 12.0, 14.0, 16.0, 18.0, 20.0]


def forward_isolated(z):
    """candidate flag of every sample: no larger power within the next T samples (strict '>', :314)"""
    zz = np.concatenate([z, np.zeros(T, z.dtype)])
    fwd = maximum_filter1d(zz, size=T, origin=-(T // 2), mode="constant")[1:len(z) + 1]  # max over (p, p+T]
    return ~(fwd > z)


def walk(z, cand, hi, thr):
    """the sequential detector on a metric: examined peaks, their median counts, detections"""
    pos = np.flatnonzero(cand[:hi])
    out, margins, r, i = [], [], 0, 0
    while True:
        i = np.searchsorted(pos, r)
        if i >= len(pos):
            break
        p = int(pos[i])
        tv = np.float32(z[p]) / np.float32(thr)
        lo = max(0, p - T)
        cnt = int(np.count_nonzero(z[lo:p + T + 1] < tv)) + (T - (p - lo))   # zero history before the stream
        margins.append(abs(2 * cnt - (2 * T + 1)))
        if 2 * cnt >= 2 * T + 1:
            out.append(p)
        r = p + T + 1
    return np.array(out, np.int64), (min(margins) if margins else None)


rows = []
metrics = {}


def oracle_metric(bins):
    o = po.SyncwordDetection(rrc, SYNCWORD, BPSK, -bins, bins, T, 9.5, fft_kind=po.FFT_RADIX2, record_metric=True)
    oc, _, _ = o.run(xh, chunk=1 << 20)
    metrics[bins] = (oc, o.metric(oc)[0])


ths = [threading.Thread(target=oracle_metric, args=(b,)) for b in (4, 8, 16)]
for t in ths:
    t.start()
for t in ths:
    t.join()

for bins in (4, 8, 16):
    oc, zo = metrics[bins]
    sd = SyncwordDetection(rrc, SYNCWORD, BPSK, -bins, bins, T, 9.5)
    c, recs, _ = sd.detect_device(x.data_ptr(), n, st)
    assert c == oc
    zg = sd.metric(c)
    rel = np.abs(zg.astype(np.float64) - zo) / np.maximum(zo, 1e-30)
    cg, co = forward_isolated(zg), forward_isolated(zo)
    hi = c - T - 1
    # how close did the closest runner-up come to a candidate (oracle metric)?
    pc = np.flatnonzero(co[:hi])
    zz = np.concatenate([zo, np.zeros(T, zo.dtype)])
    fwd = maximum_filter1d(zz, size=T, origin=-(T // 2), mode="constant")[1:len(zo) + 1]
    gap = (zo[pc] - fwd[pc]) / np.maximum(zo[pc], 1e-30)
    row = {"K": 2 * bins + 1, "samples": int(c), "metric_rel_diff_max": float(rel.max()), "metric_rel_diff_median": float(np.median(rel)),
           "candidate_flags_differing": int(np.count_nonzero(cg[:hi] != co[:hi])), "candidates": int(len(pc)),
           "smallest_relative_gap_candidate_vs_runner_up": float(gap.min()), "thresholds": []}
    for thr in THRS:
        sdt = SyncwordDetection(rrc, SYNCWORD, BPSK, -bins, bins, T, thr)
        _, r, _ = sdt.detect_device(x.data_ptr(), n, st)
        gpu_idx = r["index"].astype(np.int64)
        ref_idx, margin = walk(zo, co, hi, thr)
        ref_idx = ref_idx[ref_idx + 2 * T + 1 < c]
        row["thresholds"].append({"power_threshold": thr, "gpu_detections": int(len(gpu_idx)), "reference_detections": int(len(ref_idx)),
                                  "differing": int(len(np.setxor1d(gpu_idx, ref_idx))),
                                  "smallest_median_count_margin": margin})
    rows.append(row)

# rotator: closed form on the GPU vs the reference's float recurrence, per 2^15 window
nr = 1 << 22
y = (np.ones(nr) + 0j).astype(np.complex64)
fe = FrontEnd(phase_incr=0.005, enable_resampler=False)
_, g = fe.process_bulk(y)
ref = po.rotator(y, 0.005)
w = 1 << 15
rot = [float(np.linalg.norm(g[i:i + w].astype(np.complex128) - ref[i:i + w]) / np.linalg.norm(ref[i:i + w].astype(np.complex128)))
       for i in range(0, nr, w)]
print(json.dumps({"log2n": log2n, "esn0_db": 0.0, "cfo": 0.04, "time_threshold": T, "rows": rows,
                  "rotator_rel_l2_per_2p15_window": {"phase_incr": 0.005, "first": rot[0], "window_8": rot[8], "window_32": rot[32],
                                                     "last": rot[-1], "max": max(rot), "all": rot}}, indent=1))
