"""Development aid: run detect_device a few times on a device-generated capture (for ncu captures of one configuration).
Usage: run_detect.py [log2n] [bins] [iters] [cfo] [out] [fft_size]      out = 1: block contract, the delayed output span is written
(B200SYNC_FORCE_GENERIC=1: fft_size 2048 on the generic radix-2 path)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gr4_packet_modem_b200 import SyncwordDetection
from gr4_packet_modem_b200.firdes import unit_energy_rrc, SYNCWORD, BPSK
from gr4_packet_modem_b200.stimulus import DeviceStimulus
n = 1 << int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 28
bins = int(sys.argv[2]) if len(sys.argv) > 2 else 4
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 4
cfo = float(sys.argv[4]) if len(sys.argv) > 4 else 0.005
with_out = len(sys.argv) > 5 and sys.argv[5] == "1"
fft_size = int(sys.argv[6]) if len(sys.argv) > 6 else 2048
x = DeviceStimulus(seed=1, esn0_db=20.0, cfo=cfo).generate(n, torch.device("cuda:0"))
sd = SyncwordDetection(unit_energy_rrc(), SYNCWORD, BPSK, -bins, bins, fft_size=fft_size)
st = torch.cuda.current_stream().cuda_stream
out = torch.empty(n, dtype=torch.complex64, device="cuda:0") if with_out else None
for i in range(iters):
    c, recs, tags = sd.detect_device(x.data_ptr(), n, st, d_out_ptr=out.data_ptr() if with_out else 0)
    t = sd.last_timings()
    print(f"iter {i}: corr {t['correlate_ms']:.3f} peaks {t['peaks_ms']:.3f} refine {t['refine_ms']:.3f} ms det={len(recs)}", flush=True)
