"""Development aid: stage timings of detect_device for each variant library under build/variants/
(one subprocess per variant; B200SYNC_LIB selects the library).  Usage: variant_timing.py [log2n] [bins]"""
import glob, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
logn = sys.argv[1] if len(sys.argv) > 1 else "28"
bins = sys.argv[2] if len(sys.argv) > 2 else "4"
CHILD = r'''
import sys, os
sys.path.insert(0, %r)
import torch
from gr4_packet_modem_b200 import SyncwordDetection
from gr4_packet_modem_b200.firdes import unit_energy_rrc, SYNCWORD, BPSK
from gr4_packet_modem_b200.stimulus import DeviceStimulus
n = 1 << int(sys.argv[1]); bins = int(sys.argv[2])
x = DeviceStimulus(seed=1, esn0_db=20.0, cfo=0.005).generate(n, torch.device("cuda:0"))
sd = SyncwordDetection(unit_energy_rrc(), SYNCWORD, BPSK, -bins, bins)
st = torch.cuda.current_stream().cuda_stream
tm = []
for i in range(8):
    c, recs, tags = sd.detect_device(x.data_ptr(), n, st)
    if i >= 3: tm.append(sd.last_timings())
import statistics as S
f = lambda k: S.mean(t[k] for t in tm)
print(f"corr {f('correlate_ms'):.3f} peaks {f('peaks_ms'):.3f} refine {f('refine_ms'):.3f} ms  det={len(recs)} sum={int(recs['index'].sum()) %% 1000003}")
''' % ROOT
for lib in sorted(glob.glob(os.path.join(ROOT, "build", "variants", "lib_*.so"))):
    env = dict(os.environ, B200SYNC_LIB=lib)
    r = subprocess.run([sys.executable, "-c", CHILD, logn, bins], env=env, capture_output=True, text=True)
    print(os.path.basename(lib), (r.stdout.strip() or r.stderr.strip()[-400:]), flush=True)
