"""Ad-hoc timing of detect_device on a device-resident synthetic capture (development aid)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gr4_packet_modem_b200 import SyncwordDetection
from gr4_packet_modem_b200.firdes import unit_energy_rrc, SYNCWORD, BPSK
from gr4_packet_modem_b200.stimulus import DeviceStimulus

logn = int(sys.argv[1]) if len(sys.argv) > 1 else 26
n = 1 << logn
dev = torch.device("cuda:0")
t0 = time.time()
x = DeviceStimulus(seed=1, esn0_db=20.0, cfo=0.005).generate(n, dev)
torch.cuda.synchronize()
print(f"generated 2^{logn} samples in {time.time()-t0:.2f}s", flush=True)
for bins in (4, 0, 16):
    sd = SyncwordDetection(unit_energy_rrc(), SYNCWORD, BPSK, -bins, bins)
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(2):
        c, recs, tags = sd.detect_device(x.data_ptr(), n, st)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    reps = 3
    for _ in range(reps):
        c, recs, tags = sd.detect_device(x.data_ptr(), n, st)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"K={2*bins+1}: {ms:.2f} ms, {c/ms/1e3:.1f} Msps, detections={len(recs)}, first={recs['index'][:4].tolist()}", flush=True)
