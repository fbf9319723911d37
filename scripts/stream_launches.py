"""Development aid: a few processBulk calls of the streaming path (65536-item host spans) for an ncu launch list.
Usage: ncu --metrics gpu__time_duration.sum ... python scripts/stream_launches.py [calls]"""
import ctypes as C, sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from gr4_packet_modem_b200 import SyncwordDetection, _native
from gr4_packet_modem_b200._native import SyncwordTag
from gr4_packet_modem_b200.firdes import BPSK, SYNCWORD, unit_energy_rrc
from gr4_packet_modem_b200.stimulus import DeviceStimulus
calls = int(sys.argv[1]) if len(sys.argv) > 1 else 12
chunk = 1 << 16
x = DeviceStimulus(seed=1, esn0_db=20.0, cfo=0.005).generate(calls * chunk + 4096, torch.device("cuda", 0)).cpu().numpy()
L = _native.lib()
tags = (SyncwordTag * 4096)()
sd = SyncwordDetection(unit_energy_rrc(), SYNCWORD, BPSK, -4, 4, 768, 9.5)
out = np.zeros(chunk, np.complex64)
nc, nt = C.c_size_t(0), C.c_size_t(0)
pos = 0
t0 = time.perf_counter()
for i in range(calls):
    L.b200sync_sd_process(sd._h, x.ctypes.data + 8 * pos, chunk, out.ctypes.data, C.byref(nc), tags, 4096, C.byref(nt))
    pos += nc.value
print("consumed", pos, "us per call", 1e6 * (time.perf_counter() - t0) / calls)
