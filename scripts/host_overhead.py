"""Where does the host-side time of one offline step go?  (debug aid)"""
import ctypes as C, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from gr4_packet_modem_b200 import SyncwordDetection, _native
from gr4_packet_modem_b200.blocks import RECORD_DTYPE
from gr4_packet_modem_b200.firdes import BPSK, SYNCWORD, unit_energy_rrc
from gr4_packet_modem_b200.stimulus import DeviceStimulus
n = 1 << 30
dev = torch.device("cuda", 0)
x = DeviceStimulus(seed=1, esn0_db=20.0, cfo=0.005).generate(n, dev)
sd = SyncwordDetection(unit_energy_rrc(), SYNCWORD, BPSK, -4, 4, 768, 9.5, device=0)
L = _native.lib()
max_recs = n // 769 + 2
recs = np.empty(max_recs, RECORD_DTYPE)
nr, nc = C.c_size_t(0), C.c_size_t(0)
stream = torch.cuda.current_stream().cuda_stream
for it in range(6):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    L.b200sync_sd_detect_device(sd._h, C.c_void_p(x.data_ptr()), n, None, C.c_void_p(stream), recs.ctypes.data, max_recs, C.byref(nr), C.byref(nc))
    t1 = time.perf_counter()
    r = recs[:nr.value].copy()
    t2 = time.perf_counter()
    tags = sd.records_to_tags(r)
    t3 = time.perf_counter()
    tm = sd.last_timings()
    print(f"C call {1e3*(t1-t0):.2f} ms (stages {tm['correlate_ms']:.2f}+{tm['peaks_ms']:.2f}+{tm['refine_ms']:.2f}={sum(tm.values()):.2f}), copy {1e3*(t2-t1):.2f}, tags {1e3*(t3-t2):.2f}")
