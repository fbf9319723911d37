"""Development aid: where the host time of one bulk step goes — the C call (kernels + record copy), the copy of the
records into a fresh array, output_tag() on them.  Usage: python scripts/host_overhead_step.py   (one B200, 2^30 samples)"""
import sys, time, ctypes as C
sys.path.insert(0, ".")
import torch, numpy as np
from gr4_packet_modem_b200 import SyncwordDetection, _native
from gr4_packet_modem_b200.blocks import _copy_records
from gr4_packet_modem_b200.firdes import unit_energy_rrc, SYNCWORD, BPSK
from gr4_packet_modem_b200.stimulus import DeviceStimulus
n = 1 << 30
x = DeviceStimulus(seed=1, esn0_db=20.0, cfo=0.005).generate(n, torch.device("cuda:0"))
sd = SyncwordDetection(unit_energy_rrc(), SYNCWORD, BPSK, -4, 4)
out = torch.empty(n, dtype=torch.complex64, device="cuda:0")
st = torch.cuda.current_stream().cuda_stream
L = _native.lib()
max_recs = n // 769 + 2
recs = sd._rec_buffer(max_recs)
for i in range(4):
    nr, nc = C.c_size_t(0), C.c_size_t(0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    L.b200sync_sd_detect_device(sd._h, C.c_void_p(x.data_ptr()), n, C.c_void_p(out.data_ptr()), C.c_void_p(st), recs.ctypes.data, max_recs, C.byref(nr), C.byref(nc))
    t1 = time.perf_counter()
    r = _copy_records(recs, nr.value)
    t2 = time.perf_counter()
    tags = sd.records_to_tags(r)
    t3 = time.perf_counter()
    t = sd.last_timings()
    print(f"C call {1e3*(t1-t0):.3f} ms (kernels {t['correlate_ms']+t['peaks_ms']+t['refine_ms']:.3f}) copy {1e3*(t2-t1):.3f} tags {1e3*(t3-t2):.3f} n={nr.value}")
