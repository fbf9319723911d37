"""Development aid: run the fused front end a few times on a device-generated stream (for ncu captures).
Usage: run_frontend.py [log2n] [iters] [fp_contract]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gr4_packet_modem_b200 import FrontEnd
from gr4_packet_modem_b200.firdes import lowpass_prototype_taps
from gr4_packet_modem_b200.stimulus import DeviceStimulus
n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 28)
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
fma = len(sys.argv) > 3 and sys.argv[3] == "1"
raw = DeviceStimulus(seed=1, esn0_db=20.0, cfo=0.0).generate(n, torch.device("cuda:0"))
rate = float(np.float32(1.0) + np.float32(1e-6) * np.float32(1.2))
fe = FrontEnd(rate=rate, taps=lowpass_prototype_taps(32, 40), phase_incr=0.005, fp_contract=fma)
n_y = fe.max_output(n)
y = torch.empty(n_y, dtype=torch.complex64, device="cuda:0")
st = torch.cuda.current_stream().cuda_stream
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(iters):
    fe.restart()
    e0.record()
    fe.process_device(raw.data_ptr(), n, y.data_ptr(), n_y, st)
    e1.record()
    torch.cuda.synchronize()
    print(f"iter {i}: {e0.elapsed_time(e1):.3f} ms", flush=True)
