import sys; sys.path.insert(0, ".")
import torch
from gr4_packet_modem_b200.stimulus import DeviceStimulus
dev = torch.device("cuda", 0)
n = 1 << 30
for kw in (dict(esn0_db=20.0, cfo=0.005), dict(esn0_db=None, cfo=0.0)):
    s = DeviceStimulus(seed=1, **kw)
    out = torch.empty(n, dtype=torch.complex64, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    s.generate_device(out.data_ptr(), n, 0, st); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): s.generate_device(out.data_ptr(), n, 0, st)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(kw, f"{ms:.2f} ms  {n/ms/1e6:.1f} Gsps  {8*n/ms/1e6:.0f} GB/s written")
