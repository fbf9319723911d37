"""Throughput of the literal drop-in path: SyncwordDetection::processBulk with HOST spans of GR4's ring-chunk
size (b200sync_sd_process: H2D, kernels, D2H and tags inside every call), called the way the C++ shell does —
input and output spans are views into long-lived buffers (GR4's rings), nothing is allocated per call.
Two source modes: "array" walks one big capture (every span is DRAM-cold: the staging memcpy of a pageable span then
runs at ~9 GB/s); "ring" copies each span into a long-lived 65536-item ring slot first (untimed: that is the upstream
block's write) and times only the processBulk calls, i.e. spans that are cache-warm the way a GR4 ring is.
Usage: stream_throughput.py [log2n]"""
import ctypes as C, sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from gr4_packet_modem_b200 import SyncwordDetection, _native
from gr4_packet_modem_b200._native import SyncwordTag
from gr4_packet_modem_b200.firdes import BPSK, SYNCWORD, unit_energy_rrc
from gr4_packet_modem_b200.stimulus import DeviceStimulus

log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 25
n = 1 << log2n
x = DeviceStimulus(seed=1, esn0_db=20.0, cfo=0.005).generate(n, torch.device("cuda", 0)).cpu().numpy()
L = _native.lib()
tags = (SyncwordTag * 4096)()
# ---- ring mode: the span lives in a fixed ring slot the producer has just written (GR4's CircularBuffer)
for pinned in (False, "auto", True):
    chunk = 1 << 16
    ring = torch.empty(chunk, dtype=torch.complex64, pin_memory=(pinned is True)).numpy()
    out = torch.empty(chunk, dtype=torch.complex64, pin_memory=(pinned is True)).numpy()
    out[:] = 0
    for rep in range(2):
        sd = SyncwordDetection(unit_energy_rrc(), SYNCWORD, BPSK, -4, 4, 768, 9.5)
        if pinned == "auto":
            sd.set_auto_register(True)    # pageable ring, page-locked by the context itself as its spans arrive
        nc, nt = C.c_size_t(0), C.c_size_t(0)
        pos = calls = ntags = 0
        busy = 0.0
        while n - pos >= chunk:
            ring[:] = x[pos:pos + chunk]                       # the producer's write (untimed)
            t0 = time.perf_counter()
            L.b200sync_sd_process(sd._h, ring.ctypes.data, chunk, out.ctypes.data, C.byref(nc), tags, 4096, C.byref(nt))
            busy += time.perf_counter() - t0
            pos += nc.value
            calls += 1
            ntags += nt.value
    print(f"ring   pinned={pinned} chunk 2^16 out=True: {pos/busy/1e6:8.1f} Msps  ({ntags} tags, {1e6*busy/calls:.0f} us per call)", flush=True)
for pinned in (False, True):
    if pinned:
        hx = torch.empty(n, dtype=torch.complex64, pin_memory=True)
        hx.copy_(torch.from_numpy(x))
        src = hx.numpy()
    else:
        src = x
    for chunk in (1 << 16, 1 << 18, 1 << 20):
        out = torch.empty(chunk, dtype=torch.complex64, pin_memory=pinned).numpy()
        out[:] = 0
        for want_out in (True, False):
            for rep in range(2):
                sd = SyncwordDetection(unit_energy_rrc(), SYNCWORD, BPSK, -4, 4, 768, 9.5)
                nc, nt = C.c_size_t(0), C.c_size_t(0)
                pos = calls = ntags = 0
                t0 = time.perf_counter()
                while n - pos >= 2048:
                    m = min(chunk, n - pos)
                    L.b200sync_sd_process(sd._h, src.ctypes.data + 8 * pos, m, out.ctypes.data if want_out else None,
                                          C.byref(nc), tags, 4096, C.byref(nt))
                    if nc.value == 0:
                        break
                    pos += nc.value
                    calls += 1
                    ntags += nt.value
                dt = time.perf_counter() - t0
            print(f"pinned={pinned} chunk 2^{chunk.bit_length()-1} out={want_out}: {pos/dt/1e6:8.1f} Msps  "
                  f"({ntags} tags, {1e6*dt/calls:.0f} us per call)", flush=True)
