"""Development aid: where the SymbolFilter stage of the chain workload spends its time (host replay of the
tag state machine vs kernels).  Usage (GPU box): python scripts/sf_host_overhead.py [log2n] [n_tags]"""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from gr4_packet_modem_b200 import SymbolFilter
from gr4_packet_modem_b200.blocks import STREAM_TAG_DTYPE
from gr4_packet_modem_b200.firdes import pfb_matched_filter_taps

log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
ntags = int(sys.argv[2]) if len(sys.argv) > 2 else 43000
n = 1 << log2n
dev = torch.device("cuda", 0)
x = torch.randn(n, 2, device=dev).view(torch.complex64) if False else torch.view_as_complex(torch.randn(n, 2, device=dev))
sym = torch.empty(n // 4 + ntags + 1024, dtype=torch.complex64, device=dev)
rng = np.random.default_rng(1)
it = np.zeros(ntags, STREAM_TAG_DTYPE)
it["index"] = np.sort(rng.choice(n - 10, ntags, replace=False))
it["has_syncword"] = 1
it["sw"]["syncword_amplitude"] = 1.0
it["sw"]["syncword_freq"] = rng.uniform(-0.05, 0.05, ntags)
it["sw"]["syncword_time_est"] = rng.uniform(-0.5, 0.5, ntags)
st = torch.cuda.current_stream().cuda_stream
for fused in (None, 26):
    sf = SymbolFilter(pfb_matched_filter_taps(), 32, 4, delay=44, fused_cfc_delay=fused)
    for tags in (it[:0], it):
        walls, gpus = [], []
        for rep in range(5):
            sf.restart()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record()
            sf.process_device(x.data_ptr(), n, sym.data_ptr(), sym.numel(), tags, st)
            e1.record()
            torch.cuda.synchronize()
            walls.append((time.perf_counter() - t0) * 1e3)
            gpus.append(e0.elapsed_time(e1))
        print(f"fused_cfc={fused} tags={tags.size}: wall {min(walls):.2f} ms, events {min(gpus):.2f} ms")
