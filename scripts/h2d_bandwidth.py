import torch, time
n = 1 << 30
hx = torch.empty(n, dtype=torch.complex64, pin_memory=True)
hx.view(torch.float32).fill_(1.0)
x = torch.empty(n, dtype=torch.complex64, device="cuda")
for _ in range(2):
    x.copy_(hx, non_blocking=True); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3):
    x.copy_(hx, non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 3
print(f"single copy: {8*n/dt/1e9:.2f} GB/s")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
h = n // 2
t0 = time.perf_counter()
for _ in range(3):
    with torch.cuda.stream(s1): x[:h].copy_(hx[:h], non_blocking=True)
    with torch.cuda.stream(s2): x[h:].copy_(hx[h:], non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 3
print(f"two streams: {8*n/dt/1e9:.2f} GB/s")
