import torch, time
n = 64 * 376 * 1241
h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for name, fn in (("H2D", lambda: d.copy_(h, non_blocking=True)), ("D2H", lambda: h.copy_(d, non_blocking=True))):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20): fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 20
    print(name, f"{n/1e6:.1f} MB in {dt*1e3:.3f} ms = {n/dt/1e9:.1f} GB/s")
# chunked
for chunks in (1, 4, 8):
    per = n // chunks
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(20):
        for c in range(chunks): d[c*per:(c+1)*per].copy_(h[c*per:(c+1)*per], non_blocking=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 20
    print("H2D chunks", chunks, f"{dt*1e3:.3f} ms")
