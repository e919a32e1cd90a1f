import torch, time
n = 2 << 30
h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d = torch.empty(n, dtype=torch.uint8, device='cuda')
for name, fn in (("d2h", lambda: h.copy_(d, non_blocking=True)), ("h2d", lambda: d.copy_(h, non_blocking=True))):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(4): fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t) / 4
    print(name, n / dt / 1e9, "GB/s")
# both directions at once
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
h2 = torch.empty(n, dtype=torch.uint8, pin_memory=True); d2 = torch.empty(n, dtype=torch.uint8, device='cuda')
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(4):
    with torch.cuda.stream(s1): h.copy_(d, non_blocking=True)
    with torch.cuda.stream(s2): d2.copy_(h2, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 4
print("bidir each", n / dt / 1e9, "GB/s")
