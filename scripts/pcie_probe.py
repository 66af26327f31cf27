"""H2D / D2H bandwidth of page-locked host memory on this box (the floor of the direct route)."""
import time, torch
n = 256 << 20
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for name, fn in (("H2D", lambda: d.copy_(h, non_blocking=True)), ("D2H", lambda: h.copy_(d, non_blocking=True))):
    for _ in range(2): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): fn()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
    print(f"{name}: {n / dt / 1e9:.1f} GB/s")
h2 = torch.empty(n, dtype=torch.uint8).pin_memory(); d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5):
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
print(f"bidirectional: {n / dt / 1e9:.1f} GB/s each way")
for sz in (1 << 20, 4 << 20, 16 << 20):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(20): d[:sz].copy_(h[:sz], non_blocking=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 20
    print(f"H2D {sz >> 20} MB: {sz / dt / 1e9:.1f} GB/s ({dt * 1e6:.0f} us)")
