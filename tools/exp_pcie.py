"""PCIe ceiling of the box: H2D / D2H of the e2e step's byte counts, alone and concurrently (pinned memory)."""
import time, torch
h2d_b, d2h_b = 60_000_000, 48_000_000
src = torch.empty(h2d_b, dtype=torch.uint8).pin_memory(); dst_d = torch.empty(h2d_b, dtype=torch.uint8, device="cuda")
src_d = torch.empty(d2h_b, dtype=torch.uint8, device="cuda"); dst = torch.empty(d2h_b, dtype=torch.uint8).pin_memory()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(up, down, reps=10):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        if up:
            with torch.cuda.stream(s1): dst_d.copy_(src, non_blocking=True)
        if down:
            with torch.cuda.stream(s2): dst.copy_(src_d, non_blocking=True)
        torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3
for _ in range(2): run(True, True)
a, b, c = run(True, False), run(False, True), run(True, True)
print(f"H2D 60 MB alone {a:.3f} ms = {h2d_b/a/1e6:.1f} GB/s | D2H 48 MB alone {b:.3f} ms = {d2h_b/b/1e6:.1f} GB/s | both {c:.3f} ms")
# chunked H2D as the engine does it (1M players: 4 MB + 2 MB per chunk)
def chunked(reps=10):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        with torch.cuda.stream(s1):
            for k in range(10):
                dst_d[k*6_000_000:k*6_000_000+4_000_000].copy_(src[k*6_000_000:k*6_000_000+4_000_000], non_blocking=True)
                dst_d[k*6_000_000+4_000_000:(k+1)*6_000_000].copy_(src[k*6_000_000+4_000_000:(k+1)*6_000_000], non_blocking=True)
        torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3
print(f"H2D 60 MB as 20 chunked copies {chunked():.3f} ms")
