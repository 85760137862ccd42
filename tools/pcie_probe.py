"""PCIe copy rates of the box (pinned host memory), for reading the e2e number: python tools/dbg_pcie.py"""
import torch, time
n = 64 << 20
h = torch.empty(n, dtype=torch.uint8).pin_memory(); d = torch.empty(n, dtype=torch.uint8, device="cuda")
h2 = torch.empty(n, dtype=torch.uint8).pin_memory(); d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=10):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps
print("H2D 64MiB  %.3f ms  %.1f GB/s" % (t(lambda: d.copy_(h, non_blocking=True)) * 1e3, n / t(lambda: d.copy_(h, non_blocking=True)) / 1e9))
print("D2H 64MiB  %.3f ms  %.1f GB/s" % (t(lambda: h2.copy_(d2, non_blocking=True)) * 1e3, n / t(lambda: h2.copy_(d2, non_blocking=True)) / 1e9))
def both():
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
print("both directions at once, 64MiB each: %.3f ms" % (t(both) * 1e3))
def striped():
    for i in range(16):
        a = i * (n // 16)
        d[a:a + n // 16].copy_(h[a:a + n // 16], non_blocking=True)
print("H2D 16 x 4MiB: %.3f ms" % (t(striped) * 1e3))
