"""raw pinned-memory PCIe bandwidth on this box: H2D alone, D2H alone, both at once (what bounds bench.py's e2e number)"""
import torch

n = 256 << 20
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=8):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def both():
    with torch.cuda.stream(s1):
        d_a.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s2):
        h_out.copy_(d_b, non_blocking=True)
    s1.synchronize(); s2.synchronize()


t = timed(lambda: d_a.copy_(h_in, non_blocking=True)); print("H2D %.1f GB/s" % (n / t / 1e6))
t = timed(lambda: h_out.copy_(d_b, non_blocking=True)); print("D2H %.1f GB/s" % (n / t / 1e6))
import time
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(8):
    both()
dt = (time.perf_counter() - t0) / 8
print("both directions at once: %.1f GB/s each" % (n / dt / 1e9))
