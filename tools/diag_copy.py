"""Burst vs sustained device copy bandwidth (is a back-to-back stream of HBM-bound kernels slower than isolated ones?)."""
import time, torch
n = 1 << 29   # 2 GiB of f32 per buffer
a = torch.empty(n, dtype=torch.float32, device="cuda").normal_()
b = torch.empty_like(a)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
def run(k):
    torch.cuda.synchronize(); ev[0].record()
    for _ in range(k): b.copy_(a)
    ev[1].record(); torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]) / k
run(3)
iso = []
for _ in range(10):
    time.sleep(0.05); iso.append(run(1))
print("isolated copies  : best %.1f GB/s, median %.1f GB/s" % (2 * n * 4 / min(iso) / 1e6, 2 * n * 4 / sorted(iso)[5] / 1e6))
for k in (10, 100, 1000):
    t = run(k)
    print("back-to-back x%4d: %.1f GB/s" % (k, 2 * n * 4 / t / 1e6))
