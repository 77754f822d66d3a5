"""Microbenchmark: HBM write-only / copy bandwidth on this GPU (context for the fill kernel's ceiling)."""
import torch, time
dev = torch.device("cuda")
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(n):
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best
for mb in (512, 2048):
    n = mb * 1024 * 1024 // 8
    x = torch.empty(n, dtype=torch.float64, device=dev); y = torch.empty_like(x)
    t = timeit(lambda: x.fill_(1.5)); print(f"fill_ {mb} MB: {mb/1024/t*1e3*1.073741824:.0f} GB/s ({t*1e3:.1f} us)")
    t = timeit(lambda: x.zero_()); print(f"zero_ {mb} MB: {mb/1024/t*1e3*1.073741824:.0f} GB/s")
    t = timeit(lambda: y.copy_(x)); print(f"copy_ {mb} MB: {2*mb/1024/t*1e3*1.073741824:.0f} GB/s (read+write)")
    t = timeit(lambda: torch.mul(x, 2.0, out=y)); print(f"mul out {mb} MB: {2*mb/1024/t*1e3*1.073741824:.0f} GB/s (read+write)")
    t = timeit(lambda: x.sum()); print(f"sum (read only) {mb} MB: {mb/1024/t*1e3*1.073741824:.0f} GB/s")
