"""Measure pure-write HBM bandwidth (what bounds the dense adjacency output) next to copy bandwidth."""
import torch
dev = "cuda:0"
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e-3
for mb in (164, 512, 2048):
    n = mb * 1024 * 1024 // 4
    a = torch.empty(n, dtype=torch.float32, device=dev)
    b = torch.empty(n, dtype=torch.float32, device=dev)
    big = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)   # 256 MB L2 flusher
    t = timeit(lambda: a.zero_())
    print(f"zero_  {mb:5d} MB: {t*1e6:8.1f} us  {n*4/t/1e12:.2f} TB/s (write only)")
    t = timeit(lambda: a.fill_(1.5))
    print(f"fill_  {mb:5d} MB: {t*1e6:8.1f} us  {n*4/t/1e12:.2f} TB/s (write only)")
    t = timeit(lambda: b.copy_(a))
    print(f"copy_  {mb:5d} MB: {t*1e6:8.1f} us  {2*n*4/t/1e12:.2f} TB/s (read+write)")
    def alt():
        a.zero_(); big.zero_()
    t = timeit(alt)
    print(f"zero_ {mb} MB + 256 MB flusher: {t*1e6:8.1f} us -> {(n*4+big.numel()*4)/t/1e12:.2f} TB/s")
