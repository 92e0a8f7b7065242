"""Does a small H2D copy on one stream wait behind a bulk H2D copy on another stream (same copy engine)?
Measures the latency of a 16 KB pinned H2D copy + a trivial kernel on stream B, alone and while an
8.2 MB pinned H2D copy is in flight on stream A; and the same with the small block read by the kernel
straight from mapped pinned memory (no copy engine)."""
import time
import torch

dev = torch.device("cuda:0")
big_h = torch.empty(8_232_000, dtype=torch.uint8).pin_memory()
big_d = torch.empty_like(big_h, device=dev)
small_h = torch.zeros(16384, dtype=torch.uint8).pin_memory()
small_d = torch.empty_like(small_h, device=dev)
out_d = torch.zeros(16384, dtype=torch.uint8, device=dev)
sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize()


def trial(with_big, mode):
    ts = []
    for _ in range(30):
        torch.cuda.synchronize()
        if with_big:
            with torch.cuda.stream(sa):
                big_d.copy_(big_h, non_blocking=True)
        t0 = time.perf_counter()
        with torch.cuda.stream(sb):
            if mode == "copy":
                small_d.copy_(small_h, non_blocking=True)
                out_d.add_(small_d)
            else:
                out_d.add_(1)
        sb.synchronize()
        ts.append(time.perf_counter() - t0)
        torch.cuda.synchronize()
    ts.sort()
    return 1e6 * ts[len(ts) // 2]


for mode in ("copy", "kernel_only"):
    for with_big in (False, True):
        print(f"{mode:12s} bulk copy in flight={with_big}: small op latency {trial(with_big, mode):8.1f} us (median of 30)")
t0 = time.perf_counter()
for _ in range(20):
    big_d.copy_(big_h, non_blocking=True)
torch.cuda.synchronize()
print(f"bulk 8.2 MB H2D alone: {1e6 * (time.perf_counter() - t0) / 20:.1f} us")
