"""Host-to-device bandwidth of this box from page-locked memory: what bounds ccrs_problem_create*'s upload."""
import torch, time
dev = torch.device("cuda", 0)
for mb in (0.3, 4, 12, 20, 64):
    n = int(mb * (1 << 20))
    h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    d = torch.empty(n, dtype=torch.uint8, device=dev)
    for _ in range(3):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    e0.record()
    for _ in range(reps):
        d.copy_(h, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    t0 = time.perf_counter(); d.copy_(h, non_blocking=True); torch.cuda.synchronize(); wall = (time.perf_counter() - t0) * 1e3
    print(f"{mb:5.1f} MB: {ms*1e3:7.1f} us per copy back to back = {n/ms/1e6:5.1f} GB/s; one copy + synchronize, wall: {wall*1e3:7.1f} us")
