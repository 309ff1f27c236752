"""Read-only / write-only / copy HBM bandwidth with torch kernels (context for the ESPCN layer rooflines)."""
import torch
n = 1 << 30
a = torch.empty(n, dtype=torch.bfloat16, device="cuda"); b = torch.empty_like(a)
def tm(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
t = tm(lambda: a.zero_()); print(f"write-only  {2*n/t/1e6:8.0f} GB/s")
t = tm(lambda: a.fill_(1.5)); print(f"fill        {2*n/t/1e6:8.0f} GB/s")
t = tm(lambda: torch.sum(a.view(torch.int16).view(-1, 1 << 20), dim=1)); print(f"read-only   {2*n/t/1e6:8.0f} GB/s")
t = tm(lambda: b.copy_(a)); print(f"copy (r+w)  {4*n/t/1e6:8.0f} GB/s")
