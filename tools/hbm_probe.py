#!/usr/bin/env python
"""Write-only / read-only / copy HBM throughput of this GPU (context for the write-heavy K = 512 GEMM epilogues)."""
import torch
dev = torch.device("cuda:0")
n = 1 << 29  # 2 GB of fp32
x, y = torch.empty(n, device=dev), torch.empty(n, device=dev)
def t_of(fn, it=10):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it * 1e-3
tw = t_of(lambda: x.zero_())
tr = t_of(lambda: x.sum())
tc = t_of(lambda: y.copy_(x))
tm = t_of(lambda: torch.mul(x, 2.0, out=y))
print(f"write-only {4*n/tw/1e12:.2f} TB/s   read-only {4*n/tr/1e12:.2f} TB/s   copy {8*n/tc/1e12:.2f} TB/s   scale(r+w) {8*n/tm/1e12:.2f} TB/s")
