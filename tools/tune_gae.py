import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from safevla_b200 import ops
dev = torch.device("cuda:0")
T, N = 128, 65536
r, c = torch.randn(T, N, device=dev), torch.rand(T, N, device=dev)
v, vc = torch.randn(T + 1, N, device=dev), torch.randn(T + 1, N, device=dev)
m = (torch.rand(T + 1, N, device=dev) > 0.02).float()
out = (torch.empty_like(v), torch.empty_like(vc), torch.empty_like(r), torch.empty_like(c))
for algo in (1, 20, 40, 41, 42, 43, 44, 45, 46, 47, 48, 49):
    for _ in range(3):
        ops.gae_dual(r, c, v, vc, m, 0.99, 0.95, algo, out=out)
    e = [torch.cuda.Event(enable_timing=True) for _ in range(11)]
    e[0].record()
    for i in range(10):
        ops.gae_dual(r, c, v, vc, m, 0.99, 0.95, algo, out=out)
        e[i + 1].record()
    torch.cuda.synchronize()
    t = min(e[i].elapsed_time(e[i + 1]) for i in range(10))
    print(f"algo {algo}: {t*1e3:.1f} us  {36*T*N/t/1e6:.0f} GB/s")
