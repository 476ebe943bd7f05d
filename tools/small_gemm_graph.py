#!/usr/bin/env python
"""Device time of the small decoder-shaped GEMMs (M = 1024 ... 8192), measured inside a CUDA graph of 64 launches so
that host launch overhead does not hide it.   SVLA_SMALL_TC=0/1 toggles the small-problem tile heuristic."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from safevla_b200 import ops
dev = torch.device("cuda:0"); bf = torch.bfloat16
def bench(M, N, K, out_f32=False, res=False):
    A = torch.randn(M, K, device=dev, dtype=bf); B = torch.randn(N, K, device=dev, dtype=bf)
    C = torch.zeros(M, N, device=dev, dtype=torch.float32 if out_f32 else bf)
    r = torch.randn(M, N, device=dev, dtype=C.dtype) if res else None
    f = lambda: ops.gemm(A, B, C, residual=r, impl=2)
    f(); torch.cuda.synchronize()
    s = torch.cuda.Stream()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s):
        f()
        with torch.cuda.graph(g, stream=s):
            for _ in range(64):
                f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g.replay(); torch.cuda.synchronize()
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 64 * 1e3
    print(f"M={M:6d} N={N:5d} K={K:5d} f32={int(out_f32)} res={int(res)}: {t:7.2f} us/launch  {2*M*N*K/t/1e6:7.1f} TF/s")
for M in (1024, 2048, 8192):
    bench(M, 1536, 512); bench(M, 512, 512, True, True); bench(M, 3072, 512); bench(M, 512, 1536, True, True)
bench(1024, 2048, 512); bench(1024, 512, 2048)
