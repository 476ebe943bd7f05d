import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from safevla_b200 import ops, _lib as L
dev = torch.device("cuda:0"); bf = torch.bfloat16
def bench(name, M, N, K):
    A = torch.randn(M, K, device=dev, dtype=bf); B = torch.randn(N, K, device=dev, dtype=bf)
    C = torch.zeros(M, N, device=dev, dtype=bf)
    f = lambda: ops.gemm(A, B, C, impl=2)
    for _ in range(3): f()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(9)]
    e[0].record()
    for i in range(8):
        f(); e[i+1].record()
    torch.cuda.synchronize()
    t = min(e[i].elapsed_time(e[i+1]) for i in range(8)) * 1e-3
    print(f"dbg={os.environ.get('SVLA_TC_DBG','0')} {name:20s} {t*1e6:8.1f} us  {2*M*N*K/t/1e12:7.1f} TF/s")
M = 119808
bench("K512 N2048", M, 2048, 512)
bench("K512 N512", M, 512, 512)
