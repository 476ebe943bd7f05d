#!/usr/bin/env python
"""LayerNorm forward / backward at the fusion block's shape ([rows, 512] bf16): time and HBM throughput."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from safevla_b200 import ops
dev, bf = torch.device("cuda:0"), torch.bfloat16
R = int(sys.argv[1]) if len(sys.argv) > 1 else 479232
D = 512
def t_of(fn, n=20):
    for _ in range(4): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e-3
x, res, y = torch.randn(R, D, device=dev).to(bf), torch.randn(R, D, device=dev).to(bf), torch.empty(R, D, device=dev, dtype=bf)
dy, dx = torch.randn(R, D, device=dev).to(bf), torch.empty(R, D, device=dev, dtype=bf)
g, b = torch.ones(D, device=dev), torch.zeros(D, device=dev)
dg, db = torch.zeros(D, device=dev), torch.zeros(D, device=dev)
mean, rstd = torch.empty(R, device=dev), torch.empty(R, device=dev)
nb = R * D * 2
for name, fn, streams in (
    ("fwd", lambda: ops.layernorm_fwd(x, g, b, y, mean=mean, rstd=rstd), 2),
    ("fwd + residual", lambda: ops.layernorm_fwd(x, g, b, y, res=res, mean=mean, rstd=rstd), 3),
    ("bwd", lambda: ops.layernorm_bwd(dy, x, g, b, mean, rstd, dx, dg, db), 3),
    ("bwd + residual", lambda: ops.layernorm_bwd(dy, x, g, b, mean, rstd, dx, dg, db, res=res), 4)):
    t = t_of(fn)
    print(f"layernorm {name:16s} rows {R}  {t*1e6:8.1f} us  {streams*nb/t/1e12:5.2f} TB/s")
