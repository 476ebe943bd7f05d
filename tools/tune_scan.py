#!/usr/bin/env python
"""GAE + fused-loss kernels at the roofline-scale shape (T=128, N=65536, A=20): achieved algorithmic GB/s per variant."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from safevla_b200 import ops, _lib as L
dev = torch.device("cuda:0")
T, N, A = 128, 65536, 20

def t_of(fn, n=10):
    for _ in range(3):
        fn()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    e[0].record()
    for i in range(n):
        fn(); e[i + 1].record()
    torch.cuda.synchronize()
    return min(e[i].elapsed_time(e[i + 1]) for i in range(n)) * 1e-3

r, c = torch.randn(T, N, device=dev), torch.rand(T, N, device=dev)
v, vc = torch.randn(T + 1, N, device=dev), torch.randn(T + 1, N, device=dev)
m = (torch.rand(T + 1, N, device=dev) > 0.02).float()
out = (torch.empty_like(v), torch.empty_like(vc), torch.empty_like(r), torch.empty_like(c))
for algo in (1, 20, 40, 41, 42, 43, 44, 45, 46, 47, 48, 49):
    t = t_of(lambda: ops.gae_dual(r, c, v, vc, m, 0.99, 0.95, algo, out=out))
    print(f"gae algo {algo}: {t*1e6:.1f} us  {36*T*N/t/1e9:.0f} GB/s")
R = T * N
logits = torch.randn(R, A, device=dev)
actions = torch.randint(0, A, (R,), device=dev)
z = [torch.randn(R, device=dev) for _ in range(5)]
lam = torch.full((1,), 0.1, device=dev)
hp = L.PpoHparams(0.1, 1.0, 0.5, 0.0, 0.0, 1.0 / R, 1.0, 0, 1)
dl, dv, scal = torch.empty_like(logits), torch.empty(R, device=dev), torch.empty(16, device=dev)
lib, ctx = L.load_library(), L.get_ctx()
def loss():
    L.check(lib.svla_ppo_lag_fwd_bwd(ctx, logits.data_ptr(), actions.data_ptr(), z[0].data_ptr(), z[1].data_ptr(),
                                     z[2].data_ptr(), z[3].data_ptr(), z[4].data_ptr(), None, None, None, None,
                                     lam.data_ptr(), hp, scal.data_ptr(), dl.data_ptr(), dv.data_ptr(), None, R, A,
                                     L.stream_ptr()))
t = t_of(loss)
print(f"loss (SVLA_PPO_GENERIC={os.environ.get('SVLA_PPO_GENERIC')}): {t*1e6:.1f} us  {(8*A+44)*R/t/1e9:.0f} GB/s")
# reference points: plain copy of the same byte count
src = torch.empty(int((8 * A + 44) * R / 8), device=dev); dst = torch.empty_like(src)
t = t_of(lambda: dst.copy_(src))
print(f"torch copy of the loss's byte count: {t*1e6:.1f} us  {src.numel()*8/t/1e9:.0f} GB/s")
src = torch.empty(int(36 * R / 8), device=dev); dst = torch.empty_like(src)
t = t_of(lambda: dst.copy_(src))
print(f"torch copy of the GAE's byte count: {t*1e6:.1f} us  {src.numel()*8/t/1e9:.0f} GB/s")
