#!/usr/bin/env python
"""tcgen05 attention forward / backward timings at the towers' shapes (fusion S=117 / 201, decoder T=128 / 256)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from safevla_b200 import ops
from safevla_b200 import _lib
dev = torch.device("cuda:0"); bf = torch.bfloat16
# SVLA_ATTN_IMPL: 0 = default (warp-specialised kernels for S <= 128), 3 = the round-1 one-CTA-per-item kernels
_lib.load_library().svla_set_attn_impl(int(os.environ.get("SVLA_ATTN_IMPL", "0")))
print("attention impl", os.environ.get("SVLA_ATTN_IMPL", "0"))

def t_of(fn, n=8):
    for _ in range(3):
        fn()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    e[0].record()
    for i in range(n):
        fn(); e[i + 1].record()
    torch.cuda.synchronize()
    return sum(e[i].elapsed_time(e[i + 1]) for i in range(n)) / n * 1e-3

for mode, B, S in ((0, 4096, 117), (0, 1024, 117), (0, 4096, 128), (0, 2048, 201), (1, 64, 128), (1, 1024, 128), (1, 8, 256)):
    D, H = 512, 8
    qkv = torch.randn(B * S, 3 * D, device=dev, dtype=bf) * 0.5
    o, do = torch.empty(B * S, D, device=dev, dtype=bf), torch.randn(B * S, D, device=dev, dtype=bf)
    dqkv = torch.empty_like(qkv)
    lse = torch.empty(B * H * S, device=dev)
    traj = torch.cumsum((torch.rand(B, S, device=dev) < 0.02).long(), 1) if mode == 1 else None
    f = lambda: ops.attn_fwd(mode, qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], o, lse, B, S, traj=traj)
    g = lambda: ops.attn_bwd(mode, qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], o, do, dqkv[:, :D], dqkv[:, D:2 * D],
                             dqkv[:, 2 * D:], lse, B, S, traj=traj)
    tf, tb = t_of(f), t_of(g)
    bytes_f, bytes_b = B * S * D * 2 * 4, B * S * D * 2 * 8
    fl = 4.0 * S * S * 64 * H * B
    print(f"mode {mode} B={B:5d} S={S:3d}  fwd {tf*1e6:8.1f} us {bytes_f/tf/1e9:6.0f} GB/s {fl/tf/1e12:6.1f} TF/s   "
          f"bwd {tb*1e6:8.1f} us {bytes_b/tb/1e9:6.0f} GB/s {2.5*fl/tb/1e12:6.1f} TF/s")
