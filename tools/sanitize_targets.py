#!/usr/bin/env python
"""One small launch of every kernel family written or changed in round 2, for compute-sanitizer:

    compute-sanitizer --tool memcheck  python tools/sanitize_targets.py
    compute-sanitizer --tool racecheck python tools/sanitize_targets.py
    compute-sanitizer --tool synccheck python tools/sanitize_targets.py

(race-detection row: the tcgen05 kernels use hand-rolled mbarrier protocols and relaxed arrives;
run-to-run determinism of whole updates is tested elsewhere, this is the tool-level pass.)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from safevla_b200 import _lib as L
from safevla_b200 import ops

dev = torch.device("cuda:0")
bf = torch.bfloat16
g = torch.Generator().manual_seed(0)
H, D = 8, 512


def attn(mode, S, B, dt, **kw):
    qkv = (torch.randn(B * S, 3 * D, generator=g) * 0.5).to(dev, dt)
    o, do = torch.empty(B * S, D, device=dev, dtype=dt), torch.randn(B * S, D, generator=g).to(dev, dt)
    lse, dqkv = torch.empty(B * H * S, device=dev), torch.empty_like(qkv)
    traj = torch.cumsum((torch.rand(B, S, generator=g) < 0.1).long(), 1).to(dev) if mode == 1 else None
    ops.attn_fwd(mode, qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], o, lse, B, S, traj=traj, **kw)
    ops.attn_bwd(mode, qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], o, do, dqkv[:, :D], dqkv[:, D:2 * D], dqkv[:, 2 * D:],
                 lse, B, S, traj=traj, **kw)
    torch.cuda.synchronize()
    assert torch.isfinite(dqkv.float()).all()


attn(0, 117, 3, bf)
attn(1, 128, 2, bf)
attn(0, 33, 5, bf, drop=ops.dropout_spec(0.1, 1, 2, 3))
attn(0, 117, 2, torch.float32, split=3)
attn(1, 64, 2, torch.float32, split=3)
attn(0, 201, 2, bf)                                        # two-tile kernels (two-camera fusion block)
attn(0, 201, 2, bf, drop=ops.dropout_spec(0.1, 1, 2, 3))   # ... with dropout (256-wide mask rows)
# CLS-row attention of the last fusion layer, with and without dropout, one and two cameras
for S_, dr in ((117, None), (117, ops.dropout_spec(0.1, 3, 16, 2)), (201, ops.dropout_spec(0.1, 3, 16, 2))):
    B_ = 5
    kv = (torch.randn(B_ * S_, 2 * D, generator=g) * 0.5).to(dev, bf)
    q0, o0, do0 = [(torch.randn(B_, D, generator=g) * 0.5).to(dev, bf) for _ in range(3)]
    dq0, dkv, lse0 = torch.empty_like(q0), torch.empty_like(kv), torch.empty(B_ * H, device=dev)
    ops.attn_cls_fwd(q0, kv[:, :D], kv[:, D:], o0, lse0, B_, S_, drop=dr)
    ops.attn_cls_bwd(q0, kv[:, :D], kv[:, D:], o0, do0, dq0, dkv[:, :D], dkv[:, D:], lse0, B_, S_, drop=dr)
    torch.cuda.synchronize()
    assert torch.isfinite(dkv.float()).all()
print("attention ok")

M, N, K = 512, 512, 256
x = torch.randn(M, K, generator=g).to(dev, bf)
w = torch.randn(N, K, generator=g).to(dev, bf)
b = torch.randn(N, generator=g).to(dev)
out, bits = torch.empty(M, N, device=dev, dtype=bf), torch.empty(M, N // 32, device=dev, dtype=torch.int32)
ops.gemm(x, w, out, bias=b, epilogue=L.EPI_RELU_BITS, aux=bits)
ops.gemm(x, w, out, bias=b, epilogue=L.EPI_RELU_BITS, aux=bits, dropout=ops.dropout_spec(0.1, 5, 1, 1))
dx = torch.empty(M, N, device=dev, dtype=bf)
ops.gemm(x, torch.randn(K, N, generator=g).to(dev, bf), dx, trans_b=False, aux=bits, epilogue=L.EPI_MASK_BITS, alpha=1.1)
# ragged bit records: 18 words per row (8-byte record path), a half-empty last tile, M not a multiple of 32
for M2, N2 in ((600, 576), (1000, 640), (300 + 7, 256)):
    x2, w2 = torch.randn(M2, K, generator=g).to(dev, bf), torch.randn(N2, K, generator=g).to(dev, bf)
    o2, b2 = torch.empty(M2, N2, device=dev, dtype=bf), torch.empty(M2, N2 // 32, device=dev, dtype=torch.int32)
    ops.gemm(x2, w2, o2, bias=torch.zeros(N2, device=dev), epilogue=L.EPI_RELU_BITS, aux=b2)
    ops.gemm(x2, torch.randn(K, N2, generator=g).to(dev, bf), torch.empty_like(o2), trans_b=False, aux=b2,
             epilogue=L.EPI_MASK_BITS)
res = torch.randn(M, N, generator=g).to(dev, bf)
ops.gemm(x, w, out, bias=b, residual=res)
xf, wf = torch.randn(M, K, generator=g).to(dev), torch.randn(N, K, generator=g).to(dev)
of = torch.empty(M, N, device=dev)
ops.gemm(xf, wf, of, bias=b, split=3)
ops.gemm(xf, torch.randn(M, N, generator=g).to(dev), torch.zeros(K, N, device=dev), trans_a=True, trans_b=False, accumulate=True,
         colsum_a=torch.zeros(K, device=dev), split=3)
torch.cuda.synchronize()
print("gemm ok")

ops.dropout_rows(xf, torch.empty_like(xf), ops.dropout_spec(0.1, 1, 1, 1, 7))
ops.split_concat(xf, K, M, K, 0, (0, 1, 2))
T_, N_ = 37, 65
r, v = torch.randn(T_, N_, device=dev), torch.randn(T_ + 1, N_, device=dev)
m = (torch.rand(T_ + 1, N_, device=dev) > 0.1).float()
ops.discounted_returns_dual(r, r.abs(), v, v.abs(), m, 0.99)
ops.gae_dual(r, r.abs(), v, v.abs(), m, 0.99, 0.95)
torch.cuda.synchronize()
print("elementwise ok")
