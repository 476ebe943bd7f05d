#!/usr/bin/env python
"""Launches each roofline-relevant kernel a few times at its measurement shape, for short `ncu --set full`
captures (`ncu -k regex:<kernel> -s 2 -c 1 ... python tools/ncu_targets.py <target>`).

targets: gae | loss | adam | ln | gemm_fwd | gemm_fwd_bits | gemm_dgrad | gemm_dgrad_mask | gemm_dgrad_bits | gemm_res |
         gemm_wgrad | gemm_x3 | attn | attn_x3 | attn_drop | split | ln_fwd | attn_cls | vision
"""
from __future__ import annotations

import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from safevla_b200 import _lib as L  # noqa: E402
from safevla_b200 import ops  # noqa: E402


def main():
    target = sys.argv[1] if len(sys.argv) > 1 else "gae"
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    dev = torch.device("cuda:0")
    bf = torch.bfloat16
    if target == "gae":
        T, N = 128, 65536
        r, c = torch.randn(T, N, device=dev), torch.rand(T, N, device=dev)
        v, vc = torch.randn(T + 1, N, device=dev), torch.randn(T + 1, N, device=dev)
        m = (torch.rand(T + 1, N, device=dev) > 0.02).float()
        out = (torch.empty_like(v), torch.empty_like(vc), torch.empty_like(r), torch.empty_like(c))
        fn = lambda: ops.gae_dual(r, c, v, vc, m, 0.99, 0.95, 1, out=out)  # noqa: E731
    elif target == "loss":
        R, A = 128 * 65536, 20
        logits = torch.randn(R, A, device=dev)
        actions = torch.randint(0, A, (R,), device=dev)
        z = [torch.randn(R, device=dev) for _ in range(5)]
        lam = torch.full((1,), 0.1, device=dev)
        hp = L.PpoHparams(0.1, 1.0, 0.5, 0.0, 0.0, 1.0 / R, 1.0, 0, 1)
        fn = lambda: ops.ppo_lag_fwd_bwd(logits, actions, z[0], z[1], z[2], z[3], z[4], None, None, lam, hp)  # noqa: E731
    elif target == "adam":
        n = 62_900_000 // 64 * 64
        p, g, m, v = [torch.randn(n, device=dev) * 0.01 for _ in range(4)]
        v.abs_()
        sh = torch.empty(n, device=dev, dtype=bf)
        sq = torch.ones(1, device=dev)
        hp = L.AdamHparams(2e-5, 0.9, 0.999, 1e-8, 0.5, 1.0, 1, 1)
        fn = lambda: (ops.sq_norm(g, sq), ops.clip_adam(p, g, m, v, sh, sq, hp))  # noqa: E731
    elif target == "ln":
        rows, D = 119808, 512
        x, dy = torch.randn(rows, D, device=dev, dtype=bf), torch.randn(rows, D, device=dev, dtype=bf)
        y, dx = torch.empty_like(x), torch.empty_like(x)
        gam, bet = torch.ones(D, device=dev), torch.zeros(D, device=dev)
        mean, rstd = torch.empty(rows, device=dev), torch.empty(rows, device=dev)
        dg, db = torch.zeros(D, device=dev), torch.zeros(D, device=dev)
        fn = lambda: (ops.layernorm_fwd(x, gam, bet, y, mean=mean, rstd=rstd),  # noqa: E731
                      ops.layernorm_bwd(dy, x, gam, bet, mean, rstd, dx, dg, db))
    elif target.startswith("gemm"):
        M, N, K = 119808, 2048, 512
        if target == "gemm_fwd":
            a, b = torch.randn(M, K, device=dev, dtype=bf), torch.randn(N, K, device=dev, dtype=bf)
            out = torch.empty(M, N, device=dev, dtype=bf)
            bias = torch.zeros(N, device=dev)
            fn = lambda: ops.gemm(a, b, out, trans_b=True, bias=bias, epilogue=L.EPI_RELU)  # noqa: E731
        elif target == "gemm_fwd_bits":  # linear1 forward: ReLU + one-bit-per-element record
            a, b = torch.randn(M, K, device=dev, dtype=bf), torch.randn(N, K, device=dev, dtype=bf)
            out, bits = torch.empty(M, N, device=dev, dtype=bf), torch.empty(M, N // 32, device=dev, dtype=torch.int32)
            bias = torch.zeros(N, device=dev)
            fn = lambda: ops.gemm(a, b, out, trans_b=True, bias=bias, epilogue=L.EPI_RELU_BITS, aux=bits)  # noqa: E731
        elif target == "gemm_dgrad_bits":  # FFN-down dgrad masked by the bit record (replaces gemm_dgrad_mask on the path)
            a, b = torch.randn(M, K, device=dev, dtype=bf), torch.randn(K, N, device=dev, dtype=bf)
            out = torch.empty(M, N, device=dev, dtype=bf)
            bits = torch.randint(-2 ** 31, 2 ** 31 - 1, (M, N // 32), device=dev, dtype=torch.int32)
            fn = lambda: ops.gemm(a, b, out, trans_b=False, aux=bits, epilogue=L.EPI_MASK_BITS)  # noqa: E731
        elif target == "gemm_res":  # out-proj forward: K = 512, N = 512, bias + residual (HBM-bound: 3 KB per row)
            a, b = torch.randn(M, K, device=dev, dtype=bf), torch.randn(K, K, device=dev, dtype=bf)
            out, res = torch.empty(M, K, device=dev, dtype=bf), torch.randn(M, K, device=dev, dtype=bf)
            bias = torch.zeros(K, device=dev)
            fn = lambda: ops.gemm(a, b, out, trans_b=True, bias=bias, residual=res)  # noqa: E731
        elif target == "gemm_x3":  # parity-grade mode: fp32 operands as three split-bf16 products in one launch
            a, b = torch.randn(M, K, device=dev), torch.randn(N, K, device=dev)
            out = torch.empty(M, N, device=dev)
            bias = torch.zeros(N, device=dev)
            fn = lambda: ops.gemm(a, b, out, trans_b=True, bias=bias, epilogue=L.EPI_RELU, split=3)  # noqa: E731
        elif target == "gemm_dgrad_mask":  # FFN-down dgrad with the fused ReLU mask (side operand = hf)
            a, b = torch.randn(M, K, device=dev, dtype=bf), torch.randn(K, N, device=dev, dtype=bf)
            out, aux = torch.empty(M, N, device=dev, dtype=bf), torch.randn(M, N, device=dev, dtype=bf)
            fn = lambda: ops.gemm(a, b, out, trans_b=False, aux=aux, epilogue=L.EPI_RELU_MASK)  # noqa: E731
        elif target == "gemm_dgrad":
            a, b = torch.randn(M, N, device=dev, dtype=bf), torch.randn(N, K, device=dev, dtype=bf)
            out = torch.empty(M, K, device=dev, dtype=bf)
            fn = lambda: ops.gemm(a, b, out, trans_b=False)  # noqa: E731
        else:
            a, b = torch.randn(M, N, device=dev, dtype=bf), torch.randn(M, K, device=dev, dtype=bf)
            out = torch.zeros(N, K, device=dev)
            fn = lambda: ops.gemm(a, b, out, trans_a=True, trans_b=False, accumulate=True)  # noqa: E731
    elif target == "attn":
        B, S, D = 1024, 117, 512
        qkv = torch.randn(B * S, 3 * D, device=dev, dtype=bf) * 0.5
        o, do = torch.empty(B * S, D, device=dev, dtype=bf), torch.randn(B * S, D, device=dev, dtype=bf)
        dqkv = torch.empty_like(qkv)
        lse = torch.empty(B * 8 * S, device=dev)
        fn = lambda: (ops.attn_fwd(0, qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], o, lse, B, S),  # noqa: E731
                      ops.attn_bwd(0, qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], o, do, dqkv[:, :D],
                                   dqkv[:, D:2 * D], dqkv[:, 2 * D:], lse, B, S))
    elif target in ("attn_x3", "attn_drop"):
        B, S, D = 1024, 117, 512
        dt_ = torch.float32 if target == "attn_x3" else bf
        qkv = torch.randn(B * S, 3 * D, device=dev, dtype=dt_) * 0.5
        o, do = torch.empty(B * S, D, device=dev, dtype=dt_), torch.randn(B * S, D, device=dev, dtype=dt_)
        dqkv = torch.empty_like(qkv)
        lse = torch.empty(B * 8 * S, device=dev)
        kw = dict(split=3) if target == "attn_x3" else dict(drop=ops.dropout_spec(0.1, 1, 0, 1))
        fn = lambda: (ops.attn_fwd(0, qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], o, lse, B, S, **kw),  # noqa: E731
                      ops.attn_bwd(0, qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], o, do, dqkv[:, :D],
                                   dqkv[:, D:2 * D], dqkv[:, 2 * D:], lse, B, S, **kw))
    elif target == "ln_fwd":
        R, D = 119808, 512
        x, y = torch.randn(R, D, device=dev).to(bf), torch.empty(R, D, device=dev, dtype=bf)
        g, b = torch.ones(D, device=dev), torch.zeros(D, device=dev)
        mean, rstd = torch.empty(R, device=dev), torch.empty(R, device=dev)
        fn = lambda: ops.layernorm_fwd(x, g, b, y, mean=mean, rstd=rstd)  # noqa: E731
    elif target == "attn_cls":  # CLS-row attention of the last fusion layer (one query per sequence and head)
        B, S, D = 1024, 117, 512
        kv = torch.randn(B * S, 2 * D, device=dev).to(bf) * 0.5
        q0, o, do = [torch.randn(B, D, device=dev).to(bf) * 0.5 for _ in range(3)]
        dq0, dkv = torch.empty_like(q0), torch.empty_like(kv)
        lse = torch.empty(B * 8, device=dev)
        fn = lambda: (ops.attn_cls_fwd(q0, kv[:, :D], kv[:, D:], o, lse, B, S),  # noqa: E731
                      ops.attn_cls_bwd(q0, kv[:, :D], kv[:, D:], o, do, dq0, dkv[:, :D], dkv[:, D:], lse, B, S))
    elif target == "vision":  # rollout-side DINOv2 ViT-S/14 preprocessor (attn_flash + K = 384 GEMMs), 256 frames
        from safevla_b200.vision import B200DinoViTPreprocessor, init_hub_state_dict
        pre = B200DinoViTPreprocessor("rgb", init_hub_state_dict(0), precision="bf16", device=dev)
        fr = torch.randint(0, 256, (256, 224, 384, 3), dtype=torch.uint8, device=dev)
        fn = lambda: pre.encode(fr)  # noqa: E731
    elif target == "split":
        x = torch.randn(119808, 512, device=dev)
        fn = lambda: ops.split_concat(x, 512, 119808, 512, 1, (0, 1, 0))  # noqa: E731
    else:
        raise SystemExit(f"unknown target {target}")
    e = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    e[0].record()
    for i in range(reps):
        fn()
        e[i + 1].record()
    torch.cuda.synchronize()
    print(target, "ms per call:", [round(e[i].elapsed_time(e[i + 1]), 4) for i in range(reps)])


if __name__ == "__main__":
    main()
