#!/usr/bin/env python
"""GEMM launches of one PPO-Lagrangian update grouped by shape / operand layout: time, share, achieved TFLOP/s.

    python tools/gemm_breakdown.py [--samplers 64] [--steps 128] [--chunk-rows 4096] [--cameras 1]
"""
from __future__ import annotations

import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--samplers", type=int, default=64)
    ap.add_argument("--steps", type=int, default=128)
    ap.add_argument("--cameras", type=int, default=1)
    ap.add_argument("--chunk-rows", type=int, default=4096)
    args = ap.parse_args()
    from safevla_b200 import ops
    from safevla_b200.model import B200SafeActorCritic
    from safevla_b200.storage import B200RolloutStorage
    from safevla_b200.synthetic import RolloutSpec, make_rollout
    from safevla_b200.updater import PPOLagConfig, PPOLagUpdater

    dev = torch.device("cuda:0")
    T, N, A, C = args.steps, args.samplers, 20, args.cameras
    model = B200SafeActorCritic(A, C, precision="bf16", device=dev, extras="off", verify_dedupe=False,
                                chunk_rows=args.chunk_rows)
    upd = PPOLagUpdater(model, PPOLagConfig(update_repeats=1))
    ro = make_rollout(RolloutSpec(T, N, A, C, seed=1234))
    g = torch.Generator().manual_seed(0)
    st = B200RolloutStorage(T, dev)
    st.load_rollout(ro, torch.randn(T + 1, N, 1, generator=g), torch.randn(T + 1, N, 1, generator=g),
                    -3.0 + 0.01 * torch.randn(T, N, generator=g))
    for _ in range(2):
        model._ctx_cache = None
        upd.update(st)
    torch.cuda.synchronize()
    model._ctx_cache = None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ops.PROFILE = {}
    e0.record()
    upd.update(st)
    e1.record()
    torch.cuda.synchronize()
    prof, ops.PROFILE = ops.PROFILE, None
    total = e0.elapsed_time(e1)
    groups = {}
    for which, recs in prof.items():
        for fl, a, b, key in recs:
            g_ = groups.setdefault((which,) + key, [0.0, 0.0, 0])
            g_[0] += a.elapsed_time(b)
            g_[1] += fl
            g_[2] += 1
    gsum = sum(v[0] for v in groups.values())
    print(f"update (1 repeat) {total:.2f} ms; GEMM launches {gsum:.2f} ms ({100 * gsum / total:.1f} %)")
    print("kernel M N K tA tB epi acc dtA dtC res colsum : ms share n TF/s")
    for key, (ms, fl, n) in sorted(groups.items(), key=lambda kv: -kv[1][0]):
        print(" ".join(str(k) for k in key), f": {ms:8.3f} ms {100 * ms / total:5.2f}% n={n} {fl / ms / 1e9:7.1f} TF/s")


if __name__ == "__main__":
    main()
