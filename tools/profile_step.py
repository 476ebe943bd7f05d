#!/usr/bin/env python
"""Per-entry-point device-time breakdown of one PPO-Lagrangian update (CUDA events around every C-ABI call).

    python tools/profile_step.py [--samplers 64] [--steps 128] [--precision bf16] [--out profiles/x.json]

Complements the ncu launch list: ncu serialises and cold-caches every launch (and costs ~25 ms per kernel on the
pool's boxes), this runs the real schedule at full speed."""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--samplers", type=int, default=64)
    ap.add_argument("--steps", type=int, default=128)
    ap.add_argument("--actions", type=int, default=20)
    ap.add_argument("--cameras", type=int, default=1)
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--repeats", type=int, default=1)
    ap.add_argument("--chunk-rows", type=int, default=1024)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    from safevla_b200 import _lib as L
    from safevla_b200.model import B200SafeActorCritic
    from safevla_b200.storage import B200RolloutStorage
    from safevla_b200.synthetic import RolloutSpec, make_rollout
    from safevla_b200.updater import PPOLagConfig, PPOLagUpdater

    dev = torch.device("cuda:0")
    T, N, A, C = args.steps, args.samplers, args.actions, args.cameras
    model = B200SafeActorCritic(A, C, precision=args.precision, device=dev, extras="off", verify_dedupe=False,
                                chunk_rows=args.chunk_rows)
    upd = PPOLagUpdater(model, PPOLagConfig(update_repeats=args.repeats))
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
    L.profile_start()
    e0.record()
    upd.update(st)
    e1.record()
    prof = L.profile_stop()
    total = e0.elapsed_time(e1)
    rows = sorted(prof.items(), key=lambda kv: -kv[1][0])
    covered = sum(v[0] for v in prof.values())
    print(f"one update ({args.repeats} repeat(s), T={T} N={N} A={A} C={C} {args.precision}): {total:.2f} ms wall on device; "
          f"{covered:.2f} ms inside C-ABI calls ({len(prof)} entry points, {sum(v[1] for v in prof.values())} calls)")
    for name, (ms, n) in rows:
        print(f"  {name:28s} {ms:10.3f} ms  {100 * ms / total:6.2f} %  {n:6d} calls  {1e3 * ms / n:9.1f} us/call")
    if args.out:
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        json.dump({"total_ms": total, "config": vars(args),
                   "entry_points": {k: {"ms": v[0], "calls": v[1]} for k, v in rows}}, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
