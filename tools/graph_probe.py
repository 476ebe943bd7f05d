#!/usr/bin/env python
"""Host-side timing of the phases of PPOLagUpdater.update at a per-rank shape, eager vs CUDA-graph replay."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from safevla_b200.model import B200SafeActorCritic
from safevla_b200.storage import B200RolloutStorage
from safevla_b200.synthetic import RolloutSpec, make_rollout
from safevla_b200 import updater as U

dev = torch.device("cuda:0")
T, N, A, C = 128, int(sys.argv[1]) if len(sys.argv) > 1 else 8, 20, 1
ro = make_rollout(RolloutSpec(T, N, A, C, seed=1234))
g = torch.Generator().manual_seed(0)
vp, cvp = torch.randn(T + 1, N, 1, generator=g), torch.randn(T + 1, N, 1, generator=g).abs()
logp = -3.0 + 0.05 * torch.randn(T, N, generator=g)
for mode in (False, True):
    model = B200SafeActorCritic(A, C, precision="bf16", seed=0, device=dev, chunk_rows=4096, extras="off", verify_dedupe=False)
    upd = U.PPOLagUpdater(model, U.PPOLagConfig(update_repeats=4, cuda_graphs=mode))
    st = B200RolloutStorage(T, dev)
    st.load_rollout(ro, vp, cvp, logp)
    tim = {}
    def wrap(obj, name, key):
        fn = getattr(obj, name)
        def w(*a, **k):
            t0 = time.perf_counter(); r = fn(*a, **k); tim[key] = tim.get(key, 0.0) + time.perf_counter() - t0; return r
        setattr(obj, name, w)
    wrap(model, "prepare", "prepare")
    wrap(upd, "_repeat", "repeat_eager")
    wrap(upd, "_repeat_graphed", "repeat_graphed")
    wrap(upd, "_reduce_clip_step", "reduce_clip")
    wrap(st, "before_updates", "gae")
    for sync in (True, False):
        for it in range(8):
            if it == 3:
                tim.clear(); torch.cuda.synchronize(); t0 = time.perf_counter()
                e0 = torch.cuda.Event(enable_timing=True); e0.record()
            model._ctx_cache = None
            upd.update(st)
            if sync:
                torch.cuda.synchronize()
        e1 = torch.cuda.Event(enable_timing=True); e1.record(); torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) / 5
        print(f"graphs={mode} sync_each_step={sync}: wall {wall*1e3:.1f} ms/step, device {e0.elapsed_time(e1)/5:.1f} ms/step; host per step:",
              {k: round(v / 5 * 1e3, 1) for k, v in tim.items()})
    del model, upd, st
    torch.cuda.empty_cache()
