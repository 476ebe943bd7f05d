#!/usr/bin/env python
"""bench.py -- PPO-Lagrangian update samples/s on synthetic 64-env x 128-step rollouts (BASELINE.json).

A "step" is one whole constrained-PPO update of one rollout: GAE(reward)+GAE(cost) -> update_repeats(4) x
[3-tower forward, fused PPO-Lagrangian loss fwd+bwd, tower backwards, (N>1: one NCCL all-reduce), fused
clip+Adam] -> Lagrange-multiplier update.   samples/s = T * N_global * update_repeats / t_step.

  value : rollout already resident in HBM when the timed region starts (every step is treated as a NEW
          rollout: the observation-derived caches -- token-major features, T5 text encoding -- are rebuilt
          inside the timed region).
  e2e   : same metric through the public host API with HOST buffers: pinned-host rollout -> H2D copy into
          B200RolloutStorage -> PPOLagUpdater.update -> D2H of the loss scalars and lambda, all timed.
  --impl reference : the CPU restatement of the reference path (oracle/, torch eager fp32 on all host cores)
          on a bounded sampler-subsample of the same workload.  The reference itself is Python that needs
          /root/reference, which does not exist on the GPU box; the oracle is pinned against it by
          tests/golden/ (see DESIGN.md).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

WORKLOADS = {
    # BASELINE.json configs[1] / configs[2]: 64-env x 128-step ObjectNav rollouts, one camera, 32-token prompt
    "cfg2_64env_128step": dict(T=128, N=64, A=20, C=1, L=32, gflop_per_sample=18.4),
    # the per-rank shape of configs[2] (64 env over 8 ranks) on ONE GPU: separates small-problem efficiency from
    # communication when reading the 8-GPU number
    "cfg3_rank_8env_128step": dict(T=128, N=8, A=20, C=1, L=32, gflop_per_sample=18.4),
    # configs[0] (CPU-runnable correctness gate)
    "cfg1_1env_16step": dict(T=16, N=1, A=6, C=1, L=32, gflop_per_sample=18.4),
    # configs[3] PickupType head, two cameras (S = 201)
    "cfg4_32env_256step": dict(T=256, N=32, A=20, C=2, L=32, gflop_per_sample=31.5),
    # configs[4]: FetchType multi-task, dual cost channels (K = 2: extension beyond the reference's scalar cost)
    "cfg5_128env_128step": dict(T=128, N=128, A=20, C=2, L=32, K=2, gflop_per_sample=31.5),
}
UPDATE_REPEATS = 4


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


def ncu_traffic():
    """DRAM bytes of the dominant kernel from the committed `ncu --set full` capture (profiles/): one representative
    launch (the FFN-up forward GEMM of a 1024-row chunk), next to its algorithmic bytes."""
    for rnd in ("r02", "r01"):
        p = os.path.join(ROOT, "profiles", f"ncu_{rnd}_traffic.json")
        if not os.path.exists(p):
            continue
        d = json.load(open(p))
        d = d.get("gemm_fwd_bits") or d.get("gemm_fwd") or d.get("svla_gemm_tc2_kernel<0, 0>")
        if not d:
            continue
        return {"traffic": d["dram_bytes"], "traffic_launch": d["shape"],
                "traffic_algorithmic_bytes": d["algorithmic_bytes"], "traffic_src": f"profiles/ncu_{rnd}.md"}
    return {"traffic": None}


class ClockSampler(threading.Thread):
    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.rows, self._halt = index, [], threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._halt.wait(0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=3)
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None,
                "sm_max_mhz": int(self.rows[0][1]) if self.rows and self.rows[0][1].isdigit() else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


# ---------------------------------------------------------------------------------------------------- CPU arm
def cpu_sample(wl, steps: int, warmup: int, sample_T: int, sample_N: int, device: str = "cpu", tf32: bool = False,
               autocast=None):
    """Times the oracle's whole update on a sampler-subsample (sample_N of the N samplers, first sample_T steps, ONE
    update repeat); returns samples/s, seconds per iteration and a description.  device="cuda" runs the same torch-eager
    restatement on the GPU (fp32; optionally TF32 matmuls or bf16 autocast): the stock-library baseline of SURVEY 8d
    (`library_baseline` in the JSON line)."""
    from oracle.update_oracle import oracle_update  # the one place bench.py executes oracle/
    from safevla_b200.params import init_state_dict
    from safevla_b200.synthetic import RolloutSpec, make_rollout
    from safevla_b200.updater import PPOLagConfig

    torch.set_num_threads(os.cpu_count() or 1)
    sd = init_state_dict(wl["A"], wl["C"], seed=0)
    ro = make_rollout(RolloutSpec(sample_T, sample_N, wl["A"], wl["C"], prompt_tokens=wl["L"], seed=1234))
    g = torch.Generator().manual_seed(0)
    vp = torch.randn(sample_T + 1, sample_N, 1, generator=g)
    cvp = torch.randn(sample_T + 1, sample_N, 1, generator=g).abs()
    logp = -3.0 + 0.05 * torch.randn(sample_T, sample_N, generator=g)
    cfg = PPOLagConfig(update_repeats=1)
    on_gpu = device != "cpu"
    if on_gpu:
        def mv(x):
            if isinstance(x, dict):
                return {k: mv(v) for k, v in x.items()}
            return x.to(device) if torch.is_tensor(x) else x
        sd, ro, vp, cvp, logp = mv(sd), mv(ro), vp.to(device), cvp.to(device), logp.to(device)
        torch.set_default_device(device)  # the restatement creates its index / mask tensors with bare factories
    old_tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = bool(tf32)
    times = []
    try:
        for i in range(warmup + steps):
            if on_gpu:
                torch.cuda.synchronize()
            t0 = time.perf_counter()
            if autocast is not None:
                with torch.autocast("cuda", dtype=autocast):
                    oracle_update(sd, ro, vp, cvp, logp, cfg, wl["A"], wl["C"])
            else:
                oracle_update(sd, ro, vp, cvp, logp, cfg, wl["A"], wl["C"])
            if on_gpu:
                torch.cuda.synchronize()
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old_tf32
        if on_gpu:
            torch.set_default_device("cpu")
    t = sum(times) / len(times)
    mode = "bf16 autocast" if autocast is not None else ("TF32 matmuls" if tf32 else "fp32")
    what = (f"CPU oracle (torch eager fp32, {torch.get_num_threads()} threads)" if not on_gpu
            else f"torch-eager oracle on the GPU ({mode})")
    return sample_T * sample_N / t, t, (f"{what}: {sample_N} of {wl['N']} samplers x {sample_T} of {wl['T']} steps, "
                                        f"1 of {UPDATE_REPEATS} update repeats per step, {warmup} warm-up + {steps} "
                                        f"timed iterations")


REF_SAMPLERS = 4  # "4 of 64 samplers, full T" per reference-arm step


def run_reference(args, wl, name):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.ref_device != "cpu":  # stock-library baseline on the GPU: larger sample, same code
        v, t, desc = cpu_sample(wl, args.steps, args.warmup, sample_T=wl["T"], sample_N=min(wl["N"], args.ref_samplers),
                                device=args.ref_device, tf32=args.ref_tf32)
    else:
        v, t, desc = cpu_sample(wl, args.steps, args.warmup, sample_T=wl["T"], sample_N=min(wl["N"], REF_SAMPLERS))
    line = {"impl": "reference" if args.ref_device == "cpu" else "reference-eager-" + args.ref_device, "metric": "ppo_lagrangian_update_samples_per_sec", "value": v, "unit": "samples/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": name, "T": wl["T"], "N_global": wl["N"], "actions": wl["A"], "cameras": wl["C"],
                       "prompt_tokens": wl["L"], "update_repeats": UPDATE_REPEATS},
            "cpu_baseline": {"value": v, "unit": "samples/s", "cores": torch.get_num_threads(), "kind": "port",
                             "sample": desc},
            "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def library_baseline(wl, samplers: int = 8):
    """The same torch-eager restatement of the reference on the B200 through the stock libraries (cuBLAS / ATen
    kernels, autograd): fp32, TF32 matmuls and bf16 autocast.  This is what the reference's own code would reach on
    this GPU, i.e. the baseline the hand-written path has to beat; the CPU arm is the contract's reference arm."""
    out = {}
    for key, kw in (("fp32", {}), ("tf32", {"tf32": True}), ("bf16_autocast", {"autocast": torch.bfloat16})):
        try:
            v, t, desc = cpu_sample(wl, 2, 1, sample_T=wl["T"], sample_N=min(wl["N"], samplers), device="cuda", **kw)
            out[key] = {"value": v, "unit": "samples/s", "ms_per_iteration": t * 1e3, "sample": desc}
        except Exception as e:  # noqa: BLE001  (e.g. out of memory next to the timed model: reported, not fatal)
            out[key] = {"value": None, "error": f"{type(e).__name__}: {e}"[:200]}
        torch.cuda.empty_cache()
    return out


# ---------------------------------------------------------------------------------------------------- GPU arm
def scan_roofline(dev, pk):
    """GAE + fused loss at a roofline-scale shape (T=128, N=65536): achieved algorithmic GB/s."""
    from safevla_b200 import _lib as L
    from safevla_b200 import ops
    T, N, A = 128, 65536, 20
    r, c = torch.randn(T, N, device=dev), torch.rand(T, N, device=dev)
    v, vc = torch.randn(T + 1, N, device=dev), torch.randn(T + 1, N, device=dev)
    m = (torch.rand(T + 1, N, device=dev) > 0.02).float()
    out = (torch.empty_like(v), torch.empty_like(vc), torch.empty_like(r), torch.empty_like(c))
    logits = torch.randn(T * N, A, device=dev)
    actions = torch.randint(0, A, (T * N,), device=dev)
    oldlp = torch.full((T * N,), -3.0, device=dev)
    lam = torch.full((1,), 0.1, device=dev)
    hp = L.PpoHparams(0.1, 1.0, 0.5, 0.0, 0.0, 1.0 / (T * N), 1.0, 0, 1)
    dl, dv = torch.empty_like(logits), torch.empty(T * N, device=dev)
    scal = torch.empty(16, device=dev)

    def t_of(fn, n=20):
        for _ in range(3):
            fn()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
        ev[0].record()
        for i in range(n):
            fn()
            ev[i + 1].record()
        torch.cuda.synchronize()
        return sum(ev[i].elapsed_time(ev[i + 1]) for i in range(n)) / n * 1e-3  # MEAN launch duration

    t_gae = t_of(lambda: ops.gae_dual(r, c, v, vc, m, 0.99, 0.95, 1, out=out))
    lib, ctx = L.load_library(), L.get_ctx()

    def loss():
        L.check(lib.svla_ppo_lag_fwd_bwd(ctx, logits.data_ptr(), actions.data_ptr(), oldlp.data_ptr(),
                                         out[2].data_ptr(), out[3].data_ptr(), v.data_ptr(), out[0].data_ptr(), None,
                                         None, None, None, lam.data_ptr(), hp, scal.data_ptr(), dl.data_ptr(),
                                         dv.data_ptr(), None, T * N, A, L.stream_ptr()))
    t_loss = t_of(loss)
    gae_b, loss_b = 36.0 * T * N, (8.0 * A + 44.0) * T * N
    return {"shape": f"T={T},N={N},A={A}", "peak_gbs": pk["hbm"], "peak_src": pk["src"], "timing": "mean of 20 launches",
            "gae": {"gbs": gae_b / t_gae / 1e9, "frac": gae_b / t_gae / 1e9 / pk["hbm"], "us": t_gae * 1e6},
            "loss": {"gbs": loss_b / t_loss / 1e9, "frac": loss_b / t_loss / 1e9 / pk["hbm"], "us": t_loss * 1e6}}


def run_b200(args, wl, name):
    import torch.distributed as dist

    from safevla_b200 import _lib as L
    from safevla_b200 import ops
    from safevla_b200.model import ACTOR, COST, CRITIC, B200SafeActorCritic
    from safevla_b200.storage import B200RolloutStorage
    from safevla_b200.synthetic import RolloutSpec, make_rollout
    from safevla_b200.updater import PPOLagConfig, PPOLagUpdater

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        # NCCL_DEBUG is left as the caller set it (the driver counts ranks from NCCL's INFO lines); NCCL would print
        # them on stdout, which must carry only the JSON line, so they are routed to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    assert wl["N"] % world == 0, "samplers must divide evenly over the ranks"
    T, A, C = wl["T"], wl["A"], wl["C"]
    n_local = wl["N"] // world
    pk = peaks()

    # GAE / fused-loss roofline shapes first: each kernel timed alone (burst conditions, like the copy that
    # MEASURED_PEAKS.json's HBM figure comes from), not after seconds of power-capped tensor-core work
    scan = scan_roofline(dev, pk) if world == 1 else None
    K = wl.get("K", 1)  # cost channels
    model = B200SafeActorCritic(A, C, precision=args.precision, seed=0, device=dev, chunk_rows=args.chunk_rows,
                                extras="off", verify_dedupe=False, num_cost_channels=K, dropout=args.dropout,
                                dropout_seed=1234)
    cfg = PPOLagConfig(update_repeats=UPDATE_REPEATS, cost_limit=2.31964 if K == 1 else (2.31964,) * K,
                       cuda_graphs=bool(args.cuda_graphs))
    upd = PPOLagUpdater(model, cfg)
    ro = make_rollout(RolloutSpec(T, n_local, A, C, prompt_tokens=wl["L"], seed=1234, num_cost_channels=K), rank=rank,
                      pin=True)
    storage = B200RolloutStorage(T, dev, num_cost_channels=K)

    # ---- "collection": value / cost-value predictions and old log-probs from the model's own forward
    obs_full = {k: v.to(dev) for k, v in ro["observations"].items()}
    pa_full = torch.cat([torch.zeros(1, n_local, dtype=torch.int64), ro["actions"]], 0).to(dev)
    mk_full = ro["masks"].to(dev).view(T + 1, n_local)
    def collect(t0, t1):  # one update-mode forward of the three towers over steps [t0, t1)
        model._ctx_cache = None
        rc = model.prepare({k: v[t0:t1] for k, v in obs_full.items()}, t1 - t0, n_local)
        return {i: model.tower_forward(i, rc, pa_full[t0:t1].contiguous(), mk_full[t0:t1].contiguous(), keep=False,
                                       want_logits=(i == ACTOR), want_values=(i != ACTOR))[0]
                for i in (ACTOR, CRITIC, COST)}

    with torch.no_grad():
        if T + 1 <= 256:
            outs = collect(0, T + 1)
            vals, cvals, logits = outs[CRITIC]["values"], outs[COST]["values"], outs[ACTOR]["logits"][:T]
        else:  # the decoder's window is 256 steps (config 4: T = 256): bootstrap row from the window shifted by one
            outs, last = collect(0, T), collect(1, T + 1)
            vals = torch.cat([outs[CRITIC]["values"], last[CRITIC]["values"][-1:]], 0)
            cvals = torch.cat([outs[COST]["values"], last[COST]["values"][-1:]], 0)
            logits = outs[ACTOR]["logits"]
            del last
        logp = torch.log_softmax(logits, -1).gather(-1, ro["actions"].to(dev).unsqueeze(-1))
    vp_host = vals.cpu().pin_memory()
    cvp_host = cvals.cpu().pin_memory()
    logp_host = logp.squeeze(-1).cpu().pin_memory()
    del obs_full, outs, vals, cvals, logits
    model._ctx_cache = None
    storage.load_rollout(ro, vp_host, cvp_host, logp_host)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    host_s = []

    def step_resident():
        model._ctx_cache = None  # a new rollout every step: rebuild the observation-derived caches
        t0 = time.perf_counter()
        res = upd.update(storage)
        host_s.append(time.perf_counter() - t0)  # host time to ENQUEUE the step (no sync): launch-bound when ~ step time
        return res

    host_out = torch.empty(16 + 8, pin_memory=True)

    def step_e2e():
        model._ctx_cache = None
        storage.load_rollout(ro, vp_host, cvp_host, logp_host)  # pinned host -> HBM
        res = upd.update(storage)
        host_out[:16].copy_(res["loss_scalars"], non_blocking=True)
        host_out[16:16 + K].copy_(res["lambda"], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return res

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1) * 1e-3], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item() / steps

    lib = L.load_library()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # ---- device-resident metric (+ per-kernel event timing of the dominant GEMM kernel)
    for _ in range(args.warmup):
        step_resident()
    barrier()
    l0 = lib.svla_launch_count()
    if rank == 0:  # tools/collect_profiles.sh skips this many launches to capture exactly the timed step under ncu
        print(f"[bench] launches before the timed region: {l0}", file=sys.stderr)
    # Per-launch CUDA events (roofline of the dominant kernel) need the launches on one stream, one after the other.
    # When the update runs its towers on three streams (small per-rank rollouts) or replays CUDA
    # graphs, rank 0 times the launches of `steps` extra single-stream, eagerly launched steps right after the timed
    # region (same kernels, same shapes); otherwise the events are recorded inside the timed region itself.
    n_local_rows = T * n_local
    streams_on = (not args.no_tower_streams) and (args.tower_streams or n_local_rows <= 4096)
    upd.cfg.tower_streams = streams_on
    graphs = args.cuda_graphs
    separate = graphs or streams_on
    ops.PROFILE = {} if (rank == 0 and not separate) else None
    t_res = timed(step_resident, args.steps, 0)
    launches = (lib.svla_launch_count() - l0) // args.steps
    host_timed = list(host_s[-args.steps:])
    t_prof = t_res
    if separate:
        prev_mode = (upd.cfg.cuda_graphs, upd.cfg.tower_streams)
        upd.cfg.cuda_graphs, upd.cfg.tower_streams = False, False
        ops.PROFILE = {} if rank == 0 else None
        t_prof = timed(step_resident, args.steps, 0)
        upd.cfg.cuda_graphs, upd.cfg.tower_streams = prev_mode
    prof, ops.PROFILE = ops.profile_summary(ops.PROFILE), None
    # ---- end-to-end metric
    t_e2e = timed(step_e2e, args.steps, max(1, args.warmup // 2))
    clocks = sampler.stop() if rank == 0 else None
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    samples = T * wl["N"] * UPDATE_REPEATS
    value, e2e = samples / t_res, samples / t_e2e
    dtype_name = {"bf16": "bf16", "fp32": "f32", "bf16x3": "f32 (3 split-bf16 tcgen05 products per GEMM)",
                  "bf16x6": "f32 (6 split-bf16 tcgen05 products per GEMM)"}[args.precision]
    # roofline of the dominant kernel: summed algorithmic FLOPs / summed event time of its launches
    roof = None
    executed = None
    if prof:
        # FLOPs the step really launches (every GEMM's 2MNK + the attention kernels' S x S x dh products), next to
        # the SURVEY model's algorithmic figure (which also counts work the CLS-only last layer and the prompt
        # de-duplication legitimately skip)
        ex_flops = sum(r["flops"] for r in prof.values()) / args.steps
        ex_tf = ex_flops / t_res / 1e12  # launched FLOPs are the same whether a step is replayed or launched eagerly
        executed = {"achieved": ex_tf * world, "peak": pk["tf_sust"] * world, "frac": ex_tf / pk["tf_sust"],
                    "gflop_per_sample": ex_flops * world / samples / 1e9,
                    "by_kernel": {k: {"tflop_per_step": r["flops"] / args.steps / 1e12,
                                      "ms_per_step": r["ms"] / args.steps, "launches_per_step": r["n"] // args.steps,
                                      "tflops": r["flops"] / max(r["ms"], 1e-9) / 1e9}
                                  for k, r in sorted(prof.items())},
                    "note": "rank 0's launched FLOPs (GEMM 2MNK incl. split-operand products + attention) / step time"}
        best = max(((k, v) for k, v in prof.items() if "gemm" in k), key=lambda kv: kv[1]["ms"])
        kname, rec = best
        torch.cuda.synchronize()
        ach = rec["flops"] / (rec["ms"] * 1e-3) / 1e12
        peak = pk["tf_sust"]
        roof = {"bound": "tensor", "kernel": kname, "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                "frac": ach / peak, "traffic": None, "launches_per_step": rec["n"] // args.steps,
                "share_of_step": rec["ms"] * 1e-3 / (t_prof * args.steps), "peak_src": pk["src"] + " (sustained bf16)",
                "timing": ("CUDA events around every launch inside the timed region" if not separate else
                           "CUDA events around every launch of single-stream steps right after the timed region "
                           "(the timed steps run the towers on three streams / replay CUDA graphs)")}
        roof.update(ncu_traffic())
    step_tf = value * wl["gflop_per_sample"] / 1e3
    line = {
        "metric": "ppo_lagrangian_update_samples_per_sec", "value": value, "unit": "samples/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_res * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None,
        "dtype": dtype_name, "data": "synthetic",
        "config": {"workload": name, "T": T, "N_global": wl["N"], "N_per_rank": n_local, "actions": A, "cameras": C,
                   "prompt_tokens": wl["L"], "cost_channels": K, "update_repeats": UPDATE_REPEATS,
                   "parallelism": f"dp{world}",
                   "l2": "inputs larger than L2 (1.06 GB of observations per rank-rollout at N=64)",
                   "precision": args.precision, "chunk_rows": args.chunk_rows, "cuda_graphs": bool(graphs),
                   "tower_streams": bool(streams_on),
                   "dropout": args.dropout},
        "e2e": {"value": e2e, "unit": "samples/s", "h2d_bytes_per_step": storage.h2d_bytes(),
                "d2h_bytes_per_step": (16 + K) * 4, "ms_per_step": t_e2e * 1e3},
        "gpu_launches": int(launches),
        "host_enqueue_ms_per_step": 1e3 * sum(host_timed) / max(1, len(host_timed)),
        "clocks": clocks,
        "roofline": roof,
        "step_algorithmic_tflops": {"achieved": step_tf, "peak": pk["tf_sust"] * world,
                                    "frac": step_tf / (pk["tf_sust"] * world), "gflop_per_sample": wl["gflop_per_sample"],
                                    "note": "SURVEY 8a FLOP model x samples/s over all GPUs; includes every non-GEMM kernel"},
        "step_executed_tflops": executed,
    }
    if world == 1:
        line["roofline_scan"] = scan
        if not args.no_cpu_baseline:
            v, t, desc = cpu_sample(wl, 1, 1, sample_T=T, sample_N=min(wl["N"], 2))  # warmed: 1 + 1 iterations
            line["cpu_baseline"] = {"value": v, "unit": "samples/s", "cores": torch.get_num_threads(), "kind": "port",
                                    "sample": desc}
        del upd, storage, model  # the legs below need the HBM the stash held
        torch.cuda.empty_cache()
        if not args.no_parity_mode and args.precision == "bf16":
            # the same update in the parity-grade tensor-core mode (fp32 activations / weights, every GEMM and attention
            # product as three split-bf16 tcgen05 products: logits / values / losses within 1e-4 of the fp32 reference,
            # tests/test_fastpath_parity_gpu.py), one warm-up + one timed step, device-resident like `value`
            try:
                pm = B200SafeActorCritic(A, C, precision="bf16x3", seed=0, device=dev, chunk_rows=args.chunk_rows,
                                         extras="off", verify_dedupe=False, num_cost_channels=K)
                pu = PPOLagUpdater(pm, cfg)
                ps = B200RolloutStorage(T, dev, num_cost_channels=K)
                ps.load_rollout(ro, vp_host, cvp_host, logp_host)

                def step_parity():
                    pm._ctx_cache = None
                    return pu.update(ps)
                t_par = timed(step_parity, 1, 1)
                line["parity_mode"] = {"precision": "bf16x3", "dtype": "f32 (3 split-bf16 tcgen05 products per GEMM / attention matmul)",
                                       "value": samples / t_par, "unit": "samples/s", "ms_per_step": t_par * 1e3,
                                       "steps": 1, "warmup": 1,
                                       "tolerance": "logits / values / losses within 1e-4 rel of the fp32 reference goldens"}
                del pm, pu, ps
            except Exception as e:  # noqa: BLE001
                line["parity_mode"] = {"value": None, "error": f"{type(e).__name__}: {e}"[:200]}
            torch.cuda.empty_cache()
        if not args.no_library_baseline:
            line["library_baseline"] = library_baseline(wl)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2_64env_128step", choices=list(WORKLOADS))
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32", "bf16x3", "bf16x6"])
    ap.add_argument("--chunk-rows", type=int, default=4096)
    ap.add_argument("--dropout", type=float, default=0.0,
                    help="training-mode dropout of the fusion block (reference: 0.1); 0 = the parity configuration")
    ap.add_argument("--no-cuda-graphs", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--cuda-graphs", action="store_true",
                    help="replay the forward / loss / backward of every update repeat from a CUDA graph")
    ap.add_argument("--no-tower-streams", action="store_true", help="run the three towers on one stream")
    ap.add_argument("--tower-streams", action="store_true", help="three tower streams at any rollout size")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-library-baseline", action="store_true")
    ap.add_argument("--no-parity-mode", action="store_true", help="skip the one-step bf16x3 (parity-grade) measurement")
    ap.add_argument("--ref-device", default="cpu", help="--impl reference only: cpu (the reference arm) or cuda")
    ap.add_argument("--ref-samplers", type=int, default=2, help="--ref-device cuda: samplers in the timed sample")
    ap.add_argument("--ref-tf32", action="store_true", help="--ref-device cuda: allow TF32 tensor-core matmuls")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl, args.workload)
    else:
        run_b200(args, wl, args.workload)


if __name__ == "__main__":
    main()
