"""PPOLagUpdater: the whole constrained-PPO update as one explicit schedule of C-ABI launches.

Replaces the allenact engine's update loop (entered from
training/online/allenact_trainer.py:47-72 with the hyper-parameters of
training/online/dinov2_vits_tsfm_base.py:310-379):

    storage full -> GAE(reward) + GAE(cost) ->
      update_repeats x [ 3-tower forward -> fused PPO-Lagrangian loss fwd+bwd -> tower backwards ->
                         ONE all-reduce of the flat gradient arena (+ cost scalars in its tail) ->
                         fused global-norm clip + Adam (+ bf16 shadow refresh, + grad zeroing) ]
    -> Lagrange-multiplier update on the device.

No autograd, no host sync inside the schedule; scalars for logging are returned as device tensors.
"""
from __future__ import annotations

from dataclasses import dataclass
import ctypes as C
from typing import Dict, Optional

import torch
import torch.distributed as dist

from . import _lib as L
from . import ops
from .lagrange import Lagrange
from .misc import nvtx_range
from .model import ACTOR, COST, CRITIC, B200SafeActorCritic
from .parallel import TAIL
from .storage import B200RolloutStorage


@dataclass
class PPOLagConfig:
    # training/online/dinov2_vits_tsfm_base.py:314-347
    clip_param: float = 0.1
    value_loss_coef: float = 0.5
    entropy_coef: float = 0.0
    use_clipped_value_loss: bool = False
    normalize_advantage: bool = False  # :321 (off in the shipped config); global-batch statistics under data parallelism
    gamma: float = 0.99
    gae_lambda: float = 0.95
    update_repeats: int = 4
    max_grad_norm: float = 0.5
    lr: float = 2e-5
    betas: tuple = (0.9, 0.999)
    eps: float = 1e-8
    # stage 0 = ["ppo_value_loss", "safe_ppo_value_loss"]; stage 1 = ["ppo_log_loss"] (:348-378)
    stage: int = 1
    # omnisafe Lagrange defaults + README cost limit; a tuple = one limit per cost channel (K-channel
    # extension: the model and the storage must be built with the same num_cost_channels)
    cost_limit: object = 2.31964
    lambda_init: float = 0.001
    lambda_lr: float = 0.035
    lambda_upper_bound: Optional[float] = None
    # the reference evaluates all three towers in every forward even when a stage's losses ignore one
    # (separate_actor_critic.py:27-37); keep that by default so samples/s counts the same work
    evaluate_unused_towers: bool = True
    # Replay the forward + loss + backward of an update repeat from a CUDA graph (captured on the second repeat of a
    # given (storage, T, N, prompt length) and replayed ever after; the optimizer step stays outside: its bias
    # corrections are host-side scalars).  The schedule is ~700 C-ABI launches per repeat; once the per-rank problem is
    # small (64 env over 8 GPUs: 1 024 rows per rank) enqueueing them from Python takes longer than the GPU needs to
    # run them -- 82.7 ms of host time for an 87.6 ms step, measured -- so data-parallel runs are launch-bound without
    # it.  Needs dropout off and normalize_advantage off.  Off by default: measured at that per-rank shape the host needs
    # 39 ms per step to enqueue what the GPU runs in 80 ms, so replay changes the step by 2 % (tools/graph_probe.py);
    # what is lost there is small-kernel efficiency on the device, which `tower_streams` addresses.
    cuda_graphs: Optional[bool] = None
    # Run the three towers on three CUDA streams (they are independent until the loss and again in the backward): the
    # launch-latency-bound decoder of one tower (~100 small launches forward, ~200 backward) then hides behind the
    # large encoder GEMMs of another.  Matters when the per-rank problem is small (data-parallel runs).
    # None = on when the rollout has at most 4 096 rows per rank (above that every launch fills the GPU by itself).
    tower_streams: Optional[bool] = None


class PPOLagUpdater:
    TAIL = TAIL  # floats appended to the gradient arena for the packed scalar all-reduce

    def __init__(self, model: B200SafeActorCritic, cfg: PPOLagConfig = PPOLagConfig(),
                 process_group: Optional[dist.ProcessGroup] = None):
        self.model, self.cfg, self.pg = model, cfg, process_group
        self.world = dist.get_world_size(process_group) if (dist.is_available() and dist.is_initialized()) else 1
        dev = model.dev
        n = model.layout.total
        self.exp_avg = torch.zeros(n, device=dev)
        self.exp_avg_sq = torch.zeros(n, device=dev)
        self.sq = torch.zeros(1, device=dev)
        self.adam_step = 0
        self.tower_steps = [0, 0, 0]
        self.lagrange = Lagrange(cfg.cost_limit, cfg.lambda_init, cfg.lambda_lr, "Adam", cfg.lambda_upper_bound, dev)
        # flat [grads | tail] buffer for the single collective; model.grad_arena aliases its head
        if self.world > 1:
            self.comm = torch.zeros(n + self.TAIL, device=dev)
            model.grad_arena = self.comm[:n]
            for tw in model.towers:
                tw.W.grads = model.grad_arena
            model.attach_grads()
        self.launches = 0
        self._pending = []  # (tower, async all-reduce handle) of gradient slices already in flight
        self._graphs: Dict[tuple, dict] = {}  # CUDA graphs of one update repeat, by (storage, shapes, stage)
        self._streams = None  # one stream per tower, created on first use

    # ------------------------------------------------------------------
    def _hp(self, R: int) -> L.PpoHparams:
        c = self.cfg
        if c.stage == 0:
            return L.PpoHparams(c.clip_param, 0.0, 1.0, 0.0, 1.0, 1.0 / R, 1.0, int(c.use_clipped_value_loss), 0)
        return L.PpoHparams(c.clip_param, 1.0, c.value_loss_coef, c.entropy_coef, 0.0, 1.0 / R, 1.0,
                            int(c.use_clipped_value_loss), 1)

    def update(self, storage: B200RolloutStorage) -> Dict[str, torch.Tensor]:
        m, c = self.model, self.cfg
        T, N = storage.T, storage.N
        R = T * N
        K = storage.K
        assert K == self.lagrange.K == m.K, "storage, model and cost limits must agree on the number of cost channels"
        # the bootstrap rows (value / cost-value predictions of the step after the rollout) are already in the arena
        with nvtx_range("update/gae"):
            storage.before_updates(next_value=storage.value_preds[T], next_c_value=None,
                                   use_gae=True, gamma=c.gamma, tau=c.gae_lambda,
                                   normalize_advantage=c.normalize_advantage, process_group=self.pg)
        adv_t = storage.norm_adv_targ if c.normalize_advantage else storage.adv_targ
        c_adv_1 = storage.c_norm_adv_targ if c.normalize_advantage else storage.c_adv_targ
        c_adv_k = storage.c_norm_adv_targ_k if c.normalize_advantage else storage.c_adv_targ_k
        obs = {k: v[:T] for k, v in storage.observations.items()}
        with nvtx_range("update/prepare"):
            rc = m.prepare(obs, T, N)
        pa = storage.prev_actions[:T]
        mk = storage.masks[:T].view(T, N)
        grad_towers = (CRITIC, COST) if c.stage == 0 else (ACTOR, CRITIC)
        m.set_trainable_towers(grad_towers)
        hp = self._hp(R)
        scal = None
        use_graph = bool(c.cuda_graphs) and m.dropout == 0.0 and \
            not c.normalize_advantage and ops.PROFILE is None and L._prof_sink is None
        ctx = dict(storage=storage, pa=pa, mk=mk, hp=hp, grad_towers=grad_towers, adv_t=adv_t, c_adv_1=c_adv_1,
                   c_adv_k=c_adv_k, T=T, R=R, K=K)
        gkey = (id(storage), T, N, rc.L, K, c.stage, c.evaluate_unused_towers, m.precision)
        for rep in range(c.update_repeats):
            drop_step = None
            if m.dropout > 0.0:  # fresh masks every repeat, as every reference forward draws new ones
                m.dropout_step += 1
                drop_step = m.dropout_step
            if use_graph:
                scal = self._repeat_graphed(gkey, rc, ctx)
            else:
                scal = self._repeat(rc, ctx, rep, drop_step, overlap=True)
            with nvtx_range(f"update/rep{rep}/reduce_clip_adam"):
                self._reduce_clip_step(storage, last=(rep == c.update_repeats - 1))
        # lambda <- proj(lambda + Adam step on (Jc - d)); Jc from the (all-reduced) finished-episode costs
        cost_pair = self.comm[-self.TAIL:-self.TAIL + 2 * K] if self.world > 1 else storage.cost_sum_cnt
        self.lagrange.update_from_sum_count(cost_pair)
        return {"loss_scalars": scal, "lambda": self.lagrange.lagrangian_multiplier, "grad_sq_norm": self.sq}

    # ------------------------------------------------------------------ one repeat: forward, loss, backward
    def _repeat(self, rc, ctx, rep: int, drop_step, overlap: bool):
        """3-tower forward -> fused loss forward + backward -> tower backwards into the gradient arena.  Returns the
        loss scalars (device).  overlap: start a tower's gradient all-reduce as soon as its backward is enqueued."""
        m, c = self.model, self.cfg
        storage, pa, mk, hp, grad_towers = ctx["storage"], ctx["pa"], ctx["mk"], ctx["hp"], ctx["grad_towers"]
        adv_t, c_adv_1, c_adv_k, T, R, K = ctx["adv_t"], ctx["c_adv_1"], ctx["c_adv_k"], ctx["T"], ctx["R"], ctx["K"]
        reduce_async = self._reduce_tower_async if overlap else (lambda idx: None)
        outs, states = {}, {}
        main = torch.cuda.current_stream()
        want_streams = c.tower_streams if c.tower_streams is not None else R <= 4096
        side = None
        if want_streams and not torch.cuda.is_current_stream_capturing() and ops.PROFILE is None and L._prof_sink is None:
            if self._streams is None:
                self._streams = [torch.cuda.Stream(device=m.dev) for _ in range(3)]
            side = self._streams  # (per-launch event timing wants one stream: kernels of two towers would overlap)

        def on_tower_stream(idx, fn):
            """Run fn with tower idx's stream current (ordered after everything enqueued on the main stream so far)."""
            if side is None:
                return fn()
            side[idx].wait_stream(main)
            with torch.cuda.stream(side[idx]):
                return fn()

        def join(towers):
            if side is not None:
                for idx in towers:
                    main.wait_stream(side[idx])

        fwd_towers = [idx for idx in (ACTOR, CRITIC, COST) if idx in grad_towers or c.evaluate_unused_towers]
        for idx in fwd_towers:
            with nvtx_range(f"update/rep{rep}/fwd/tower{idx}"):
                o, st = on_tower_stream(idx, lambda idx=idx: m.tower_forward(
                    idx, rc, pa, mk, keep=idx in grad_towers, want_logits=(idx == ACTOR), want_values=(idx != ACTOR),
                    dropout_step=drop_step))
            outs[idx], states[idx] = o, st
        join(fwd_towers)
        if c.stage == 0 and K == 1:
            scal, _, dv, dcv = ops.ppo_lag_fwd_bwd(
                None, None, None, None, None, outs[CRITIC]["values"], storage.returns[:T],
                outs[COST]["values"], storage.c_returns[:T], None, hp,
                old_values=storage.value_preds[:T], old_c_values=storage.c_value_preds[:T])
            on_tower_stream(CRITIC, lambda: (m.tower_backward(CRITIC, states[CRITIC], None, dv), reduce_async(CRITIC)))
            on_tower_stream(COST, lambda: m.tower_backward(COST, states[COST], None, dcv))
            join((CRITIC, COST))
        elif c.stage == 0:
            # K cost channels: the cost critic's head emits [T, N, K]; its loss is the SUM over channels of the
            # per-channel SafePPOValue means (inv_count stays 1 / R), evaluated over the R * K flattened entries
            scal, _, dv, _ = ops.ppo_lag_fwd_bwd(None, None, None, None, None, outs[CRITIC]["values"],
                                                 storage.returns[:T], None, None, None, hp,
                                                 old_values=storage.value_preds[:T])
            tnk = lambda x: x.reshape(K, R).t().contiguous()  # noqa: E731  channel-major -> the head's [R, K]
            scal_c, _, _, dcv = ops.ppo_lag_fwd_bwd(None, None, None, None, None, None, None, outs[COST]["values"],
                                                    tnk(storage.c_returns_k[:, :T]), None, hp,
                                                    old_c_values=tnk(storage.c_value_preds_k[:, :T]))
            on_tower_stream(CRITIC, lambda: (m.tower_backward(CRITIC, states[CRITIC], None, dv), reduce_async(CRITIC)))
            on_tower_stream(COST, lambda: m.tower_backward(COST, states[COST], None, dcv))
            join((CRITIC, COST))
            scal = torch.cat([scal[0:4], scal_c[4:5], scal[5:]])  # [0] value-critic total, [4] cost-critic total
        else:
            c_adv, lam = c_adv_1, self.lagrange.lagrangian_multiplier
            if K > 1:  # fold the K (advantage, multiplier) pairs into the one pair the fused loss takes
                c_adv, lam = ops.combine_cost_advantages(c_adv_k.reshape(K, R), lam)
            scal, dl, dv, _ = ops.ppo_lag_fwd_bwd(
                outs[ACTOR]["logits"], storage.actions, storage.action_log_probs, adv_t,
                c_adv, outs[CRITIC]["values"], storage.returns[:T], None, None,
                lam, hp, old_values=storage.value_preds[:T])
            with nvtx_range(f"update/rep{rep}/bwd/tower{ACTOR}"):
                # the actor's slice crosses NVLink while the critic's backward runs
                on_tower_stream(ACTOR, lambda: (m.tower_backward(ACTOR, states[ACTOR], dl, None), reduce_async(ACTOR)))
            with nvtx_range(f"update/rep{rep}/bwd/tower{CRITIC}"):
                on_tower_stream(CRITIC, lambda: m.tower_backward(CRITIC, states[CRITIC], None, dv))
            join((ACTOR, CRITIC))
        return scal

    _RC_FIELDS = ("text_idx", "time_step", "in_hand", "traj_nt", "perm_tn", "perm_nt")

    def _repeat_graphed(self, key, rc, ctx):
        """The same repeat replayed from a CUDA graph.  First call for a key: eager (it also sets the kernels' launch
        attributes, which cannot be captured).  Second call: the rollout context is copied into buffers that live as
        long as the graph, the repeat is captured over them and replayed.  Later calls: refresh the buffers, replay."""
        lib = L.load_library()
        g = self._graphs.get(key)
        if g is None:
            self._graphs[key] = {"graph": None}
            return self._repeat(rc, ctx, 0, None, overlap=True)
        if g["graph"] is None:
            R = rc.T * rc.N
            st = type(rc)(rc.T, rc.N, rc.L, [v.clone() for v in rc.vis],
                          torch.zeros(R * rc.L, rc.text_u.shape[1], device=rc.text_u.device, dtype=rc.text_u.dtype),
                          *[(getattr(rc, f).clone() if getattr(rc, f) is not None else None) for f in self._RC_FIELDS],
                          key=("static",) + tuple(key))
            g["rc"], g["src_key"] = st, None
            self._refresh_static(g, rc)
            cap = torch.cuda.Stream(device=self.model.dev)
            with torch.cuda.stream(cap):
                L.get_ctx()  # the library context of the capture stream allocates its scratch: not capturable
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            n0 = lib.svla_launch_count()
            with torch.cuda.graph(graph, stream=cap):
                g["scal"] = self._repeat(st, ctx, 0, None, overlap=False)
            g["launches"] = lib.svla_launch_count() - n0  # counted at capture; nothing ran yet
            lib.svla_launch_count_add(C.c_ulonglong(-g["launches"] & 0xFFFFFFFFFFFFFFFF))
            g["graph"] = graph
        self._refresh_static(g, rc)
        g["graph"].replay()
        lib.svla_launch_count_add(g["launches"])
        return g["scal"]

    def _refresh_static(self, g, rc):
        """New rollout -> copy its observation-derived context into the buffers the graph reads."""
        if g["src_key"] == rc.key:
            return
        st = g["rc"]
        for a, b in zip(st.vis, rc.vis):
            a.copy_(b)
        st.text_u[: rc.text_u.shape[0]].copy_(rc.text_u)
        for f in self._RC_FIELDS:
            if getattr(rc, f) is not None:
                getattr(st, f).copy_(getattr(rc, f))
        g["src_key"] = rc.key

    # ------------------------------------------------------------------ resume state
    def state_dict(self) -> Dict:
        """Everything a resumed run needs besides the model's own state_dict (checkpoint row: the
        reference engine saves `optimizer_state_dict` next to `model_state_dict`, training/online/
        dinov2_vits_tsfm_base.py:79,329): Adam moments as flat fp32 arenas in ParamLayout order, the per-tower step
        counts (torch.optim.Adam keeps one per parameter; towers start counting when they first train) and the
        Lagrange multiplier with its own Adam state."""
        return {"exp_avg": self.exp_avg.detach().cpu().clone(), "exp_avg_sq": self.exp_avg_sq.detach().cpu().clone(),
                "adam_step": self.adam_step, "tower_steps": list(self.tower_steps),
                "layout_total": self.model.layout.total,
                "lagrange": {k: v.detach().cpu() for k, v in self.lagrange.state_dict().items()}}

    def load_state_dict(self, sd: Dict) -> None:
        if sd["layout_total"] != self.model.layout.total:
            raise ValueError("optimizer state belongs to a model with a different parameter layout")
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        self.adam_step, self.tower_steps = int(sd["adam_step"]), [int(x) for x in sd["tower_steps"]]
        self.lagrange.load_state_dict({k: v.to(self.model.dev) for k, v in sd["lagrange"].items()})

    def named_optimizer_state(self) -> Dict[str, Dict]:
        """torch.optim.Adam-style view of the flat state: {parameter name: {exp_avg, exp_avg_sq, step}} (views, no
        copies) -- what a tool that inspects or converts a reference checkpoint's optimizer state expects."""
        from .params import TOWERS
        lay, out = self.model.layout, {}
        for name in lay.slots:
            ti = max(i for i, pre in enumerate(TOWERS) if name.startswith(pre))
            out[name] = {"exp_avg": lay.view(self.exp_avg, name), "exp_avg_sq": lay.view(self.exp_avg_sq, name),
                         "step": self.tower_steps[ti]}
        return out

    def _reduce_tower_async(self, idx: int):
        """Data parallel only: starts the all-reduce of ONE tower's gradient slice on NCCL's stream as soon as that
        tower's backward has been enqueued, so it overlaps the next tower's backward; `_reduce_clip_step` waits for it
        and reduces what is left.  The towers are independent, so nothing later writes into the slice."""
        if self.world == 1:
            return
        from .params import TOWERS
        a, b = self.model.layout.tower_range[TOWERS[idx]]
        self._pending.append((idx, dist.all_reduce(self.comm[a:b], op=dist.ReduceOp.SUM, group=self.pg, async_op=True)))

    def _reduce_clip_step(self, storage: B200RolloutStorage, last: bool):
        m, c = self.model, self.cfg
        n = m.layout.total
        prescale = 1.0
        if self.world > 1:
            # Only the towers this stage trains carry gradients (the third tower's slice is all zeros: a third of the
            # arena is never sent); slices already in flight are waited for, the rest + the packed cost scalars of the
            # last repeat go in one more collective.
            from .params import TOWERS
            done = {i for i, _ in self._pending}
            rest = [i for i in m.trainable_towers if i not in done]
            lo = min(m.layout.tower_range[TOWERS[i]][0] for i in rest)
            hi = max(m.layout.tower_range[TOWERS[i]][1] for i in rest)
            if last:  # tail: [sum of finished-episode costs, count] per cost channel -> lambda identical on all ranks
                k = storage.cost_sum_cnt.numel()
                self.comm[n:n + self.TAIL].zero_()
                self.comm[n:n + k].copy_(storage.cost_sum_cnt)
                if hi == n:
                    hi = n + self.TAIL  # the cost tower's slice is adjacent to the tail: one collective
                else:
                    self._pending.append((-1, dist.all_reduce(self.comm[n:n + self.TAIL], op=dist.ReduceOp.SUM,
                                                              group=self.pg, async_op=True)))
            dist.all_reduce(self.comm[lo:hi], op=dist.ReduceOp.SUM, group=self.pg)
            for _, work in self._pending:
                work.wait()
            self._pending.clear()
            prescale = 1.0 / self.world
        self.adam_step += 1
        hp = L.AdamHparams(c.lr, c.betas[0], c.betas[1], c.eps, c.max_grad_norm, prescale, self.adam_step, 1)
        # torch.optim.Adam skips parameters whose .grad is None: only the towers this stage trains are
        # stepped (they are adjacent in the arena: stage 0 = critic|cost, stages 1-2 = actor|critic)
        from .params import TOWERS
        lo = min(m.layout.tower_range[TOWERS[i]][0] for i in m.trainable_towers)
        hi = max(m.layout.tower_range[TOWERS[i]][1] for i in m.trainable_towers)
        g = m.grad_arena[lo:hi]
        if c.max_grad_norm > 0:
            ops.sq_norm(g, self.sq)
            if prescale != 1.0:
                # clip_adam expects the norm of the pre-scaled gradient
                ops.scale_by(self.sq, torch.full((1,), prescale * prescale, device=m.dev))
        for i in m.trainable_towers:  # per-parameter Adam step counts start when a tower first trains
            self.tower_steps[i] += 1
            hp.step = self.tower_steps[i]
            a, b = m.layout.tower_range[TOWERS[i]]
            ops.clip_adam(m.param_arena[a:b], m.grad_arena[a:b], self.exp_avg[a:b], self.exp_avg_sq[a:b],
                          m.shadow_arena[a:b] if m.shadow_arena is not None else None,
                          self.sq if c.max_grad_norm > 0 else None, hp)
        ops.split_cache_clear()  # split-operand copies of the weights (bf16x3 / bf16x6 modes) are stale now
