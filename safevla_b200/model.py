"""B200SafeActorCritic: drop-in for SafeDinoLLAMATxNavActorCriticSeparate
(architecture/models/allenact_transformer_models/separate_actor_critic.py:22-37 on top of
allenact_dino_transformer.py:47-475): three independent towers (actor / reward critic / cost critic),
same `forward(observations, memory, prev_actions, masks)` signature, same `state_dict` keys, same
output struct -- but every FLOP runs in libsafevla_b200 (sm_100a) through the explicit schedule in
tower.py.  Autograd only sees one opaque node per tower output.

Deliberate, documented differences from the reference (results identical with dropout off):
  * the frozen T5 encoder runs ONCE per rollout on the de-duplicated prompts and is shared by the
    three towers (reference: per tower, per row, per forward -- SURVEY.md fact 6);
  * the last fusion layer only evaluates the CLS row it returns (fact 7);
  * heads that the separate-tower wrapper discards are not evaluated (fact 4);
  * dropout is off by default (the parity setting; the reference trains with p = 0.1, fact 8): `dropout=0.1` turns the
    fusion block's four dropout sites on (counter-based masks, bf16 precision).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn as nn

from . import ops
from .misc import CategoricalDistr, Memory, SafeActorCriticOutput
from .params import D, TOWERS, ParamLayout, T5Layout, init_state_dict, t5_spec, tower_spec, tower_values
from .tower import TOK, EncStash, T5Encoder, Tower, TowerWeights

ACTOR, CRITIC, COST = 0, 1, 2


class _Node(nn.Module):
    """Empty container used to reproduce the reference's dotted parameter names."""


def default_synthetic_tokenizer(goals: Sequence[str]) -> Tuple[torch.Tensor, torch.Tensor]:
    """Synthetic goal codec (safevla_b200.synthetic.encode_goal_ids): blank-separated decimal ids,
    EOS = 1 appended, right-padded with 0 to the batch longest, mask marks real tokens -- the
    contract of `self.text_tokenizer(goals, return_tensors="pt", padding=True)` at
    allenact_dino_transformer.py:600-602.  The real t5-small sentencepiece model is not available
    offline; pass `tokenizer=` (any callable with this contract) to use it."""
    rows = [[int(tok) for tok in g.split()] + [1] for g in goals]
    L = max(len(r) for r in rows)
    ids = torch.zeros(len(rows), L, dtype=torch.int64)
    am = torch.zeros(len(rows), L, dtype=torch.int64)
    for i, r in enumerate(rows):
        ids[i, : len(r)] = torch.tensor(r, dtype=torch.int64)
        am[i, : len(r)] = 1
    return ids, am


@dataclass
class RolloutContext:
    """Everything derived from the observations alone; shared by towers and update repeats."""
    T: int
    N: int
    L: int
    vis: List[torch.Tensor]          # per camera [R*84, 384] token-major, activation dtype
    text_u: torch.Tensor             # [U*L, 512] T5 last_hidden_state of the unique prompts (act dtype)
    text_idx: torch.Tensor           # int64 [R*L] gather index into text_u rows
    time_step: torch.Tensor          # int64 [T, N]
    in_hand: Optional[torch.Tensor]  # int64 [T, N] or None
    traj_nt: torch.Tensor            # int64 [N, T]
    perm_tn: torch.Tensor            # int64 [T*N]: row t*N+n <- n*T+t
    perm_nt: torch.Tensor            # int64 [N*T]: row n*T+t <- t*N+n
    key: tuple = ()


class _TowerOutput(torch.autograd.Function):
    """Opaque autograd node: forward already ran; backward launches the tower's backward schedule and
    accumulates straight into the flat gradient arena (parameters' .grad are views of it)."""

    @staticmethod
    def forward(ctx, anchor, model, idx, state, out):
        ctx.model, ctx.idx, ctx.state = model, idx, state
        return out.view_as(out)

    @staticmethod
    def backward(ctx, grad_out):
        model, idx, state = ctx.model, ctx.idx, ctx.state
        g = grad_out.contiguous()
        model._tower_backward_autograd(idx, state, g)
        return None, None, None, None, None


class _TowerOutputPair(torch.autograd.Function):
    """Discrete-critic towers return (values, bin logits); the losses differentiate the logits
    (customized_loss.py:364-370), the value read-out carries no gradient."""

    @staticmethod
    def forward(ctx, anchor, model, idx, state, values, full_logits):
        ctx.model, ctx.idx, ctx.state = model, idx, state
        ctx.set_materialize_grads(False)
        return values.view_as(values), full_logits.view_as(full_logits)

    @staticmethod
    def backward(ctx, g_values, g_logits):
        if g_values is not None:
            raise NotImplementedError("discrete critic: only the bin logits are differentiated")
        if g_logits is not None:
            ctx.model._tower_backward_autograd(ctx.idx, ctx.state, None, dfull=g_logits.contiguous())
        return None, None, None, None, None, None


class B200SafeActorCritic(nn.Module):
    def __init__(self, num_actions: int, num_cameras: int = 1, *, precision: str = "bf16",
                 device: Optional[torch.device] = None, state_dict: Optional[Dict[str, torch.Tensor]] = None,
                 seed: int = 0, tokenizer: Optional[Callable] = None, chunk_rows: int = 1024,
                 stash_budget_bytes: int = 100 << 30, cls_only_last_layer: bool = True,
                 goal_sensor_uuid: str = "natural_language_spec", rgb_uuid: str = "rgb_dinov2",
                 manip_uuid: str = "manipulation_rgb_dinov2", in_hand_uuid: str = "an_object_is_in_hand",
                 time_step_uuid: str = "time_step", traj_idx_uuid: str = "traj_index", extras: str = "eager",
                 verify_dedupe: bool = True, max_steps: int = 1000, num_cost_channels: int = 1,
                 critic_type: str = "linear", dropout: float = 0.0, dropout_seed: int = 0):
        super().__init__()
        # "bf16": bf16 operands on the tcgen05 kernels (the fast path); "fp32": fp32 FMA kernels (CUDA cores);
        # "bf16x3" / "bf16x6": fp32 activations and weights, every tensor-core-shaped product evaluated on the tcgen05
        # kernels as 3 / 6 split-bf16 products (parity-grade tensor-core mode, see csrc/split.cu)
        assert precision in ("bf16", "fp32", "bf16x3", "bf16x6")
        self.split = {"bf16x3": 3, "bf16x6": 6}.get(precision, 0)
        if not torch.cuda.is_available():
            raise RuntimeError("B200SafeActorCritic needs a CUDA device (sm_100a); there is no CPU fallback")
        self.dev = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.A, self.C = num_actions, num_cameras
        # K cost channels (extension beyond the reference's single scalar cost, tasks/abstract_task.py:333): the cost
        # critic's head predicts K values; K = 1 is the reference model, key for key
        self.K = int(num_cost_channels)
        self.precision = precision
        self.adt = torch.bfloat16 if precision == "bf16" else torch.float32
        # training-mode dropout of the fusion block (reference: p = 0.1 live); 0 = the parity setting.
        # Counter-based masks: (dropout_seed, dropout_step, site, row, column); dropout_step advances with every
        # update-mode forward, the rollout-side step (collection under no_grad) runs without dropout.  The frozen T5
        # encoder runs without dropout (its noise would be an input perturbation shared by the three towers).
        if not 0.0 <= dropout < 1.0:
            raise ValueError("dropout must be in [0, 1)")
        if dropout > 0.0 and precision != "bf16":
            raise NotImplementedError("dropout > 0 is built on the bf16 tensor-core kernels (precision='bf16')")
        self.dropout, self.dropout_seed, self.dropout_step = float(dropout), int(dropout_seed), 0
        self.uu = dict(goal=goal_sensor_uuid, rgb=rgb_uuid, manip=manip_uuid, hand=in_hand_uuid,
                       time=time_step_uuid, traj=traj_idx_uuid)
        self.tokenizer = tokenizer or default_synthetic_tokenizer
        self.chunk_rows, self.stash_budget = chunk_rows, stash_budget_bytes
        self.extras_mode, self.verify_dedupe = extras, verify_dedupe
        self.trainable_towers: Tuple[int, ...] = (ACTOR, CRITIC, COST)

        # "linear" (shipped, allenact_dino_transformer.py:148-149) or "discrete" (HL-Gauss DiscreteCriticHead, :152-159)
        self.critic_type = critic_type
        self.dc_loss = None
        if critic_type == "discrete":
            from .losses import HLGaussLoss
            from .params import DC_BINS, DC_MAX, DC_MIN, DC_SIGMA
            self.dc_loss = HLGaussLoss(min_value=DC_MIN, max_value=DC_MAX, num_bins=DC_BINS, sigma=DC_SIGMA)
        self.layout, self.t5_layout = ParamLayout(num_actions, num_cameras, self.K, critic_type), T5Layout()
        f32 = dict(device=self.dev, dtype=torch.float32)
        self.param_arena = torch.zeros(self.layout.total, **f32)
        self.grad_arena = torch.zeros(self.layout.total, **f32)
        self.shadow_arena = (torch.zeros(self.layout.total, device=self.dev, dtype=torch.bfloat16)
                             if precision == "bf16" else None)
        self.t5_arena = torch.zeros(self.t5_layout.total, **f32)
        self._register_names()
        self.load_state_dict(state_dict if state_dict is not None
                             else init_state_dict(num_actions, num_cameras, seed, num_cost_channels=self.K,
                                                  critic_type=critic_type), strict=True)

        self.t5 = T5Encoder(self.t5_layout, self.t5_arena, self.adt, self.split)
        if self.split:
            ops.split_cache_open()
        self.towers: List[Tower] = []
        for pre in TOWERS:
            tw = Tower(TowerWeights(self.layout, pre, self.param_arena, self.grad_arena, self.shadow_arena),
                       num_actions, num_cameras, self.adt, cls_only_last_layer, self.split)
            tw.div_term = self.get_buffer(pre + "time_encoder.div_term")
            tw.tower_idx = len(self.towers)
            self.towers.append(tw)
        self._anchor = torch.zeros(1, device=self.dev, requires_grad=True)
        self._ctx_cache: Optional[RolloutContext] = None
        self._tok_cache: Dict[bytes, Tuple[torch.Tensor, torch.Tensor]] = {}
        # rollout-side (T = 1) state: position in the KV caches (allenact_dino_transformer.py:376-406) and the
        # per-tower, per-layer caches [N, max_steps, 512] (llama/model.py:224-247), allocated on the first step
        self.max_steps = max_steps
        self.time_step_counter = 0
        self._kv: Optional[List[List[Tuple[torch.Tensor, torch.Tensor]]]] = None
        self.train()

    # ------------------------------------------------------------------ naming / state dict
    def _node(self, path: List[str]) -> nn.Module:
        m: nn.Module = self
        for part in path:
            if part not in m._modules:
                m.add_module(part, _Node())
            m = m._modules[part]
        return m

    def _register_names(self):
        for ti, pre in enumerate(TOWERS):
            for k, shape, _ in tower_spec(self.A, self.C, tower_values(ti, self.K), self.critic_type):
                name = pre + k
                parts = name.split(".")
                p = nn.Parameter(self.layout.view(self.param_arena, name), requires_grad=True)
                p.grad = self.layout.view(self.grad_arena, name)
                self._node(parts[:-1]).register_parameter(parts[-1], p)
            self._node((pre + "time_encoder").split(".")).register_buffer(
                "div_term", torch.zeros(D // 2, device=self.dev))
            # frozen T5: one arena, exposed under every tower prefix as in the reference state_dict
            te = pre + "visual_encoder.text_encoder."
            for k, shape, _ in t5_spec():
                parts = (te + k).split(".")
                node = self._node(parts[:-1])
                if pre == TOWERS[0]:
                    par = nn.Parameter(self.t5_layout.view(self.t5_arena, k), requires_grad=False)
                else:
                    par = self.get_parameter(TOWERS[0] + "visual_encoder.text_encoder." + k)
                node.register_parameter(parts[-1], par)
                if k == "shared.weight":
                    self._node((te + "encoder.embed_tokens").split(".")).register_parameter("weight", par)

    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        res = super().load_state_dict(state_dict, strict=strict)
        self.refresh_shadow()
        return res

    def refresh_shadow(self):
        """bf16 GEMM-operand copy of the fp32 master weights (refreshed by the fused Adam kernel)."""
        ops.split_cache_clear()  # staged split-operand copies of the weights are stale too
        if self.shadow_arena is not None:
            ops.cast_bf16(self.param_arena, self.shadow_arena)
        if getattr(self, "t5", None) is not None:  # frozen T5: bf16 operand copy + relative-position bias table
            self.t5.refresh()

    def attach_grads(self):
        """(Re)point every trainable parameter's .grad at its slice of the gradient arena."""
        for name, slot in self.layout.slots.items():
            p = self.get_parameter(name)
            if p.grad is None or p.grad.data_ptr() != self.grad_arena.data_ptr() + 4 * slot.offset:
                p.grad = self.layout.view(self.grad_arena, name)

    def set_trainable_towers(self, towers: Sequence[int]):
        """Which towers will receive a gradient this stage (stage 0: critic + cost critic; stages 1-2:
        actor + critic -- training/online/dinov2_vits_tsfm_base.py:348-378).  Others skip the stash."""
        self.trainable_towers = tuple(towers)

    # ------------------------------------------------------------------ allenact ActorCriticModel surface
    def _recurrent_memory_specification(self):
        return None

    @property
    def recurrent_memory_specification(self):
        return None

    def sampler_select(self, keep: list):
        """Keeps the KV-cache rows of the surviving samplers (llama/model.py:238-244; called by the engine when
        task samplers finish, allenact_dino_transformer.py:197-199)."""
        if self._kv is None:
            return
        for tw in self._kv:
            for l, (k, v) in enumerate(tw):
                if k.shape[0] == 1 and len(keep) > 1:
                    tw[l] = (k.repeat(len(keep), 1, 1) * 0, v.repeat(len(keep), 1, 1) * 0)
                else:
                    tw[l] = (k[keep].contiguous(), v[keep].contiguous())

    def _step_caches(self, N: int):
        if self._kv is None or self._kv[0][0][0].shape[0] != N:
            mk = lambda: torch.zeros(N, self.max_steps, D, device=self.dev, dtype=self.adt)  # noqa: E731
            self._kv = [[(mk(), mk()) for _ in range(3)] for _ in TOWERS]
        return self._kv

    # ------------------------------------------------------------------ observation-side preparation
    def _decode_rows(self, rows_u8: np.ndarray) -> List[str]:
        return [r.tobytes().rstrip(b"\x00").decode() for r in rows_u8]

    def prepare(self, observations: Dict[str, torch.Tensor], T: int, N: int) -> RolloutContext:
        rgb = observations[self.uu["rgb"]]
        goal = observations[self.uu["goal"]]
        # every consumed observation tensor takes part (an in-place change of time_step / traj_index / the manipulation
        # frames / in_hand alone must not reuse a stale context)
        key = tuple((k, v.data_ptr(), v._version) for k, v in sorted(observations.items())
                    if k in self.uu.values()) + (T, N)
        if self._ctx_cache is not None and self._ctx_cache.key == key:
            return self._ctx_cache
        R = T * N
        dev = self.dev
        if T > 1:  # an update-mode forward restarts the rollout-side cache position (:376-377)
            self.time_step_counter = 0

        def tokens(x):
            x = x.to(dev, non_blocking=True).reshape(R, 384, TOK).contiguous()
            assert x.dtype == torch.float32
            return ops.nchw_to_tokens(x, torch.empty(R * TOK, 384, device=dev, dtype=self.adt))

        vis = [tokens(rgb)]
        if self.C == 2:
            vis.append(tokens(observations[self.uu["manip"]]))
        # ---- prompt de-duplication: hash rows on the device, tokenise + encode unique prompts only
        g = goal.to(dev, non_blocking=True).reshape(R, -1).contiguous()
        assert g.dtype == torch.uint8
        hashes = ops.hash_rows(g)
        uniq, inverse = torch.unique(hashes, return_inverse=True)  # index bookkeeping, not path arithmetic
        U = uniq.numel()
        first = torch.full((U,), R, device=dev, dtype=torch.int64)
        first.scatter_reduce_(0, inverse, torch.arange(R, device=dev), reduce="amin")
        if self.verify_dedupe and not bool((g == g[first[inverse]]).all()):
            raise RuntimeError("goal-byte hash collision: rows with equal hash differ")
        reps = g[first].cpu().numpy()
        strs = self._decode_rows(reps)
        ids, am = self.tokenizer(strs)
        L = ids.shape[1]
        text_u = self.t5.forward(ids.to(dev), am.to(dev))  # [U*L, 512] fp32
        if self.adt != torch.float32:
            text_u = ops.copy_rows(text_u, torch.empty(U * L, D, device=dev, dtype=self.adt), U * L, D)
        text_idx = (inverse[:, None] * L + torch.arange(L, device=dev)[None, :]).reshape(-1).contiguous()
        ar = torch.arange(R, device=dev)
        perm_tn = ((ar % N) * T + ar // N).contiguous()  # dst row t*N+n reads n*T+t
        perm_nt = ((ar % T) * N + ar // T).contiguous()
        traj = observations[self.uu["traj"]].to(dev).reshape(T, N)
        in_hand = None
        if self.C == 2:
            in_hand = observations[self.uu["hand"]].to(dev).reshape(T, N).contiguous()
        ctx = RolloutContext(T, N, L, vis, text_u, text_idx,
                             observations[self.uu["time"]].to(dev).reshape(T, N).contiguous(), in_hand,
                             traj.t().contiguous(), perm_tn, perm_nt, key)
        self._ctx_cache = ctx
        return ctx

    # ------------------------------------------------------------------ tower schedules
    def _chunks(self, R: int):
        c = max(1, min(self.chunk_rows, R))
        return [(r0, min(R, r0 + c)) for r0 in range(0, R, c)]

    def _chunk_inputs(self, rc: RolloutContext, r0: int, r1: int):
        vis = [v[r0 * TOK: r1 * TOK] for v in rc.vis]
        n = (r1 - r0) * rc.L
        th = ops.copy_rows(rc.text_u, torch.empty(n, D, device=self.dev, dtype=self.adt), n, D,
                           idx=rc.text_idx[r0 * rc.L: r1 * rc.L])
        return vis, th

    def stash_bytes_per_row(self, L: int) -> int:
        S = 1 + TOK * self.C + L
        e = 2 if self.adt == torch.bfloat16 else 4
        full = S * (D * 6 + 3 * D + 2048) * e + S * 40  # x, qkv, ao, s1, x1, hf, s2 + stats
        pre = (TOK * self.C * 3 + L) * D * e + S * D * e
        return 3 * full + pre

    def tower_forward(self, idx: int, rc: RolloutContext, prev_actions, masks, *, keep: bool,
                      want_logits: bool, want_values: bool, dropout_step: Optional[int] = None):
        """Returns (outputs dict, state for tower_backward or None).  dropout_step: the mask counter of this
        forward (None: no dropout, e.g. value collection); the backward regenerates the masks from the stored step."""
        tw = self.towers[idx]
        use_drop = self.dropout > 0.0 and dropout_step is not None
        tw.drop_p, tw.drop_seed, tw.drop_step = (self.dropout if use_drop else 0.0), self.dropout_seed, dropout_step or 0
        R = rc.T * rc.N
        obs_embed = torch.empty(R, D, device=self.dev, dtype=self.adt)
        chunks = self._chunks(R)
        stash_all = keep and (self.stash_bytes_per_row(rc.L) * R * len(self.trainable_towers) <= self.stash_budget)
        stashes: List[Optional[EncStash]] = []
        for (r0, r1) in chunks:
            vis, th = self._chunk_inputs(rc, r0, r1)
            cls, st = tw.encoder_fwd(vis, th, rc.L, keep=stash_all, row_off=r0)
            obs_embed[r0:r1].copy_(cls)
            stashes.append(st)
        out, dstash = tw.decoder_fwd(obs_embed, prev_actions, masks, rc.in_hand, rc.time_step, rc.traj_nt,
                                     rc.perm_tn, rc.T, rc.N, want_logits, want_values, keep)
        state = dict(rc=rc, dec=dstash, enc=stashes, chunks=chunks, prev=prev_actions, masks=masks,
                     stash_all=stash_all, drop=(tw.drop_p, tw.drop_step)) if keep else None
        return out, state

    def tower_backward(self, idx: int, state, dlogits, dvalues, dfull=None):
        tw, rc = self.towers[idx], state["rc"]
        tw.drop_p, tw.drop_step = state["drop"]  # the masks of THIS state's forward
        d_obs = tw.decoder_bwd(dlogits, dvalues, state["dec"], state["prev"], state["masks"], rc.in_hand,
                               rc.traj_nt, rc.perm_nt, rc.T, rc.N, dfull=dfull)
        state["dec"] = None
        for ci, (r0, r1) in enumerate(state["chunks"]):
            vis, th = self._chunk_inputs(rc, r0, r1)
            st = state["enc"][ci]
            if st is None:  # recompute mode
                _, st = tw.encoder_fwd(vis, th, rc.L, keep=True, row_off=r0)
            tw.encoder_bwd(d_obs[r0:r1], vis, th, rc.L, st, row_off=r0)
            state["enc"][ci] = None

    def _tower_backward_autograd(self, idx: int, state, grad_out: Optional[torch.Tensor], dfull=None):
        pre = TOWERS[idx]
        sentinel = self.get_parameter(pre + "decoder.norm.weight")
        if sentinel.grad is None:  # optimizer.zero_grad(set_to_none=True) happened: arena slice is stale
            lo, hi = self.layout.tower_range[pre]
            self.grad_arena[lo:hi].zero_()
        self.tower_backward(idx, state, grad_out if idx == ACTOR else None, grad_out if idx != ACTOR else None,
                            dfull=dfull)
        self.attach_grads()

    # ------------------------------------------------------------------ nn.Module forward (drop-in)
    def forward(self, observations: Dict[str, torch.Tensor], memory: Optional[Memory], prev_actions: torch.Tensor,
                masks: torch.Tensor):
        T, N = prev_actions.shape
        if T == 1:
            return self._forward_step(observations, memory, prev_actions, masks)
        ops.split_cache_clear()  # the weights may have been stepped by any optimizer since the last call
        rc = self.prepare({k: v[:T] for k, v in observations.items()}, T, N)
        pa = prev_actions.to(self.dev).contiguous()
        mk = masks.to(self.dev, dtype=torch.float32).reshape(T, N).contiguous()
        grad = torch.is_grad_enabled()
        outs = {}
        step = None
        if self.dropout > 0.0 and self.training and grad:  # nn.Module semantics: dropout in train mode only
            self.dropout_step += 1
            step = self.dropout_step
        for idx in (ACTOR, CRITIC, COST):
            keep = grad and idx in self.trainable_towers
            o, state = self.tower_forward(idx, rc, pa, mk, keep=keep, want_logits=(idx == ACTOR),
                                          want_values=(idx != ACTOR), dropout_step=step)
            t = o["logits"] if idx == ACTOR else o["values"]
            fl = o.get("full_logits")
            if keep and fl is not None:
                t, fl = _TowerOutputPair.apply(self._anchor, self, idx, state, t, fl)
            elif keep:
                t = _TowerOutput.apply(self._anchor, self, idx, state, t)
            outs[idx] = t
            if idx == COST:
                cost_logits = fl
        extras = self._extras(outs[COST], cost_logits)
        aco = SafeActorCriticOutput(distributions=CategoricalDistr(logits=outs[ACTOR]), values=outs[CRITIC],
                                    c_values=outs[COST], extras=extras)
        return aco, memory

    @torch.no_grad()
    def _forward_step(self, observations, memory, prev_actions, masks):
        """Rollout-side single step (T = 1): encoder on the N current observations, KV-cache decoder step
        (allenact_dino_transformer.py:376-406).  No autograd graph: the engine collects under no_grad."""
        N = prev_actions.shape[1]
        if self.time_step_counter >= self.max_steps:
            self.time_step_counter = 0
        pos = self.time_step_counter
        self._ctx_cache = None  # rollout observations are fresh every step: never reuse a cached context
        rc = self.prepare({k: v[:1] for k, v in observations.items()}, 1, N)
        pa = prev_actions.to(self.dev).reshape(1, N).contiguous()
        mk = masks.to(self.dev, dtype=torch.float32).reshape(1, N).contiguous()
        kv = self._step_caches(N)
        outs = {}
        for idx in (ACTOR, CRITIC, COST):
            tw = self.towers[idx]
            obs_embed = torch.empty(N, D, device=self.dev, dtype=self.adt)
            for (r0, r1) in self._chunks(N):
                vis, th = self._chunk_inputs(rc, r0, r1)
                cls, _ = tw.encoder_fwd(vis, th, rc.L, keep=False)
                obs_embed[r0:r1].copy_(cls)
            o = tw.decoder_step(obs_embed, pa, mk, rc.in_hand, rc.time_step, kv[idx], pos, N,
                                want_logits=(idx == ACTOR), want_values=(idx != ACTOR))
            outs[idx] = o["logits"] if idx == ACTOR else o["values"]
            if idx == COST:
                cost_logits = o.get("full_logits")
        self.time_step_counter += 1
        aco = SafeActorCriticOutput(distributions=CategoricalDistr(logits=outs[ACTOR]), values=outs[CRITIC],
                                    c_values=outs[COST], extras=self._extras(outs[COST], cost_logits))
        return aco, memory

    def _extras(self, c_values: torch.Tensor, cost_logits: Optional[torch.Tensor] = None):
        """Logging extras with the reference's quirks: they describe the COST tower
        (separate_actor_critic.py:35) and are 1-element CPU tensors (allenact_dino_transformer.py:431-455).
        Discrete critics (:434-439): `full_logits` / `stop_grad_logits` / `loss_func` are the COST tower's too -- which
        is what SafePPOLogGrad's value term then trains (customized_loss.py:364-370), a quirk kept as it is."""
        disc = {}
        if self.critic_type == "discrete" and cost_logits is not None:
            disc = {"full_logits": cost_logits, "stop_grad_logits": cost_logits.detach(), "loss_func": self.dc_loss}
        if self.extras_mode == "off":
            return disc
        pre = TOWERS[COST]
        lo, hi = self.layout.tower_range[pre]
        buf = torch.empty(4, device=self.dev)
        ops.sq_norm(self.grad_arena[lo:hi], buf[0:1])
        head = "critic.fc.2." if self.critic_type == "discrete" else "critic.fc."  # fc[-1] of the MLP heads (:456-468)
        wslot, bslot = self.layout.slots[pre + head + "weight"], self.layout.slots[pre + head + "bias"]
        ops.sq_norm(self.param_arena[wslot.offset: wslot.offset + wslot.numel], buf[1:2])
        nb = (bslot.numel + 3) // 4 * 4  # slots are padded with zeros to 64 elements
        ops.sq_norm(self.param_arena[bslot.offset: bslot.offset + nb], buf[2:3])
        ops.sq_norm(self.grad_arena[wslot.offset: wslot.offset + wslot.numel], buf[3:4])
        host = buf.cpu().sqrt()
        ex = {"total_norm": host[0:1].clone(), "weight_norm": host[1:2].clone(), "bias_norm": host[2:3].clone(),
              "weight_grad_norm": host[3:4].clone()}
        if disc:
            ex.update(disc)
        else:
            ex["stop_grad_values"] = c_values.detach()
        return ex
