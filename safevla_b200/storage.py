"""B200RolloutStorage: device-resident rollout arena with the allenact (fork) RolloutBlockStorage
surface the reference drives -- `initialize`, `add(observations=, memory=, actions=, action_log_probs=,
value_preds=, rewards=, costs=, c_value_preds=, masks=)`, `agent_input_for_next_step`, `after_updates`
(witnessed at architecture/models/allenact_transformer_models/inference_agent.py:172-175,246-269) and,
per upstream allenact, `before_updates`, `batched_experience_generator`, `to`.

Layout: one contiguous tensor per stream, time-major [T(+1), N, ...] fp32/int64, so GAE reads are
coalesced across samplers and a `num_mini_batch == 1` batch is a zero-copy view.  GAE for the reward
and the cost stream is ONE kernel launch (ops.gae_dual).
"""
from __future__ import annotations

from typing import Dict, Iterator, Optional

import torch

from . import ops
from .misc import Memory


class B200RolloutStorage:
    def __init__(self, num_steps: int, device: Optional[torch.device] = None, num_cost_channels: int = 1):
        if not torch.cuda.is_available():
            raise RuntimeError("B200RolloutStorage keeps rollouts in HBM; no CUDA device is visible")
        self.T = num_steps
        # K cost channels (extension; the reference has one): every cost stream is kept channel-major [K, T(+1), N, 1]
        # so that each channel is the contiguous [T, N] plane the GAE kernel marches over; the un-suffixed attributes
        # (`costs`, `c_value_preds`, `c_returns`, `c_adv_targ`) are channel 0, i.e. exactly the K = 1 surface
        self.K = int(num_cost_channels)
        self.dev = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.step = 0
        self.N = 0
        self.observations: Dict[str, torch.Tensor] = {}
        self._initialized = False

    # ------------------------------------------------------------------ allocation
    def to(self, device):
        self.dev = torch.device(device)
        return self

    def initialize(self, *, observations: Dict[str, torch.Tensor], num_samplers: int,
                   recurrent_memory_specification=None, action_space=None, **kwargs):
        T, N, dev = self.T, num_samplers, self.dev
        self.N = N
        f = dict(device=dev, dtype=torch.float32)
        self.observations = {k: torch.zeros(T + 1, N, *v.shape[1:], device=dev, dtype=v.dtype)
                             for k, v in observations.items()}
        K = self.K
        self.rewards = torch.zeros(T, N, 1, **f)
        self.costs_k = torch.zeros(K, T, N, 1, **f)
        self.value_preds = torch.zeros(T + 1, N, 1, **f)
        self.c_value_preds_k = torch.zeros(K, T + 1, N, 1, **f)
        self.returns = torch.zeros(T + 1, N, 1, **f)
        self.c_returns_k = torch.zeros(K, T + 1, N, 1, **f)
        self.adv_targ = torch.zeros(T, N, 1, **f)
        self.c_adv_targ_k = torch.zeros(K, T, N, 1, **f)
        self.costs, self.c_value_preds = self.costs_k[0], self.c_value_preds_k[0]
        self.c_returns, self.c_adv_targ = self.c_returns_k[0], self.c_adv_targ_k[0]
        self.action_log_probs = torch.zeros(T, N, 1, **f)
        self.actions = torch.zeros(T, N, device=dev, dtype=torch.int64)
        self.prev_actions = torch.zeros(T + 1, N, device=dev, dtype=torch.int64)
        self.masks = torch.zeros(T + 1, N, 1, **f)
        self.episode_cost = torch.zeros(K, N, **f)
        self.cost_sum_cnt = torch.zeros(2 * K, **f)  # per channel: [sum of finished-episode costs, finished episodes]
        self.norm_adv_targ = None
        for k, v in observations.items():
            self.observations[k][0].copy_(v.to(dev), non_blocking=True)
        self.step = 0
        self._initialized = True

    # ------------------------------------------------------------------ collection
    def add(self, *, observations: Dict[str, torch.Tensor], memory: Optional[Memory], actions: torch.Tensor,
            action_log_probs: torch.Tensor, value_preds: torch.Tensor, rewards: torch.Tensor,
            costs: Optional[torch.Tensor] = None, c_value_preds: Optional[torch.Tensor] = None,
            masks: torch.Tensor = None, **kwargs):
        """One environment step for all samplers (inference_agent.py:255-267 keyword set)."""
        assert self._initialized and self.step < self.T, "storage full: call after_updates()"
        t, dev, N = self.step, self.dev, self.N
        for k, v in observations.items():
            self.observations[k][t + 1].copy_(v.to(dev).reshape(self.observations[k][t + 1].shape), non_blocking=True)
        a = actions.to(dev).reshape(N)
        self.actions[t].copy_(a)
        self.prev_actions[t + 1].copy_(a)
        self.action_log_probs[t].copy_(action_log_probs.to(dev).reshape(N, 1))
        self.value_preds[t].copy_(value_preds.to(dev).reshape(N, 1))
        self.rewards[t].copy_(rewards.to(dev).reshape(N, 1))
        K = self.K
        if costs is not None:  # [N, 1] (K = 1) or [N, K]
            self.costs_k[:, t].copy_(costs.to(dev).reshape(N, K).t().reshape(K, N, 1))
        if c_value_preds is not None:
            self.c_value_preds_k[:, t].copy_(c_value_preds.to(dev).reshape(N, K).t().reshape(K, N, 1))
        self.masks[t + 1].copy_(masks.to(dev).reshape(N, 1))
        if costs is not None:  # Jc bookkeeping: totals of the episodes that ended at this step
            for k in range(K):
                ops.episode_cost_step(self.costs_k[k, t].view(N), self.masks[t + 1].view(N), self.episode_cost[k],
                                      self.cost_sum_cnt[2 * k: 2 * k + 2])
        self.step += 1

    def load_rollout(self, ro: Dict, value_preds, c_value_preds, action_log_probs):
        """Bulk fill from a synthetic rollout dict (safevla_b200.synthetic.make_rollout): host buffers in,
        one H2D copy per stream."""
        dev = self.dev
        N = ro["actions"].shape[1]
        first = {k: v[0] for k, v in ro["observations"].items()}
        if not self._initialized or self.N != N:
            self.initialize(observations=first, num_samplers=N)
        for k, v in ro["observations"].items():
            self.observations[k].copy_(v, non_blocking=True)
        K = self.K
        self.rewards.copy_(ro["rewards"], non_blocking=True)
        if K == 1:
            self.costs.copy_(ro["costs"], non_blocking=True)
        else:  # host [T, N, K] -> channel-major planes
            self.costs_k.copy_(ro["costs"].permute(2, 0, 1).unsqueeze(-1), non_blocking=True)
        self.masks.copy_(ro["masks"], non_blocking=True)
        self.actions.copy_(ro["actions"], non_blocking=True)
        self.prev_actions[1:].copy_(ro["actions"], non_blocking=True)
        self.prev_actions[0].zero_()
        self.value_preds.copy_(value_preds.reshape(self.T + 1, N, 1), non_blocking=True)
        if K == 1:
            self.c_value_preds.copy_(c_value_preds.reshape(self.T + 1, N, 1), non_blocking=True)
        else:
            self.c_value_preds_k.copy_(c_value_preds.reshape(self.T + 1, N, K).permute(2, 0, 1).unsqueeze(-1),
                                       non_blocking=True)
        self.action_log_probs.copy_(action_log_probs.reshape(self.T, N, 1), non_blocking=True)
        pair = torch.stack([ro["episode_cost_sum"].reshape(K), ro["episode_count"].reshape(1).expand(K)], 1)
        self.cost_sum_cnt.copy_(pair.reshape(2 * K), non_blocking=True)
        self.step = self.T

    def h2d_bytes(self) -> int:
        n = sum(v.numel() * v.element_size() for v in self.observations.values())
        for t in (self.rewards, self.costs_k, self.masks, self.actions, self.value_preds, self.c_value_preds_k,
                  self.action_log_probs):
            n += t.numel() * t.element_size()
        return n

    def agent_input_for_next_step(self) -> Dict:
        t = self.step
        return {"observations": {k: v[t:t + 1] for k, v in self.observations.items()}, "memory": None,
                "prev_actions": self.prev_actions[t:t + 1], "masks": self.masks[t:t + 1]}

    # ------------------------------------------------------------------ update side
    def before_updates(self, *, next_value: torch.Tensor, next_c_value: Optional[torch.Tensor] = None,
                       use_gae: bool = True, gamma: float = 0.99, tau: float = 0.95, adv_stats_callback=None,
                       normalize_advantage: bool = False, gae_algo: int = 0, **kwargs):
        """Bootstrap + GAE(reward) + GAE(cost) + advantages, one launch.  `use_gae=False` (not used
        by the shipped config) computes the plain discounted returns of the same upstream `compute_returns`."""
        N = self.N
        self.value_preds[self.T].copy_(next_value.to(self.dev).reshape(N, 1))
        K = self.K
        if next_c_value is not None and K == 1:
            self.c_value_preds[self.T].copy_(next_c_value.to(self.dev).reshape(N, 1))
        elif next_c_value is not None:  # [N, K] as the cost critic emits it -> channel-major planes
            self.c_value_preds_k[:, self.T].copy_(next_c_value.to(self.dev).reshape(N, K).t().reshape(K, N, 1).clone())
        if use_gae:
            march = lambda r, c, v, cv, out: ops.gae_dual(r, c, v, cv, self.masks, gamma, tau, gae_algo, out=out)  # noqa: E731
        else:
            march = lambda r, c, v, cv, out: ops.discounted_returns_dual(r, c, v, cv, self.masks, gamma, out=out)  # noqa: E731
        march(self.rewards, self.costs, self.value_preds, self.c_value_preds,
              (self.returns, self.c_returns, self.adv_targ, self.c_adv_targ))
        for k in range(1, K):  # further cost channels: the same march, one stream per launch
            march(self.costs_k[k], None, self.c_value_preds_k[k], None, (self.c_returns_k[k], None, self.c_adv_targ_k[k], None))
        if normalize_advantage:
            self.norm_adv_targ = self._normalized(self.adv_targ, kwargs.get("process_group"))
            self.c_norm_adv_targ_k = torch.stack([self._normalized(self.c_adv_targ_k[k], kwargs.get("process_group"))
                                                  for k in range(K)])  # each cost channel with its own statistics
            self.c_norm_adv_targ = self.c_norm_adv_targ_k[0]

    @staticmethod
    def _normalized(adv: torch.Tensor, group=None) -> torch.Tensor:
        """(adv - mean) / (std + 1e-5) with the statistics of the WHOLE batch: under data parallelism the three sums
        {sum, sum of squares, count} are all-reduced first, so every rank normalises exactly as one process would."""
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            sums = ops.advantage_sums(adv)
            dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
            return ops.normalize_advantage_from_sums(adv, sums)[0]
        return ops.normalize_advantage(adv)[0]

    def batched_experience_generator(self, num_mini_batch: int = 1) -> Iterator[Dict]:
        T, N = self.T, self.N
        assert N % num_mini_batch == 0 or num_mini_batch == 1
        per = N // num_mini_batch
        for b in range(num_mini_batch):
            sl = slice(b * per, (b + 1) * per) if num_mini_batch > 1 else slice(0, N)
            c = (lambda x: x[:, sl].contiguous()) if num_mini_batch > 1 else (lambda x: x)
            batch = {
                "observations": {k: c(v[:T]) for k, v in self.observations.items()},
                "memory": None,
                "prev_actions": c(self.prev_actions[:T]),
                "masks": c(self.masks[:T]),
                "actions": c(self.actions),
                "old_action_log_probs": c(self.action_log_probs).squeeze(-1),
                "values": c(self.value_preds[:T]),
                "c_values": c(self.c_value_preds[:T]),
                "returns": c(self.returns[:T]),
                "c_returns": c(self.c_returns[:T]),
                "adv_targ": c(self.adv_targ),
                "c_adv_targ": c(self.c_adv_targ),
            }
            if self.K > 1:  # channel-major [K, T, n, 1]
                ck = (lambda x: x[:, :, sl].contiguous()) if num_mini_batch > 1 else (lambda x: x)
                batch.update({"c_values": ck(self.c_value_preds_k[:, :T]), "c_returns": ck(self.c_returns_k[:, :T]),
                              "c_adv_targ": ck(self.c_adv_targ_k)})
            if self.norm_adv_targ is not None:
                batch["norm_adv_targ"] = c(self.norm_adv_targ)
                # K > 1: channel-major like c_adv_targ (each channel normalised with its own statistics)
                batch["c_norm_adv_targ"] = ck(self.c_norm_adv_targ_k) if self.K > 1 else c(self.c_norm_adv_targ)
            yield batch

    def after_updates(self, **kwargs):
        """Roll the last step into slot 0 for the next rollout."""
        for v in self.observations.values():
            v[0].copy_(v[self.T])
        self.masks[0].copy_(self.masks[self.T])
        self.prev_actions[0].copy_(self.prev_actions[self.T])
        self.cost_sum_cnt.zero_()  # Jc counts the episodes finished inside ONE rollout; running episode totals carry over
        self.step = 0
