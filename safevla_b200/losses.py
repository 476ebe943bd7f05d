"""allenact loss plugins on the fused sm_100a kernel.

Same constructor signatures, `loss(step_count, batch, actor_critic_output, *args, **kwargs)` contract
and info-dict keys as the reference (training/online/loss/customized_loss.py: SafePPOLogGrad :301-449,
PPOLogGrad :163-298; allenact-fork PPOValue / SafePPOValue wired at
training/online/dinov2_vits_tsfm_base.py:336-343).  One kernel launch computes the loss AND its
gradients; the returned 0-d tensor is an autograd node whose backward just hands those gradients to
the tower outputs.  The info dict is filled from ONE device->host copy (the reference does 4-5
`.item()` syncs, customized_loss.py:443-444).
"""
from __future__ import annotations

from typing import Callable, Dict, Optional

import torch

from . import _lib as L
from . import ops


class AbstractActorCriticLoss:
    def __init__(self, *args, **kwargs):
        pass

    def loss(self, *args, **kwargs):
        raise NotImplementedError


class _FusedLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, total, grads, *inputs):
        ctx.grads = grads
        return total.view(())

    @staticmethod
    def backward(ctx, grad_out):
        # the saved gradient buffers stay d loss / d input: a second backward through this node (retain_graph, the
        # loss used in two sums) must not compound the upstream scale, so the scaling happens on a copy
        scale = grad_out.reshape(1).to(torch.float32).contiguous()
        outs = []
        for g in ctx.grads:
            outs.append(None if g is None else ops.scale_by(g.clone(), scale))
        return (None, None, *outs)


def _flat(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    return None if t is None else t.contiguous()


def _lambda_dev(lm, device) -> torch.Tensor:
    if not torch.is_tensor(lm):
        lm = torch.tensor(lm, dtype=torch.float32)
    return lm.detach().to(device=device, dtype=torch.float32).reshape(-1).contiguous()  # [1], or [K] cost channels


def fused_ppo_loss(*, logits, values, c_values, batch, hp: L.PpoHparams, lagrangian_multiplier=None,
                   adv_key="adv_targ", c_adv_key="c_adv_targ", c_returns_key="c_returns", need_grad=True):
    dev = (logits if logits is not None else (values if values is not None else c_values)).device
    lam = _lambda_dev(lagrangian_multiplier, dev) if lagrangian_multiplier is not None else None
    g = lambda k: _flat(batch[k].to(dev)) if (k in batch and batch[k] is not None) else None  # noqa: E731
    actions = g("actions") if logits is not None else None
    c_adv = g(c_adv_key) if (logits is not None and hp.use_lagrangian) else None
    if lam is not None and lam.numel() > 1 and c_adv is not None:
        # K cost channels (extension): c_adv is channel-major [K, T, N, 1]; fold (A_c,k, lambda_k) into one pair
        R = logits.numel() // logits.shape[-1]
        assert c_adv.numel() == lam.numel() * R, \
            f"cost advantages must be channel-major [K={lam.numel()}, T, N, 1] ({lam.numel() * R} values), got {tuple(c_adv.shape)}"
        c_adv, lam = ops.combine_cost_advantages(c_adv.view(lam.numel(), -1), lam)
    scal, dlogits, dvalues, dcvalues = ops.ppo_lag_fwd_bwd(
        _flat(logits.detach()) if logits is not None else None, actions,
        g("old_action_log_probs") if logits is not None else None,
        g(adv_key) if logits is not None else None,
        c_adv,
        _flat(values.detach()) if values is not None else None, g("returns") if values is not None else None,
        _flat(c_values.detach()) if c_values is not None else None,
        g(c_returns_key) if c_values is not None else None, lam, hp,
        old_values=g("values") if (values is not None and hp.use_clipped_value_loss) else None,
        old_c_values=g("c_values") if (c_values is not None and hp.use_clipped_value_loss) else None,
        want_grads=need_grad)
    inputs = [t for t in (logits, values, c_values)]
    grads = [dlogits, dvalues, dcvalues]
    live = [(i, gr) for i, gr in zip(inputs, grads) if i is not None and i.requires_grad]
    if need_grad and live:
        total = _FusedLoss.apply(scal[0:1], [gr for _, gr in live], *[i for i, _ in live])
    else:
        total = scal[0].clone()
    return total, scal


class PPO(AbstractActorCriticLoss):
    """Constructor of allenact's PPO."""

    def __init__(self, clip_param, value_loss_coef, entropy_coef, use_clipped_value_loss=True, clip_decay=None,
                 entropy_method_name="entropy", normalize_advantage=True, show_ratios=False, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.clip_param, self.value_loss_coef, self.entropy_coef = clip_param, value_loss_coef, entropy_coef
        self.use_clipped_value_loss = use_clipped_value_loss
        self.clip_decay = clip_decay if clip_decay is not None else (lambda x: 1.0)
        self.entropy_method_name, self.show_ratios = entropy_method_name, show_ratios
        self.adv_key = "norm_adv_targ" if normalize_advantage else "adv_targ"


class _LogGradBase(PPO):
    _lagrangian = False

    def __init__(self, discrete_critics: bool, action_loss_schedule: Optional[Callable], *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.discrete_critics = discrete_critics
        self.action_loss_schedule = action_loss_schedule if action_loss_schedule is not None else (lambda x: 1.0)
        self.c_adv_key = "c_" + self.adv_key

    def loss(self, step_count: int, batch: Dict, actor_critic_output, *args, **kwargs):
        dist = actor_critic_output.distributions
        logits = getattr(dist, "raw_logits", dist.logits)
        values = actor_critic_output.values
        n = logits.numel() // logits.shape[-1]
        action_weight = self.action_loss_schedule(step_count)
        hp = L.PpoHparams(self.clip_param * self.clip_decay(step_count), action_weight, self.value_loss_coef,
                          self.entropy_coef, 0.0, 1.0 / n, 1.0, int(self.use_clipped_value_loss),
                          int(self._lagrangian))
        lm = kwargs["lagrangian_multiplier"] if self._lagrangian else None
        dc_value = None
        if self.discrete_critics:
            # customized_loss.py:364-370: value_loss = 0.5 * HLGauss(extras["full_logits"], returns); the fused kernel
            # evaluates the action / entropy terms, the HL-Gauss kernel the value term (loss + d / d logits in one pass)
            ex = actor_critic_output.extras
            fl = ex["full_logits"]
            dc_value = 0.5 * ex["loss_func"](fl.reshape(-1, fl.shape[-1]), batch["returns"].to(fl.device).reshape(-1))
            values = None
        total, scal = fused_ppo_loss(logits=logits, values=values, c_values=None, batch=batch, hp=hp,
                                     lagrangian_multiplier=lm, adv_key=self.adv_key, c_adv_key=self.c_adv_key)
        s = scal.tolist()  # the one host sync of this loss
        if dc_value is not None:
            total = total + self.value_loss_coef * dc_value
            s[1] = float(dc_value)
            s[0] = float(total)
        if s[10] != 0.0:
            raise ValueError(f"{int(s[10])} action indices outside [0, {logits.shape[-1]}): batch['actions'] is corrupted "
                             "or mis-shaped")
        ex = actor_critic_output.extras
        info = {
            "ppo_total": s[0], "value": s[1], "action": s[2], "entropy": s[3],
            "bias_norm": ex.get("bias_norm", torch.tensor([0.0])),
            "weight_norm": ex.get("weight_norm", torch.tensor([0.0])),
            "weight_grad": ex.get("weight_grad_norm", torch.tensor([0.0])),
            "action_weight": action_weight,
            # extra diagnostics the fused kernel produces for free
            "approx_kl": s[5], "clip_fraction": s[6], "ratio_mean": s[7],
        }
        return total, info


class SafePPOLogGrad(_LogGradBase):
    """customized_loss.py:301-449: clipped surrogate on (A - lambda*A_c)/(1 + lambda)."""
    _lagrangian = True


class PPOLogGrad(_LogGradBase):
    """customized_loss.py:163-298."""
    _lagrangian = False


class PPOValue(AbstractActorCriticLoss):
    """allenact PPOValue: 0.5 * mean((returns - values)^2) (optionally clipped)."""
    _cost = False

    def __init__(self, clip_param: float, use_clipped_value_loss=True, clip_decay=None, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.clip_param, self.use_clipped_value_loss = clip_param, use_clipped_value_loss
        self.clip_decay = clip_decay if clip_decay is not None else (lambda x: 1.0)

    def loss(self, step_count: int, batch: Dict, actor_critic_output, *args, **kwargs):
        v = actor_critic_output.c_values if self._cost else actor_critic_output.values
        K = v.shape[-1] if self._cost else 1
        R = v.numel() // K
        if K > 1:
            # K cost channels (extension): the head emits [T, N, K] row-major while the storage keeps the targets
            # channel-major [K, T, N, 1]; pair them as [R, K] and sum the per-channel means (1 / R), exactly as
            # PPOLagUpdater's stage-0 path does
            tnk = lambda x: x.to(v.device).reshape(K, R).t().contiguous()  # noqa: E731
            batch = dict(batch, c_returns=tnk(batch["c_returns"]),
                         c_values=tnk(batch["c_values"]) if batch.get("c_values") is not None else None)
        hp = L.PpoHparams(self.clip_param * self.clip_decay(step_count), 0.0, 0.0 if self._cost else 1.0, 0.0,
                          1.0 if self._cost else 0.0, 1.0 / R, 1.0, int(self.use_clipped_value_loss), 0)
        total, scal = fused_ppo_loss(logits=None, values=None if self._cost else v, c_values=v if self._cost else None,
                                     batch=batch, hp=hp)
        return total, {"value": scal[4 if self._cost else 1].item()}


class SafePPOValue(PPOValue):
    """Fork's cost-critic twin of PPOValue on (c_values, c_returns)."""
    _cost = True


class Imitation(AbstractActorCriticLoss):
    """Expert-imitation plugin (customized_loss.py:17-83): binary cross-entropy between the policy's normalised logit of
    ONE action (`distributions.logits[:, :, action_idx]`) and the expert observation `batch["observations"][uuid]`.
    Not part of the PPO-Lagrangian update (the shipped pipeline never instantiates it); a [T, N] torch expression on the
    tower output, whose gradient reaches the actor tower through the same autograd node the fused loss uses."""

    def __init__(self, uuid: str = "expert_pickupable", action_idx: int = 8, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.uuid, self.action_idx = uuid, action_idx

    def loss(self, step_count: int, batch: Dict, actor_critic_output, *args, **kwargs):
        observations = batch["observations"]
        if self.uuid not in observations:
            raise NotImplementedError("Imitation loss requires either `expert_action` or `expert_policy`"
                                      " sensor to be active.")
        logit = actor_critic_output.distributions.logits[:, :, self.action_idx]
        target = observations[self.uuid].to(device=logit.device, dtype=logit.dtype)
        total = torch.nn.functional.binary_cross_entropy_with_logits(logit, target)
        return total, {"expert_cross_entropy": total.item()}


class _HLGaussFn(torch.autograd.Function):
    """One fused launch computes the loss and d loss / d logits; backward only scales by the upstream gradient."""

    @staticmethod
    def forward(ctx, logits, target, support, sigma):
        loss, dl, _ = ops.hl_gauss_fwd_bwd(logits.contiguous(), target, support, sigma)
        ctx.save_for_backward(dl)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        (dl,) = ctx.saved_tensors
        return dl * g, None, None, None


class HLGaussLoss(torch.nn.Module):
    """utils/loss_functions.py:7-30 -- the discrete-critic loss behind `critic_type="discrete"`
    (allenact_dino_transformer.py:743-766).  Same constructor, `forward(logits, target)`, `transform_to_probs`
    and `transform_from_probs`; `forward` and `values_from_logits` run the fused sm_100a kernel
    (`svla_hl_gauss_fwd_bwd`), the two transform_* helpers are the reference's tensor expressions (they are not on
    the update path: the fused kernel builds the soft targets itself)."""

    def __init__(self, min_value: float, max_value: float, num_bins: int, sigma: float):
        super().__init__()
        self.min_value, self.max_value, self.num_bins, self.sigma = min_value, max_value, num_bins, sigma
        self.support = torch.linspace(min_value, max_value, num_bins + 1, dtype=torch.float32)

    def _support(self, device):
        if self.support.device != device:
            self.support = self.support.to(device)
        return self.support

    def forward(self, logits: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        lg = logits.reshape(-1, logits.shape[-1]).float()
        return _HLGaussFn.apply(lg, target.reshape(-1).float(), self._support(logits.device), self.sigma)

    def values_from_logits(self, logits: torch.Tensor) -> torch.Tensor:
        """DiscreteCriticHead's value read-out: transform_from_probs(softmax(logits)) in the same fused kernel."""
        lg = logits.reshape(-1, logits.shape[-1]).float().contiguous()
        _, _, vals = ops.hl_gauss_fwd_bwd(lg, torch.zeros(lg.shape[0], device=lg.device), self._support(lg.device),
                                          self.sigma, want_grad=False, want_values=True)
        return vals.view(logits.shape[:-1])

    def transform_to_probs(self, target: torch.Tensor) -> torch.Tensor:
        sup = self._support(target.device)
        cdf = torch.special.erf((sup - target.unsqueeze(-1)) / (torch.sqrt(torch.tensor(2.0)) * self.sigma))
        z = cdf[..., -1] - cdf[..., 0]
        return (cdf[..., 1:] - cdf[..., :-1]) / z.unsqueeze(-1)

    def transform_from_probs(self, probs: torch.Tensor) -> torch.Tensor:
        sup = self._support(probs.device)
        return torch.sum(probs * ((sup[:-1] + sup[1:]) / 2), dim=-1)
