"""TEST INFRASTRUCTURE ONLY.  CPU restatement of one whole constrained-PPO update (the schedule the
allenact fork's engine runs): GAE on both streams, then
`update_repeats` x [3-tower forward, SafePPOLogGrad (stage 1) or PPOValue+SafePPOValue (stage 0),
backward, clip_grad_norm_, Adam], then the Lagrange-multiplier step.  Uses torch autograd over the
restated forward of oracle/torch_oracle.py.  Also the `cpu_baseline` ("port") leg of bench.py."""
from __future__ import annotations

from typing import Dict

import torch

from . import torch_oracle as TO


def oracle_update(sd: Dict[str, torch.Tensor], ro: Dict, value_preds, c_value_preds, old_logp, cfg,
                  num_actions: int, num_cameras: int, share_t5: bool = True):
    """Returns (new state dict, new lambda, info).  `cfg` is a safevla_b200.updater.PPOLagConfig-like
    object (only plain attributes are read)."""
    T = ro["actions"].shape[0]
    obs = {k: v[:T] for k, v in ro["observations"].items()}
    prev = torch.cat([torch.zeros_like(ro["actions"][:1]), ro["actions"][:-1]], 0)
    masks = ro["masks"][:T]
    ret, adv = TO.gae_returns(ro["rewards"], value_preds, ro["masks"], cfg.gamma, cfg.gae_lambda)
    # K cost channels (K = 1: the reference; K > 1: the extension of DESIGN.md section 7): one GAE per channel
    K = ro["costs"].shape[-1]
    c_value_preds = c_value_preds.reshape(T + 1, -1, K)
    chan = [TO.gae_returns(ro["costs"][..., k:k + 1], c_value_preds[..., k:k + 1], ro["masks"], cfg.gamma,
                           cfg.gae_lambda) for k in range(K)]
    cret, cadv = chan[0]
    limits = list(cfg.cost_limit) if isinstance(cfg.cost_limit, (list, tuple)) else [cfg.cost_limit]
    assert len(limits) == K
    params = {k: v.clone() for k, v in sd.items()}
    trainable = [k for k in params if "text_encoder" not in k and not k.endswith("div_term")]
    m = {k: torch.zeros_like(params[k]) for k in trainable}
    v = {k: torch.zeros_like(params[k]) for k in trainable}
    lags = [TO.LagrangeOracle(l, cfg.lambda_init, cfg.lambda_lr, cfg.lambda_upper_bound) for l in limits]
    lag = lags[0]
    info = {}
    # the frozen T5 does not change inside an update: evaluate it once (identical result)
    ids, am = TO.decode_goal_ids(obs["natural_language_spec"].reshape(T * masks.shape[1], -1))
    with torch.no_grad():
        th = TO.t5_encoder(params, "visual_encoder.text_encoder.", ids, am)
    for rep in range(cfg.update_repeats):
        leaf = {k: (params[k].clone().requires_grad_(True) if k in trainable else params[k]) for k in params}
        outs = {}
        for pre in TO.TOWER_PREFIXES:
            outs[pre] = TO.tower_forward(leaf, pre, obs, th, prev, masks, num_actions, num_cameras)
        logits, values, c_values = outs[""][0], outs["critic_tsfm."][1], outs["c_critic_tsfm."][1]
        if cfg.stage == 0:
            total = TO.ppo_value_loss(values, ret[:T])
            for k in range(K):  # sum over channels of the per-channel SafePPOValue means
                total = total + TO.ppo_value_loss(c_values[..., k:k + 1], chan[k][0][:T])
        else:
            lam_sum = sum(l.lam for l in lags)
            cadv_eff = cadv if K == 1 else (sum(l.lam * chan[k][1] for k, l in enumerate(lags)) / lam_sum
                                            if lam_sum > 0 else torch.zeros_like(cadv))
            # (A - sum_k lam_k A_c,k) / (1 + sum_k lam_k)
            total, _ = TO.safe_ppo_log_grad(logits, ro["actions"], old_logp, adv, cadv_eff, values, ret[:T],
                                            lag.lam if K == 1 else lam_sum,
                                            clip_param=cfg.clip_param, value_loss_coef=cfg.value_loss_coef,
                                            entropy_coef=cfg.entropy_coef)
        total.backward()
        info["last_total"] = float(total.detach())
        with_grad = [k for k in trainable if leaf[k].grad is not None]
        grads, norm = TO.clip_grad_norm([leaf[k].grad for k in with_grad], cfg.max_grad_norm)
        info["grad_norm"] = float(norm)
        step = rep + 1
        for k, g in zip(with_grad, grads):
            params[k], m[k], v[k] = TO.adam_step(params[k], g, m[k], v[k], step, cfg.lr, cfg.betas[0], cfg.betas[1],
                                                 cfg.eps)
    cnt = float(ro["episode_count"])
    sums = ro["episode_cost_sum"].reshape(-1).tolist()
    # no finished episode in the rollout -> no Jc estimate: the multiplier and its Adam state are left untouched
    lams = [(l.update(sums[k] / cnt) if cnt >= 1.0 else l.lam) for k, l in enumerate(lags)]
    return params, (lams[0] if K == 1 else lams), info
