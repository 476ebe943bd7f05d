#!/usr/bin/env python
"""Per-tensor gradient cosine / norm error of the bf16 path against the CPU oracle, with and without dropout
(the oracle gets the device-generated masks).  Diagnostic for tests/test_dropout_gpu.py."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import torch_oracle as TO
from safevla_b200.losses import SafePPOLogGrad
from safevla_b200.model import B200SafeActorCritic
from safevla_b200.params import init_state_dict
from safevla_b200.synthetic import RolloutSpec, make_rollout, prev_actions_from
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_dropout_gpu import _oracle_masks

dev = torch.device("cuda:0")
T, N, A, C = 6, 2, 6, 1
sd = init_state_dict(A, C, seed=17, actor_gain=1.0)
ro = make_rollout(RolloutSpec(T, N, A, C, episode_end_prob=0.25, seed=3))
obs = {k: v[:-1] for k, v in ro["observations"].items()}
prev, masks = prev_actions_from(ro["actions"]), ro["masks"][:-1]
for p in (0.0, 0.1):
    model = B200SafeActorCritic(A, C, precision="bf16", state_dict=sd, device=dev, dropout=p, dropout_seed=77, extras="off")
    model.set_trainable_towers((0, 1))
    out, _ = model({k: v.to(dev) for k, v in obs.items()}, None, prev.to(dev), masks.to(dev))
    drop = _oracle_masks(dev, model, T * N, 117, step=1) if p > 0 else None
    leaf = {k: (v.clone().requires_grad_(True) if "text_encoder" not in k and not k.endswith("div_term") else v) for k, v in sd.items()}
    ref = TO.safe_model_forward(leaf, obs, prev, masks, A, C, drop=drop)
    g = torch.Generator().manual_seed(1)
    vp, cvp = torch.randn(T + 1, N, 1, generator=g), torch.randn(T + 1, N, 1, generator=g).abs()
    ret, adv = TO.gae_returns(ro["rewards"], vp, ro["masks"], 0.99, 0.95)
    _, cadv = TO.gae_returns(ro["costs"], cvp, ro["masks"], 0.99, 0.95)
    old_logp = torch.log_softmax(ref["logits"].detach(), -1).gather(-1, ro["actions"].unsqueeze(-1)).squeeze(-1) + 0.1
    loss = SafePPOLogGrad(clip_param=0.1, value_loss_coef=0.5, entropy_coef=0.01, use_clipped_value_loss=False,
                          action_loss_schedule=None, discrete_critics=False, normalize_advantage=False)
    batch = {"actions": ro["actions"].to(dev), "old_action_log_probs": old_logp.to(dev), "adv_targ": adv.to(dev),
             "c_adv_targ": cadv.to(dev), "values": vp[:-1].to(dev), "returns": ret[:-1].to(dev)}
    total, _ = loss.loss(0, batch, out, lagrangian_multiplier=torch.tensor(0.3))
    total.backward()
    ref_total, _ = TO.safe_ppo_log_grad(ref["logits"], ro["actions"], old_logp, adv, cadv, ref["values"], ret[:-1], 0.3, entropy_coef=0.01)
    ref_total.backward()
    rows = []
    for k, v in leaf.items():
        if not (torch.is_tensor(v) and v.requires_grad) or v.grad is None or k.startswith("c_critic_tsfm."):
            continue
        gm, gr = model.get_parameter(k).grad.detach().cpu().double().reshape(-1), v.grad.double().reshape(-1)
        if gr.norm() < 1e-12:
            continue
        rows.append(((gm @ gr / (gm.norm() * gr.norm())).item(), abs(gm.norm().item() - gr.norm().item()) / gr.norm().item(), k))
    rows.sort()
    print(f"p = {p}: loss {total.item():.5f} vs {ref_total.item():.5f}")
    for cos, err, k in rows[:8]:
        print(f"   cos {cos:.5f} norm err {err:.4f} {k}")
