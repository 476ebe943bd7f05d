"""TEST INFRASTRUCTURE ONLY.  Mints tests/golden/*.pt by running the reference's own code
(unmodified, from /root/reference, through oracle/ref_shim.py) on seeded synthetic rollouts.

    python -m oracle.make_golden            # in the build container (needs /root/reference)

Each fixture stores only the *recipe* for its inputs (seeds + RolloutSpec; weights come from
safevla_b200.params.init_state_dict, rollouts from safevla_b200.synthetic.make_rollout, both
deterministic CPU generators) and the reference's outputs: logits / values / c_values, the
SafePPOLogGrad (stage 1-2) and value-only (stage 0) losses, d loss / d logits, the L2 norm of
the gradient of EVERY trainable tensor, and a few small gradient tensors in full.
"""
from __future__ import annotations

import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import ref_shim, torch_oracle as TO  # noqa: E402
from safevla_b200.params import init_state_dict  # noqa: E402
from safevla_b200.synthetic import RolloutSpec, make_rollout, prev_actions_from  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

CASES = {
    # BASELINE.json configs[0]: 1 env x 16 steps, 6 actions, one camera, 32-token prompt
    "cfg1_T16_N1_A6_C1": dict(T=16, N=1, A=6, C=1, end_prob=1 / 6, wseed=11, rseed=1234, lam=0.37),
    # two cameras + in-hand sensor, 20 actions, two samplers (S = 201)
    "T12_N2_A20_C2": dict(T=12, N=2, A=20, C=2, end_prob=1 / 5, wseed=12, rseed=4321, lam=0.05),
}
# Pinned on the CPU only (tests/test_cpu.py): the oracle at the BASELINE trajectory length -- a full 128-step decoder
# window with several episode boundaries inside it (trajectory mask, time encoding, 20-action head)
ORACLE_ONLY_CASES = {
    "T128_N1_A20_C1": dict(T=128, N=1, A=20, C=1, end_prob=1 / 40, wseed=13, rseed=777, lam=0.2),
}
# critic_type="discrete": HL-Gauss DiscreteCriticHead in every tower, SafePPOLogGrad(discrete_critics=True)
# (allenact_dino_transformer.py:152-159,434-439,743-766; customized_loss.py:364-370)
DISCRETE_CASES = {
    "disc_T10_N2_A6_C1": dict(T=10, N=2, A=6, C=1, end_prob=1 / 5, wseed=14, rseed=555, lam=0.3, critic_type="discrete"),
}
FULL_GRADS = ["actor.linear.weight", "actor.linear.bias", "critic_tsfm.critic.fc.weight",
              "visual_encoder.fusion_token", "last_actions_embed.weight",
              "critic_tsfm.visual_encoder.visual_sensor_token_raw_navigation_camera",
              "decoder.norm.weight", "visual_encoder.fusion_xformer.layers.0.norm1.weight",
              "visual_encoder.text_adapter.1.bias", "critic_tsfm.decoder.layers.2.ffn_norm.weight"]


def build_inputs(case: dict):
    """Everything the update consumes, derived from seeds only (runs on the GPU box too)."""
    spec = RolloutSpec(case["T"], case["N"], case["A"], case["C"], episode_end_prob=case["end_prob"],
                       seed=case["rseed"])
    ro = make_rollout(spec)
    g = torch.Generator().manual_seed(case["rseed"] + 99)
    T, N = case["T"], case["N"]
    extra = {
        "value_preds": torch.randn(T + 1, N, 1, generator=g),
        "c_value_preds": torch.randn(T + 1, N, 1, generator=g).abs(),
        "logp_noise": 0.15 * torch.randn(T, N, generator=g),
    }
    return spec, ro, extra


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    ref_loss, _, _ = ref_shim.reference_modules()
    only = set(sys.argv[1:])
    for name, case in {**CASES, **ORACLE_ONLY_CASES, **DISCRETE_CASES}.items():
        if only and name not in only:
            continue
        T, N, A, C = case["T"], case["N"], case["A"], case["C"]
        ctype = case.get("critic_type", "linear")
        sd = init_state_dict(A, C, case["wseed"], actor_gain=1.0, critic_type=ctype)
        model = ref_shim.build_reference_model(A, C, seed=0, num_samplers=N, critic_type=ctype)
        missing = model.load_state_dict(sd, strict=True)
        spec, ro, extra = build_inputs(case)
        obs = {k: v[:-1] for k, v in ro["observations"].items()}
        prev = prev_actions_from(ro["actions"])
        masks = ro["masks"][:-1]

        raw = {}

        def keep_raw(_m, _i, o):  # un-normalised actor output (Categorical.logits is log-softmaxed)
            o.retain_grad()
            raw["logits"] = o

        model.actor.linear.register_forward_hook(keep_raw)

        def fwd():
            model.zero_grad(set_to_none=True)
            out, _ = model(obs, None, prev, masks)
            return out

        out = fwd()
        logits = raw["logits"]
        with torch.no_grad():
            old_logp = out.distributions.log_prob(ro["actions"]) + extra["logp_noise"]
        ret, adv = TO.gae_returns(ro["rewards"], extra["value_preds"], ro["masks"], 0.99, 0.95)
        cret, cadv = TO.gae_returns(ro["costs"], extra["c_value_preds"], ro["masks"], 0.99, 0.95)
        batch = {"actions": ro["actions"], "old_action_log_probs": old_logp, "adv_targ": adv,
                 "c_adv_targ": cadv, "values": extra["value_preds"][:-1], "returns": ret[:-1]}
        loss_fn = ref_loss.SafePPOLogGrad(clip_param=0.1, value_loss_coef=0.5, entropy_coef=0.01,
                                          use_clipped_value_loss=False, action_loss_schedule=None,
                                          discrete_critics=(ctype == "discrete"), normalize_advantage=False)
        total, info = loss_fn.loss(step_count=0, batch=batch, actor_critic_output=out,
                                   lagrangian_multiplier=torch.tensor(case["lam"]))
        total.backward()
        gold = {
            "case": case, "logits": logits.detach().clone(),
            "log_probs": out.distributions.logits.detach().clone(), "values": out.values.detach().clone(),
            "c_values": out.c_values.detach().clone(), "old_logp": old_logp,
            "loss_total": total.detach().clone(),
            "info": {k: (float(v) if not torch.is_tensor(v) else v.clone()) for k, v in info.items()},
            "dlogits": logits.grad.clone(),
            "grad_norms": {k: (p.grad.norm().item() if p.grad is not None else None)
                           for k, p in model.named_parameters() if "text_encoder" not in k},
            "grads": {k: p.grad.clone() for k, p in model.named_parameters()
                      if p.grad is not None and (k in FULL_GRADS or ".critic.fc." in k)},
        }
        if ctype == "discrete":
            gold["full_logits"] = out.extras["full_logits"].detach().clone()
            with torch.no_grad():
                mine = TO.safe_model_forward(sd, obs, prev, masks, A, C)
            for k in ("logits", "values", "c_values", "full_logits"):
                err = (mine[k] - gold[k]).abs().max().item()
                print(f"{name}: oracle vs reference {k}: max abs err {err:.3e}")
                assert err < 2e-5, (k, err)
            torch.save(gold, os.path.join(GOLDEN_DIR, name + ".pt"))
            print(f"wrote {name}.pt  loss={float(total):.6f}")
            continue
        # lambda == 0 KAT: SafePPOLogGrad == PPOLogGrad bit-for-bit
        out = fwd()
        t0, _ = loss_fn.loss(0, batch, out, lagrangian_multiplier=torch.tensor(0.0))
        ppo = ref_loss.PPOLogGrad(clip_param=0.1, value_loss_coef=0.5, entropy_coef=0.01,
                                  use_clipped_value_loss=False, action_loss_schedule=None,
                                  discrete_critics=False, normalize_advantage=False)
        t1, _ = ppo.loss(0, batch, out)
        assert torch.equal(t0, t1)
        gold["loss_lambda0"] = t0.detach().clone()

        # stage 0 (value-only) losses: PPOValue + SafePPOValue are fork code -> restated oracle on
        # the reference model's outputs (dinov2_vits_tsfm_base.py:337-343,350)
        out = fwd()
        l0 = TO.ppo_value_loss(out.values, ret[:-1]) + TO.ppo_value_loss(out.c_values, cret[:-1])
        l0.backward()
        gold["stage0_loss"] = l0.detach().clone()
        gold["stage0_grad_norms"] = {k: (p.grad.norm().item() if p.grad is not None else None)
                                     for k, p in model.named_parameters() if "text_encoder" not in k}

        # cross-check the restated oracle against the reference right here
        with torch.no_grad():
            mine = TO.safe_model_forward(sd, obs, prev, masks, A, C)
        for k in ("logits", "values", "c_values"):
            err = (mine[k] - gold[k]).abs().max().item()
            print(f"{name}: oracle vs reference {k}: max abs err {err:.3e}")
            assert err < 2e-5, (k, err)
        torch.save(gold, os.path.join(GOLDEN_DIR, name + ".pt"))
        print(f"wrote {name}.pt  loss={float(total):.6f} info={ {k: v for k, v in gold['info'].items() if isinstance(v, float)} }")


if __name__ == "__main__":
    main()
