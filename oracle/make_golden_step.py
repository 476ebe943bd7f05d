"""TEST INFRASTRUCTURE ONLY.  Mints tests/golden/step_*.pt: the reference model (unmodified, from /root/reference,
through oracle/ref_shim.py) driven one step at a time (T = 1, KV-cache decoder, episode-start mask), the way the
rollout loop calls it (inference_agent.py:229-296; allenact_dino_transformer.py:376-406).

    python -m oracle.make_golden_step        # in the build container (needs /root/reference)

Schedule per case: K single steps of a seeded synthetic rollout -> one update-mode forward over the first
T_upd steps (resets the cache position, :376-377) -> K2 more single steps (stale cache rows beyond the position
must not be read).  Stored: the reference's raw actor logits, values and c_values of every single step.
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
    "step_N3_A6_C1": dict(T=14, N=3, A=6, C=1, end_prob=1 / 4, wseed=21, rseed=777, K=9, T_upd=5, K2=5, max_steps=8),
    "step_N2_A20_C2": dict(T=10, N=2, A=20, C=2, end_prob=1 / 3, wseed=22, rseed=778, K=6, T_upd=4, K2=4, max_steps=16),
}


def schedule(case):
    """[(kind, t0, t1)]: 'step' uses rollout row t0; 'update' uses rows [t0, t1)."""
    sched = [("step", t, t + 1) for t in range(case["K"])]
    sched.append(("update", 0, case["T_upd"]))
    sched += [("step", t, t + 1) for t in range(case["K"], case["K"] + case["K2"])]
    return sched


def build_inputs(case):
    spec = RolloutSpec(case["T"], case["N"], case["A"], case["C"], episode_end_prob=case["end_prob"], seed=case["rseed"])
    ro = make_rollout(spec)
    return ro, prev_actions_from(ro["actions"])


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    for name, case in CASES.items():
        A, C, N = case["A"], case["C"], case["N"]
        sd = init_state_dict(A, C, case["wseed"], actor_gain=1.0)
        model = ref_shim.build_reference_model(A, C, seed=0, num_samplers=N, max_steps=case["max_steps"])
        model.load_state_dict(sd, strict=True)
        raw = {}
        model.actor.linear.register_forward_hook(lambda _m, _i, o: raw.__setitem__("logits", o))
        ro, prev = build_inputs(case)
        st = TO.StepState(case["max_steps"])
        gold = {"case": case, "steps": []}
        with torch.no_grad():
            for kind, t0, t1 in schedule(case):
                obs = {k: v[t0:t1] for k, v in ro["observations"].items()}
                out, _ = model(obs, None, prev[t0:t1], ro["masks"][t0:t1])
                if kind == "update":
                    st.reset_positions()
                    continue
                rec = {"t": t0, "logits": raw["logits"].clone(), "values": out.values.clone(),
                       "c_values": out.c_values.clone()}
                mine = TO.safe_model_step(sd, obs, prev[t0:t1], ro["masks"][t0:t1], A, C, st)
                for k in ("logits", "values", "c_values"):
                    err = (mine[k] - rec[k]).abs().max().item()
                    assert err < 2e-5, (name, t0, k, err)
                gold["steps"].append(rec)
        torch.save(gold, os.path.join(GOLDEN_DIR, name + ".pt"))
        print(f"wrote {name}.pt: {len(gold['steps'])} single steps, oracle restatement within 2e-5 of the reference")


if __name__ == "__main__":
    main()
