"""CPU suite (no GPU): oracle vs the reference's golden vectors, host logic, and that the C-ABI library
loads and exports every symbol include/safevla_b200.h declares."""
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

from oracle import torch_oracle as TO  # noqa: E402
from oracle.make_golden import CASES, GOLDEN_DIR, ORACLE_ONLY_CASES, build_inputs  # noqa: E402
from safevla_b200.params import ParamLayout, T5Layout, init_state_dict, tower_spec  # noqa: E402
from safevla_b200.synthetic import RolloutSpec, make_rollout, prev_actions_from  # noqa: E402


def relerr(a, b):
    return ((a.double() - b.double()).abs().max() / (b.double().abs().max() + 1e-30)).item()


# ------------------------------------------------------------------ oracle pinned on the reference's outputs
@pytest.mark.parametrize("name", list(CASES) + list(ORACLE_ONLY_CASES))
def test_oracle_reproduces_reference_golden(name):
    torch.set_num_threads(os.cpu_count() or 1)
    gold = torch.load(os.path.join(GOLDEN_DIR, name + ".pt"), weights_only=False)
    case = gold["case"]
    T, N, A, C = case["T"], case["N"], case["A"], case["C"]
    sd = init_state_dict(A, C, case["wseed"], actor_gain=1.0)
    spec, ro, extra = build_inputs(case)
    obs = {k: v[:-1] for k, v in ro["observations"].items()}
    prev, masks = prev_actions_from(ro["actions"]), ro["masks"][:-1]
    leaf = {k: (v.clone().requires_grad_(True) if "text_encoder" not in k and not k.endswith("div_term") else v)
            for k, v in sd.items()}
    out = TO.safe_model_forward(leaf, obs, prev, masks, A, C)
    assert relerr(out["logits"], gold["logits"]) < 2e-5
    assert relerr(torch.log_softmax(out["logits"], -1), gold["log_probs"]) < 2e-5
    assert relerr(out["values"], gold["values"]) < 2e-5 and relerr(out["c_values"], gold["c_values"]) < 2e-5
    ret, adv = TO.gae_returns(ro["rewards"], extra["value_preds"], ro["masks"], 0.99, 0.95)
    cret, cadv = TO.gae_returns(ro["costs"], extra["c_value_preds"], ro["masks"], 0.99, 0.95)
    out["logits"].retain_grad()
    total, info = TO.safe_ppo_log_grad(out["logits"], ro["actions"], gold["old_logp"], adv, cadv, out["values"],
                                       ret[:-1], case["lam"], entropy_coef=0.01)
    assert abs(total.item() - gold["loss_total"].item()) < 1e-5 * max(1, abs(gold["loss_total"].item()))
    for k in ("value", "action", "entropy"):
        assert abs(info[k].item() - gold["info"][k]) < 1e-5 * max(1, abs(gold["info"][k]))
    total.backward()
    assert relerr(out["logits"].grad, gold["dlogits"]) < 1e-4
    for k, gn in gold["grad_norms"].items():
        g = leaf[k].grad
        if gn is None:
            assert g is None or g.abs().max() == 0
        else:
            assert abs(g.norm().item() - gn) / max(gn, 1e-8) < 1e-3, k
    for k, gref in gold["grads"].items():
        assert relerr(leaf[k].grad, gref) < 1e-3, k
    # lambda = 0 reduces SafePPOLogGrad to PPOLogGrad (reference KAT)
    t0, _ = TO.safe_ppo_log_grad(out["logits"].detach(), ro["actions"], gold["old_logp"], adv, cadv,
                                 out["values"].detach(), ret[:-1], 0.0, entropy_coef=0.01)
    assert abs(t0.item() - gold["loss_lambda0"].item()) < 1e-5 * max(1, abs(t0.item()))


def test_oracle_reproduces_discrete_critic_golden():
    """critic_type="discrete" (DiscreteCriticHead + HL-Gauss, allenact_dino_transformer.py:152-159,434-439,743-766)
    with SafePPOLogGrad(discrete_critics=True) (customized_loss.py:364-370): the restated oracle against the reference's
    own outputs -- incl. the quirk that the value term trains the COST tower's head on the reward returns."""
    from oracle.make_golden import DISCRETE_CASES
    torch.set_num_threads(os.cpu_count() or 1)
    (name, case), = DISCRETE_CASES.items()
    gold = torch.load(os.path.join(GOLDEN_DIR, name + ".pt"), weights_only=False)
    T, N, A, C = case["T"], case["N"], case["A"], case["C"]
    sd = init_state_dict(A, C, case["wseed"], actor_gain=1.0, critic_type="discrete")
    assert sd["c_critic_tsfm.critic.fc.0.weight"].shape == (256, 512) and sd["critic.fc.2.weight"].shape == (101, 256)
    spec, ro, extra = build_inputs(case)
    obs = {k: v[:-1] for k, v in ro["observations"].items()}
    prev, masks = prev_actions_from(ro["actions"]), ro["masks"][:-1]
    leaf = {k: (v.clone().requires_grad_(True) if "text_encoder" not in k and not k.endswith("div_term") else v)
            for k, v in sd.items()}
    out = TO.safe_model_forward(leaf, obs, prev, masks, A, C)
    for k in ("logits", "values", "c_values", "full_logits"):
        assert relerr(out[k], gold[k]) < 2e-5, k
    ret, adv = TO.gae_returns(ro["rewards"], extra["value_preds"], ro["masks"], 0.99, 0.95)
    _, cadv = TO.gae_returns(ro["costs"], extra["c_value_preds"], ro["masks"], 0.99, 0.95)
    # action + entropy terms as usual (value weight 0), value term = 0.5 * HLGauss(cost-tower logits, reward returns)
    pol, info = TO.safe_ppo_log_grad(out["logits"], ro["actions"], gold["old_logp"], adv, cadv, out["values"].detach(),
                                     ret[:-1], case["lam"], entropy_coef=0.01, value_loss_coef=0.0)
    vl = 0.5 * TO.hl_gauss_loss(out["full_logits"].reshape(T * N, -1), ret[:-1].reshape(-1),
                                TO.hl_gauss_support(-5.0, 15.0, 101), 0.15)
    total = pol + 0.5 * vl
    assert abs(vl.item() - gold["info"]["value"]) < 1e-5 * max(1, abs(gold["info"]["value"]))
    assert abs(total.item() - gold["loss_total"].item()) < 1e-5 * max(1, abs(gold["loss_total"].item()))
    total.backward()
    for k, gn in gold["grad_norms"].items():
        g = leaf[k].grad
        if gn is None:
            assert g is None or g.abs().max() == 0, k
        else:
            assert abs(g.norm().item() - gn) / max(gn, 1e-8) < 1e-3, k
    # the reward critic tower receives NO gradient, the cost tower does
    assert gold["grad_norms"]["critic_tsfm.decoder.norm.weight"] is None
    assert gold["grad_norms"]["c_critic_tsfm.critic.fc.2.weight"] > 0


@pytest.mark.parametrize("name", ["step_N3_A6_C1", "step_N2_A20_C2"])
def test_oracle_step_mode_reproduces_reference_golden(name):
    """Rollout-side T = 1 path (KV-cache decoder, episode-start mask, position wrap at max_steps, reset by an
    update-mode forward): the restated oracle against the reference's per-step outputs."""
    from oracle.make_golden_step import build_inputs as step_inputs, schedule
    torch.set_num_threads(os.cpu_count() or 1)
    gold = torch.load(os.path.join(GOLDEN_DIR, name + ".pt"), weights_only=False)
    case = gold["case"]
    A, C = case["A"], case["C"]
    sd = init_state_dict(A, C, case["wseed"], actor_gain=1.0)
    ro, prev = step_inputs(case)
    st = TO.StepState(case["max_steps"])
    it = iter(gold["steps"])
    for kind, t0, t1 in schedule(case):
        if kind == "update":
            st.reset_positions()
            continue
        rec = next(it)
        obs = {k: v[t0:t1] for k, v in ro["observations"].items()}
        out = TO.safe_model_step(sd, obs, prev[t0:t1], ro["masks"][t0:t1], A, C, st)
        for k in ("logits", "values", "c_values"):
            assert relerr(out[k], rec[k]) < 2e-5, (t0, k)


def test_oracle_matches_reference_live():
    """Runs the unmodified reference (only where /root/reference exists, i.e. the build container)."""
    from oracle import ref_shim
    if not ref_shim.reference_available():
        pytest.skip("reference tree not present on this box")
    torch.set_num_threads(os.cpu_count() or 1)
    A, C, T, N = 6, 1, 5, 2
    sd = init_state_dict(A, C, 31, actor_gain=1.0)
    model = ref_shim.build_reference_model(A, C, seed=0, num_samplers=N)
    model.load_state_dict(sd, strict=True)
    ro = make_rollout(RolloutSpec(T, N, A, C, episode_end_prob=0.3, seed=9))
    obs = {k: v[:-1] for k, v in ro["observations"].items()}
    prev, masks = prev_actions_from(ro["actions"]), ro["masks"][:-1]
    with torch.no_grad():
        ref, _ = model(obs, None, prev, masks)
        mine = TO.safe_model_forward(sd, obs, prev, masks, A, C)
    assert relerr(torch.log_softmax(mine["logits"], -1), ref.distributions.logits) < 2e-5
    assert relerr(mine["values"], ref.values) < 2e-5 and relerr(mine["c_values"], ref.c_values) < 2e-5
    # KAT (iv): perturbing a finished trajectory leaves later steps / other samplers untouched
    assert torch.equal(ref.extras["stop_grad_values"], ref.c_values)


# ------------------------------------------------------------------ restated pieces: closed forms / KATs
def test_gae_closed_form_and_masks():
    T, N = 30, 3
    r, z, m = torch.ones(T, N, 1), torch.zeros(T + 1, N, 1), torch.ones(T + 1, N, 1)
    ret, adv = TO.gae_returns(r, z, m, 0.99, 0.95)
    gl = 0.99 * 0.95
    assert torch.allclose(ret[:T, 0, 0], torch.tensor([(1 - gl ** (T - t)) / (1 - gl) for t in range(T)]), rtol=1e-5)
    m[7] = 0
    ret2, _ = TO.gae_returns(r, z, m, 0.99, 0.95)
    assert torch.allclose(ret2[6], torch.ones(N, 1))
    # permuting samplers permutes outputs
    g = torch.Generator().manual_seed(0)
    r, v = torch.randn(T, N, 1, generator=g), torch.randn(T + 1, N, 1, generator=g)
    perm = torch.tensor([2, 0, 1])
    a, _ = TO.gae_returns(r, v, m, 0.99, 0.95)
    b, _ = TO.gae_returns(r[:, perm], v[:, perm], m[:, perm], 0.99, 0.95)
    assert torch.equal(a[:, perm], b)


def test_lagrange_first_step_is_sign_only():
    for jc, sign in ((10.0, +1), (0.0, -1)):
        lag = TO.LagrangeOracle(2.0, init=0.5, lr=0.035)
        assert abs(lag.update(jc) - (0.5 + sign * 0.035)) < 1e-7
    lag = TO.LagrangeOracle(2.0, init=0.01)
    assert lag.update(0.0) == 0.0  # projection onto [0, inf)


def test_adam_and_clip_oracle_match_torch():
    g = torch.Generator().manual_seed(1)
    p0, gr = torch.randn(1000, generator=g), torch.randn(1000, generator=g)
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([ref], lr=2e-5)
    p, m, v = p0.clone(), torch.zeros(1000), torch.zeros(1000)
    for step in (1, 2, 3):
        ref.grad = gr.clone() * step
        torch.nn.utils.clip_grad_norm_([ref], 0.5)
        opt.step()
        (gc,), _ = TO.clip_grad_norm([gr * step], 0.5)
        p, m, v = TO.adam_step(p, gc, m, v, step)
    assert (p - ref.detach()).abs().max() < 1e-7


def test_t5_buckets_match_oracle():
    from safevla_b200.t5_buckets import relative_position_bucket
    for L in (1, 5, 32, 200):
        pos = torch.arange(L)
        assert torch.equal(relative_position_bucket(L), TO.t5_relative_position_bucket(pos[None, :] - pos[:, None]))


# ------------------------------------------------------------------ host logic
def test_param_layout_and_state_dict_contract():
    for A, C, n_keys in ((6, 1, 411), (20, 2, 417)):
        sd = init_state_dict(A, C, 0)
        assert len(sd) == n_keys  # 414 / 417 entries (411 = 414 - the 3 in-hand tables at C=1)
        lay = ParamLayout(A, C)
        offs = [(s.offset, s.numel) for s in lay.slots.values()]
        assert all(o % 64 == 0 for o, _ in offs)
        assert all(offs[i][0] + offs[i][1] <= offs[i + 1][0] for i in range(len(offs) - 1))
        n_train = sum(s.numel for s in lay.slots.values())
        assert n_train == sum(v.numel() for k, v in sd.items() if "text_encoder" not in k and "div_term" not in k)
        # fused-GEMM adjacency the tower relies on
        for pre in ("", "critic_tsfm.", "c_critic_tsfm."):
            for i in range(3):
                q, k_, v_ = [lay.slots[pre + f"decoder.layers.{i}.attention.w{x}.weight"] for x in "qkv"]
                assert k_.offset == q.offset + q.numel and v_.offset == k_.offset + k_.numel
                w1, w3 = [lay.slots[pre + f"decoder.layers.{i}.feed_forward.w{x}.weight"] for x in "13"]
                assert w3.offset == w1.offset + w1.numel
    assert 62_000_000 < ParamLayout(6, 1).total < 63_500_000
    t5 = T5Layout()
    q, k_ = t5.slots["encoder.block.3.layer.0.SelfAttention.q.weight"], t5.slots["encoder.block.3.layer.0.SelfAttention.k.weight"]
    assert k_.offset == q.offset + q.numel


def test_synthetic_rollout_contract_and_determinism():
    spec = RolloutSpec(16, 3, 20, 2, episode_end_prob=0.2, seed=5)
    a, b = make_rollout(spec), make_rollout(spec)
    o = a["observations"]
    assert o["rgb_dinov2"].shape == (17, 3, 384, 7, 12) and o["rgb_dinov2"].dtype == torch.float32
    assert o["natural_language_spec"].shape == (17, 3, 1000) and o["natural_language_spec"].dtype == torch.uint8
    assert o["time_step"].dtype == torch.int64 and o["traj_index"].dtype == torch.int64
    assert o["an_object_is_in_hand"].shape == (17, 3, 1)
    assert all(torch.equal(a["observations"][k], b["observations"][k]) for k in o)
    assert torch.equal(a["rewards"], b["rewards"]) and torch.equal(a["actions"], b["actions"])
    # episode boundary <=> mask 0 <=> time_step reset and traj_index bump
    m = a["masks"][1:, :, 0] == 0
    assert torch.all(o["time_step"][1:][m] == 0)
    assert torch.all((o["traj_index"][1:] - o["traj_index"][:-1])[m] % 2048 == 1)
    ids, am = TO.decode_goal_ids(o["natural_language_spec"][0])
    assert ids.shape == (3, 32) and am.all() and (ids[:, -1] == 1).all()
    from safevla_b200.model import default_synthetic_tokenizer
    s = [r.numpy().tobytes().rstrip(b"\x00").decode() for r in o["natural_language_spec"][0]]
    ids2, am2 = default_synthetic_tokenizer(s)
    assert torch.equal(ids, ids2) and torch.equal(am, am2)


def test_loss_plugin_constructors_match_reference_signatures():
    from safevla_b200.losses import PPOLogGrad, PPOValue, SafePPOLogGrad, SafePPOValue
    cfg = dict(clip_param=0.1, value_loss_coef=0.5, entropy_coef=0.0, use_clipped_value_loss=False,
               action_loss_schedule=None, discrete_critics=False, normalize_advantage=False)  # dinov2_vits_tsfm_base.py:314-322
    s = SafePPOLogGrad(**cfg)
    assert s.adv_key == "adv_targ" and s.c_adv_key == "c_adv_targ" and s.action_loss_schedule(123) == 1.0
    assert PPOLogGrad(**dict(cfg, normalize_advantage=True)).adv_key == "norm_adv_targ"
    PPOValue(clip_param=0.1, use_clipped_value_loss=False)
    SafePPOValue(clip_param=0.1, use_clipped_value_loss=False)


def test_imitation_plugin_matches_reference_class():
    """customized_loss.py:17-83 on the same logits / expert observation: the unmodified reference class (through the
    shim, when the tree is present) and the closed form."""
    from safevla_b200.losses import Imitation
    from safevla_b200.misc import CategoricalDistr
    g = torch.Generator().manual_seed(0)
    raw = torch.randn(5, 3, 12, generator=g, requires_grad=True)
    expert = (torch.rand(5, 3, generator=g) < 0.4).float()

    class Out:
        distributions = CategoricalDistr(logits=raw)
    batch = {"observations": {"expert_pickupable": expert}}
    total, info = Imitation().loss(0, batch, Out)
    x = torch.log_softmax(raw.detach(), -1)[:, :, 8]
    closed = (torch.clamp(x, min=0) - x * expert + torch.log1p(torch.exp(-x.abs()))).mean()
    assert abs(total.item() - closed.item()) < 1e-6 and abs(info["expert_cross_entropy"] - closed.item()) < 1e-6
    total.backward()
    assert raw.grad is not None and raw.grad.abs().sum() > 0
    with pytest.raises(NotImplementedError):
        Imitation(uuid="absent").loss(0, batch, Out)
    from oracle import ref_shim
    if ref_shim.reference_available():
        ref_loss, _, _ = ref_shim.reference_modules()
        raw2 = raw.detach().clone().requires_grad_(True)

        class RefOut:
            distributions = torch.distributions.Categorical(logits=raw2)
        rt, rinfo = ref_loss.Imitation().loss(0, batch, RefOut)
        rt.backward()
        assert abs(rt.item() - total.item()) < 1e-6 and torch.allclose(raw2.grad, raw.grad, atol=1e-7)
        assert set(rinfo) == set(info)


def test_no_cpu_fallback():
    """Product entry points must fail loudly without a GPU instead of computing on the CPU."""
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from safevla_b200 import _lib
    from safevla_b200.model import B200SafeActorCritic
    from safevla_b200.storage import B200RolloutStorage
    with pytest.raises(RuntimeError):
        _lib.get_ctx()
    with pytest.raises(RuntimeError):
        B200SafeActorCritic(6, 1)
    with pytest.raises(RuntimeError):
        B200RolloutStorage(16)


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "safevla_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f
                assert "/root/reference" not in src, f


# ------------------------------------------------------------------ the C-ABI library
def test_library_exports_every_declared_symbol():
    from safevla_b200 import _lib
    lib = _lib.load_library()
    header = open(os.path.join(ROOT, "include", "safevla_b200.h")).read()
    declared = set(re.findall(r"\b(svla_[a-z0-9_]+)\s*\(", header))
    declared -= {"svla_ctx", "svla_stream"}
    assert len(declared) >= 30
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
        assert name in _lib.PROTOTYPES, f"{name} has no ctypes prototype"
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\bT (svla_[a-z0-9_]+)", out))
    assert declared <= exported
    assert lib.svla_version() == 100
    assert b"" == lib.svla_last_error() or isinstance(lib.svla_last_error(), bytes)


def test_library_is_sm100a_native():
    from safevla_b200 import _lib
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out[:500]


# ------------------------------------------------------------------ checkpoint formats
def test_checkpoint_format_conversion():
    from safevla_b200 import checkpoint as CK
    sd = init_state_dict(6, 1, seed=3)
    single = {k: v for k, v in sd.items() if "critic_tsfm" not in k}
    # 1. Lightning IL checkpoint: "model." prefix, bare-Linear actor head, plus frozen DINO weights to ignore
    il = {"state_dict": {("model." + k).replace("actor.linear.", "actor."): v + 1.0 for k, v in single.items()
                         if "text_encoder" not in k}}
    il["state_dict"]["model.visual_encoder.image_encoder.model.blocks.0.attn.qkv.weight"] = torch.zeros(3)
    il["state_dict"]["model.some_aux_head.weight"] = torch.zeros(2)
    assert CK.detect_format(il) == "lightning"
    conv = CK.to_model_state_dict(il)
    assert "actor.linear.weight" in conv and "actor.weight" not in conv
    assert torch.equal(conv["decoder.norm.weight"], sd["decoder.norm.weight"] + 1.0)
    new, rep = CK.merge_il_checkpoint(sd, il)
    for pre in CK.TOWER_PREFIXES:  # every tower is initialised from the IL policy
        assert torch.equal(new[pre + "decoder.norm.weight"], sd["decoder.norm.weight"] + 1.0)
        assert torch.equal(new[pre + "actor.linear.bias"], sd["actor.linear.bias"] + 1.0)
    assert torch.equal(new["visual_encoder.text_encoder.shared.weight"], sd["visual_encoder.text_encoder.shared.weight"])
    assert "some_aux_head.weight" in rep.unexpected_in_checkpoint
    assert not any("image_encoder" in k for k in rep.unexpected_in_checkpoint)
    assert any("text_encoder" in k for k in rep.missing_in_checkpoint) and "decoder.norm.weight" in rep.loaded
    # 2. allenact RL checkpoint round trip, 3. bare state dict
    ck = CK.allenact_checkpoint(sd, total_steps=7)
    assert CK.detect_format(ck) == "allenact" and ck["total_steps"] == 7
    back = CK.to_model_state_dict(ck)
    assert set(back) == set(sd) and all(torch.equal(back[k], sd[k]) for k in sd)
    assert CK.detect_format(sd) == "bare" and set(CK.to_model_state_dict(sd)) == set(sd)
    assert not any("critic_tsfm" in k for k in CK.strip_critic_towers(sd))
    with pytest.raises(ValueError):
        CK.detect_format({"weights": 1})


def test_il_checkpoint_merge_equals_reference_loader(tmp_path):
    """f-3: `checkpoint.merge_il_checkpoint` against the reference's OWN `load_pl_ckpt_allenact`
    (training/offline/train_utils.py:6-68) executed through the shim on a synthetic PyTorch-Lightning file: IL head
    names (`actor.weight`), the `model.` prefix, frozen DINO weights to be ignored, keys missing from the checkpoint
    (an IL policy has no critic head) and an unexpected key.  Every tower's constructor receives `prev_checkpoint`
    (allenact_dino_transformer.py:169-176, separate_actor_critic.py:8-11,23-25), so the reference loader is applied to
    each of the three tower modules."""
    from oracle import ref_shim
    if not ref_shim.reference_available():
        pytest.skip("reference tree not present on this box")
    import contextlib
    import io
    from safevla_b200 import checkpoint as CK
    ref_shim.reference_modules()
    from training.offline.train_utils import load_pl_ckpt_allenact
    A, C = 6, 1
    sd0 = init_state_dict(A, C, 41, actor_gain=1.0)
    il = init_state_dict(A, C, 42, actor_gain=1.0)  # a different single-tower policy plays the IL model
    ckpt_sd = {}
    for k, shape, _ in tower_spec(A, C):
        if k.startswith("critic.") or k == "decoder.layers.1.ffn_norm.weight":
            continue  # absent from the IL checkpoint: the model keeps its own value
        name = {"actor.linear.weight": "actor.weight", "actor.linear.bias": "actor.bias"}.get(k, k)
        ckpt_sd["model." + name] = il[k].clone()
    ckpt_sd["model.visual_encoder.image_encoder.model.blocks.0.attn.qkv.weight"] = torch.randn(8, 8)  # frozen DINO
    ckpt_sd["model.some_il_only_head.weight"] = torch.randn(3, 3)  # unexpected
    for k, v in il.items():
        if k.startswith("visual_encoder.text_encoder."):
            ckpt_sd["model." + k] = v.clone() + 0.5  # the IL file carries its own T5 copy
    ckpt = {"state_dict": ckpt_sd, "epoch": 3}
    path = str(tmp_path / "il.ckpt")
    torch.save(ckpt, path)
    # ---- the reference loader, tower by tower
    model = ref_shim.build_reference_model(A, C, seed=0)
    model.load_state_dict(sd0, strict=True)
    with contextlib.redirect_stdout(io.StringIO()):
        for tower in (model, model.critic_tsfm, model.c_critic_tsfm):
            load_pl_ckpt_allenact(tower, path, ckpt_prefix="model.")
    ref_state = model.state_dict()
    # ---- ours
    new, report = CK.merge_il_checkpoint(sd0, torch.load(path, weights_only=False))
    assert set(new) == set(ref_state)
    for k, v in ref_state.items():
        assert torch.equal(new[k], v), k
    assert "actor.linear.weight" in report.loaded and "decoder.layers.1.ffn_norm.weight" in report.missing_in_checkpoint
    assert "critic.fc.weight" in report.missing_in_checkpoint
    assert report.unexpected_in_checkpoint == ["some_il_only_head.weight"]
    assert torch.equal(new["critic_tsfm.visual_encoder.text_adapter.0.weight"], il["visual_encoder.text_adapter.0.weight"])
    assert torch.equal(new["c_critic_tsfm.critic.fc.weight"], sd0["c_critic_tsfm.critic.fc.weight"])


def test_hl_gauss_oracle_reproduces_reference_golden():
    from oracle.make_golden_hlgauss import inputs
    for rec in torch.load(os.path.join(GOLDEN_DIR, "hl_gauss.pt"), weights_only=False):
        c = rec["case"]
        logits, target = inputs(c)
        sup = TO.hl_gauss_support(c["vmin"], c["vmax"], c["bins"])
        assert torch.equal(sup, rec["support"])
        lg = logits.clone().requires_grad_(True)
        loss = TO.hl_gauss_loss(lg, target, sup, c["sigma"])
        loss.backward()
        assert abs(loss.item() - rec["loss"].item()) < 1e-6 * max(1, abs(rec["loss"].item()))
        assert relerr(lg.grad, rec["dlogits"]) < 1e-5 and relerr(TO.hl_gauss_value(logits, sup), rec["values"]) < 1e-5
        assert relerr(TO.hl_gauss_probs(sup, target, c["sigma"]), rec["probs"]) < 1e-6


# ------------------------------------------------------------------ vision preprocessor
def test_vit_oracle_reproduces_hf_golden_and_layouts_agree():
    from oracle import vit_oracle as VO
    from oracle.make_golden_vit import frames
    from safevla_b200.vision import hub_to_canonical, interpolate_pos_embed
    torch.set_num_threads(os.cpu_count() or 1)
    for rec in torch.load(os.path.join(GOLDEN_DIR, "dinov2_vits14.pt"), weights_only=False):
        c = rec["case"]
        sd = VO.init_hub_state_dict(c["wseed"])
        out = VO.dino_preprocess(sd, frames(c), c["crop"])
        assert out.shape == rec["out"].shape == (c["N"], 384, 7, 12)
        assert relerr(out, rec["out"]) < 1e-4
    # the product's weight loader sees the same canonical tensors from the hub and the HuggingFace layouts
    sd = VO.init_hub_state_dict(5)
    a, b = hub_to_canonical(sd), hub_to_canonical(VO.hub_to_hf(sd))
    assert set(a) == set(b) and all(torch.equal(a[k], b[k]) for k in a)
    pos = interpolate_pos_embed(a["pos"], 16, 27)
    assert pos.shape == (1 + 16 * 27, 384) and torch.equal(pos[0], a["pos"][0])
    assert torch.equal(interpolate_pos_embed(a["pos"], 37, 37), a["pos"])


def test_goal_byte_codec_matches_reference_encoding():
    """convert_string_to_byte / convert_byte_to_string restate utils/string_utils.py:11-18 (numpy `S{max_len}` view)."""
    import numpy as np
    from safevla_b200.ingest import SafeRLStepResult, convert_byte_to_string, convert_string_to_byte
    for s in ["navigate to a mug", "", "pick up the apple and bring it to the kitchen counter, in that order",
              "x" * 1000, "y" * 1200, "café table"]:
        ref = np.array([s.encode()], dtype="S1000").view("uint8")  # the reference's own expression on the utf-8 bytes
        mine = convert_string_to_byte(s, 1000)
        assert mine.dtype == np.uint8 and mine.shape == (1000,) and np.array_equal(mine, ref)
        assert convert_byte_to_string(mine) == ref.view("S1000")[0].decode(errors="ignore") or len(s.encode()) > 1000
    r = SafeRLStepResult({"a": 1}, 1.0, 0.0, False, {})
    assert r._fields == ("observation", "reward", "cost", "done", "info")  # tasks/abstract_task.py:369-380


def test_cost_channel_extension_contract():
    """K cost channels only widen the cost critic's head and the cost arrays; K = 1 is the reference key for key."""
    sd1, sd2 = init_state_dict(6, 1, 3), init_state_dict(6, 1, 3, num_cost_channels=2)
    assert list(sd1) == list(sd2)
    for k in sd1:
        if k.startswith("c_critic_tsfm.critic.fc"):
            assert sd2[k].shape[0] == 2 and sd1[k].shape[0] == 1
        else:
            assert torch.equal(sd1[k], sd2[k]), k
    assert ParamLayout(6, 1, 2).total - ParamLayout(6, 1).total == 512
    from safevla_b200.synthetic import RolloutSpec, make_rollout
    r1 = make_rollout(RolloutSpec(6, 2, 6, 1, seed=5))
    r2 = make_rollout(RolloutSpec(6, 2, 6, 1, seed=5, num_cost_channels=2))
    assert r1["costs"].shape == (6, 2, 1) and r2["costs"].shape == (6, 2, 2) and r2["episode_cost_sum"].shape == (2,)
    assert torch.equal(r1["masks"], r2["masks"]) and torch.equal(r1["rewards"], r2["rewards"])


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the driver's reference arm) runs on host cores only and prints ONE JSON line with
    the base-contract keys, `impl`, a `cpu_baseline` describing the run and a zero-copy `e2e`."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline", "gpu_launches"):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "ppo_lagrangian_update_samples_per_sec" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"] == "cfg2_64env_128step" and d["vs_baseline"] is None


# ------------------------------------------------------------------ property tests of the restated (unpinned) pieces
from hypothesis import given, settings, strategies as hst  # noqa: E402


@settings(max_examples=25, deadline=None)
@given(T=hst.integers(1, 24), N=hst.integers(1, 4), seed=hst.integers(0, 10_000), p_done=hst.floats(0.0, 0.5),
       gamma=hst.floats(0.5, 1.0), lam=hst.floats(0.0, 1.0))
def test_gae_recursion_equals_explicit_double_sum(T, N, seed, p_done, gamma, lam):
    """adv_t = sum_{k >= t} (gamma lam)^{k-t} (prod_{j=t+1..k} m_j) delta_k, evaluated directly in fp64, against the
    sequential recursion the oracle (and the march kernel, bit for bit) runs."""
    g = torch.Generator().manual_seed(seed)
    r, v = torch.randn(T, N, 1, generator=g), torch.randn(T + 1, N, 1, generator=g)
    m = (torch.rand(T + 1, N, 1, generator=g) >= p_done).float()
    ret, adv = TO.gae_returns(r, v, m, gamma, lam)
    rd, vd, md = r.double(), v.double(), m.double()
    delta = rd + gamma * vd[1:] * md[1:] - vd[:-1]
    exp = torch.zeros_like(rd)
    for t in range(T):
        w = torch.ones_like(rd[0])
        for k in range(t, T):
            if k > t:
                w = w * gamma * lam * md[k]
            exp[t] += w * delta[k]
    assert torch.allclose(adv.double(), exp, atol=2e-5, rtol=1e-5)
    assert torch.allclose(ret[:T].double(), exp + vd[:-1], atol=2e-5, rtol=1e-5) and torch.equal(ret[T], v[T])
    # lambda = 1 telescopes to the plain discounted return (`use_gae=False`)
    ret1, _ = TO.gae_returns(r, v, m, gamma, 1.0)
    ret_plain, _ = TO.gae_returns(r, v, m, gamma, 1.0, use_gae=False)
    assert torch.allclose(ret1, ret_plain, atol=1e-4, rtol=1e-5)


@settings(max_examples=25, deadline=None)
@given(R=hst.integers(1, 40), A=hst.integers(2, 20), K=hst.integers(1, 4), seed=hst.integers(0, 10_000))
def test_loss_identities_lambda_zero_and_channel_folding(R, A, K, seed):
    g = torch.Generator().manual_seed(seed)
    logits, actions = torch.randn(R, 1, A, generator=g), torch.randint(0, A, (R, 1), generator=g)
    old = torch.log_softmax(logits, -1).gather(-1, actions.unsqueeze(-1)).squeeze(-1) + 0.2 * torch.randn(R, 1, generator=g)
    adv, vals, rets = (torch.randn(R, 1, 1, generator=g) for _ in range(3))
    cadv = torch.randn(K, R, 1, 1, generator=g)
    # lambda = 0: the cost advantage drops out (SafePPOLogGrad == PPOLogGrad, customized_loss.py:208-212 vs :350-362)
    t0, _ = TO.safe_ppo_log_grad(logits, actions, old, adv, cadv[0], vals, rets, 0.0)
    t1, _ = TO.safe_ppo_log_grad(logits, actions, old, adv, torch.zeros_like(adv), vals, rets, 0.0)
    assert torch.equal(t0, t1)
    # K channels: (A - sum_k l_k A_k) / (1 + sum_k l_k) == (A - L A_eff) / (1 + L) with the folded pair
    lam = torch.rand(K, generator=g)
    L = float(lam.sum())
    eff = (lam.view(K, 1, 1, 1) * cadv).sum(0) / L if L > 0 else torch.zeros_like(adv)
    direct = (adv - (lam.view(K, 1, 1, 1) * cadv).sum(0)) / (1.0 + L)
    folded = (adv - L * eff) / (1.0 + L)
    assert torch.allclose(direct, folded, atol=1e-6, rtol=1e-5)


@settings(max_examples=20, deadline=None)
@given(n0=hst.integers(1, 200), n1=hst.integers(1, 200), seed=hst.integers(0, 10_000))
def test_two_phase_normalisation_identity(n0, n1, seed):
    """{sum, sum^2, n} of the shards added up reproduce the statistics of the concatenated batch (what the data-parallel
    advantage normalisation relies on)."""
    g = torch.Generator().manual_seed(seed)
    a, b = torch.randn(n0, generator=g) * 2 + 0.3, torch.randn(n1, generator=g) * 2 + 0.3
    s = torch.stack([torch.stack([x.double().sum(), (x.double() ** 2).sum(), torch.tensor(float(x.numel()), dtype=torch.float64)])
                     for x in (a, b)]).sum(0)
    mean = s[0] / s[2]
    var = (s[1] - s[2] * mean * mean) / (s[2] - 1) if s[2] > 1 else torch.tensor(0.0, dtype=torch.float64)
    full = torch.cat([a, b])
    exp = TO.normalize_advantage(full) if full.numel() > 1 else None
    if exp is not None:
        got = (full.double() - mean) / (var.clamp_min(0).sqrt() + 1e-5)
        assert torch.allclose(got.float(), exp, atol=1e-4, rtol=1e-4)


def test_gae_c_restatement_is_bit_identical():
    """oracle/c/gae_ref.c (plain C, no FMA contraction) and oracle/torch_oracle.py::gae_returns produce the same bits:
    two independent restatements of the recursion the CUDA march kernel is held to bit for bit."""
    import ctypes as C
    from oracle.c import build as oracle_c
    lib = C.CDLL(oracle_c.build())
    lib.gae_returns_f32.argtypes = [C.c_void_p] * 5 + [C.c_int, C.c_int, C.c_double, C.c_double]
    lib.discounted_returns_f32.argtypes = [C.c_void_p] * 4 + [C.c_int, C.c_int, C.c_double]
    g = torch.Generator().manual_seed(0)
    for T, N, gamma, lam in ((128, 64, 0.99, 0.95), (1, 1, 0.9, 0.5), (37, 5, 0.97, 1.0), (16, 3, 1.0, 0.0)):
        r, v = torch.randn(T, N, 1, generator=g) * 3, torch.randn(T + 1, N, 1, generator=g) * 3
        m = (torch.rand(T + 1, N, 1, generator=g) > 0.1).float()
        ret, adv = torch.empty_like(v), torch.empty_like(r)
        lib.gae_returns_f32(r.data_ptr(), v.data_ptr(), m.data_ptr(), ret.data_ptr(), adv.data_ptr(), T, N, gamma, lam)
        ret_t, adv_t = TO.gae_returns(r, v, m, gamma, lam)
        assert torch.equal(ret, ret_t) and torch.equal(adv, adv_t), (T, N)
        lib.discounted_returns_f32(r.data_ptr(), v.data_ptr(), m.data_ptr(), ret.data_ptr(), T, N, gamma)
        assert torch.equal(ret, TO.gae_returns(r, v, m, gamma, lam, use_gae=False)[0])
