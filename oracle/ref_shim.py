"""TEST INFRASTRUCTURE ONLY — never imported by the product package.

Runs the *unmodified* SafeVLA hot-path sources from ``/root/reference`` on CPU
in this container, behind stand-ins for the packages that are not installed
(``allenact`` fork, ``gym``, ``omnisafe``, ``open_clip`` ...).  Used by
``oracle/make_golden.py`` to mint the golden vectors under ``tests/golden/``
and by ``tests/test_oracle_vs_reference.py`` (skipped when ``/root/reference``
is absent, i.e. on the GPU box).

What is executed verbatim from the reference tree:
  * training/online/loss/customized_loss.py            (SafePPOLogGrad, PPOLogGrad, ...)
  * architecture/models/allenact_transformer_models/allenact_dino_transformer.py
  * architecture/models/allenact_transformer_models/separate_actor_critic.py
  * training/online/third_party_models/llama/model.py
  * utils/loss_functions.py
  * PositionalEncoder from architecture/models/transformer_models/text_cond_visual_encoder.py
    (AST-extracted: the module itself cannot be imported on Python >= 3.11).

What is a stand-in (the allenact fork / omnisafe are NOT in the reference tree,
SURVEY.md section 8c): ``PPO`` ctor, ``CategoricalDistr``, ``LinearActorHead``,
``LinearCriticHead``, ``ActorCriticOutput``/``SafeActorCriticOutput``,
``VisualNavActorCritic``.  They follow upstream allenact semantics as recalled;
the pretrained T5 / tokenizer are replaced by a seeded ``T5Config()`` random
init and the synthetic id tokenizer below (no weights are available offline).
"""
from __future__ import annotations

import ast
import contextlib
import io
import logging
import math
import os
import sys
import types
from typing import Any, Dict, Generic, List, Optional, TypeVar

import numpy as np
import torch
import torch.nn as nn

REFERENCE_ROOT = os.environ.get("SAFEVLA_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "training", "online", "loss"))


# --------------------------------------------------------------------------------------
# synthetic goal codec shared by the reference rig, the oracle and the product
# --------------------------------------------------------------------------------------
class SyntheticIdTokenizer:
    """Stand-in for the t5-small sentencepiece tokenizer (not available offline).

    A goal string is a blank-separated list of decimal token ids; EOS (1) is
    appended, rows are right-padded with 0 to the longest row and the attention
    mask marks real tokens -- the same output contract as the HF tokenizer call
    at allenact_dino_transformer.py:600-602.
    """

    class _Enc(dict):
        def to(self, device):
            return SyntheticIdTokenizer._Enc({k: v.to(device) for k, v in self.items()})

    def __call__(self, goals: List[str], return_tensors="pt", padding=True):
        rows = [[int(tok) for tok in g.split()] + [1] for g in goals]
        L = max(len(r) for r in rows)
        ids = torch.zeros(len(rows), L, dtype=torch.int64)
        am = torch.zeros(len(rows), L, dtype=torch.int64)
        for i, r in enumerate(rows):
            ids[i, : len(r)] = torch.tensor(r, dtype=torch.int64)
            am[i, : len(r)] = 1
        return SyntheticIdTokenizer._Enc(input_ids=ids, attention_mask=am)


def _byte_to_string(bytes_to_decode: np.ndarray, max_len: Optional[int] = None) -> str:
    raw = np.ascontiguousarray(bytes_to_decode).astype(np.uint8).tobytes()
    if max_len is not None:
        raw = raw[:max_len]
    return raw.rstrip(b"\x00").decode()


def _string_to_byte(s: str, max_len: int) -> np.ndarray:
    out = np.zeros(max_len, dtype=np.uint8)
    b = s.encode()[:max_len]
    out[: len(b)] = np.frombuffer(b, dtype=np.uint8)
    return out


# --------------------------------------------------------------------------------------
# stand-in modules
# --------------------------------------------------------------------------------------
def _mod(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _pkg(name: str, **attrs) -> types.ModuleType:
    m = _mod(name, **attrs)
    m.__path__ = []  # type: ignore[attr-defined]
    return m


_INSTALLED = False


def install_stubs() -> None:
    global _INSTALLED
    if _INSTALLED:
        return
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")

    # ---- gym ----
    class Space:
        pass

    class Discrete(Space):
        def __init__(self, n):
            self.n = int(n)

    class Box(Space):
        def __init__(self, low=0.0, high=1.0, shape=(), dtype=np.float32):
            self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), dtype

    class DictSpace(Space):
        def __init__(self, spaces=None, **kw):
            self.spaces = dict(spaces or {}, **kw)

    spaces = _mod("gym.spaces", Discrete=Discrete, Box=Box, Dict=DictSpace, Space=Space)
    _pkg("gym", spaces=spaces, Space=Space)

    # ---- allenact ----
    class Distr:
        pass

    class CategoricalDistr(torch.distributions.Categorical, Distr):
        def mode(self):
            return self._param.argmax(dim=-1, keepdim=False)

        def log_prob(self, value):
            if value.shape == self.logits.shape[:-1]:
                return super().log_prob(value)
            if value.shape == self.logits.shape[:-1] + (1,):
                return super().log_prob(value.squeeze(-1)).unsqueeze(-1)
            raise NotImplementedError(f"bad action shape {value.shape}")

    T = TypeVar("T")

    class ActorCriticOutput(Generic[T]):
        def __init__(self, distributions, values, extras):
            self.distributions, self.values, self.extras = distributions, values, extras

    class SafeActorCriticOutput(Generic[T]):
        def __init__(self, distributions, values, c_values, extras):
            self.distributions, self.values = distributions, values
            self.c_values, self.extras = c_values, extras

    class Memory(dict):
        pass

    class AbstractActorCriticLoss:
        def __init__(self, *args, **kwargs):
            pass

        def loss(self, *args, **kwargs):
            raise NotImplementedError

    class PPO(AbstractActorCriticLoss):
        def __init__(self, clip_param, value_loss_coef, entropy_coef, use_clipped_value_loss=True,
                     clip_decay=None, entropy_method_name="entropy", normalize_advantage=True,
                     show_ratios=False, *args, **kwargs):
            super().__init__(*args, **kwargs)
            self.clip_param = clip_param
            self.value_loss_coef = value_loss_coef
            self.entropy_coef = entropy_coef
            self.use_clipped_value_loss = use_clipped_value_loss
            self.clip_decay = clip_decay if clip_decay is not None else (lambda x: 1.0)
            self.entropy_method_name = entropy_method_name
            self.show_ratios = show_ratios
            self.adv_key = "norm_adv_targ" if normalize_advantage else "adv_targ"

    class LinearCriticHead(nn.Module):
        def __init__(self, input_size: int):
            super().__init__()
            self.fc = nn.Linear(input_size, 1)
            nn.init.orthogonal_(self.fc.weight)
            nn.init.constant_(self.fc.bias, 0)

        def forward(self, x):
            return self.fc(x).view(*x.shape[:2], -1)

    class LinearActorHead(nn.Module):
        def __init__(self, num_inputs: int, num_outputs: int):
            super().__init__()
            self.linear = nn.Linear(num_inputs, num_outputs)
            nn.init.orthogonal_(self.linear.weight, gain=0.01)
            nn.init.constant_(self.linear.bias, 0)

        def forward(self, x):
            return CategoricalDistr(logits=self.linear(x))

    class VisualNavActorCritic(nn.Module):
        def __init__(self, action_space, observation_space, hidden_size=512, multiple_beliefs=False,
                     beliefs_fusion=None, auxiliary_uuids=None, **kwargs):
            super().__init__()
            self.action_space = action_space
            self.observation_space = observation_space
            self._hidden_size = hidden_size
            self.multiple_beliefs = multiple_beliefs
            self.beliefs_fusion = beliefs_fusion
            self.auxiliary_uuids = auxiliary_uuids
            self.aux_models = nn.ModuleDict()

        def create_aux_models(self, obs_embed_size, action_embed_size):
            return None

    class MultiAuxTaskNegEntropyLoss:
        UUID = "multitask_entropy"

    class Sensor:
        def __init__(self, *a, **k):
            pass

    class Lagrange:  # import-only at customized_loss.py:14
        pass

    _pkg("allenact")
    _pkg("allenact.algorithms")
    _pkg("allenact.algorithms.onpolicy_sync")
    _pkg("allenact.algorithms.onpolicy_sync.losses", PPO=PPO)
    _mod("allenact.algorithms.onpolicy_sync.losses.abstract_loss",
         AbstractActorCriticLoss=AbstractActorCriticLoss, ObservationType=Dict[str, Any])
    _mod("allenact.algorithms.onpolicy_sync.policy", LinearActorHead=LinearActorHead,
         LinearCriticHead=LinearCriticHead, DistributionType=Any, ObservationType=Dict[str, Any])
    _pkg("allenact.base_abstractions")
    _mod("allenact.base_abstractions.distributions", Distr=Distr, CategoricalDistr=CategoricalDistr)
    _mod("allenact.base_abstractions.misc", ActorCriticOutput=ActorCriticOutput,
         SafeActorCriticOutput=SafeActorCriticOutput, Memory=Memory)
    _mod("allenact.base_abstractions.sensor", Sensor=Sensor)
    _pkg("allenact.embodiedai")
    _pkg("allenact.embodiedai.aux_losses")
    _mod("allenact.embodiedai.aux_losses.losses", MultiAuxTaskNegEntropyLoss=MultiAuxTaskNegEntropyLoss)
    _pkg("allenact.embodiedai.models")
    _mod("allenact.embodiedai.models.visual_nav_models", VisualNavActorCritic=VisualNavActorCritic,
         FusionType=Any)
    _pkg("allenact.utils")
    _mod("allenact.utils.system", get_logger=lambda: logging.getLogger("allenact-stub"))
    _pkg("omnisafe")
    _pkg("omnisafe.common")
    _mod("omnisafe.common.lagrange", Lagrange=Lagrange)

    # ---- reference modules that cannot be imported here ----
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _pkg("architecture.models.transformer_models").__path__ = []  # do not run its __init__
    src_path = os.path.join(REFERENCE_ROOT, "architecture", "models", "transformer_models",
                            "text_cond_visual_encoder.py")
    with open(src_path) as f:
        tree = ast.parse(f.read())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "PositionalEncoder")
    ns: Dict[str, Any] = {"math": math, "torch": torch, "nn": nn}
    exec(compile(ast.Module(body=[cls], type_ignores=[]), src_path, "exec"), ns)
    _mod("architecture.models.transformer_models.text_cond_visual_encoder",
         PositionalEncoder=ns["PositionalEncoder"])
    _mod("utils.string_utils", convert_byte_to_string=_byte_to_string,
         convert_string_to_byte=_string_to_byte)
    _mod("utils.bbox_utils", get_best_of_two_bboxes=None)
    _mod("utils.nn_utils", debug_model_info=lambda *a, **k: None)
    _INSTALLED = True


# --------------------------------------------------------------------------------------
# builders
# --------------------------------------------------------------------------------------
def reference_modules():
    """Returns (loss_module, model_module, separate_module) imported from the reference tree."""
    install_stubs()
    import transformers

    # offline: T5 geometry from T5Config() defaults (= t5-small), random init under the caller's seed
    def _t5_from_pretrained(name, *a, **k):
        return transformers.T5EncoderModel(transformers.T5Config())

    transformers.T5EncoderModel.from_pretrained = staticmethod(_t5_from_pretrained)  # type: ignore
    transformers.AutoTokenizer.from_pretrained = staticmethod(lambda *a, **k: SyntheticIdTokenizer())  # type: ignore
    import training.online.loss.customized_loss as ref_loss  # noqa
    import architecture.models.allenact_transformer_models.allenact_dino_transformer as ref_model  # noqa
    import architecture.models.allenact_transformer_models.separate_actor_critic as ref_sep  # noqa
    return ref_loss, ref_model, ref_sep


def build_reference_model(num_actions: int, num_cameras: int, seed: int, max_steps: int = 500,
                          num_samplers: int = 1, dropout_off: bool = True, critic_type: str = "linear"):
    """SafeDinoLLAMATxNavActorCriticSeparate with the kwargs of
    training/online/dinov2_vits_tsfm_base.py:233-270 (C = 1 drops the manipulation camera
    and the in-hand sensor, as `full_sensor=False` does at :225-231)."""
    _, _, ref_sep = reference_modules()
    import gym

    spaces = {
        "rgb_dinov2": gym.spaces.Box(shape=(7, 12, 384)),
        "natural_language_spec": gym.spaces.Box(shape=(1000,), dtype=np.uint8),
    }
    if num_cameras == 2:
        spaces["manipulation_rgb_dinov2"] = gym.spaces.Box(shape=(7, 12, 384))
    torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        model = ref_sep.SafeDinoLLAMATxNavActorCriticSeparate(
            action_space=gym.spaces.Discrete(num_actions),
            observation_space=gym.spaces.Dict(spaces),
            goal_sensor_uuid="natural_language_spec",
            rgb_dino_preprocessor_uuid="rgb_dinov2",
            manipulation_rgb_dino_preprocessor_uuid="manipulation_rgb_dinov2" if num_cameras == 2 else None,
            an_object_is_in_hand_uuid="an_object_is_in_hand" if num_cameras == 2 else None,
            num_tx_layers=3, num_tx_heads=8, hidden_size=512, goal_dims=512,
            add_prev_actions=True, add_prev_action_null_token=True, auxiliary_uuids=[],
            max_steps=max_steps, time_step_uuid="time_step",
            initial_tgt_cache_shape=(max_steps, num_samplers, 512),
            traj_idx_uuid="traj_index", traj_max_idx=2048,
            relevant_object_box_uuid=None, accurate_object_box_uuid=None, prev_checkpoint=None,
            critic_type=critic_type,
        )
    if dropout_off:
        # parity setting: the model forces train(); switch every Dropout off instead
        for m in model.modules():
            if isinstance(m, nn.Dropout):
                m.p = 0.0
            if isinstance(m, nn.MultiheadAttention):
                m.dropout = 0.0
            if isinstance(m, nn.TransformerEncoderLayer):
                pass
        for name, m in model.named_modules():
            if name.endswith("text_encoder"):
                m.eval()
    return model
