"""Parameter inventory of the three-tower SafeVLA actor-critic and its flat-arena layout.

Key names and shapes are the reference's ``state_dict`` contract (SURVEY.md App. B.3;
architecture/models/allenact_transformer_models/allenact_dino_transformer.py:47-195,478-569,
separate_actor_critic.py:8-37, training/online/third_party_models/llama/model.py:425-437),
so reference checkpoints load key-for-key.  Trainable tensors of all three towers live in one
contiguous fp32 arena (`ParamLayout`), which is what the fused clip+Adam kernel and the single
NCCL all-reduce operate on; the frozen T5-small encoder has its own arena, shared by the towers.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from dataclasses import dataclass
from typing import Dict, List, Tuple

import torch

TOWERS = ("", "critic_tsfm.", "c_critic_tsfm.")  # actor, reward critic, cost critic
D = 512
FF = 2048
DEC_FF = 1536  # 256*ceil((2*4*512/3)/256), llama/model.py:349-353
T5_VOCAB = 32128
ALIGN = 64  # elements; keeps every tensor 256 B (fp32) / 128 B (bf16) aligned for TMA


DC_HIDDEN, DC_BINS = 256, 101  # DiscreteCriticHead: 512 -> 256 -> 101 bins (allenact_dino_transformer.py:152-159,743-766)
DC_MIN, DC_MAX, DC_SIGMA = -5.0, 15.0, 0.15


def tower_spec(num_actions: int, num_cameras: int, num_values: int = 1,
               critic_type: str = "linear") -> List[Tuple[str, Tuple[int, ...], str]]:
    """(name, shape, init) for the trainable tensors of ONE tower, in reference state_dict order.  `num_values` is the
    width of the critic head: 1 in the reference; K for the cost tower of the K-cost-channel extension.
    critic_type "linear" (shipped) or "discrete" (HL-Gauss DiscreteCriticHead: critic.fc.0 / critic.fc.2)."""
    ve = "visual_encoder."
    s: List[Tuple[str, Tuple[int, ...], str]] = [
        (ve + "fusion_token", (D,), "token"),
        (ve + "visual_sensor_token_raw_navigation_camera", (D,), "token"),
    ]
    if num_cameras == 2:
        s.append((ve + "visual_sensor_token_raw_manipulation_camera", (D,), "token"))
    s += [
        (ve + "text_adapter.0.weight", (D, D), "linear"),
        (ve + "text_adapter.0.bias", (D,), "bias:512"),
        (ve + "text_adapter.1.weight", (D,), "ones"),
        (ve + "text_adapter.1.bias", (D,), "zeros"),
        (ve + "visual_compressor.0.weight", (D, 384, 1, 1), "linear"),
        (ve + "visual_compressor.0.bias", (D,), "bias:384"),
        (ve + "visual_compressor.2.weight", (D, D, 1, 1), "linear"),
        (ve + "visual_compressor.2.bias", (D,), "bias:512"),
        (ve + "visual_adapter.0.weight", (D, D), "linear"),
        (ve + "visual_adapter.0.bias", (D,), "bias:512"),
        (ve + "visual_adapter.1.weight", (D,), "ones"),
        (ve + "visual_adapter.1.bias", (D,), "zeros"),
    ]
    for i in range(3):
        p = ve + f"fusion_xformer.layers.{i}."
        s += [
            (p + "self_attn.in_proj_weight", (3 * D, D), "xavier"),
            (p + "self_attn.in_proj_bias", (3 * D,), "zeros"),
            (p + "self_attn.out_proj.weight", (D, D), "linear"),
            (p + "self_attn.out_proj.bias", (D,), "zeros"),
            (p + "linear1.weight", (FF, D), "linear"),
            (p + "linear1.bias", (FF,), "bias:512"),
            (p + "linear2.weight", (D, FF), "linear"),
            (p + "linear2.bias", (D,), "bias:2048"),
            (p + "norm1.weight", (D,), "ones"),
            (p + "norm1.bias", (D,), "zeros"),
            (p + "norm2.weight", (D,), "ones"),
            (p + "norm2.bias", (D,), "zeros"),
        ]
    if num_cameras == 2:
        s.append(("object_in_hand_embed.weight", (3, D), "embed"))
    s.append(("last_actions_embed.weight", (num_actions + 2, D), "embed"))
    for i in range(3):
        p = f"decoder.layers.{i}."
        s += [
            (p + "attention.wq.weight", (D, D), "linear"),
            (p + "attention.wk.weight", (D, D), "linear"),
            (p + "attention.wv.weight", (D, D), "linear"),
            (p + "attention.wo.weight", (D, D), "linear"),
            (p + "feed_forward.w1.weight", (DEC_FF, D), "linear"),  # w1|w3 adjacent: one fused GEMM
            (p + "feed_forward.w3.weight", (DEC_FF, D), "linear"),
            (p + "feed_forward.w2.weight", (D, DEC_FF), "linear"),
            (p + "attention_norm.weight", (D,), "ones"),
            (p + "ffn_norm.weight", (D,), "ones"),
        ]
    s += [
        ("decoder.norm.weight", (D,), "ones"),
        ("decoder.output.weight", (D, D), "linear"),
        ("actor.linear.weight", (num_actions, D), "actor"),
        ("actor.linear.bias", (num_actions,), "zeros"),
    ]
    if critic_type == "discrete":
        assert num_values == 1, "the discrete critic head predicts one value"
        s += [("critic.fc.0.weight", (DC_HIDDEN, D), "critic"), ("critic.fc.0.bias", (DC_HIDDEN,), "zeros"),
              ("critic.fc.2.weight", (DC_BINS, DC_HIDDEN), "critic"), ("critic.fc.2.bias", (DC_BINS,), "zeros")]
    elif critic_type == "linear":
        s += [("critic.fc.weight", (num_values, D), "critic"), ("critic.fc.bias", (num_values,), "zeros")]
    else:
        raise NotImplementedError(f"critic_type={critic_type!r}: 'linear' and 'discrete' are built")
    return s


def t5_spec() -> List[Tuple[str, Tuple[int, ...], str]]:
    """Frozen T5-small encoder (HF T5Config() defaults), keys relative to '...text_encoder.'."""
    s: List[Tuple[str, Tuple[int, ...], str]] = [("shared.weight", (T5_VOCAB, D), "normal:1.0")]
    for i in range(6):
        a = f"encoder.block.{i}.layer.0."
        s += [
            (a + "SelfAttention.q.weight", (D, D), f"normal:{(D * 64) ** -0.5}"),
            (a + "SelfAttention.k.weight", (D, D), f"normal:{D ** -0.5}"),
            (a + "SelfAttention.v.weight", (D, D), f"normal:{D ** -0.5}"),
            (a + "SelfAttention.o.weight", (D, D), f"normal:{D ** -0.5}"),
        ]
        if i == 0:
            s.append((a + "SelfAttention.relative_attention_bias.weight", (32, 8), f"normal:{D ** -0.5}"))
        s.append((a + "layer_norm.weight", (D,), "ones"))
        f = f"encoder.block.{i}.layer.1."
        s += [
            (f + "DenseReluDense.wi.weight", (FF, D), f"normal:{D ** -0.5}"),
            (f + "DenseReluDense.wo.weight", (D, FF), f"normal:{FF ** -0.5}"),
            (f + "layer_norm.weight", (D,), "ones"),
        ]
    s.append(("encoder.final_layer_norm.weight", (D,), "ones"))
    return s


def _init(shape, kind: str, g: torch.Generator, actor_gain: float) -> torch.Tensor:
    if kind == "ones":
        return torch.ones(shape)
    if kind == "zeros":
        return torch.zeros(shape)
    if kind == "token":  # 0.1 * U(0,1), allenact_dino_transformer.py:515,526
        return 0.1 * torch.rand(shape, generator=g)
    if kind == "embed":  # U(-0.01, 0.01), :132,220
        return (torch.rand(shape, generator=g) * 2 - 1) * 0.01
    if kind == "linear":  # nn.Linear / Conv2d default: U(+-1/sqrt(fan_in))
        fan_in = int(torch.Size(shape[1:]).numel())
        return (torch.rand(shape, generator=g) * 2 - 1) / math.sqrt(fan_in)
    if kind.startswith("bias:"):
        return (torch.rand(shape, generator=g) * 2 - 1) / math.sqrt(int(kind[5:]))
    if kind == "xavier":
        a = math.sqrt(6.0 / (shape[0] + shape[1]))
        return (torch.rand(shape, generator=g) * 2 - 1) * a
    if kind.startswith("normal:"):
        return torch.randn(shape, generator=g) * float(kind[7:])
    if kind in ("actor", "critic"):  # orthogonal rows, gain 0.01 / 1.0 (allenact heads)
        w = torch.randn(shape[1], shape[0], generator=g)
        q, r = torch.linalg.qr(w)
        q = q * torch.sign(torch.diagonal(r)).unsqueeze(0)
        return q.T.contiguous() * (actor_gain if kind == "actor" else 1.0)
    raise ValueError(kind)


def tower_values(tower_index: int, num_cost_channels: int) -> int:
    """Critic-head width of tower `tower_index`: the cost critic (index 2) predicts one value per cost channel."""
    return num_cost_channels if tower_index == 2 else 1


def init_state_dict(num_actions: int, num_cameras: int, seed: int, actor_gain: float = 0.01,
                    num_cost_channels: int = 1, critic_type: str = "linear") -> "OrderedDict[str, torch.Tensor]":
    """Deterministic (CPU generator) random init with the reference's key set; the three towers
    get independent trainable weights and the SAME frozen T5 weights (as `from_pretrained` gives)."""
    g = torch.Generator().manual_seed(seed)
    t5 = OrderedDict((k, _init(shape, kind, g, actor_gain)) for k, shape, kind in t5_spec())
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    div_term = torch.exp(torch.arange(0, D, 2) * (-math.log(10000.0) / D))
    for ti, pre in enumerate(TOWERS):
        for k, shape, kind in tower_spec(num_actions, num_cameras, tower_values(ti, num_cost_channels), critic_type):
            sd[pre + k] = _init(shape, kind, g, actor_gain)
            if k == "last_actions_embed.weight":
                sd[pre + k][num_actions + 1].zero_()  # padding_idx row
                sd[pre + "time_encoder.div_term"] = div_term.clone()
        te = pre + "visual_encoder.text_encoder."
        for k, v in t5.items():
            sd[te + k] = v
            if k == "shared.weight":
                sd[te + "encoder.embed_tokens.weight"] = v
    return sd


@dataclass
class Slot:
    name: str
    shape: Tuple[int, ...]
    offset: int  # in elements
    numel: int


class ParamLayout:
    """Offsets of every trainable tensor inside the flat arena (all three towers)."""

    def __init__(self, num_actions: int, num_cameras: int, num_cost_channels: int = 1, critic_type: str = "linear"):
        self.num_actions, self.num_cameras, self.num_cost_channels = num_actions, num_cameras, num_cost_channels
        self.critic_type = critic_type
        self.slots: "OrderedDict[str, Slot]" = OrderedDict()
        self.tower_range: Dict[str, Tuple[int, int]] = {}
        off = 0
        for ti, pre in enumerate(TOWERS):
            start = off
            for k, shape, _ in tower_spec(num_actions, num_cameras, tower_values(ti, num_cost_channels), critic_type):
                n = int(torch.Size(shape).numel())
                self.slots[pre + k] = Slot(pre + k, shape, off, n)
                off += (n + ALIGN - 1) // ALIGN * ALIGN
            self.tower_range[pre] = (start, off)
        self.total = off

    def view(self, arena: torch.Tensor, name: str) -> torch.Tensor:
        s = self.slots[name]
        return arena[s.offset: s.offset + s.numel].view(s.shape)


class T5Layout:
    def __init__(self):
        self.slots: "OrderedDict[str, Slot]" = OrderedDict()
        off = 0
        for k, shape, _ in t5_spec():
            n = int(torch.Size(shape).numel())
            self.slots[k] = Slot(k, shape, off, n)
            off += (n + ALIGN - 1) // ALIGN * ALIGN
        self.total = off

    def view(self, arena: torch.Tensor, name: str) -> torch.Tensor:
        s = self.slots[name]
        return arena[s.offset: s.offset + s.numel].view(s.shape)
