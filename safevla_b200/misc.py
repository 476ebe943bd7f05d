"""Plain-data mirrors of the allenact types the reference path exchanges
(allenact.base_abstractions.misc / .distributions in the un-vendored fork; fields witnessed at
architecture/models/allenact_transformer_models/separate_actor_critic.py:31-36 and
training/online/loss/customized_loss.py:329-332)."""
from __future__ import annotations

from typing import Any, Dict, Generic, Optional, TypeVar

import torch

T = TypeVar("T")


class Memory(dict):
    """The towers are KV-cache transformers: `recurrent_memory_specification` is None
    (allenact_dino_transformer.py:277-278) and Memory passes through untouched."""


class ActorCriticOutput(Generic[T]):
    def __init__(self, distributions, values, extras: Dict[str, Any]):
        self.distributions, self.values, self.extras = distributions, values, extras


class SafeActorCriticOutput(Generic[T]):
    def __init__(self, distributions, values, c_values, extras: Dict[str, Any]):
        self.distributions, self.values, self.c_values, self.extras = distributions, values, c_values, extras


class CategoricalDistr(torch.distributions.Categorical):
    """Same surface as allenact's CategoricalDistr.  `raw_logits` keeps the un-normalised actor
    output (torch's Categorical stores log-softmaxed logits) so the fused loss kernel can consume it
    and route the gradient back to the actor head.  Sampling stays the stock torch op on our logits,
    which is what makes sampled indices bit-identical to the reference given identical logits and
    seed (hard part 4)."""

    def __init__(self, logits: torch.Tensor, **kw):
        self.raw_logits = logits
        super().__init__(logits=logits.detach() if not logits.requires_grad else logits, **kw)

    def mode(self):
        return self._param.argmax(dim=-1, keepdim=False)

    def log_prob(self, value: torch.Tensor):
        if value.shape == self.logits.shape[:-1]:
            return super().log_prob(value)
        if value.shape == self.logits.shape[:-1] + (1,):
            return super().log_prob(value.squeeze(-1)).unsqueeze(-1)
        raise NotImplementedError(f"bad action shape {tuple(value.shape)}")


class nvtx_range:
    """NVTX range around a phase of the update (tracing row): shows up in Nsight Systems / ncu
    range filters (`--nvtx --nvtx-include "update/"`); a no-op cost of two driver calls when no tool is attached."""

    def __init__(self, name: str):
        self.name = name

    def __enter__(self):
        import torch
        torch.cuda.nvtx.range_push(self.name)
        return self

    def __exit__(self, *exc):
        import torch
        torch.cuda.nvtx.range_pop()
        return False
