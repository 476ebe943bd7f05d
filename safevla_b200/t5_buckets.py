"""T5 relative-position bucketing (bidirectional, 32 buckets, max distance 128) as integer index
math -- HF transformers modeling_t5.py `_relative_position_bucket`, SURVEY.md A.7."""
import math

import torch


def relative_position_bucket(L: int, num_buckets: int = 32, max_distance: int = 128) -> torch.Tensor:
    """int64 [L(query), L(key)] bucket of (key - query)."""
    pos = torch.arange(L)
    rel = pos[None, :] - pos[:, None]
    half = num_buckets // 2
    out = (rel > 0).long() * half
    rel = rel.abs()
    exact = half // 2
    big = exact + (torch.log(rel.float().clamp(min=1) / exact) / math.log(max_distance / exact) * (half - exact)).long()
    big = big.clamp(max=half - 1)
    return out + torch.where(rel < exact, rel, big)
