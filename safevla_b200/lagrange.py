"""Device-resident Lagrange multiplier with the omnisafe 0.5.0 `Lagrange` surface
(omnisafe/common/lagrange.py, imported by the reference at training/online/loss/customized_loss.py:14;
`cost_limit` plumbed at training/online/allenact_trainer.py:22,71).  The fork's actual hyper-parameters are
unknown: OmniSafe PPOLag defaults are used and every one is a constructor argument.
lambda and its Adam state live in HBM and are updated by one tiny kernel, so the multiplier never
round-trips to the host except for logging."""
from __future__ import annotations

from typing import Optional

import torch

from . import ops


class Lagrange:
    """`cost_limit` may be a sequence (K cost channels, an extension beyond the reference's single cost): then there
    are K multipliers, each with its own limit and Adam state, updated independently and projected onto
    [0, upper_bound]; with a float it is the omnisafe object."""

    def __init__(self, cost_limit, lagrangian_multiplier_init: float = 0.001, lambda_lr: float = 0.035,
                 lambda_optimizer: str = "Adam", lagrangian_upper_bound: Optional[float] = None,
                 device: Optional[torch.device] = None):
        if lambda_optimizer != "Adam":
            raise NotImplementedError("only lambda_optimizer='Adam' (the OmniSafe default) is implemented")
        if not torch.cuda.is_available():
            raise RuntimeError("Lagrange keeps its state in HBM; no CUDA device is visible")
        dev = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        limits = [float(c) for c in cost_limit] if isinstance(cost_limit, (list, tuple)) else [float(cost_limit)]
        self.K = len(limits)
        self.cost_limits = limits
        self.cost_limit, self.lambda_lr = limits[0], float(lambda_lr)
        self.lagrangian_upper_bound = lagrangian_upper_bound
        self.lagrangian_multiplier = torch.full((self.K,), max(float(lagrangian_multiplier_init), 0.0), device=dev)
        self.state = torch.zeros(self.K, 4, device=dev)  # per channel: Adam m, v, step, last Jc

    def update_lagrange_multiplier(self, Jc) -> None:
        """omnisafe API: Jc = mean episode cost (python float or tensor; K values for K channels)."""
        jc = Jc if torch.is_tensor(Jc) else torch.tensor(Jc, dtype=torch.float32)
        jc = jc.to(self.state.device, torch.float32).reshape(self.K)
        self.update_from_sum_count(torch.stack([jc, torch.ones_like(jc)], 1).reshape(-1))

    def update_from_sum_count(self, cost_sum_cnt: torch.Tensor) -> None:
        """Per channel [sum of finished-episode costs, finished-episode count] (already all-reduced), device tensor
        of 2 K floats."""
        ub = -1.0 if self.lagrangian_upper_bound is None else float(self.lagrangian_upper_bound)
        pairs = cost_sum_cnt.contiguous()
        for k in range(self.K):
            ops.lagrange_update(self.lagrangian_multiplier[k:k + 1], self.state[k], pairs[2 * k: 2 * k + 2],
                                self.cost_limits[k], self.lambda_lr, ub)

    def state_dict(self):
        return {"lagrangian_multiplier": self.lagrangian_multiplier.clone(), "state": self.state.clone()}

    def load_state_dict(self, sd):
        self.lagrangian_multiplier.copy_(sd["lagrangian_multiplier"].reshape(self.K))
        self.state.copy_(sd["state"].reshape(self.K, 4))
