"""TEST INFRASTRUCTURE ONLY.  Mints tests/golden/hl_gauss.pt from the reference's own HLGaussLoss
(utils/loss_functions.py, imported unmodified from /root/reference):  python -m oracle.make_golden_hlgauss"""
from __future__ import annotations

import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.environ.get("SAFEVLA_REFERENCE", "/root/reference")

CASES = [dict(R=64, bins=101, vmin=-10.0, vmax=10.0, sigma=0.75 * 20.0 / 101, seed=5),
         dict(R=200, bins=51, vmin=0.0, vmax=25.0, sigma=0.4, seed=6),
         dict(R=7, bins=300, vmin=-1.0, vmax=1.0, sigma=0.01, seed=7)]


def inputs(c):
    g = torch.Generator().manual_seed(c["seed"])
    logits = torch.randn(c["R"], c["bins"], generator=g) * 2
    target = c["vmin"] + (c["vmax"] - c["vmin"]) * torch.rand(c["R"], generator=g)
    return logits, target


def main():
    sys.path.insert(0, REF)
    from utils.loss_functions import HLGaussLoss  # the reference's class
    out = []
    for c in CASES:
        logits, target = inputs(c)
        lg = logits.clone().requires_grad_(True)
        m = HLGaussLoss(c["vmin"], c["vmax"], c["bins"], c["sigma"])
        loss = m(lg, target)
        loss.backward()
        with torch.no_grad():
            vals = m.transform_from_probs(torch.softmax(logits, -1))
        out.append({"case": c, "loss": loss.detach().clone(), "dlogits": lg.grad.clone(), "values": vals,
                    "probs": m.transform_to_probs(target), "support": m.support.clone()})
        print(c, float(loss))
    torch.save(out, os.path.join(ROOT, "tests", "golden", "hl_gauss.pt"))


if __name__ == "__main__":
    main()
