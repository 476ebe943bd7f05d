"""TEST INFRASTRUCTURE ONLY.  Mints tests/golden/dinov2_vits14.pt: the vision-preprocessor pipeline with the ViT
evaluated by HuggingFace transformers' Dinov2Model (independent implementation of the DINOv2 architecture) on seeded
frames and seeded weights; also checks oracle/vit_oracle.py against it.     python -m oracle.make_golden_vit"""
from __future__ import annotations

import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import vit_oracle as VO  # noqa: E402

CASES = [dict(name="frames_224x384", N=2, H=224, W=384, crop=(3, 3), wseed=31, fseed=41),
         dict(name="frames_224x224", N=1, H=224, W=224, crop=(0, 0), wseed=32, fseed=42)]


def frames(c):
    g = torch.Generator().manual_seed(c["fseed"])
    return torch.randint(0, 256, (c["N"], c["H"], c["W"], 3), generator=g, dtype=torch.uint8)


def main():
    from transformers import Dinov2Config, Dinov2Model
    torch.set_num_threads(os.cpu_count() or 1)
    out = []
    for c in CASES:
        sd = VO.init_hub_state_dict(c["wseed"])
        cfg = Dinov2Config(hidden_size=384, num_hidden_layers=12, num_attention_heads=6, mlp_ratio=4, image_size=518,
                           patch_size=14, hidden_act="gelu", layer_norm_eps=1e-6, qkv_bias=True, layerscale_value=1.0,
                           use_swiglu_ffn=False, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0,
                           drop_path_rate=0.0)
        model = Dinov2Model(cfg).eval()
        missing = model.load_state_dict(VO.hub_to_hf(sd), strict=True)
        fr = frames(c)
        with torch.no_grad():
            x = VO.normalize_frames(fr)
            x = x[:, :, :, c["crop"][0]: x.shape[-1] - c["crop"][1]]
            tok = model(pixel_values=x).last_hidden_state[:, 1:]
            ph, pw = x.shape[-2] // 14, x.shape[-1] // 14
            ref = F.adaptive_avg_pool2d(tok.permute(0, 2, 1).reshape(c["N"], 384, ph, pw), (7, 12))
            mine = VO.dino_preprocess(sd, fr, c["crop"])
        err = ((mine - ref).abs().max() / ref.abs().max()).item()
        print(c["name"], "restated oracle vs HF Dinov2Model: rel err", err, "| out absmax", ref.abs().max().item())
        assert err < 1e-4
        out.append({"case": c, "out": ref.clone(), "tokens_head": tok[:, :4].clone()})
    torch.save(out, os.path.join(ROOT, "tests", "golden", "dinov2_vits14.pt"))


if __name__ == "__main__":
    main()
