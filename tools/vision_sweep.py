#!/usr/bin/env python
"""Throughput of the rollout-side vision preprocessor (uint8 224 x 384 frames -> [N, 384, 7, 12] DINOv2 features)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from safevla_b200 import _lib as L
from safevla_b200.vision import B200DinoViTPreprocessor, init_hub_state_dict
dev = torch.device("cuda:0")
pre = B200DinoViTPreprocessor("rgb", init_hub_state_dict(0), precision="bf16", device=dev)
GF = 12 * (2 * 433 * 384 * (1152 + 384 + 2 * 1536) + 4 * 433 * 433 * 64 * 6) / 1e9 + 2 * 432 * 588 * 384 / 1e9
for N in (8, 64, 256, 1024):
    fr = torch.randint(0, 256, (N, 224, 384, 3), dtype=torch.uint8, device=dev)
    for _ in range(2):
        pre.encode(fr)
    e = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
    e[0].record()
    for i in range(5):
        pre.encode(fr); e[i + 1].record()
    torch.cuda.synchronize()
    t = min(e[i].elapsed_time(e[i + 1]) for i in range(5)) * 1e-3
    print(f"N={N:5d} frames: {t*1e3:8.2f} ms  {N/t:9.0f} frames/s  {N*GF/t/1e3:7.1f} TF/s ({GF:.1f} GF/frame)")
L.profile_start()
pre.encode(fr)
prof = L.profile_stop()
tot = sum(v[0] for v in prof.values())
for k, (ms, n) in sorted(prof.items(), key=lambda kv: -kv[1][0]):
    print(f"  {k:24s} {ms:8.3f} ms {100*ms/tot:5.1f} %  {n:4d} calls")
