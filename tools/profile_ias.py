#!/usr/bin/env python
"""Small driver for ncu: a few IAS windows (16 maps of 19x1024x2048) through phases A/B/C.

    ncu --set full --clock-control none --import-source on -k regex:k_softmax_hist -s 2 -c 1 \
        -o gpurun_out/phase_a python tools/profile_ias.py --mode 6
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hiast_b200.ias_engine import IASEngine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--mode', type=int, default=0)
ap.add_argument('--images', type=int, default=16)
ap.add_argument('--windows', type=int, default=3)
ap.add_argument('--dist', default='mixed')
args = ap.parse_args()
C, H, W = 19, 1024, 2048
g = torch.Generator(device='cuda').manual_seed(1234)
pool = torch.empty(args.images, C, H, W, device='cuda')
for i in range(args.images):
    if args.dist == 'diffuse' or (args.dist == 'mixed' and i % 2 == 0):
        pool[i] = torch.randn(C, H, W, generator=g, device='cuda') * 3
    else:
        scale = 60 if args.dist == 'saturated' else 4
        low = torch.randn(1, C, 32, 64, generator=g, device='cuda') * scale
        pool[i] = torch.nn.functional.interpolate(low, size=(H, W), mode='bilinear', align_corners=True)[0]
        pool[i] += torch.randn(C, H, W, generator=g, device='cuda') * 0.5
eng = IASEngine(C, H, W, 2, 0.5, 0.9, 8.0, 0.99, args.images, hist_mode=args.mode)
for _ in range(args.windows):
    eng.process(pool)
torch.cuda.synchronize()
print('ok', eng.thr_state[:3].tolist())
