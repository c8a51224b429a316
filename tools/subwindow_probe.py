#!/usr/bin/env python
"""Does the conf/label spill of phase A survive in L2 until phase C when windows are small?  (development probe)

    python tools/subwindow_probe.py            # times 64 maps processed in windows of S = 2..64 images
    ncu ... -k regex:k_select python tools/subwindow_probe.py --only 4 --reps 2   # DRAM bytes of phase C at S = 4
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hiast_b200.ias_engine import IASEngine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--images', type=int, default=64)
ap.add_argument('--only', type=int, default=0)
ap.add_argument('--reps', type=int, default=5)
ap.add_argument('--mode', type=int, default=0)
args = ap.parse_args()
C, H, W = 19, 1024, 2048
g = torch.Generator(device='cuda').manual_seed(1234)
pool = torch.empty(args.images, C, H, W, device='cuda')
for i in range(args.images):
    if i % 2 == 0:
        pool[i] = torch.randn(C, H, W, generator=g, device='cuda') * 3
    else:
        low = torch.randn(1, C, 32, 64, generator=g, device='cuda') * 4
        pool[i] = torch.nn.functional.interpolate(low, size=(H, W), mode='bilinear', align_corners=True)[0]
        pool[i] += torch.randn(C, H, W, generator=g, device='cuda') * 0.5
sizes = [args.only] if args.only else [2, 4, 8, 16, 32, 64]
for S in sizes:
    eng = IASEngine(C, H, W, 2, 0.5, 0.9, 8.0, 0.99, S, hist_mode=args.mode)

    def run():
        for i in range(0, args.images, S):
            eng.process(pool[i:i + S])
    run()
    torch.cuda.synchronize()
    ts = []
    for _ in range(args.reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        run()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = sorted(ts)[len(ts) // 2]
    print('window %3d images: %.3f ms per %d images  (%.1f us/img, %.0f img/s)' % (S, ms, args.images, ms * 1e3 / args.images,
                                                                                 args.images / ms * 1e3), flush=True)
