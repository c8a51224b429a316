#!/usr/bin/env python
"""Window size vs phase-A time per image, interleaved (development probe)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hiast_b200 import ops  # noqa: E402
from tools.bench_kernels import time_variants  # noqa: E402

C, H, W, B = 19, 1024, 2048, 2
g = torch.Generator(device='cuda').manual_seed(1234)
N = 148
pool = torch.empty(N, C, H, W, device='cuda')
for i in range(N):
    if i % 2 == 0:
        pool[i].normal_(generator=g).mul_(3)
    else:
        low = torch.randn(1, C, 32, 64, generator=g, device='cuda') * 4
        pool[i] = torch.nn.functional.interpolate(low, size=(H, W), mode='bilinear', align_corners=True)[0]
        pool[i] += torch.randn(C, H, W, generator=g, device='cuda') * 0.5
key_lo = ops.ias_key_lo(C)
conf = torch.empty(N, H, W, device='cuda')
label = torch.empty(N, H, W, dtype=torch.uint8, device='cuda')
hist = ops.ias_new_hist(N // B, C, key_lo, 'cuda')
variants = {}
for n in (16, 64, 74, 128):
    for mode in (80, 83):
        variants['%d/m%d' % (n, mode)] = (lambda n=n, mode=mode: ops.ias_softmax_hist(pool[:n], B, key_lo, conf[:n], label[:n],
                                                                                  hist[:n // B], hist_mode=mode))
for rep in range(2):
    res = time_variants(variants, rounds=6, inner=2)
    print('  '.join('%s: %.2f' % (n, ms * 1e3 / int(n.split('/')[0])) for n, ms in res.items()), 'us/img', flush=True)
