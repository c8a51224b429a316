#!/usr/bin/env python
"""Window size vs throughput of the three-kernel pipeline (development probe)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hiast_b200 import ops  # noqa: E402
from hiast_b200.ias_engine import IASEngine  # noqa: E402

C, H, W, B = 19, 1024, 2048, 2
g = torch.Generator(device='cuda').manual_seed(1234)
N = 296
pool = torch.empty(N, C, H, W, device='cuda')
for i in range(N):
    if i % 2 == 0:
        pool[i].normal_(generator=g).mul_(3)
    else:
        low = torch.randn(1, C, 32, 64, generator=g, device='cuda') * 4
        pool[i] = torch.nn.functional.interpolate(low, size=(H, W), mode='bilinear', align_corners=True)[0]
        pool[i] += torch.randn(C, H, W, generator=g, device='cuda') * 0.5


def timeit(fn, iters=5):
    fn(); fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


for n in (64, 74, 128, 148, 222, 296):
    eng = IASEngine(C, H, W, B, 0.5, 0.9, 8.0, 0.99, n, fused=False)
    x = pool[:n]
    a = timeit(lambda: eng.phase_a(x))
    b = timeit(lambda: (eng.phase_a(x), eng.phase_b(0, n))) - a
    c = timeit(lambda: eng.phase_c(0, n))
    full = timeit(lambda: eng.process(x))
    print('window %3d: A %.3f ms (%.1f us/img)  B %.3f  C %.3f  process %.3f ms -> %.0f img/s (%.1f us/img)' %
          (n, a, a * 1e3 / n, b, c, full, n / full * 1e3, full * 1e3 / n), flush=True)
    del eng
