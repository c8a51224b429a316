#!/usr/bin/env python
"""Timing of the loss kernels on BASELINE configs[2] (2x19x512x1024, int64 labels, 50 % ignored).  Development tool.

Two protocols: `flushed` = median of single calls with a 256 MB memset between them (what bench.py's extra leg reports) and
`rotating` = 30 calls back to back over three input sets (3 x 247 MB, more than L2) inside one pair of CUDA events."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hiast_b200 import ops  # noqa: E402


def main():
    dev = torch.device('cuda')
    B, C, H, W = 2, 19, 512, 1024
    sets = []
    for k in range(3):
        g = torch.Generator(device=dev).manual_seed(k)
        z = torch.randn(B, C, H, W, generator=g, device=dev) * 3
        t = torch.softmax(torch.randn(B, C, H, W, generator=g, device=dev) * 3, dim=1)
        y = torch.randint(0, C, (B, H, W), generator=g, device=dev)
        y[torch.rand(B, H, W, generator=g, device=dev) < 0.5] = 255
        sets.append((z, t, y, torch.empty_like(z)))
    gw = torch.tensor([1.0, 0.1, 1.0, 0.5], device=dev)
    scales = torch.full((4,), 1e-6, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    calls = {
        'one_pass': lambda s: ops.st_loss_fused(s[0], s[1], s[2], gw, 'ignored', grad=s[3]),
        'fwd': lambda s: ops.st_loss_fwd(s[0], s[1], s[2], 'ignored'),
        'bwd': lambda s: ops.st_loss_bwd(s[0], s[1], s[2], scales, 'ignored', grad=s[3]),
    }
    out = {}
    for name, fn in calls.items():
        for _ in range(5):
            fn(sets[0])
        ts = []
        for i in range(15):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn(sets[i % 3])
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        n = 30
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            fn(sets[i % 3])
        e1.record()
        torch.cuda.synchronize()
        out[name] = {'flushed_median_us': 1e3 * ts[len(ts) // 2], 'flushed_min_us': 1e3 * ts[0],
                     'rotating_us': 1e3 * e0.elapsed_time(e1) / n}
    px = B * H * W
    for name in out:
        out[name]['gbs_on_236B_per_px_rotating'] = px * 236 / (out[name]['rotating_us'] * 1e-6) / 1e9
    print(json.dumps(out, indent=1))
    if len(sys.argv) > 1:
        json.dump(out, open(sys.argv[1], 'w'), indent=1)


if __name__ == '__main__':
    main()
