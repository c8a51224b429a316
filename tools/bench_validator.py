#!/usr/bin/env python
"""Validator post-logit chain (workflows/validator.py:34-55,92-93): fused kernels vs the reference's torch ops.

    python tools/bench_validator.py [--batch 4] [--out gpurun_out/validator.json]
"""

import argparse
import json
import os
import sys

import torch
from torch.nn import functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hiast_b200 import ops  # noqa: E402
from tools.bench_kernels import PEAK, time_variants  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=4)
    ap.add_argument('--out', default='gpurun_out/validator.json')
    args = ap.parse_args()
    B, C, H, W = args.batch, 19, 1024, 2048
    res = {}
    for name, sizes, flip in (('val_yaml_768x1536', [(768, 1536)], False), ('768x1536_flip', [(768, 1536)], True),
                              ('3_scales_flip', [(512, 1024), (768, 1536), (1024, 2048)], True)):
        g = torch.Generator(device='cuda').manual_seed(1)
        zs = [(torch.randn(B, C, h, w, generator=g, device='cuda') * 4,
               torch.randn(B, C, h, w, generator=g, device='cuda') * 4 if flip else None) for h, w in sizes]

        def torch_chain():
            out = []
            for z0, z1 in zs:
                p = F.softmax(z0, dim=1)
                if flip:
                    p += torch.flip(F.softmax(z1, dim=1), dims=[3])
                out.append(F.interpolate(p, (H, W), mode='bilinear', align_corners=True))
            return sum(out).argmax(dim=1)

        def fused():
            return ops.probs_upsample_argmax([ops.softmax_flip_sum(z0, z1) for z0, z1 in zs], (H, W))

        def k1_only():
            return [ops.softmax_flip_sum(z0, z1) for z0, z1 in zs]

        probs = k1_only()

        def k2_only():
            return ops.probs_upsample_argmax(probs, (H, W))

        assert torch.equal(fused().long(), torch_chain())
        from hiast_b200._lib import lib

        def k2_direct():
            lib().hiast_debug_validate_direct(1)
            r = ops.probs_upsample_argmax(probs, (H, W))
            lib().hiast_debug_validate_direct(0)
            return r

        ms = time_variants({'torch': torch_chain, 'fused': fused, 'k1': k1_only, 'k2': k2_only, 'k2_direct': k2_direct})
        px = sum(h * w for h, w in sizes)
        k1_bytes = B * C * px * 4 * (3 if flip else 2)
        k2_bytes = B * (C * px * 4 + H * W)
        res[name] = dict(batch=B, torch_ms=ms['torch'], fused_ms=ms['fused'], speedup=ms['torch'] / ms['fused'],
                         k1_ms=ms['k1'], k1_gbs=k1_bytes / ms['k1'] / 1e6, k1_frac_of_measured_peak=k1_bytes / ms['k1'] / 1e-3 / PEAK,
                         k2_ms=ms['k2'], k2_direct_ms=ms['k2_direct'], k2_gbs=k2_bytes / ms['k2'] / 1e6, k2_frac_of_measured_peak=k2_bytes / ms['k2'] / 1e-3 / PEAK,
                         images_per_s_fused=B / ms['fused'] * 1e3, images_per_s_torch=B / ms['torch'] * 1e3)
    os.makedirs(os.path.dirname(args.out) or '.', exist_ok=True)
    json.dump(res, open(args.out, 'w'), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == '__main__':
    main()
