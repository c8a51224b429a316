#!/usr/bin/env python
"""Device PNG writer micro-benchmark: images/s and file sizes on pseudo-label maps of BASELINE size, with the host
encoder of the reference (cv2.imencode, what cv2.imwrite runs) timed beside it.

    python tools/bench_png.py [--images 64] [--out gpurun_out/png.json]
"""

import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hiast_b200.ops import PngEncoder  # noqa: E402
from tools.bench_kernels import make_logits, time_variants  # noqa: E402


def pseudo_labels(n, dist, keep):
    out = torch.empty(n, 1024, 2048, dtype=torch.uint8, device='cuda')
    for i0 in range(0, n, 8):
        logits = make_logits(min(8, n - i0), dist)
        conf, lbl = torch.softmax(logits, 1).max(1)
        lbl = lbl.to(torch.uint8)
        thr = torch.quantile(conf.flatten()[::97].float(), 1.0 - keep)
        lbl[conf < thr] = 255
        out[i0:i0 + lbl.shape[0]] = lbl
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--images', type=int, default=64)
    ap.add_argument('--out', default='gpurun_out/png.json')
    ap.add_argument('--only', default=None, help='run one data set only (e.g. saturated_keep90)')
    args = ap.parse_args()
    n = args.images
    import cv2
    res = {}
    enc = PngEncoder(1024, 2048, max_images=n, expect_ratio=1.0)
    for name, dist, keep in (('peaked_keep70', 'peaked', 0.7), ('saturated_keep90', 'saturated', 0.9), ('diffuse_keep30', 'diffuse', 0.3)):
        if args.only and name != args.only:
            continue
        lbl = pseudo_labels(n, dist, keep)
        files = enc.encode_to_host(lbl)
        sizes = [len(f) for f in files]
        host = lbl[:4].cpu().numpy()
        for i in range(4):
            a = cv2.imdecode(np.frombuffer(bytes(files[i]), np.uint8), cv2.IMREAD_UNCHANGED)
            assert np.array_equal(a, host[i]), 'decode mismatch'
        from hiast_b200._lib import lib

        def launch_v(v):
            lib().hiast_debug_png_variant(v)
            enc._launch(lbl)
            lib().hiast_debug_png_variant(1)

        ms = time_variants({'launch': lambda: enc._launch(lbl), 'to_host': lambda: enc.encode_to_host(lbl),
                            'launch_v0': lambda: launch_v(0), 'launch_v1': lambda: launch_v(1)})
        t0 = time.perf_counter()
        cv_sizes = [len(cv2.imencode('.png', host[i])[1]) for i in range(4)]
        cv_ms = (time.perf_counter() - t0) / 4 * 1e3
        res[name] = dict(images=n, device_ms=ms['launch'], device_images_per_s=n / ms['launch'] * 1e3,
                         to_host_ms=ms['to_host'], device_ms_v0=ms['launch_v0'], device_ms_v1=ms['launch_v1'], to_host_images_per_s=n / ms['to_host'] * 1e3,
                         mean_file_bytes=float(np.mean(sizes)), ratio=1024 * 2048 / float(np.mean(sizes)),
                         label_read_gbs=n * 1024 * 2048 * 3 / ms['launch'] / 1e6,
                         cv2_imencode_ms_per_image=cv_ms, cv2_mean_file_bytes=float(np.mean(cv_sizes)))
    os.makedirs(os.path.dirname(args.out) or '.', exist_ok=True)
    json.dump(res, open(args.out, 'w'), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == '__main__':
    main()
