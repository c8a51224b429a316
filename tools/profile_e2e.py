#!/usr/bin/env python
"""Host-side profile (cProfile) of the generator's from-stride-8 e2e loop: where the time between kernels goes.

    python tools/profile_e2e.py [--steps 8] [--png]
"""
import argparse
import cProfile
import os
import pstats
import sys
import tempfile
from types import SimpleNamespace

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hiast_b200.pseudo_label_generator import IASPseudoGenerator  # noqa: E402

C, H, W, GROUP, WINDOW = 19, 1024, 2048, 2, 64


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--steps', type=int, default=8)
    ap.add_argument('--png', action='store_true')
    args = ap.parse_args()
    h_lr, w_lr = H // 8 + 1, W // 8 + 1
    host = torch.empty((8, C, h_lr, w_lr)).pin_memory()
    host.copy_(torch.randn(8, C, h_lr, w_lr, generator=torch.Generator().manual_seed(5)) * 4)

    class LowRes:
        def __call__(self, x):
            return {'logits_lr': x, 'size': (H, W)}

    def loader(n):
        for i in range(0, n, GROUP):
            j = i % 8
            yield {'images': host[j:j + GROUP], 'image_paths': ['img_%06d.png' % (i + k) for k in range(GROUP)]}

    cfg = SimpleNamespace(dataset=SimpleNamespace(num_classes=C),
                          pseudo_policy=SimpleNamespace(type='IAS', batch_size=GROUP, ias=SimpleNamespace(alpha=0.5, beta=0.9, gamma=8.0)),
                          preprocessor=SimpleNamespace(copy_paste=SimpleNamespace(gamma=0.99)))

    class Gen(IASPseudoGenerator):
        def save_data(self):
            pass

    class GenStub(Gen):
        def save_pseudo_label(self, plbl, img_path):
            pass

    def run(n):
        cls = Gen if args.png else GenStub
        g = cls(cfg, model=LowRes(), loader=loader(n), save_dir=os.path.join(tempfile.mkdtemp(), 'pl'), window_batches=WINDOW // GROUP)
        g.run()

    import time
    run(WINDOW)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    run(args.steps * WINDOW)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print('unprofiled: %.1f images/s (%.2f ms per %d-image window)' % (args.steps * WINDOW / dt, dt / args.steps * 1e3, WINDOW))
    pr = cProfile.Profile()
    pr.enable()
    run(args.steps * WINDOW)
    torch.cuda.synchronize()
    pr.disable()
    st = pstats.Stats(pr)
    st.sort_stats('cumulative').print_stats(28)
    st.sort_stats('tottime').print_stats(18)


if __name__ == '__main__':
    main()
