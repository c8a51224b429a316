#!/usr/bin/env python
"""Kernel micro-benchmarks (CUDA events, L2-exceeding inputs).  Development tool, not bench.py.

    python tools/bench_kernels.py [--images 16] [--out gpurun_out/kernels.json]
"""

import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hiast_b200 import ops  # noqa: E402

PEAK = 6459.3e9
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), '..', 'MEASURED_PEAKS.json')))['hbm_gbs'] * 1e9
except Exception:
    pass


def timeit(fn, iters=5, warmup=2):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    evs = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2]


def time_variants(variants, rounds=6, inner=3):
    """Round-robin timing of several callables so that clock / thermal drift hits all of them alike.
    Returns {name: median ms per call}."""
    for fn in variants.values():
        fn()
    torch.cuda.synchronize()
    samples = {k: [] for k in variants}
    for _ in range(rounds):
        for name, fn in variants.items():
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(inner):
                fn()
            b.record()
            torch.cuda.synchronize()
            samples[name].append(a.elapsed_time(b) / inner)
    return {k: sorted(v)[len(v) // 2] for k, v in samples.items()}


def make_logits(n, dist, C=19, H=1024, W=2048):
    g = torch.Generator(device='cuda').manual_seed(1234)
    out = torch.empty(n, C, H, W, device='cuda')
    for i in range(n):
        if dist == 'diffuse':
            out[i] = torch.randn(C, H, W, generator=g, device='cuda') * 3
        else:
            scale = 4 if dist == 'peaked' else 60          # 'saturated': most pixels have conf == 1.0 in fp16
            low = torch.randn(1, C, 32, 64, generator=g, device='cuda') * scale
            out[i] = torch.nn.functional.interpolate(low, size=(H, W), mode='bilinear', align_corners=True)[0]
            out[i] += torch.randn(C, H, W, generator=g, device='cuda') * 0.5
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--images', type=int, default=16)
    ap.add_argument('--out', default='gpurun_out/kernels.json')
    args = ap.parse_args()
    n, C, H, W, B = args.images, 19, 1024, 2048, 2
    res = {'peak_gbs': PEAK / 1e9, 'images': n}
    key_lo = ops.ias_key_lo(C)
    G = (n + B - 1) // B
    warm = torch.empty(1 << 28, device='cuda')
    for _ in range(200):
        warm.add_(1.0)
    torch.cuda.synchronize()
    del warm
    # plain copy of the same bytes for reference
    src = torch.empty(n * C * H * W, device='cuda')
    dst = torch.empty_like(src)
    ms = timeit(lambda: dst.copy_(src))
    res['copy_gbs'] = 2 * src.numel() * 4 / ms / 1e6
    del src, dst
    for dist in ('diffuse', 'peaked', 'saturated'):
        logits = make_logits(n, dist)
        conf = torch.empty(n, H, W, device='cuda')
        label = torch.empty(n, H, W, dtype=torch.uint8, device='cuda')
        hist = ops.ias_new_hist(G, C, key_lo, 'cuda')
        alg_a = n * H * W * (4 * C + 1)
        modes = (36, 56, 80)
        variants = {m: (lambda m=m: ops.ias_softmax_hist(logits, B, key_lo, conf, label, hist, hist_mode=m)) for m in modes}
        for mode, ms in time_variants(variants).items():
            res['A_%s_mode%d' % (dist, mode)] = dict(ms=ms, img_s=n / ms * 1e3, alg_gbs=alg_a / ms / 1e6,
                                                     frac=alg_a / ms / 1e6 / (PEAK / 1e9))
        ops.ias_softmax_hist(logits, B, key_lo, conf, label, hist)
        thr_state = torch.full((C,), 0.9, dtype=torch.float64, device='cuda')
        hist_keep = hist.clone()

        def scan():
            hist.copy_(hist_keep)
            ops.ias_threshold_scan(hist, G, C, key_lo, 0.5, 0.9, 8.0, thr_state)
        ms_scan = timeit(scan)
        ms_copy = timeit(lambda: hist.copy_(hist_keep))
        res['B_%s' % dist] = dict(ms=ms_scan - ms_copy, us_per_group=(ms_scan - ms_copy) * 1e3 / G)
        thr_groups, _ = ops.ias_threshold_scan(hist, G, C, key_lo, 0.5, 0.9, 8.0, thr_state)
        plbl = torch.empty_like(label)
        counts = torch.zeros(n, C, dtype=torch.int64, device='cuda')
        confsum = torch.zeros(G, C, dtype=torch.int64, device='cuda')
        ms = timeit(lambda: ops.ias_select(conf, label, thr_groups, C, B, plbl, counts, confsum))
        res['C_%s' % dist] = dict(ms=ms, gbs=n * H * W * 6 / ms / 1e6, kept=float((plbl != 255).float().mean()),
                                  top_bin=float((conf >= 0.99976).float().mean()))
        from hiast_b200.ias_engine import IASEngine
        eng = IASEngine(C, H, W, B, 0.5, 0.9, 8.0, 0.99, n, fused=True)
        fv = {}
        for gif in (2, 4):
            for keep in (False,):
                def f(gif=gif, keep=keep):
                    eng.groups_in_flight, eng.keep_spill = gif, keep
                    eng.process_fused(logits)
                fv['gif%d%s' % (gif, '_keep' if keep else '')] = f
        for name, ms in time_variants(fv, rounds=4, inner=2).items():
            res['fused_%s_%s' % (dist, name)] = dict(ms=ms, img_s=n / ms * 1e3, frac=alg_a / ms / 1e6 / (PEAK / 1e9))
        assert eng.check_errors() is not None
        del eng
        tot = min(res['A_%s_mode%d' % (dist, m)]['ms'] for m in (36, 56, 80)) + res['B_%s' % dist]['ms'] + res['C_%s' % dist]['ms']
        res['pipeline_%s' % dist] = dict(ms=tot, img_s=n / tot * 1e3, frac=alg_a / tot / 1e6 / (PEAK / 1e9))
        # torch reference chain for context (what the reference launches on the GPU for a1 only)
        def torch_a1():
            p = torch.softmax(logits[:4], dim=1)
            return p.max(dim=1)
        ms = timeit(torch_a1)
        res['torch_a1_%s' % dist] = dict(ms=ms, img_s=4 / ms * 1e3)
        del logits
    # fused up-sampling: stride-8 logits [n,19,129,257] -> IAS phase A at 1024x2048, vs interpolate + phase A
    lr = torch.randn(n, C, 129, 257, device='cuda') * 4
    conf = torch.empty(n, H, W, device='cuda')
    label = torch.empty(n, H, W, dtype=torch.uint8, device='cuda')
    hist = ops.ias_new_hist(G, C, key_lo, 'cuda')
    full = torch.nn.functional.interpolate(lr, size=(H, W), mode='bilinear', align_corners=True).contiguous()

    def two_step():
        torch.nn.functional.interpolate(lr, size=(H, W), mode='bilinear', align_corners=True, out=None)
        ops.ias_softmax_hist(full, B, key_lo, conf, label, hist)
    ms_fused = timeit(lambda: ops.ias_upsample_softmax_hist(lr, (H, W), B, key_lo, conf, label, hist), iters=5)
    ms_interp = timeit(lambda: torch.nn.functional.interpolate(lr, size=(H, W), mode='bilinear', align_corners=True), iters=5)
    ms_a = timeit(lambda: ops.ias_softmax_hist(full, B, key_lo, conf, label, hist), iters=5)
    from hiast_b200 import _lib as _l
    _l.lib().hiast_debug_upsample_v1(1)
    ms_fused_v1 = timeit(lambda: ops.ias_upsample_softmax_hist(lr, (H, W), B, key_lo, conf, label, hist), iters=5)
    _l.lib().hiast_debug_upsample_v1(0)
    res['upsample_fused_v1'] = dict(ms=ms_fused_v1, img_s=n / ms_fused_v1 * 1e3)
    res['upsample_fused'] = dict(ms_fused=ms_fused, img_s_fused=n / ms_fused * 1e3, ms_interpolate=ms_interp, ms_phase_a=ms_a,
                                 img_s_two_step=n / (ms_interp + ms_a) * 1e3)
    del lr, full
    # loss fwd/bwd, config 3
    z = torch.randn(2, 19, 512, 1024, device='cuda') * 3
    t = torch.softmax(torch.randn(2, 19, 512, 1024, device='cuda') * 3, dim=1)
    plbl = torch.randint(0, 19, (2, 512, 1024), device='cuda')
    plbl[torch.rand(2, 512, 1024, device='cuda') < 0.5] = 255
    scales = torch.full((4,), 1e-6, device='cuda')
    grad = torch.empty_like(z)
    flush = torch.empty(64 * 1024 * 1024, device='cuda')

    def loss_step():
        flush.zero_()
        ops.st_loss_fwd(z, t, plbl, 'ignored')
        ops.st_loss_bwd(z, t, plbl, scales, 'ignored', grad=grad)
    ms_all = timeit(loss_step, iters=10)
    ms_flush = timeit(lambda: flush.zero_(), iters=10)
    px = 2 * 512 * 1024
    ms_f = timeit(lambda: ops.st_loss_fwd(z, t, plbl, 'ignored'), iters=10)
    ms_b = timeit(lambda: ops.st_loss_bwd(z, t, plbl, scales, 'ignored', grad=grad), iters=10)
    res['loss'] = dict(ms_fwd_bwd_cold=ms_all - ms_flush, ms_fwd_warm=ms_f, ms_bwd_warm=ms_b,
                       gbs_cold=px * (160 + 236) / (ms_all - ms_flush) / 1e6,
                       fwd_gbs=px * 160 / ms_f / 1e6, bwd_gbs=px * 236 / ms_b / 1e6)
    from hiast_b200 import _lib
    _lib.lib().hiast_debug_loss_scalar(1)
    ms_f = timeit(lambda: ops.st_loss_fwd(z, t, plbl, 'ignored'), iters=10)
    ms_b = timeit(lambda: ops.st_loss_bwd(z, t, plbl, scales, 'ignored', grad=grad), iters=10)
    _lib.lib().hiast_debug_loss_scalar(0)
    res['loss_scalar_kernels'] = dict(ms_fwd_warm=ms_f, ms_bwd_warm=ms_b, fwd_gbs=px * 160 / ms_f / 1e6, bwd_gbs=px * 236 / ms_b / 1e6)
    # confusion, int64 2x1024x2048
    pred = torch.randint(0, 19, (8, 1024, 2048), device='cuda')
    tgt = torch.randint(0, 19, (8, 1024, 2048), device='cuda')
    cm = torch.zeros(20, 20, dtype=torch.int64, device='cuda')
    ms = timeit(lambda: ops.confusion_matrix(pred, tgt, 19, cm=cm), iters=10)
    res['confusion_random'] = dict(ms=ms, gbs=pred.numel() * 16 / ms / 1e6)
    pred2 = (torch.arange(8 * 1024 * 2048, device='cuda') // 100000 % 19).view(8, 1024, 2048)
    ms = timeit(lambda: ops.confusion_matrix(pred2, pred2, 19, cm=cm), iters=10)
    res['confusion_coherent'] = dict(ms=ms, gbs=pred.numel() * 16 / ms / 1e6)
    # copy-paste, uint8 1024x2048x3, 16 images
    img = torch.randint(0, 256, (16, 1024, 2048, 3), dtype=torch.uint8, device='cuda')
    lbl = torch.randint(0, 19, (16, 1024, 2048), dtype=torch.uint8, device='cuda')
    mask = torch.full((16, 1024, 2048), 255, dtype=torch.uint8, device='cuda')
    dimg = torch.randint(0, 256, (16, 1024, 2048, 3), dtype=torch.uint8, device='cuda')
    dlbl = (torch.arange(16 * 1024 * 2048, device='cuda') // 5000 % 19).to(torch.uint8).view(16, 1024, 2048)
    ms = timeit(lambda: ops.copy_paste(img, lbl, mask, dimg, dlbl, list(range(14))), iters=10)
    res['copy_paste'] = dict(ms=ms, img_s=16 / ms * 1e3, gbs=16 * 1024 * 2048 * 13 / ms / 1e6)
    # EMA teacher update: a ResNet-101-sized parameter set (about 300 tensors, 44.5 M float32 parameters)
    from hiast_b200.ema import update_ema_model

    class Bag(torch.nn.Module):
        def __init__(self):
            super().__init__()
            sizes = [64 * 3 * 49] + [c * c * 9 for c in (64, 128, 256, 512) for _ in range(12)] + \
                    [c * 4 * c for c in (64, 128, 256, 512) for _ in range(24)] + [2048 * 19 * 9 * 4] + [256] * 150
            self.ps = torch.nn.ParameterList([torch.nn.Parameter(torch.randn(n, device='cuda')) for n in sizes])
    student, teacher = Bag(), Bag()
    n_par = sum(p.numel() for p in student.parameters())
    ms = timeit(lambda: update_ema_model(teacher, student, 0.999), iters=10)

    def torch_ema():       # the reference's eager loop, utils/utils.py:117-119
        for pq, pk in zip(student.parameters(), teacher.parameters()):
            pk.data = pk.data.clone() * 0.999 + pq.data.clone() * (1 - 0.999)
    ms_ref = timeit(torch_ema, iters=5)
    res['ema_update'] = dict(ms=ms, gbs=n_par * 12 / ms / 1e6, params=n_par, tensors=len(student.ps), ms_reference_loop=ms_ref)
    os.makedirs(os.path.dirname(args.out) or '.', exist_ok=True)
    json.dump(res, open(args.out, 'w'), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == '__main__':
    main()
