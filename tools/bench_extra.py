"""Extra legs of bench.py: the other BASELINE.json configs, short, after the headline.

Each function returns a plain dict that bench.py puts under ``extra`` in its JSON line.

* ``loss_leg``        configs[2]  region-adaptive regularisation + consistency loss fwd/bwd, 2x19x512x1024, with its own roofline
* ``copy_paste_leg``  the masked-gather kernel (13 B/px) at 1024x2048x3
* ``confusion_leg``   the privatised bincount (16 B/px, int64 as the reference calls it)
* ``distributions_leg`` configs[1] on each of the two logit distributions of SURVEY 8d alone (the headline pool alternates them)
* ``synthia_leg``     configs[3]  SYNTHIA 16-class IAS + run_batch copy-paste (batch-level donor sampler) over the ranks,
                      with a sharded-vs-single-rank parity check
* ``full_round_leg``  configs[4]  random-init DeepLabv2-ResNet101 forward -> IAS pseudo-labels -> confusion matrix / mIoU

Timing: CUDA events on the launching stream, an L2 flush (a 256 MB memset) between timed iterations where the working set is
smaller than a few L2s, medians over the iterations.
"""

from __future__ import annotations

import os
import sys
import time
from types import SimpleNamespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _median_ms(fn, iters=10, warm=3, flush=None):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def _roof(bytes_, ms, peak):
    gbs = bytes_ / (ms / 1e3) / 1e9
    return {'bound': 'hbm', 'achieved': gbs, 'peak': peak, 'unit': 'GB/s', 'frac': gbs / peak, 'launch_ms': ms,
            'algorithmic_bytes_per_launch': bytes_}


def loss_leg(device, peak):
    """BASELINE configs[2]: SURVEY 8d config 3 inputs; forward + backward of the four-term loss through the kernels the
    segmentor calls (hiast_st_loss_fwd / _bwd, or the one-pass kernel when available)."""
    from hiast_b200 import ops
    B, C, H, W = 2, 19, 512, 1024
    g = torch.Generator(device=device)
    z = torch.randn(B, C, H, W, generator=g.manual_seed(0), device=device) * 3
    t = torch.softmax(torch.randn(B, C, H, W, generator=g.manual_seed(1), device=device) * 3, dim=1)
    plbl = torch.randint(0, C, (B, H, W), generator=g.manual_seed(2), device=device)
    plbl[torch.rand(B, H, W, generator=g, device=device) < 0.5] = 255
    scales = torch.full((4,), 1e-6, device=device)
    grad = torch.empty_like(z)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)
    px = B * H * W
    out = {'workload': 'configs[2]: CE + KLD + ENT + SoftCE(ignored) fwd/bwd, 2x19x512x1024, int64 pseudo-labels, 50 % ignored',
           'l2': 'flushed between iterations (256 MB memset)'}
    ms_f = _median_ms(lambda: ops.st_loss_fwd(z, t, plbl, 'ignored'), flush=flush)
    ms_b = _median_ms(lambda: ops.st_loss_bwd(z, t, plbl, scales, 'ignored', grad=grad), flush=flush)
    out['two_pass'] = {'fwd_ms': ms_f, 'bwd_ms': ms_b, 'fwd_bwd_ms': ms_f + ms_b,
                       'roofline': _roof(px * 236, ms_f + ms_b, peak), 'bytes_moved_per_px': 160 + 236}
    if hasattr(ops, 'st_loss_fused'):
        w = torch.tensor([1.0, 0.1, 1.0, 0.5], device=device)
        ms = _median_ms(lambda: ops.st_loss_fused(z, t, plbl, w, 'ignored', grad=grad), flush=flush)
        out['one_pass'] = {'fwd_bwd_ms': ms, 'roofline': _roof(px * 236, ms, peak), 'bytes_moved_per_px': 236 + 1}
    best = min(v['fwd_bwd_ms'] for k, v in out.items() if isinstance(v, dict) and 'fwd_bwd_ms' in v)
    out['steps_per_s'] = 1e3 / best
    # through the reference-facing API: SelfTrainingSegmentor.compute_loss + backward of the summed dict (base_trainer.py:129-133)
    from hiast_b200.segmentor import SelfTrainingSegmentor
    cfg = SimpleNamespace(
        model=SimpleNamespace(predictor=SimpleNamespace(seg_loss=SimpleNamespace(type='CE', target_pseudo_weight=1.0),
                                                        kld_loss=SimpleNamespace(weight=0.1), ent_loss=SimpleNamespace(weight=1.0))),
        cst_training=SimpleNamespace(is_enabled=True, cst_loss=SimpleNamespace(type='SoftCE', weight=0.5, region='ignored')))
    seg = SelfTrainingSegmentor(cfg)
    zz = z.clone().requires_grad_(True)

    def api():
        zz.grad = None
        losses = seg.compute_loss(zz, plbl, t)
        sum(losses.values()).backward()
    out['compute_loss_backward_ms'] = _median_ms(api, flush=flush)
    return out


def copy_paste_leg(device, peak):
    from hiast_b200 import ops
    n, H, W = 16, 1024, 2048
    g = torch.Generator(device=device).manual_seed(3)
    img = torch.randint(0, 256, (n, H, W, 3), dtype=torch.uint8, generator=g, device=device)
    lbl = torch.randint(0, 19, (n, H, W), dtype=torch.uint8, generator=g, device=device)
    mask = torch.full((n, H, W), 255, dtype=torch.uint8, device=device)
    dimg = torch.randint(0, 256, (n, H, W, 3), dtype=torch.uint8, generator=g, device=device)
    dlbl = (torch.arange(n * H * W, device=device) // 5000 % 19).to(torch.uint8).view(n, H, W)
    ms = _median_ms(lambda: ops.copy_paste(img, lbl, mask, dimg, dlbl, list(range(14))))
    return {'workload': 'hard-aware copy-paste, 16 images 1024x2048x3 uint8, 14 hard classes, coherent donor labels',
            'images_per_s': n / ms * 1e3, 'roofline': _roof(n * H * W * 13, ms, peak)}


def confusion_leg(device, peak):
    from hiast_b200 import ops
    n, H, W = 8, 1024, 2048
    g = torch.Generator(device=device).manual_seed(4)
    pred = torch.randint(0, 19, (n, H, W), generator=g, device=device)
    tgt = torch.randint(0, 19, (n, H, W), generator=g, device=device)
    cm = torch.zeros(20, 20, dtype=torch.int64, device=device)
    ms = _median_ms(lambda: ops.confusion_matrix(pred, tgt, 19, cm=cm))
    return {'workload': '19x19 confusion matrix, 8 maps 1024x2048, int64 prediction and target (as metrics.py:6-19 is called), random labels',
            'images_per_s': n / ms * 1e3, 'roofline': _roof(n * H * W * 16, ms, peak)}


def distributions_leg(device, peak, make_pool, reserve_sms, steps=10, window=64):
    """SURVEY 8d config 2 asks for BOTH logit distributions: D1 "diffuse" (randn * 3: uniform class mix, confidences spread over
    many fp16 keys) and D2 "peaked" (a low-resolution field up-sampled x32 + noise: coherent classes, most confidences near 1.0,
    the worst case for the histogram tables).  The headline pool alternates them; here each one alone, same pipeline, one GPU."""
    from hiast_b200.ias_engine import IASEngine
    from hiast_b200.sharded import ShardedIAS
    C, H, W, B = 19, 1024, 2048, 2
    engine = IASEngine(C, H, W, B, 0.5, 0.9, 8.0, 0.99, 3 * window, device=device)
    engine.reserve_sms = int(reserve_sms)
    out = {'workload': 'configs[1] per distribution: %d windows of %d maps, one GPU, device-resident' % (steps, window)}
    for name in ('diffuse', 'peaked'):
        pool = make_pool(device, name, window)

        def job(k):
            engine.thr_state.fill_(0.9)
            engine.mean_state.zero_()
            return ShardedIAS(engine, window, k * window, 0, 1).run(lambda w: pool)
        job(3)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _, _, statics = job(steps)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        ips = window / ms * 1e3
        out[name] = {'images_per_s': ips, 'ms_per_step': ms, 'hbm_frac_of_peak': ips * H * W * (4 * C + 1) / 1e9 / peak,
                     'kept_pixel_fraction': float(statics.sum().item()) / (steps * window * H * W)}
        del pool
    engine.check_errors()
    return out


class _PoolDonors:
    """The three methods CopyPaste needs from its donor dataset (preprocessor.py:26,96-97), over uint8 host arrays."""

    def __init__(self, imgs, lbls, C):
        self.imgs, self.lbls = imgs, lbls
        self.names = ['donor_%03d.png' % i for i in range(len(imgs))]
        self.samples = {c: [self.names[i] for i in range(len(imgs)) if (lbls[i] == c).any()] or [self.names[0]] for c in range(C)}

    def get_samples_with_class(self):
        return self.samples

    def get_file_to_idx(self, name):
        return int(name[6:9])

    def load_data(self, idx):
        return self.imgs[idx], self.lbls[idx], self.names[idx]


def synthia_leg(device, rank, world, barrier, steps=6, window=64):
    """BASELINE configs[3]: 16-class SYNTHIA (classes 9, 14, 16 never predicted: their logit planes are -1e4) IAS over the
    ranks with the NCCL threshold hand-off, followed per window by the hard-aware copy-paste of the window's images with
    the pseudo-labels just produced (``CopyPaste.run_batch`` with the batch-level donor sampler: donors drawn from the global
    np.random stream, loaded from a host donor set, uploaded once each).  Parity: thresholds and label hashes of the sharded
    job equal the same job replayed on one rank."""
    import hashlib
    import torch.distributed as dist
    from hiast_b200.ias_engine import IASEngine
    from hiast_b200.preprocessor import CopyPaste
    from hiast_b200.sharded import ShardedIAS
    C, H, W, B = 19, 1024, 2048, 2
    g = torch.Generator(device=device).manual_seed(77)
    pool = torch.randn(window, C, H, W, generator=g, device=device) * 3
    low = torch.randn(window // 2, C, 32, 64, generator=g, device=device) * 4
    pool[1::2] = torch.nn.functional.interpolate(low, size=(H, W), mode='bilinear', align_corners=True) + pool[1::2] / 6
    pool[:, [9, 14, 16]] = -1e4
    engine = IASEngine(C, H, W, B, 0.5, 0.9, 8.0, 0.99, 3 * window, device=device)
    imgs = torch.randint(0, 256, (window, H, W, 3), dtype=torch.uint8, generator=g, device=device)
    # donor set on the host: 16 images with coherent labels over the 16 valid classes
    rs = np.random.RandomState(5)
    valid = [c for c in range(C) if c not in (9, 14, 16)]
    d_imgs = [rs.randint(0, 256, size=(H, W, 3)).astype(np.uint8) for _ in range(16)]
    d_lbls = [np.kron(rs.choice(valid, size=(H // 64, W // 64)), np.ones((64, 64), dtype=np.int64)).astype(np.uint8) for _ in range(16)]
    cfg = SimpleNamespace(dataset=SimpleNamespace(source=SimpleNamespace(type='SYNTHIA'), num_classes=C),
                          preprocessor=SimpleNamespace(copy_paste=SimpleNamespace(selected_num_classes=14, mode='original')))
    class_value = rs.uniform(0.6, 0.99, size=C)
    cp = CopyPaste(cfg, _PoolDonors(d_imgs, d_lbls, C), class_value, device=device)
    work_img = imgs.clone()
    pasted = torch.zeros((), dtype=torch.int64, device=device)

    def job(drv, with_paste, capture=None):
        engine.thr_state.fill_(0.9)
        engine.mean_state.zero_()
        np.random.seed(1234 + rank)

        def on_window(w, plbl, counts, thr_groups):
            if capture is not None:
                capture[w] = (thr_groups.clone(), plbl.clone())
            if with_paste:
                lbl = plbl.clone()
                _, _, masks = cp.run_batch(work_img[:plbl.shape[0]], lbl)
                pasted.add_((masks != 255).sum())
        return drv.run(lambda w: pool, on_window)

    job(ShardedIAS(engine, window, 2 * world * window, rank, world), True)          # warm-up (donor cache, NCCL pairs)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    job(ShardedIAS(engine, window, steps * world * window, rank, world), True)
    e1.record()
    torch.cuda.synchronize()
    secs = max(time.perf_counter() - t0, e0.elapsed_time(e1) / 1e3)
    t = torch.tensor([secs], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    secs = float(t[0])
    out = {'workload': 'configs[3]: SYNTHIA 16-class IAS (planes 9/14/16 = -1e4) + CopyPaste.run_batch with the batch-level donor '
                       'sampler per 64-image window, %d windows per rank' % steps,
           'n_gpus': world, 'images_per_s': steps * world * window / secs, 'ms_per_window': secs / steps * 1e3,
           'pasted_pixels_per_image': int(pasted.item()) / max(1, (steps + 2) * window)}
    # parity: sharded vs the same job on one rank (IAS part; the paste is per image and tested against the reference fixture)
    cap = {}
    thr, mean, statics = job(ShardedIAS(engine, window, 2 * world * window, rank, world), False, cap)
    torch.cuda.synchronize()
    mine = ({w: (tg.cpu().numpy().tobytes(), hashlib.sha256(pl.cpu().numpy().tobytes()).hexdigest()) for w, (tg, pl) in cap.items()},
            thr.cpu().numpy().tobytes())
    parts = [mine]
    if world > 1:
        parts = [None] * world
        dist.all_gather_object(parts, mine)
    if rank == 0:
        ref = {}
        ref_thr = job(ShardedIAS(engine, window, 2 * world * window, 0, 1), False, ref)[0]
        torch.cuda.synchronize()
        merged = {}
        for p in parts:
            merged.update(p[0])
        out['parity'] = {'windows': 2 * world,
                         'thr_equal': all(merged[w][0] == ref[w][0].cpu().numpy().tobytes() for w in ref)
                         and all(p[1] == ref_thr.cpu().numpy().tobytes() for p in parts),
                         'plbl_sha_equal': all(merged[w][1] == hashlib.sha256(ref[w][1].cpu().numpy().tobytes()).hexdigest() for w in ref),
                         'never_predicted_classes_absent': bool(statics[[9, 14, 16]].sum().item() == 0)}
    barrier()
    return out


def full_round_leg(device, rank, world, barrier, images_per_rank=8, window=4):
    """BASELINE configs[4]: random-init DeepLabv2-ResNet101 (stock torchvision / cuDNN, bf16 autocast as the stand-in for the
    reference's apex O1) forward on synthetic 1024x2048 images -> full-resolution logits -> IAS pseudo-labels (windows striped
    over the ranks) -> 19x19 confusion matrix / mIoU against random ground truth.  Reports end-to-end images/s and the share
    of the device time spent in this library's kernels (expected << the backbone)."""
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(ROOT, 'examples'))
    from full_round import DeepLabV2, make_cfg
    from hiast_b200.ias_engine import IASEngine
    from hiast_b200.metrics import ConfusionMeter
    from hiast_b200.segmentor import SelfTrainingSegmentor
    from hiast_b200.sharded import ShardedIAS, window_images
    C, H, W, B = 19, 1024, 2048, 2
    torch.manual_seed(0)
    seg = SelfTrainingSegmentor(make_cfg(C), seg_model=DeepLabV2(C)).to(device).eval()
    engine = IASEngine(C, H, W, B, 0.5, 0.9, 8.0, 0.99, 3 * window, device=device)
    meter = ConfusionMeter(C, device=device)
    bufs = [torch.empty((window, C, H, W), device=device) for _ in range(2)]
    n_total = images_per_rank * world
    hot = []
    turn = [0]

    def window_logits(w):
        i0, n = window_images(w, window, n_total)
        buf = bufs[turn[0] % 2][:n]
        turn[0] += 1
        g = torch.Generator(device=device).manual_seed(1000 + w)
        with torch.no_grad(), torch.autocast('cuda', dtype=torch.bfloat16):
            for k in range(0, n, B):
                imgs = torch.randn(min(B, n - k), 3, H, W, generator=g, device=device)
                buf[k:k + B] = seg(imgs)['logits'].float()
        gt = torch.randint(0, C, (n, H, W), generator=g, device=device)
        gt[torch.rand((n, H, W), generator=g, device=device) < 0.1] = 255
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        meter.update_from_logits(buf, gt)
        e1.record()
        hot.append((e0, e1))
        return buf

    class Timed:
        def __getattr__(self, name):
            return getattr(engine, name)

        def phase_a(self, logits, first_image=0):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            engine.phase_a(logits, first_image)
            e1.record()
            hot.append((e0, e1))

        def phase_c(self, first_image, n_images):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            engine.phase_c(first_image, n_images)
            e1.record()
            hot.append((e0, e1))

    def run():
        engine.thr_state.fill_(0.9)
        engine.mean_state.zero_()
        meter.cm.zero_()
        del hot[:]
        turn[0] = 0
        out = ShardedIAS(Timed(), window, n_total, rank, world).run(window_logits)
        meter.all_reduce()
        return out

    run()                                          # warm-up: cuDNN autotune, NCCL pairs
    barrier()
    t0 = time.perf_counter()
    thr, mean, statics = run()
    torch.cuda.synchronize()
    barrier()
    secs = time.perf_counter() - t0
    hot_ms = sum(a.elapsed_time(b) for a, b in hot)
    t = torch.tensor([secs, hot_ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    res = meter.result(exact=True)
    return {'workload': 'configs[4]: DeepLabv2-ResNet101 (random init, bf16 autocast) on synthetic 1024x2048 images -> IAS -> 19x19 '
                        'confusion / mIoU, %d images per rank' % images_per_rank,
            'n_gpus': world, 'images_per_s': n_total / float(t[0]), 'seconds': float(t[0]),
            'hot_path_share_of_time': float(t[1]) / 1e3 / float(t[0]),
            'hot_path_ms_per_image': float(t[1]) / images_per_rank,
            'miou_vs_random_gt': float(res['miou']), 'kept_pixels': int(statics.sum().item()),
            'note': 'hot path = phase A + phase C + confusion-from-logits kernels (the threshold chain runs beside them on its own stream)'}
