#!/usr/bin/env python
"""BASELINE.json configs[4]: one full self-training round on synthetic data.

random-init DeepLabv2-ResNet101 (stock torchvision / cuDNN -- the logit producer, out of this package's scope)
  -> full-resolution logits -> IAS pseudo-labels (hiast_b200, windows striped over ranks, NCCL threshold hand-off)
  -> 19x19 confusion matrix / mIoU of the arg-max against a random ground truth (all-reduced).

    python examples/full_round.py --images 8 --height 1024 --width 2048
    torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 examples/full_round.py --images 64

Prints one JSON line: end-to-end images/s and the share of the time spent in the hot path.
"""

import argparse
import json
import os
import sys
import time
from types import SimpleNamespace

import torch
import torch.distributed as dist
from torch import nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hiast_b200.ias_engine import IASEngine          # noqa: E402
from hiast_b200.metrics import ConfusionMeter        # noqa: E402
from hiast_b200.segmentor import SelfTrainingSegmentor  # noqa: E402
from hiast_b200.sharded import ShardedIAS, window_images  # noqa: E402


class DeepLabV2(nn.Module):
    """Dilated ResNet-101 (output stride 8) + ASPP(6,12,18,24) summed -> (logits at stride 8, features);
    same topology as the reference's sseg/models/modules/seg_models/deeplab_v2.py:27-64, random init."""

    def __init__(self, num_classes=19):
        super().__init__()
        import torchvision
        r = torchvision.models.resnet101(weights=None, replace_stride_with_dilation=[False, True, True])
        self.backbone = nn.Sequential(r.conv1, r.bn1, r.relu, r.maxpool, r.layer1, r.layer2, r.layer3, r.layer4)
        self.aspp = nn.ModuleList([nn.Conv2d(2048, num_classes, 3, padding=d, dilation=d) for d in (6, 12, 18, 24)])

    def forward(self, x):
        f = self.backbone(x)
        out = self.aspp[0](f)
        for conv in self.aspp[1:]:
            out = out + conv(f)
        return out, f


def make_cfg(C):
    return SimpleNamespace(
        model=SimpleNamespace(predictor=SimpleNamespace(seg_loss=SimpleNamespace(type='CE', target_pseudo_weight=1.0),
                                                        kld_loss=SimpleNamespace(weight=0.1),
                                                        ent_loss=SimpleNamespace(weight=1.0))),
        cst_training=SimpleNamespace(is_enabled=True, cst_loss=SimpleNamespace(type='SoftCE', weight=0.5, region='ignored')))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--images', type=int, default=8, help='images in the whole job')
    ap.add_argument('--height', type=int, default=1024)
    ap.add_argument('--width', type=int, default=2048)
    ap.add_argument('--batch', type=int, default=2)
    ap.add_argument('--window', type=int, default=4, help='images per window (multiple of --batch)')
    ap.add_argument('--classes', type=int, default=19)
    ap.add_argument('--amp', action='store_true', help='bf16 autocast for the backbone')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('TORCH_NCCL_SHOW_EAGER_INIT_P2P_SERIALIZATION_WARNING', 'false')
        dist.init_process_group('nccl', device_id=dev)
    C, H, W, B = args.classes, args.height, args.width, args.batch
    torch.manual_seed(0)
    seg = SelfTrainingSegmentor(make_cfg(C), seg_model=DeepLabV2(C)).to(dev).eval()
    engine = IASEngine(C, H, W, B, 0.5, 0.9, 8.0, 0.99, 2 * args.window, device=dev)
    meter = ConfusionMeter(C, device=dev)
    win_logits = [torch.empty((args.window, C, H, W), device=dev) for _ in range(2)]
    t_hot = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    hot_ms = [0.0]
    slot = [0]

    def window_logits(w):
        i0, n = window_images(w, args.window, args.images)
        buf = win_logits[slot[0] % 2][:n]
        slot[0] += 1
        g = torch.Generator(device=dev).manual_seed(1000 + w)
        with torch.no_grad(), torch.autocast('cuda', dtype=torch.bfloat16, enabled=args.amp):
            for k in range(0, n, B):
                imgs = torch.randn(min(B, n - k), 3, H, W, generator=g, device=dev)
                buf[k:k + B] = seg(imgs)['logits'].float()
        gt = torch.randint(0, C, (n, H, W), generator=g, device=dev)
        gt[torch.rand((n, H, W), generator=g, device=dev) < 0.1] = 255
        meter.update_from_logits(buf, gt)                         # validation metric on the same logits
        return buf

    kept = torch.zeros((), dtype=torch.int64, device=dev)

    def on_window(w, plbl, counts, thr_groups):
        kept.add_(counts.sum())

    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    thr, mean, statics = ShardedIAS(engine, args.window, args.images, rank, world).run(window_logits, on_window)
    meter.all_reduce()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dt = time.perf_counter() - t0
    res = meter.result(exact=True)
    if rank == 0:
        print(json.dumps({'images': args.images, 'n_gpus': world, 'seconds': dt, 'images_per_s': args.images / dt,
                          'miou': float(res['miou']), 'kept_pixels': int(statics.sum().item()),
                          'class_threshold': [round(float(x), 6) for x in thr.tolist()][:4],
                          'pow_rounding_certified': engine.check_errors()}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
