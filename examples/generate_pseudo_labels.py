#!/usr/bin/env python
"""The reference's pseudo-label stage (code/generate_pseudo_labels.py -> PSEUDO_POLICY[...](cfg).run()) through this
package's reference-facing API, on synthetic images, with every widened piece in the loop:

  random-init DeepLabv2-ResNet101 (stock torchvision) returning its stride-8 logits
    -> PSEUDO_POLICY['IAS'] (one GPU) or ['IAS_SHARDED'] (torchrun, one rank per GPU): fused up-sampling + IAS on the device,
       PNG files encoded on the device and written by hiast_write_files, npy / json statistics as the reference writes them
    -> the reader side: stat_samples_with_class + load_pseudo_labels (PIL decode, device nearest resize)
    -> Validator (fused multi-scale / flip prediction + confusion matrix) against random ground truth.

    python examples/generate_pseudo_labels.py --images 8 --height 512 --width 1024
    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 examples/generate_pseudo_labels.py --images 16

Prints one JSON line.
"""

import argparse
import json
import os
import shutil
import sys
import tempfile
import time
from types import SimpleNamespace

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import hiast_b200                                        # noqa: E402
from full_round import DeepLabV2, make_cfg              # noqa: E402
from hiast_b200 import PSEUDO_POLICY, pseudo_store      # noqa: E402
from hiast_b200.pseudo_label_generator import striped_batch_order  # noqa: E402
from hiast_b200.segmentor import SelfTrainingSegmentor  # noqa: E402
from hiast_b200.validator import Validator              # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--images', type=int, default=8)
    ap.add_argument('--height', type=int, default=512)
    ap.add_argument('--width', type=int, default=1024)
    ap.add_argument('--batch', type=int, default=2)
    ap.add_argument('--window_batches', type=int, default=2)
    ap.add_argument('--classes', type=int, default=19)
    ap.add_argument('--save_dir', default=None)
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (('RANK', '0'), ('WORLD_SIZE', '1'), ('LOCAL_RANK', '0')))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)
    hiast_b200.register_all()
    C, H, W, B, N = args.classes, args.height, args.width, args.batch, args.images
    root = args.save_dir or os.path.join(tempfile.gettempdir(), 'hiast_b200_pseudo_%d' % os.getppid())
    save_dir = os.path.join(root, 'pseudo_labels')
    if rank == 0:
        shutil.rmtree(root, ignore_errors=True)
    if world > 1:
        dist.barrier()
    torch.manual_seed(0)
    seg = SelfTrainingSegmentor(make_cfg(C), seg_model=DeepLabV2(C)).to(dev).eval()
    seg.fused_upsample = True                          # forward() hands over the stride-8 logits ('logits_lr')

    def image_batch(idx):                              # the "dataset": seeded synthetic images, pinned host memory
        g = torch.Generator().manual_seed(1000 + idx[0])
        return {'images': torch.randn(len(idx), 3, H, W, generator=g).pin_memory(),
                'image_paths': ['/data/target/img_%05d.png' % i for i in idx]}

    order = striped_batch_order(N, args.window_batches * B, B, rank, world)
    cfg = SimpleNamespace(dataset=SimpleNamespace(num_classes=C, source=SimpleNamespace(type='GTA5')),
                          pseudo_policy=SimpleNamespace(type='IAS', batch_size=B, ias=SimpleNamespace(alpha=0.5, beta=0.9, gamma=8.0)),
                          preprocessor=SimpleNamespace(copy_paste=SimpleNamespace(gamma=0.99)),
                          validate=SimpleNamespace(resize_sizes=[[H * 3 // 4, W * 3 // 4]], is_flip=True, batch_size=B,
                                                   color_mask_dir_path=None))
    gen = PSEUDO_POLICY['IAS_SHARDED' if world > 1 else 'IAS'](
        cfg, model=seg, loader=(image_batch(idx) for idx in order), dataset_len=N, save_dir=save_dir,
        window_batches=args.window_batches, device=dev)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    gen.run()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dt = time.perf_counter() - t0

    out = {}
    if rank == 0:                                      # reader side + validator on one rank
        files = sorted(os.listdir(save_dir))
        assert len(files) == N, (len(files), N)
        samples = pseudo_store.stat_samples_with_class(root, C)
        paths = ['/data/target/img_%05d.png' % i for i in range(min(N, 4))]
        back = pseudo_store.load_pseudo_labels(save_dir, paths, (H, W), device=dev)
        bigger = pseudo_store.load_pseudo_labels(save_dir, paths, (H * 4 // 3, W * 4 // 3), device=dev)
        kept = float((back != 255).float().mean())

        def val_batches():
            g = torch.Generator().manual_seed(7)
            for k in range(2):
                lbl = torch.randint(0, C, (B, H, W), generator=g)
                lbl[torch.rand(B, H, W, generator=g) < 0.1] = 255
                yield {'images': torch.randn(B, 3, H, W, generator=g), 'labels': lbl,
                       'image_paths': ['/data/val/v_%d_%d.png' % (k, j) for j in range(B)]}

        class FullRes(torch.nn.Module):                # the validator wants logits at the size of its input
            def forward(self, x):
                with torch.no_grad():
                    seg.fused_upsample = False
                    return seg(x)

        res = Validator(cfg, model=FullRes(), loader=val_batches(), device=dev).run()
        out = {'images': N, 'n_gpus': world, 'seconds': round(dt, 3), 'images_per_s': round(N / dt, 2), 'png_files': len(files),
               'mean_file_bytes': int(np.mean([os.path.getsize(os.path.join(save_dir, f)) for f in files])),
               'kept_fraction_first_images': round(kept, 4), 'resized_shape': list(bigger.shape),
               'classes_with_donor_images': sum(1 for v in samples.values() if v),
               'class_threshold_head': [round(float(x), 6) for x in gen.class_threshold[:4]],
               'pow_rounding_certified': bool(gen.pow_rounding_certified), 'val_miou': float(res['miou'])}
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
