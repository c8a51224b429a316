#!/usr/bin/env python
"""Per-rank host trace and device timeline of the sharded from-stride-8 e2e call (development).

    torchrun --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 tools/e2e_sharded_probe.py [--windows 46] [--png]

Every rank prints its pipeline trace; with HIAST_PIPE_EVENTS=<dir> the generator also dumps its device timeline there.
"""
import argparse
import json
import os
import shutil
import sys
import tempfile
import time
from types import SimpleNamespace

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hiast_b200.pseudo_label_generator import IASPseudoGenerator, ShardedIASPseudoGenerator  # noqa: E402

C, H, W, GROUP, WINDOW = 19, 1024, 2048, 2, 64


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--windows', type=int, default=46)
    ap.add_argument('--png', action='store_true')
    ap.add_argument('--profile', action='store_true', help='cProfile of rank 0')
    ap.add_argument('--lazy-init', action='store_true', help='init_process_group without device_id (per-pair P2P communicators)')
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (('RANK', '0'), ('WORLD_SIZE', '1'), ('LOCAL_RANK', '0')))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', **({} if args.lazy_init else {'device_id': dev}))
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
    Base = ShardedIASPseudoGenerator if world > 1 else IASPseudoGenerator

    class Gen(Base):
        def save_data(self):
            pass

    class GenStub(Gen):
        def save_pseudo_label(self, plbl, img_path):
            pass

    def run(n):
        root = '/dev/shm' if os.path.isdir('/dev/shm') else None
        d = tempfile.mkdtemp(dir=root)
        try:
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            g = (Gen if args.png else GenStub)(cfg, model=LowRes(), loader=loader(n), dataset_len=n * world if world > 1 else None,
                                               save_dir=os.path.join(d, 'pl'), window_batches=WINDOW // GROUP, device=dev)
            t1 = time.perf_counter()
            g.run()
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            return t1 - t0, t2 - t1, g.pipeline_trace
        finally:
            shutil.rmtree(d, ignore_errors=True)

    ev_dir = os.environ.pop('HIAST_PIPE_EVENTS', None)
    run(WINDOW)
    if ev_dir:
        os.environ['HIAST_PIPE_EVENTS'] = ev_dir
    if args.profile and rank == 0:
        import cProfile
        import pstats
        pr = cProfile.Profile()
        pr.enable()
        t_init, t_run, trace = run(args.windows * WINDOW)
        pr.disable()
        pstats.Stats(pr).sort_stats('tottime').print_stats(16)
    else:
        t_init, t_run, trace = run(args.windows * WINDOW)
    print(json.dumps({'rank': rank, 'init_s': round(t_init, 4), 'run_s': round(t_run, 4), 'images_per_s_this_rank': round(args.windows * WINDOW / (t_init + t_run)),
                      'trace': {k: (round(v, 4) if isinstance(v, float) else v) for k, v in trace.items()}}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
