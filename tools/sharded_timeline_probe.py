#!/usr/bin/env python
"""Device timeline of the device-resident sharded driver (what bench.py times), per rank: when phase A, the scan (token hand-off
inside) and phase C of every window END, for the concurrent schedule (SMs reserved from phase A) and the serial one, with the
peer-memory ring and with NCCL send / recv.  Development tool; run under torchrun.

    python -m torch.distributed.run --nproc-per-node 4 --master-addr 127.0.0.1 tools/sharded_timeline_probe.py [--steps 12]
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from hiast_b200.ias_engine import IASEngine  # noqa: E402
from hiast_b200.sharded import ShardedIAS, TokenRing, device_stream  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--steps', type=int, default=12)
    ap.add_argument('--out', default='gpurun_out/sharded_timeline.json')
    ap.add_argument('--only-first', action='store_true', help='only the concurrent schedule with the peer-memory ring')
    args = ap.parse_args()
    rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    device = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=device)
    pool = bench.make_pool(device, 'mixed', bench.WINDOW)
    engine = IASEngine(bench.C, bench.H, bench.W, bench.GROUP, bench.ALPHA, bench.BETA, bench.GAMMA, bench.CP_GAMMA,
                       3 * bench.WINDOW, device=device)
    side = device_stream(device, 'chain')
    marks = []

    class Marked:
        def __getattr__(self, name):
            return getattr(engine, name)

        def _mark(self, what):
            ev = torch.cuda.Event(enable_timing=True)
            ev.record(torch.cuda.current_stream(device))
            marks.append((what, ev))

        def phase_a(self, logits, first_image=0):
            self._mark('a0')
            engine.phase_a(logits, first_image)
            self._mark('a1')

        def phase_b(self, first_image, n_images, **kw):
            self._mark('b0')
            engine.phase_b(first_image, n_images, **kw)
            self._mark('b1')

        def phase_c(self, first_image, n_images):
            engine.phase_c(first_image, n_images)
            self._mark('c1')

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def mark(what):
        ev = torch.cuda.Event(enable_timing=True)
        ev.record(torch.cuda.current_stream(device))
        marks.append((what, ev))

    class Probed(ShardedIAS):
        """finish_state with marks on the main stream: fs0 = main reaches it (the last phase A has ended), ag0 / ag1 = around the
        all-gather (ag0 comes after the chain stream has drained and the packing kernels), end = everything queued has run."""
        def finish_state(self):
            mark('fs0')
            real = dist.all_gather

            def timed_all_gather(*a, **k):
                mark('ag0')
                out = real(*a, **k)
                mark('ag1')
                return out
            if world > 1:
                dist.all_gather = timed_all_gather
            try:
                return super().finish_state()
            finally:
                dist.all_gather = real
                mark('end')

    def job(k, eng):
        drv = Probed(eng, bench.WINDOW, k * world * bench.WINDOW, rank, world)
        drv.warm_collective()
        return drv, drv.run(lambda w: pool, None)

    results = {}
    modes = (('concurrent', 12, 'peer'), ('serial', 0, 'peer'), ('concurrent', 12, 'nccl'), ('serial', 0, 'nccl'))
    if args.only_first:
        modes = modes[:1]
    for mode, reserve, ringmode in modes:
        if world == 1 and ringmode == 'nccl':
            continue
        os.environ['HIAST_RING'] = ringmode
        TokenRing._cache.clear()
        engine.reserve_sms = reserve
        job(4, engine)
        engine.thr_state.fill_(0.9)
        engine.mean_state.zero_()
        barrier()
        del marks[:]
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        drv, _ = job(args.steps, Marked())
        t1.record()
        barrier()
        ms = t0.elapsed_time(t1)
        tl = [(what, round(t0.elapsed_time(ev), 3)) for what, ev in marks]
        res = {'ms_per_step': ms / args.steps, 'ring': 'peer' if drv.ring is not None else 'nccl', 'timeline': tl[:60],
               'tail': tl[-12:], 'total_ms': ms}
        allres = [None] * world
        if world > 1:
            dist.all_gather_object(allres, res)
        else:
            allres = [res]
        results['%s/%s' % (mode, ringmode)] = allres
        if rank == 0:
            print(mode, ringmode, 'ring used:', [r['ring'] for r in allres], 'ms/step per rank:', [round(r['ms_per_step'], 3) for r in allres], flush=True)
    if rank == 0:
        os.makedirs(os.path.dirname(args.out), exist_ok=True)
        json.dump(results, open(args.out, 'w'))
        for key, allres in results.items():
            print('==', key)
            for r, res in enumerate(allres):
                print(' rank', r, ' '.join('%s@%.2f' % (w, t) for w, t in res['timeline'][:30]))
                print('   tail', ' '.join('%s@%.2f' % (w, t) for w, t in res['tail']), 'total %.2f' % res['total_ms'])
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
