"""Multi-GPU IAS over NCCL (needs >= 2 GPUs; skipped otherwise): sharded result == single-GPU result == oracle."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import golden_inputs as gi
from oracle import ias as oias

pytestmark = pytest.mark.gpu

SPEC = dict(C=19, H=32, W=64, N=19, B=2, alpha=0.5, beta=0.9, gamma=8.0, cp_gamma=0.99, seed=41, dist='mixed', absent=())


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    import sys
    sys.path.insert(0, os.path.dirname(__file__))
    sys.path.insert(0, os.path.dirname(os.path.dirname(__file__)))
    from hiast_b200.ias_engine import IASEngine
    from hiast_b200.sharded import ShardedIAS, window_images
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    try:
        s = SPEC
        logits = torch.cat([lg for lg, _ in gi.ias_batches(s)]).cuda()
        window = 2 * s['B']
        eng = IASEngine(s['C'], s['H'], s['W'], s['B'], s['alpha'], s['beta'], s['gamma'], s['cp_gamma'], 2 * window,
                        device=torch.device('cuda', rank))
        got = {}

        def on_window(w, plbl, counts, thr_groups):
            got[w] = (plbl.cpu().numpy(), thr_groups.cpu().numpy())

        def window_logits(w):
            i0, n = window_images(w, window, s['N'])
            return logits[i0:i0 + n]

        thr, mean, statics = ShardedIAS(eng, window, s['N']).run(window_logits, on_window)
        torch.cuda.synchronize()
        np.savez(os.path.join(out_dir, 'rank%d.npz' % rank), thr=thr.cpu().numpy(), mean=mean.cpu().numpy(),
                 statics=statics.cpu().numpy(), windows=np.array(sorted(got)),
                 **{'plbl_%d' % w: v[0] for w, v in got.items()}, **{'thr_%d' % w: v[1] for w, v in got.items()})
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')
def test_nccl_threshold_handoff_is_bit_identical(tmp_path):
    world = min(torch.cuda.device_count(), 4)
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    s = SPEC
    oracle = oias.IASOracle(s['C'], s['alpha'], s['beta'], s['gamma'], s['cp_gamma'])
    oracle.run([(lg.cuda(), p) for lg, p in gi.ias_batches(s)])
    ranks = [np.load(os.path.join(str(tmp_path), 'rank%d.npz' % r)) for r in range(world)]
    by_window = {}
    for r in ranks:
        assert np.array_equal(r['thr'], oracle.class_threshold)
        assert np.array_equal(r['statics'], oracle.statics_class)
        assert np.array_equal(r['mean'], ranks[0]['mean'])
        np.testing.assert_allclose(r['mean'], oracle.class_mean_probs, rtol=1e-6)
        for w in r['windows']:
            by_window[int(w)] = (r['plbl_%d' % w], r['thr_%d' % w])
    n_win = (s['N'] + 2 * s['B'] - 1) // (2 * s['B'])
    assert sorted(by_window) == list(range(n_win))
    assert np.array_equal(np.concatenate([by_window[w][0] for w in range(n_win)]), np.stack(oracle.labels))
    assert np.array_equal(np.concatenate([by_window[w][1] for w in range(n_win)]), np.stack(oracle.threshold_trace))
