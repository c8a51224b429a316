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


def _worker(rank, world, port, out_dir, slots=3, ring='peer', reserve=0):
    import sys
    sys.path.insert(0, os.path.dirname(__file__))
    sys.path.insert(0, os.path.dirname(os.path.dirname(__file__)))
    from hiast_b200.ias_engine import IASEngine
    from hiast_b200.sharded import ShardedIAS, window_images
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    os.environ['HIAST_RING'] = ring
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    try:
        s = SPEC
        logits = torch.cat([lg for lg, _ in gi.ias_batches(s)]).cuda()
        window = 2 * s['B']
        eng = IASEngine(s['C'], s['H'], s['W'], s['B'], s['alpha'], s['beta'], s['gamma'], s['cp_gamma'], slots * window,
                        device=torch.device('cuda', rank))
        eng.reserve_sms = reserve
        got = {}

        def on_window(w, plbl, counts, thr_groups):
            got[w] = (plbl.cpu().numpy(), thr_groups.cpu().numpy())

        def window_logits(w):
            i0, n = window_images(w, window, s['N'])
            return logits[i0:i0 + n]

        drv = ShardedIAS(eng, window, s['N'])
        assert (drv.ring is not None) == (ring == 'peer'), 'hand-off mode: wanted %s' % ring
        for rep in range(2):                              # twice: the ring's sequence numbers carry over from job to job
            got.clear()
            eng.thr_state.fill_(0.9)
            eng.mean_state.zero_()
            thr, mean, statics = drv.run(window_logits, on_window)
        assert eng.check_errors()
        torch.cuda.synchronize()
        np.savez(os.path.join(out_dir, 'rank%d.npz' % rank), thr=thr.cpu().numpy(), mean=mean.cpu().numpy(),
                 statics=statics.cpu().numpy(), windows=np.array(sorted(got)),
                 **{'plbl_%d' % w: v[0] for w, v in got.items()}, **{'thr_%d' % w: v[1] for w, v in got.items()})
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')
@pytest.mark.parametrize('slots,ring,reserve', [(3, 'peer', 0), (3, 'nccl', 0), (2, 'peer', 0), (3, 'peer', 12), (3, 'nccl', 12)])
def test_threshold_handoff_is_bit_identical(slots, ring, reserve, tmp_path):
    """Windows striped over 2-4 GPUs == the single-process oracle, with the state handed over inside the scan kernel through
    peer memory ('peer': hiast_ias_threshold_scan_ring over CUDA IPC mailboxes) and through NCCL send / recv ('nccl');
    reserve > 0: the concurrent schedule (scan and phase C on the SMs phase A leaves free)."""
    world = min(torch.cuda.device_count(), 4)
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), slots, ring, reserve), nprocs=world, join=True)
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


def _gen_worker(rank, world, port, out_dir):
    import json
    import sys
    from types import SimpleNamespace
    sys.path.insert(0, os.path.dirname(__file__))
    sys.path.insert(0, os.path.dirname(os.path.dirname(__file__)))
    import hiast_b200
    hiast_b200.register_all()
    from hiast_b200 import PSEUDO_POLICY
    from hiast_b200.pseudo_label_generator import striped_batch_order
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    try:
        s = SPEC
        batches = gi.ias_batches(s)
        cfg = SimpleNamespace(
            dataset=SimpleNamespace(num_classes=s['C']),
            pseudo_policy=SimpleNamespace(type='IAS_SHARDED', batch_size=s['B'],
                                          ias=SimpleNamespace(alpha=s['alpha'], beta=s['beta'], gamma=s['gamma'])),
            preprocessor=SimpleNamespace(copy_paste=SimpleNamespace(gamma=s['cp_gamma'])))
        window_batches = 2
        order = striped_batch_order(s['N'], window_batches * s['B'], s['B'], rank, world)
        loader = [{'images': batches[idx[0] // s['B']][0], 'image_paths': batches[idx[0] // s['B']][1]} for idx in order]

        class Identity:
            def __call__(self, x):
                return {'logits': x}

        gen = PSEUDO_POLICY['IAS_SHARDED'](cfg, model=Identity(), loader=loader, dataset_len=s['N'],
                                           save_dir=os.path.join(out_dir, 'run', 'pseudo_labels'),
                                           window_batches=window_batches, device=torch.device('cuda', rank))
        gen.run()
        torch.cuda.synchronize()
        np.savez(os.path.join(out_dir, 'gen_rank%d.npz' % rank), thr=gen.class_threshold, mean=gen.class_mean_probs,
                 statics=gen.statics_class, trace=np.concatenate(gen.threshold_trace),
                 stats=np.array(json.dumps(gen.sample_stats)), certified=np.array(gen.pow_rounding_certified))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')
def test_sharded_generator_over_nccl_writes_device_png_files(tmp_path):
    """PSEUDO_POLICY['IAS_SHARDED'] on 2-4 GPUs: thresholds, statistics and the PNG files (encoded on each rank's GPU,
    byte-compared with the oracle stream) equal the single-process oracle run."""
    import json
    from oracle import png as opng
    world = min(torch.cuda.device_count(), 4)
    mp.spawn(_gen_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    s = SPEC
    batches = gi.ias_batches(s)
    oracle = oias.IASOracle(s['C'], s['alpha'], s['beta'], s['gamma'], s['cp_gamma'])
    oracle.run([(lg.cuda(), p) for lg, p in batches])
    for r in range(world):
        got = np.load(os.path.join(str(tmp_path), 'gen_rank%d.npz' % r))
        assert np.array_equal(got['thr'], oracle.class_threshold)
        assert np.array_equal(got['statics'], oracle.statics_class)
        assert np.array_equal(got['trace'], np.stack(oracle.threshold_trace))
        np.testing.assert_allclose(got['mean'], oracle.class_mean_probs, rtol=1e-6)
        assert json.loads(str(got['stats'])) == json.loads(json.dumps(oracle.sample_stats))
        assert bool(got['certified'])
    save_dir = os.path.join(str(tmp_path), 'run', 'pseudo_labels')
    paths = [p for _, ps in batches for p in ps]
    assert len(os.listdir(save_dir)) == len(paths)
    for i, p in enumerate(paths):
        blob = open(os.path.join(save_dir, os.path.splitext(p)[0] + '_pseudo_label.png'), 'rb').read()
        assert blob == opng.encode_png(oracle.labels[i].astype(np.uint8))
        assert np.array_equal(opng.decode_png(blob), oracle.labels[i])
    assert np.array_equal(np.load(os.path.join(save_dir, '..', 'class_threshold.npy')), oracle.class_threshold)
