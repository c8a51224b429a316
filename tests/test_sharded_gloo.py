"""Host-side logic of the multi-GPU IAS path with world_size 2 and 3 over gloo on CPU: group partitioning,
the rank-to-rank threshold hand-off order, the all-gather replay of the mean-prob EMA.  The kernels are
replaced by a CPU stand-in engine (tests/host_engine.py); sharded == unsharded == oracle, bit for bit."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import golden_inputs as gi
from hiast_b200.sharded import ShardedIAS, local_windows, window_images
from oracle import ias as oias

SPEC = dict(C=7, H=12, W=20, N=13, B=2, alpha=0.5, beta=0.9, gamma=8.0, cp_gamma=0.99, seed=31, dist='mixed', absent=())


def test_window_striping():
    assert local_windows(5, 0, 2) == [0, 2, 4] and local_windows(5, 1, 2) == [1, 3]
    assert local_windows(2, 3, 4) == []
    assert window_images(2, 4, 9) == (8, 1)            # last window holds one image
    assert window_images(1, 4, 9) == (4, 4)
    assert window_images(3, 4, 9) == (12, 0)


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir, slots=2):
    import sys
    sys.path.insert(0, os.path.dirname(__file__))
    from host_engine import HostEngine
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        s = SPEC
        logits = torch.cat([lg for lg, _ in gi.ias_batches(s)])
        window = 2 * s['B']                               # 2 groups per window -> 4 windows for 13 images
        eng = HostEngine(s['C'], s['H'], s['W'], s['B'], s['alpha'], s['beta'], s['gamma'], s['cp_gamma'], slots * window)
        drv = ShardedIAS(eng, window, s['N'])
        drv.warm_collective()                             # collective with the job's shapes; must not disturb the job
        got = {}

        def on_window(w, plbl, counts, thr_groups):
            got[w] = (np.array(plbl), np.array(thr_groups))

        def window_logits(w):
            i0, n = window_images(w, window, s['N'])
            return logits[i0:i0 + n]

        thr, mean, statics = drv.run(window_logits, on_window)
        np.savez(os.path.join(out_dir, 'rank%d.npz' % rank), thr=thr.numpy(), mean=mean.numpy(),
                 statics=statics.numpy(), windows=np.array(sorted(got)),
                 **{'plbl_%d' % w: v[0] for w, v in got.items()}, **{'thr_%d' % w: v[1] for w, v in got.items()})
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world,slots', [(2, 3), (3, 3), (2, 2)])
def test_sharded_equals_unsharded_equals_oracle(world, slots, tmp_path):
    """slots = 3: chain of window j-1 behind phase A of window j, outputs of window j-2 (the pipelined schedule);
    slots = 2: outputs directly behind the chain (the round-1 schedule)."""
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path), slots), nprocs=world, join=True)
    s = SPEC
    oracle = oias.IASOracle(s['C'], s['alpha'], s['beta'], s['gamma'], s['cp_gamma'])
    oracle.run(gi.ias_batches(s))
    ranks = [np.load(os.path.join(str(tmp_path), 'rank%d.npz' % r)) for r in range(world)]
    for r in ranks:                                         # every rank ends with the same global state
        assert np.array_equal(r['thr'], oracle.class_threshold)
        assert np.array_equal(r['statics'], oracle.statics_class)
        np.testing.assert_allclose(r['mean'], oracle.class_mean_probs, rtol=1e-6)
        assert np.array_equal(r['mean'], ranks[0]['mean'])
    n_win = (s['N'] + 2 * s['B'] - 1) // (2 * s['B'])
    by_window = {}
    for r in ranks:
        for w in r['windows']:
            by_window[int(w)] = (r['plbl_%d' % w], r['thr_%d' % w])
    assert sorted(by_window) == list(range(n_win))
    plbl = np.concatenate([by_window[w][0] for w in range(n_win)])
    assert np.array_equal(plbl, np.stack(oracle.labels))
    thr_groups = np.concatenate([by_window[w][1] for w in range(n_win)])
    assert np.array_equal(thr_groups, np.stack(oracle.threshold_trace))


# ----------------------------------------------------------------- the reference-facing sharded generator
def _cfg(s):
    from types import SimpleNamespace
    return SimpleNamespace(
        dataset=SimpleNamespace(num_classes=s['C']),
        pseudo_policy=SimpleNamespace(type='IAS_SHARDED', batch_size=s['B'],
                                      ias=SimpleNamespace(alpha=s['alpha'], beta=s['beta'], gamma=s['gamma'])),
        preprocessor=SimpleNamespace(copy_paste=SimpleNamespace(gamma=s['cp_gamma'])))


class _Identity:
    def __call__(self, x):
        return {'logits': x}


def _gen_worker(rank, world, port, out_dir, window_batches=2):
    import json
    import sys
    sys.path.insert(0, os.path.dirname(__file__))
    from host_engine import HostEngine
    import hiast_b200
    hiast_b200.register_all()
    from hiast_b200 import PSEUDO_POLICY
    from hiast_b200.pseudo_label_generator import striped_batch_order
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        s = SPEC
        batches = gi.ias_batches(s)                         # global order, batch k = images 2k, 2k+1
        order = striped_batch_order(s['N'], window_batches * s['B'], s['B'], rank, world)
        loader = [{'images': batches[idx[0] // s['B']][0], 'image_paths': batches[idx[0] // s['B']][1]} for idx in order]
        assert all(len(idx) == len(b['image_paths']) for idx, b in zip(order, loader))
        gen = PSEUDO_POLICY['IAS_SHARDED'](_cfg(s), model=_Identity(), loader=loader, dataset_len=s['N'],
                                           save_dir=os.path.join(out_dir, 'run', 'pseudo_labels'),
                                           window_batches=window_batches, device='cpu', engine_factory=HostEngine)
        gen.run()
        np.savez(os.path.join(out_dir, 'gen_rank%d.npz' % rank), thr=gen.class_threshold, mean=gen.class_mean_probs,
                 statics=gen.statics_class, trace=np.concatenate(gen.threshold_trace),
                 stats=np.array(json.dumps(gen.sample_stats)), samples=np.array(json.dumps(gen.samples_class)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world,window_batches', [(2, 2), (3, 2), (3, 4)])
def test_sharded_generator_equals_oracle_and_writes_the_reference_files(world, window_batches, tmp_path):
    """(3, 4): 13 images in windows of 8 = two windows for three ranks -- the rank without a window must still take part in
    the collectives and end with the global results."""
    import json
    cv2 = pytest.importorskip('cv2')
    port = _free_port()
    mp.spawn(_gen_worker, args=(world, port, str(tmp_path), window_batches), nprocs=world, join=True)
    s = SPEC
    oracle = oias.IASOracle(s['C'], s['alpha'], s['beta'], s['gamma'], s['cp_gamma'])
    batches = gi.ias_batches(s)
    oracle.run(batches)
    for r in range(world):                                  # every rank holds the same global results
        got = np.load(os.path.join(str(tmp_path), 'gen_rank%d.npz' % r))
        assert np.array_equal(got['thr'], oracle.class_threshold)
        assert np.array_equal(got['statics'], oracle.statics_class)
        assert np.array_equal(got['trace'], np.stack(oracle.threshold_trace))
        np.testing.assert_allclose(got['mean'], oracle.class_mean_probs, rtol=1e-6)
        assert json.loads(str(got['stats'])) == json.loads(json.dumps(oracle.sample_stats))
        assert json.loads(str(got['samples'])) == json.loads(json.dumps(oracle.samples_class))
    save_dir = os.path.join(str(tmp_path), 'run', 'pseudo_labels')
    paths = [p for _, ps in batches for p in ps]
    assert len(os.listdir(save_dir)) == len(paths)
    for i, p in enumerate(paths):
        png = cv2.imread(os.path.join(save_dir, os.path.splitext(p)[0] + '_pseudo_label.png'), cv2.IMREAD_UNCHANGED)
        assert np.array_equal(png, oracle.labels[i])
    root = os.path.join(save_dir, '..')                     # written once, by rank 0
    assert np.array_equal(np.load(os.path.join(root, 'class_threshold.npy')), oracle.class_threshold)
    assert json.load(open(os.path.join(root, 'samples_with_class.json'))) == json.loads(json.dumps(oracle.samples_class))


def test_striped_batch_order():
    from hiast_b200.pseudo_label_generator import striped_batch_order
    assert striped_batch_order(9, 4, 2, 0, 2) == [[0, 1], [2, 3], [8]]
    assert striped_batch_order(9, 4, 2, 1, 2) == [[4, 5], [6, 7]]
    assert striped_batch_order(3, 4, 2, 1, 2) == []
