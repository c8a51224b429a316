"""Host-side logic of the generators that needs no GPU: the window pipeline's schedule (phase A of window j, threshold chain
of j-1, outputs of j-2), the stride-8 batch queue (one phase-A launch per window) and the window striping helper."""

from types import SimpleNamespace

import numpy as np
import torch

from hiast_b200.pseudo_label_generator import LowResLogits, _WindowPipeline, striped_batch_order


class FakeEngine:
    """Records the order of the phase calls; holds just enough state for the eager (host) path of the pipeline."""

    def __init__(self, B=2, window=4, C=3, H=4, W=4):
        self.B, self.C, self.H, self.W = B, C, H, W
        self.max_images = 3 * window
        g = self.max_images // B
        self.plbl = np.zeros((self.max_images, H, W), dtype=np.uint8)
        self.counts = torch.zeros((self.max_images, C), dtype=torch.int64)
        self.confsum = torch.zeros((g, C), dtype=torch.int64)
        self.thr_groups = torch.zeros((g, C), dtype=torch.float64)
        self.thr_state = torch.zeros(C, dtype=torch.float64)
        self.calls = []

    def phase_a_lowres(self, lr, first_image=0):
        self.calls.append(('A', first_image, tuple(lr.shape), float(lr.sum())))

    def phase_a(self, logits, first_image=0):
        self.calls.append(('A', first_image, tuple(logits.shape), None))

    def phase_b(self, first_image, n):
        self.calls.append(('B', first_image, n))

    def phase_c(self, first_image, n):
        self.calls.append(('C', first_image, n))


class FakeGen:
    _stager = None

    def __init__(self):
        self.saved = []

    def _save_async(self, plbl, path):
        self.saved.append(path)

    def _release_staged(self, slots):
        pass


def lr(n, fill, size=(4, 4), hw=(2, 3)):
    return LowResLogits(torch.full((n, 3) + hw, float(fill)), size)


def test_schedule_keeps_three_windows_in_flight():
    e, gen = FakeEngine(), FakeGen()
    pipe = _WindowPipeline(gen, e, scan=True)
    for k in range(7):                                     # 7 batches of 2 -> windows of 4, 4, 4 and a trailing 2
        pipe.add(torch.zeros(2, 3, 4, 4), ['img%d_a.png' % k, 'img%d_b.png' % k])
    pipe.finish()
    order = [(c[0], c[1]) for c in e.calls]
    slot = {0: 0, 1: 4, 2: 8, 3: 0}
    want = [('A', 0), ('A', 2), ('A', 4), ('A', 6), ('B', 0),                 # window 1 closed -> chain of window 0
            ('A', 8), ('A', 10), ('B', 4), ('C', 0),                           # window 2 closed -> chain 1, outputs 0
            ('A', 0), ('B', 8), ('C', 4),                                       # window 3 (2 images) closed at finish
            ('B', 0), ('C', 8), ('C', 0)]
    assert order == want, order
    assert [r['w'] for r in pipe.results] == [0, 1, 2, 3]
    assert [len(r['paths']) for r in pipe.results] == [4, 4, 4, 2]
    assert len(gen.saved) == 14 and slot[3] == 0


def test_constant_threshold_policies_emit_a_window_as_soon_as_it_closes():
    e, gen = FakeEngine(), FakeGen()
    pipe = _WindowPipeline(gen, e, scan=False)
    for k in range(4):
        pipe.add(torch.zeros(2, 3, 4, 4), ['a%d' % k, 'b%d' % k])
    pipe.finish()
    assert [(c[0], c[1]) for c in e.calls] == [('A', 0), ('A', 2), ('C', 0), ('A', 4), ('A', 6), ('C', 4)]


def test_stride8_batches_of_a_window_become_one_launch():
    e, gen = FakeEngine(B=2, window=8), FakeGen()
    pipe = _WindowPipeline(gen, e, scan=True)
    for k in range(4):
        pipe.add(lr(2, k + 1), ['p%d' % k, 'q%d' % k])
        assert e.calls == [] or k == 3                     # only queued until the window closes
    a = [c for c in e.calls if c[0] == 'A']
    assert len(a) == 1
    _, first, shape, total = a[0]
    assert first == 0 and shape == (8, 3, 2, 3)
    assert total == 2 * 3 * 2 * 3 * (1 + 2 + 3 + 4)        # concatenated in order
    pipe.add(lr(2, 5), ['x', 'y'])
    pipe.add(lr(2, 6, size=(8, 8)), ['z', 'w'])            # different target size -> the run so far is launched
    a = [c for c in e.calls if c[0] == 'A']
    assert len(a) == 2 and a[1][1] == 8 and a[1][2][0] == 2
    pipe.finish()
    a = [c for c in e.calls if c[0] == 'A']
    assert len(a) == 3 and a[2][1] == 10 and a[2][2] == (2, 3, 2, 3)


def test_sharded_window_sizes_are_enforced():
    import pytest
    e, gen = FakeEngine(), FakeGen()
    pipe = _WindowPipeline(gen, e, scan=True, rank=1, world=2, n_total=10)     # windows of 4: rank 1 owns window 1 only
    pipe.world = 1                                         # no process group in this test: skip the token hops
    pipe.rank, pipe._global = 1, (lambda j: 2 * j + 1)
    pipe.add(torch.zeros(2, 3, 4, 4), ['a', 'b'])
    pipe.add(torch.zeros(2, 3, 4, 4), ['c', 'd'])
    with pytest.raises(ValueError):
        pipe.add(torch.zeros(2, 3, 4, 4), ['e', 'f'])      # window 3 does not exist (10 images = windows 0, 1, 2)


def test_striping_covers_every_image_once():
    n, window, b = 37, 8, 2
    for world in (1, 2, 3, 5):
        seen = []
        for r in range(world):
            for batch in striped_batch_order(n, window, b, r, world):
                assert 1 <= len(batch) <= b
                seen += batch
        assert sorted(seen) == list(range(n))
    per_rank = [striped_batch_order(n, window, b, r, 2) for r in range(2)]
    assert per_rank[0][0] == [0, 1] and per_rank[1][0] == [8, 9] and per_rank[0][4] == [16, 17]
    assert per_rank[0][-1] == [36]                         # the trailing 1-image batch belongs to window 4 -> rank 0
