"""Host-side logic of the generators that needs no GPU: the stride-8 batch queue (one phase-A launch per contiguous run
of images) and the window striping helper."""

import torch

from hiast_b200.pseudo_label_generator import LowResLogits, _flush_lowres, _phase_a, striped_batch_order


class FakeEngine:
    def __init__(self):
        self.calls = []

    def phase_a_lowres(self, lr, first_image=0):
        self.calls.append((first_image, tuple(lr.shape), float(lr.sum())))

    def phase_a(self, logits, first_image=0):
        self.calls.append(('full', first_image, tuple(logits.shape)))


def lr(n, fill, size=(64, 128), hw=(9, 17)):
    return LowResLogits(torch.full((n, 3) + hw, float(fill)), size)


def test_contiguous_batches_become_one_launch():
    e = FakeEngine()
    for k in range(4):
        _phase_a(e, lr(2, k + 1), 2 * k)
    assert e.calls == []                                   # only queued
    _flush_lowres(e)
    assert len(e.calls) == 1
    first, shape, total = e.calls[0]
    assert first == 0 and shape == (8, 3, 9, 17)
    assert total == 2 * 3 * 9 * 17 * (1 + 2 + 3 + 4)        # concatenated in order
    _flush_lowres(e)                                       # nothing left
    assert len(e.calls) == 1


def test_gap_size_change_and_full_resolution_flush_or_bypass():
    e = FakeEngine()
    _phase_a(e, lr(2, 1), 0)
    _phase_a(e, lr(2, 2), 2)
    _phase_a(e, lr(2, 3), 64)                              # other window slot: not contiguous -> the run so far is launched
    assert [c[0] for c in e.calls] == [0] and e.calls[0][1][0] == 4
    _phase_a(e, lr(1, 4), 66)                              # trailing 1-image batch joins the run
    _phase_a(e, lr(2, 5, size=(32, 64)), 67)               # different target size -> flush, new run
    assert [c[0] for c in e.calls] == [0, 64] and e.calls[1][1][0] == 3
    _phase_a(e, torch.zeros(2, 3, 64, 128), 10)            # full-resolution logits are consumed immediately
    assert e.calls[-1] == ('full', 10, (2, 3, 64, 128))
    _flush_lowres(e)
    assert e.calls[-1][0] == 67 and e.calls[-1][1] == (2, 3, 9, 17)


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
