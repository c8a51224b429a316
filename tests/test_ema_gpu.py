"""EMA teacher update (hiast_ema_update / hiast_multi_copy) against the reference fixture and the oracle.

Reference: utils/utils.py:115-123 (update_ema_model)."""
import os

import numpy as np
import pytest
import torch

from oracle import ema as oema

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), 'golden')


def small_net():
    return torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3), torch.nn.BatchNorm2d(8), torch.nn.Conv2d(8, 5, 1, bias=False),
                               torch.nn.Linear(7, 1031))


def test_ema_update_equals_reference_fixture():
    from hiast_b200.ema import update_ema_model
    g = np.load(os.path.join(GOLD, 'ema_update.npz'))
    student, teacher = small_net().cuda(), small_net().cuda()
    for i, p in enumerate(student.parameters()):
        p.data.copy_(torch.from_numpy(g['q%d' % i]))
    for i, p in enumerate(teacher.parameters()):
        p.data.copy_(torch.from_numpy(g['k%d' % i]))
    for i, b in enumerate(student.buffers()):
        b.data.copy_(torch.from_numpy(g['bq%d' % i]))
    out = update_ema_model(teacher, student, float(g['gamma']))
    assert out is teacher
    for i, p in enumerate(teacher.parameters()):
        assert np.array_equal(p.data.cpu().numpy(), g['new%d' % i]), i
    for i, b in enumerate(teacher.buffers()):
        assert np.array_equal(b.data.cpu().numpy(), g['bnew%d' % i]), i


@pytest.mark.parametrize('gamma', [0.999, 0.99, 0.5, 0.0, 1.0])
def test_ema_update_vs_oracle_odd_sizes_and_repeated_calls(gamma):
    """Unaligned views, sizes around the chunk size, repeated calls (cached tables) and re-allocated parameters."""
    from hiast_b200.ema import update_ema_model, CHUNK_ELEMS

    class Bag(torch.nn.Module):
        def __init__(self, seed):
            super().__init__()
            g = torch.Generator().manual_seed(seed)
            sizes = [1, 3, 4, 5, 1023, CHUNK_ELEMS - 1, CHUNK_ELEMS, CHUNK_ELEMS + 1, 3 * CHUNK_ELEMS + 7]
            self.ps = torch.nn.ParameterList([torch.nn.Parameter(torch.randn(n, generator=g)) for n in sizes])
            base = torch.randn(4096 + 1, generator=g)
            self.odd = torch.nn.Parameter(base[1:])             # 4-byte aligned only
            self.register_buffer('steps', torch.randint(0, 99, (1,), generator=g))
            self.register_buffer('stat', torch.rand(37, generator=g))

    student, teacher = Bag(1).cuda(), Bag(2).cuda()
    for rep in range(3):
        ks = [p.data.cpu().numpy().copy() for p in teacher.parameters()]
        qs = [p.data.cpu().numpy().copy() for p in student.parameters()]
        want = oema.ema_update(ks, qs, gamma)
        update_ema_model(teacher, student, gamma)
        for w, p in zip(want, teacher.parameters()):
            assert np.array_equal(p.data.cpu().numpy(), w)
        for bk, bq in zip(teacher.buffers(), student.buffers()):
            assert torch.equal(bk, bq)
        with torch.no_grad():                                    # an optimizer step; then re-allocate one parameter
            for p in student.parameters():
                p.add_(0.01)
            student.ps[2].data = student.ps[2].data.clone()
            student.steps += 1


def test_ema_update_rejects_cpu_models():
    from hiast_b200.ema import update_ema_model
    from hiast_b200._lib import HiastError
    with pytest.raises(HiastError):
        update_ema_model(small_net(), small_net(), 0.99)


def test_ema_cache_follows_the_modules_not_their_ids():
    """the table cache is keyed by weak references -- a student that is dropped and replaced (possibly at the
    same id()) is walked again, and entries die with their modules."""
    import gc
    from hiast_b200 import ema
    ema.reset()
    teacher = small_net().cuda()
    for rep in range(4):
        student = small_net().cuda()
        with torch.no_grad():
            for p in student.parameters():
                p.fill_(float(rep + 1))
            for p in teacher.parameters():
                p.zero_()
        ema.update_ema_model(teacher, student, 0.5)
        for p in teacher.parameters():
            assert torch.all(p == 0.5 * (rep + 1)), rep
        del student
        gc.collect()
    assert len(ema._cache) == 1
    # a replaced FIRST Parameter object is noticed without reset()
    student = small_net().cuda()
    ema.update_ema_model(teacher, student, 0.5)
    student[0].weight = torch.nn.Parameter(torch.full_like(student[0].weight, 8.0))
    with torch.no_grad():
        teacher[0].weight.zero_()
    ema.update_ema_model(teacher, student, 0.5)
    assert torch.all(teacher[0].weight == 4.0)
    del teacher
    gc.collect()
    assert len(ema._cache) == 0
