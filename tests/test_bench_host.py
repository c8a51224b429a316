"""Host-side pieces of bench.py that need neither a GPU nor the reference: argument defaults of the two arms, the clock line built
from nvidia-smi samples, the non-zero ranks of the reference arm."""

import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location('bench_under_test', os.path.join(ROOT, 'bench.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_argument_defaults_per_arm(monkeypatch):
    b = _bench()
    monkeypatch.setattr(sys, 'argv', ['bench.py'])
    a = b.parse_args()
    assert (a.gpus, a.steps, a.warmup, a.impl) == (1, 46, 3, 'ours')
    monkeypatch.setattr(sys, 'argv', ['bench.py', '--impl', 'reference'])
    a = b.parse_args()
    assert (a.steps, a.warmup) == (2, 1)                       # a CPU step takes seconds
    monkeypatch.setattr(sys, 'argv', ['bench.py', '--impl', 'reference', '--gpus', '8', '--steps', '20', '--warmup', '5'])
    a = b.parse_args()
    assert (a.gpus, a.steps, a.warmup) == (8, 20, 5)           # the driver's flags are taken as given by both arms


def test_clock_line_from_samples():
    b = _bench()
    s = b.ClockSampler(0)
    assert s.stop()['reasons'] == ['nvidia-smi unavailable']   # never started

    class Done:
        def terminate(self):
            pass

        def wait(self, timeout=None):
            return 0
    s = b.ClockSampler(0)
    s.proc = Done()
    s.samples = ['210, 1965, Not Active, Not Active, Not Active, Not Active',          # idle before the load
                 '1905, 1965, Not Active, Not Active, Not Active, Active',
                 '1620, 1965, Not Active, Not Active, Not Active, Active',
                 '1755, 1965, Not Active, Not Active, Not Active, Not Active',
                 'garbage line']
    c = s.stop()
    assert c == {'sm_mhz': 1755.0, 'sm_max_mhz': 1965.0, 'reasons': ['sw_power_cap'], 'samples': 4}
    assert s.wait_first(timeout=0.01) is True                  # samples are there already


def test_reference_arm_runs_on_rank_zero_only(monkeypatch, capsys):
    b = _bench()
    monkeypatch.setenv('RANK', '3')
    monkeypatch.setattr(sys, 'argv', ['bench.py', '--impl', 'reference', '--gpus', '8'])
    b.reference_arm(b.parse_args())
    assert capsys.readouterr().out == ''                       # the other ranks exit without work and without a line
