#!/usr/bin/env python
"""Timeline of one fused IAS window (development): which CTA ran which unit when.

    python tools/fused_trace.py [--images 64] [--gif 2]
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hiast_b200 import _lib  # noqa: E402
from hiast_b200.ias_engine import IASEngine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--images', type=int, default=64)
ap.add_argument('--gif', type=int, default=2)
args = ap.parse_args()
C, H, W = 19, 1024, 2048
g = torch.Generator(device='cuda').manual_seed(1234)
pool = torch.randn(args.images, C, H, W, generator=g, device='cuda') * 3
eng = IASEngine(C, H, W, 2, 0.5, 0.9, 8.0, 0.99, args.images, fused=True)
eng.groups_in_flight = args.gif
for _ in range(2):
    eng.process_fused(pool)
torch.cuda.synchronize()
sms = _lib.lib().hiast_device_sm_count()
trace = torch.zeros(sms, 256, 6, dtype=torch.int64, device='cuda')
_lib.check(_lib.lib().hiast_debug_set_fused_trace(_lib.ptr(trace)), 'trace')
eng.process_fused(pool)
torch.cuda.synchronize()
_lib.lib().hiast_debug_set_fused_trace(None)
t = trace.cpu().numpy().astype(np.uint64)
kind = (t[:, :, 0] >> np.uint64(32)).astype(np.int64)
idx = (t[:, :, 0] & np.uint64(0xffffffff)).astype(np.int64)
valid = t[:, :, 1] > 0
t0 = t[:, :, 1][valid].min()
beg = (t[:, :, 1].astype(np.int64) - int(t0)) / 1e3
end = (t[:, :, 2].astype(np.int64) - int(t0)) / 1e3
print('window: %.1f us total, %d units logged' % (end[valid].max(), valid.sum()))
for k, name in ((1, 'A'), (2, 'C')):
    m = valid & (kind == k)
    d = (end - beg)[m]
    print('%s-units: n=%d  duration us: mean %.1f  p50 %.1f  p95 %.1f  max %.1f   total SM-time %.0f us' %
          (name, m.sum(), d.mean(), np.median(d), np.percentile(d, 95), d.max(), d.sum()))
closer = valid & (t[:, :, 5] > 0)
wb = (t[:, :, 3].astype(np.int64) - int(t0)) / 1e3
we = (t[:, :, 4].astype(np.int64) - int(t0)) / 1e3
pub = (t[:, :, 5].astype(np.int64) - int(t0)) / 1e3
print('closers: n=%d  wait us: mean %.1f max %.1f ; threshold step us: mean %.1f max %.1f' %
      (closer.sum(), (we - wb)[closer].mean(), (we - wb)[closer].max(), (pub - we)[closer].mean(), (pub - we)[closer].max()))
order = np.argsort(pub[closer])
gs = (idx[closer] // max(1, (idx[valid & (kind == 1)].max() + 1) // (args.images // 2)))[order]
print('group publish times us (first 12):', np.round(np.sort(pub[closer])[:12], 1).tolist())
print('group publish times us (last 6):', np.round(np.sort(pub[closer])[-6:], 1).tolist())
busy = np.array([(end - beg)[c][valid[c]].sum() for c in range(sms)])
span = np.array([end[c][valid[c]].max() - beg[c][valid[c]].min() for c in range(sms)])
print('per-CTA busy/span: mean %.3f min %.3f' % ((busy / span).mean(), (busy / span).min()))
c0 = 0
print('CTA 0 timeline (kind idx begin end):')
for e in range(min(valid[c0].sum(), 14)):
    print('   %s %5d  %8.1f %8.1f  %s' % ('?AC'[kind[c0, e]], idx[c0, e], beg[c0, e], end[c0, e],
                                         'closer wait %.1f step %.1f' % (we[c0, e] - wb[c0, e], pub[c0, e] - we[c0, e]) if t[c0, e, 5] else ''))
