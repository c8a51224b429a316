#!/usr/bin/env python
"""Kernel shares from an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file x.csv ...`).

    python tools/launch_shares.py gpurun_out/launches_r02.csv > profiles/launches_r02.md
"""
import collections
import csv
import sys


def main(path):
    rows = []
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    rd = csv.reader(lines)
    hdr = next(rd)
    ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    scale = {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}
    for r in rd:
        if len(r) <= vi:
            continue
        try:
            rows.append((r[ki], float(r[vi].replace(',', '')) * scale.get(r[ui], 1.0)))
        except ValueError:
            pass
    agg = collections.OrderedDict()
    for k, us in rows:
        name = k.split('(')[0].replace('void ', '').replace('hiast::', '')
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += us
    total = sum(a[1] for a in agg.values())
    print('# ncu launch list: %s (%d launches, %.1f ms of kernel time; per-launch times under ncu are cold-cache and serialised -- '
          'the SHARES are what matters)\n' % (path.split('/')[-1], len(rows), total / 1e3))
    print('| kernel | launches | total us | mean us | share |')
    print('|---|---|---|---|---|')
    for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print('| `%s` | %d | %.1f | %.1f | %.1f %% |' % (name[:70], n, us, us / n, 100 * us / total))


if __name__ == '__main__':
    main(sys.argv[1])
