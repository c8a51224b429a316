#!/usr/bin/env python
"""Summarise .ncu-rep captures (read with `ncu -i ... --page raw --csv`) into a small markdown table + JSON.

    python tools/ncu_summary.py gpurun_out/phase_a_mode6.ncu-rep [...] --out profiles/ncu_phase_a_r01
"""
import argparse
import csv
import io
import json
import subprocess

KEYS = [
    ('gpu__time_duration.sum', 'duration'),
    ('dram__bytes_read.sum', 'dram_read'),
    ('dram__bytes_write.sum', 'dram_write'),
    ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram_pct_of_ncu_peak'),
    ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue_active_pct'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps_active_pct'),
    ('smsp__warps_eligible.avg.per_cycle_active', 'eligible_warps_per_cycle'),
    ('launch__registers_per_thread', 'registers'),
    ('launch__grid_size', 'grid'),
    ('launch__block_size', 'block'),
    ('smsp__inst_executed.sum', 'warp_instructions'),
    ('sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'xu_pipe_pct'),
    ('sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'fma_pipe_pct'),
    ('sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'alu_pipe_pct'),
    ('l1tex__throughput.avg.pct_of_peak_sustained_active', 'l1tex_pct'),
    ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l2_pct'),
    ('smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'stall_long_scoreboard'),
    ('smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio', 'stall_not_selected'),
    ('smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'stall_wait'),
    ('smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'stall_barrier'),
    ('smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', 'stall_math_pipe'),
    ('smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio', 'stall_mio_throttle'),
    ('smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio', 'stall_lg_throttle'),
    ('smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'stall_short_scoreboard'),
]

UNIT_SCALE = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0, 'us': 1e-6, 'ms': 1e-3, 'ns': 1e-9, 's': 1.0}


def read(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    res = []
    for vals in rows[2:]:
        d = {'kernel': vals[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else '?'}
        for key, name in KEYS:
            if key in hdr:
                i = hdr.index(key)
                try:
                    v = float(vals[i].replace(',', ''))
                except ValueError:
                    continue
                d[name] = v * UNIT_SCALE.get(units[i], 1.0) if units[i] in UNIT_SCALE else v
        res.append(d)
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('reports', nargs='+')
    ap.add_argument('--out', required=True)
    ap.add_argument('--images', type=int, default=16, help='images per profiled launch')
    ap.add_argument('--note', default='')
    args = ap.parse_args()
    allres = {}
    for rep in args.reports:
        allres[rep.split('/')[-1]] = read(rep)
    json.dump(allres, open(args.out + '.json', 'w'), indent=1)
    cols = [(rep, i, r) for rep, rs in allres.items() for i, r in enumerate(rs)]       # one column per captured launch
    with open(args.out + '.md', 'w') as f:
        f.write('# ncu summary (`ncu --set full --clock-control none --import-source on`; per-launch times under ncu are cold-cache '
                'and serialised)\n\n%s\n\n' % args.note)
        names = [n for _, n in KEYS]
        f.write('| metric | ' + ' | '.join('%s #%d' % (rep.replace('.ncu-rep', ''), i) for rep, i, _ in cols) + ' |\n|---|' + '---|' * len(cols) + '\n')
        f.write('| kernel | ' + ' | '.join(r['kernel'].split('(')[0][:40] for _, _, r in cols) + ' |\n')
        for n in names:
            cells = []
            for _, _, r in cols:
                v = r.get(n)
                if v is None:
                    cells.append('')
                elif n == 'duration':
                    cells.append('%.1f us' % (v * 1e6))
                elif n in ('dram_read', 'dram_write'):
                    cells.append('%.1f MB' % (v / 1e6))
                else:
                    cells.append('%.3g' % v)
            f.write('| %s | ' % n + ' | '.join(cells) + ' |\n')
    print(open(args.out + '.md').read())


if __name__ == '__main__':
    main()
