#!/usr/bin/env python
"""Per-kernel SASS mnemonic counts of libhiast_b200.so + the hot loop of the roofline kernel (evidence in profiles/).

    python tools/sass_summary.py > profiles/sass_r02.txt

Uses `cuobjdump -sass` on the in-tree library (built with -lineinfo for sm_100a).  Counts are static instruction counts.
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'hiast_b200', 'libhiast_b200.so')
WATCH = ['FFMA2', 'FADD2', 'FMUL2', 'FFMA', 'FADD', 'FMUL', 'MUFU.EX2', 'MUFU.RCP', 'MUFU.LG2', 'LDGSTS', 'LDG', 'STG', 'LDS', 'STS', 'ATOMS',
         'ATOMG', 'RED', 'UBLKCP', 'UTMALDG', 'HMMA', 'UTCHMMA', 'BAR', 'DEPBAR', 'LDGDEPBAR', 'SHFL', 'MATCH', 'DADD', 'DMUL', 'DFMA']


def demangle(names):
    out = subprocess.run(['c++filt'], input='\n'.join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    arch = set(re.findall(r'arch = (sm_\w+)', sass))
    for line in sass.splitlines():
        m = re.match(r'\s*Function : (\S+)', line)
        if m:
            cur = kernels.setdefault(m.group(1), [])
            continue
        m = re.match(r'\s*/\*[0-9a-f]{4,}\*/\s+(.*?);', line)
        if m and cur is not None:
            cur.append(m.group(1).strip())
    names = demangle(list(kernels))
    print('# SASS summary of hiast_b200/libhiast_b200.so  (cuobjdump -sass; cubin architectures: %s)' % ', '.join(sorted(arch)))
    print('# static instruction counts per kernel; columns: total, then the watched mnemonics that occur')
    print()
    for k, ins in kernels.items():
        cnt = collections.Counter()
        for i in ins:
            op = i.split()[1] if i.startswith('@') and len(i.split()) > 1 else i.split()[0]
            for w in WATCH:
                if op == w or op.startswith(w + '.'):
                    cnt[w] += 1
                    break
        short = re.sub(r'\(.*', '', names.get(k, k).replace('(anonymous namespace)::', ''))
        print('%-64s total %5d  %s' % (short[:64], len(ins), '  '.join('%s %d' % (w, cnt[w]) for w in WATCH if cnt[w])))
    # hot loop of the roofline kernel: from the first LDS.128 of the tile loop to the backward branch
    target = next((k for k in kernels if 'k_softmax_hist_grs' in k and 'Li19ELi0' in k), None)
    if target:
        ins = kernels[target]
        print()
        print('# %s : %d instructions; excerpt around the packed exponential / first-index bookkeeping' % (names[target][:90], len(ins)))
        first = next((i for i, s in enumerate(ins) if 'MUFU.EX2' in s), 0)
        for s in ins[max(0, first - 30):first + 50]:
            print('    ' + s)


if __name__ == '__main__':
    sys.exit(main())
