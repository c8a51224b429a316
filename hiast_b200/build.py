"""Build libhiast_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m hiast_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels with the repo snapshot to
the GPU box; `hiast_b200._lib` loads it with ctypes and fails loudly if it is missing.
"""

from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libhiast_b200.so')
SOURCES = ['api.cu', 'ias_phase_a.cu', 'ias_upsample.cu', 'ias_scan_select.cu', 'ias_fused.cu', 'cbst.cu', 'loss.cu', 'confusion.cu', 'copy_paste.cu', 'ema.cu', 'png.cu', 'resize.cu', 'validate.cu', 'ce_general.cu', 'host_pipeline.cu']
HEADERS = ['common.cuh', 'ias_common.cuh', 'packed_math.cuh', 'scan_math.h', os.path.join('..', '..', 'include', 'hiast_b200.h'),
           os.path.join('..', '..', 'include', 'hiast_b200_dev.h')]

NVCC_FLAGS = [
    '-O3', '-std=c++17',
    '-gencode', 'arch=compute_100a,code=sm_100a',
    '-lineinfo',
    '-Xcompiler', '-fPIC,-O3,-ffp-contract=off,-fvisibility=hidden',
    '--cudart', 'static',
]


def find_nvcc():
    for cand in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', shutil.which('nvcc')):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found (looked at $NVCC, /usr/local/cuda/bin/nvcc, PATH)')


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def _dev_flag():
    """HIAST_DEV_VARIANTS=1 in the environment compiles the measured-and-dropped kernel variants (phase-A hist_modes other than
    1 / 83, the TMA-staged kernel, the fused persistent window kernel) into the library; the product build leaves them out."""
    return os.environ.get('HIAST_DEV_VARIANTS', '0') not in ('', '0')


def build(force=False, verbose=False):
    stamp = os.path.join(HERE, 'build', 'flavour')
    want = 'dev' if _dev_flag() else 'product'
    have = open(stamp).read().strip() if os.path.exists(stamp) else None
    if not force and not _stale() and have == want:
        return LIB
    nvcc = find_nvcc()
    objdir = os.path.join(HERE, 'build')
    os.makedirs(objdir, exist_ok=True)
    procs = []
    objs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace('.cu', '.o'))
        objs.append(obj)
        cmd = [nvcc] + NVCC_FLAGS + (['-DHIAST_DEV_VARIANTS'] if _dev_flag() else []) + (['-Xptxas', '-v'] if verbose else []) + ['-c', os.path.join(CSRC, src), '-o', obj]
        procs.append((src, cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, cmd, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write('FAILED: %s\n%s\n' % (' '.join(cmd), out))
        elif verbose or out.strip():
            sys.stderr.write(out)
    if failed:
        raise RuntimeError('nvcc failed')
    link = [nvcc, '-shared', '-o', LIB + '.tmp'] + objs + ['--cudart', 'static',
                                                         '-gencode', 'arch=compute_100a,code=sm_100a']
    subprocess.check_call(link)
    os.replace(LIB + '.tmp', LIB)
    with open(stamp, 'w') as f:
        f.write(want)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='--verbose' in sys.argv))
