#!/usr/bin/env python
"""How fast can this host create pseudo-label-sized files?  (The e2e leg with the PNG files written ends in open / write / close
of ~330 KB files; this probe separates the file system from the GPU pipeline.)

    python tools/fs_probe.py [--dirs /tmp /dev/shm] [--files 2048] [--kb 328] [--procs 1 2 4 8]

Uses the library's own writer (hiast_write_files: POSIX threads, no interpreter lock).  Prints one JSON line.
"""
import argparse
import json
import multiprocessing as mp
import os
import shutil
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def work(args):
    root, n_files, kb, threads, barrier_path = args
    from hiast_b200 import ops
    blob = np.random.default_rng(0).integers(0, 256, size=kb * 1024 * 64, dtype=np.uint8)
    d = tempfile.mkdtemp(dir=root)
    offs = [kb * 1024 * i for i in range(65)]
    t0 = time.perf_counter()
    done = 0
    while done < n_files:
        paths = [os.path.join(d, 'f%06d_pseudo_label.png' % (done + i)) for i in range(64)]
        ops.write_files(paths, blob, offs, threads)
        done += 64
    dt = time.perf_counter() - t0
    shutil.rmtree(d, ignore_errors=True)
    return dt


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--dirs', nargs='*', default=['/tmp', '/dev/shm'])
    ap.add_argument('--files', type=int, default=2048)
    ap.add_argument('--kb', type=int, default=328)
    ap.add_argument('--procs', nargs='*', type=int, default=[1, 2, 4, 8])
    ap.add_argument('--threads', nargs='*', type=int, default=[1, 3, 8])
    args = ap.parse_args()
    res = {'cpus': os.cpu_count(), 'file_kb': args.kb}
    for root in args.dirs:
        if not os.path.isdir(root):
            continue
        for p in args.procs:
            for t in args.threads:
                with mp.get_context('spawn').Pool(p) as pool:
                    dts = pool.map(work, [(root, args.files, args.kb, t, None)] * p)
                rate = p * args.files / max(dts)
                res['%s procs=%d threads=%d' % (root, p, t)] = {'files_per_s': round(rate), 'gb_per_s': round(rate * args.kb * 1024 / 1e9, 2)}
    try:
        res['mounts'] = [l.split()[:3] for l in open('/proc/mounts') if l.split()[1] in ('/', '/tmp', '/dev/shm')]
    except Exception:
        pass
    print(json.dumps(res))


if __name__ == '__main__':
    main()
