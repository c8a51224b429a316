#!/usr/bin/env python
"""Benchmark of the IAS pseudo-labelling hot path (BASELINE.json metric) -- one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): instance-adaptive pseudo-labelling of 19x1024x2048 float32 logit
maps, batch (group) size 2, alpha .5 / beta .9 / gamma 8.  A *step* is one window of 64 images
(32 groups) per GPU through phase A (softmax + arg-max + confidence + key histograms), phase B (the
sequential threshold scan), phase C (threshold-and-mask, counts) and the mean-prob EMA; the threshold
state carries over from step to step, so K steps are one K*64-image job (46 steps ~ the 2975-image
Cityscapes train set).  474 GB of logits do not fit a GPU, so the maps cycle through a resident pool
of 64 synthetic maps (10.2 GB: every step streams 80x the L2, no flush needed).

* ``value``     images/s, whole job over all N GPUs, inputs resident in HBM, CUDA-event timed, max over ranks.
* ``e2e``       same metric through the reference-facing API (``PSEUDO_POLICY['IAS'](cfg).run()``) with HOST
                (pinned) logits: H2D of every batch and D2H of every label map inside the timed region.
* ``roofline``  phase A (the dominant kernel): algorithmic bytes (77 B/px) / its mean launch time, measured
                with CUDA events inside the timed steps, against MEASURED_PEAKS.json's HBM copy bandwidth.
* ``cpu_baseline`` / ``--impl reference``: the UNMODIFIED reference (oracle/_ref/code, ``IASPseudoGenerator.run``) on the host
                cores, on a bounded sample of the same workload; the oracle port only if that copy is absent.

Multi-GPU (torchrun, one rank per GPU): windows are striped over ranks and the 19-double threshold state is
handed rank to rank INSIDE the scan kernel through CUDA-IPC peer memory (NCCL send/recv as the fallback;
hiast_b200/sharded.py); one NCCL all-gather at the end of the job; weak scaling, per-GPU work fixed.  After the
timed job the same global job is replayed on rank 0 alone for two windows per rank: `parity`.

Inside the warm-up the schedule is calibrated (`config.schedule_calibration_ms_per_step`): SMs reserved from phase A
for the scan and the mask pass of the window before, or the serial order -- whichever is faster on this box.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

C, H, W, GROUP = 19, 1024, 2048, 2
ALPHA, BETA, GAMMA, CP_GAMMA = 0.5, 0.9, 8.0, 0.99
WINDOW = 64                      # images per step per GPU
RESERVE_SMS = 12                 # SMs phase A leaves to the threshold chain and phase C (see gpu_arm)
ALG_BYTES_PER_IMAGE = H * W * (4 * C + 1)     # read logits + write uint8 label = 161 480 704 B
METRIC = 'pseudo-labelled 19x1024x2048 images/s'
UNIT = 'images/s'


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=None, help='timed steps (default: 46 windows for the GPU arm, 2 runs for --impl reference)')
    ap.add_argument('--warmup', type=int, default=None, help='untimed steps (default: 3 for the GPU arm, 1 for --impl reference)')
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--dist', default='mixed', choices=['mixed', 'diffuse', 'peaked'])
    ap.add_argument('--e2e-steps', type=int, default=4)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true', help='development: skip the end-to-end legs (the line then has no e2e)')
    ap.add_argument('--no-extra', action='store_true', help='skip the legs for BASELINE configs[2], [3], [4]')
    ap.add_argument('--extra-timeout', type=int, default=240, help='seconds after which the extra legs are abandoned (the headline line is printed without them)')
    ap.add_argument('--cpu-images', type=int, default=8)
    args = ap.parse_args()
    if args.steps is None:
        args.steps = 2 if args.impl == 'reference' else 46
    if args.warmup is None:
        args.warmup = 1 if args.impl == 'reference' else 3
    return args


def peaks():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md, 6.65 TB/s)'


def ncu_traffic():
    """dram bytes per phase-A launch from the committed ncu capture, scaled to this window; else None."""
    try:
        with open(os.path.join(ROOT, 'profiles', 'roofline_phase_a.json')) as f:
            d = json.load(f)
        return float(d['dram_bytes_per_image']) * WINDOW
    except Exception:
        return None


# ------------------------------------------------------------------------- CPU arm
def synth_logits_cpu(n, seed=1234):
    """The bench distributions on the host (diffuse / peaked alternating), float32 [n,C,H,W]."""
    import torch
    from torch.nn import functional as F
    g = torch.Generator().manual_seed(seed)
    out = torch.empty(n, C, H, W)
    for i in range(n):
        if i % 2 == 0:
            out[i] = torch.randn(C, H, W, generator=g) * 3
        else:
            low = torch.randn(1, C, 32, 64, generator=g) * 4
            out[i] = F.interpolate(low, size=(H, W), mode='bilinear', align_corners=True)[0]
            out[i] += torch.randn(C, H, W, generator=g) * 0.5
    return out


def run_cpu_arm(n_images, steps=1, warmup=0):
    """The reference's CPU path on a bounded sample: images/s, best step seconds, torch threads, kind.

    kind 'reference': the UNMODIFIED ``IASPseudoGenerator.run`` (workflows/pseudo_label_generator.py:181-213) from
    ``oracle/_ref/code`` (the hot-path files copied there by ``__graft_entry__.build()`` in the build container; git-ignored,
    travels with the snapshot) through the import shim of ``oracle/ref.py``.  kind 'port': the oracle's faithful restatement
    (same op sequence, pinned bit for bit against the reference's outputs) when that copy is absent."""
    import torch
    from oracle import ias as oias
    from oracle import ref as oref
    # torchrun exports OMP_NUM_THREADS=1; only one rank runs this leg, so it may use every host core
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    logits = synth_logits_cpu(n_images)
    batches = [(logits[i:i + GROUP], ['img_%05d.png' % (i + j) for j in range(min(GROUP, n_images - i))])
               for i in range(0, n_images, GROUP)]
    kind = 'reference' if oref.available() else 'port'
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        if kind == 'reference':
            oref.run_ias(batches, C, ALPHA, BETA, GAMMA, CP_GAMMA, keep_labels=False)
        else:
            oias.IASOracle(C, ALPHA, BETA, GAMMA, CP_GAMMA, faithful=True, keep_labels=False).run(batches)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    mean = sum(times) / len(times)                  # every timed step counts, as in the GPU arm (total time / steps)
    return n_images / mean, mean, torch.get_num_threads(), kind


CPU_NOTE = {'reference': 'the unmodified reference (oracle/_ref/code: workflows/pseudo_label_generator.py IASPseudoGenerator.run) on the '
                         'host cores, identity model on synthetic logits, cv2.imwrite stubbed',
            'port': 'oracle/_ref is absent on this box: the oracle port (oracle/ias.py, faithful op sequence, pinned bit for bit '
                    'against the reference) on the host cores'}


def reference_arm(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    # The driver launches this arm with the GPU arm's --steps K --warmup W: exactly K timed and W untimed steps, each step one
    # run() of the reference over a bounded sample of the workload -- 4 full-resolution maps (2 groups, ~2.5 s on 16 host
    # cores), so that 20 + 5 steps end in about a minute; with few steps the sample is --cpu-images maps (default 8).
    steps = max(1, args.steps)
    warm = max(0, args.warmup)
    n = args.cpu_images if steps + warm <= 4 else min(args.cpu_images, 4)
    value, secs, threads, kind = run_cpu_arm(n, steps=steps, warmup=warm)
    sample = '%d maps of 19x1024x2048, batch 2 (%d groups), per step; PNG write excluded' % (n, (n + 1) // 2)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': steps,
        'warmup': warm, 'ms_per_step': secs * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'GTA5->Cityscapes IAS pseudo-labelling, 19x1024x2048 logit maps, batch 2 (configs[1])',
                   'images_per_step': n, 'alpha': ALPHA, 'beta': BETA, 'gamma': GAMMA, 'distribution': 'mixed',
                   'note': CPU_NOTE[kind]},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': threads, 'kind': kind, 'sample': sample,
                         'host_cpus': os.cpu_count(),
                         'note': 'only softmax/max use all torch threads; the numpy/Python part is single-threaded '
                                 'by construction, as in the reference'},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = 'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        self.index, self.samples, self.proc, self.thread = index, [], None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '5'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def wait_first(self, timeout=8.0):
        """Blocks until nvidia-smi has delivered its first sample.  Its start-up (NVML attaches to every GPU of the box, under
        driver-wide locks) takes 0.1-0.5 s on a loaded host and stalls CUDA calls of other processes while it lasts: it must be
        over BEFORE the timed region starts, not inside it (seen once at N = 4: 14 samples instead of 68 and one 43 ms gap in a
        36 ms timed region)."""
        t_end = time.perf_counter() + timeout
        while self.proc is not None and not self.samples and time.perf_counter() < t_end and self.proc.poll() is None:
            time.sleep(0.01)
        return bool(self.samples)

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for s in self.samples:
            parts = [p.strip() for p in s.split(',')]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx = float(parts[1])
            except ValueError:
                continue
            for name, val in zip(names, parts[2:6]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        sm.sort()
        busy = [x for x in sm if mx and x > 0.5 * mx] or sm
        return {'sm_mhz': busy[len(busy) // 2] if busy else None, 'sm_max_mhz': mx, 'reasons': sorted(reasons),
                'samples': len(sm)}


# ------------------------------------------------------------------------- GPU arm
def make_pool(device, dist_name, n):
    import torch
    from torch.nn import functional as F
    g = torch.Generator(device=device).manual_seed(1234)
    pool = torch.empty(n, C, H, W, device=device)
    for i in range(n):
        diffuse = dist_name == 'diffuse' or (dist_name == 'mixed' and i % 2 == 0)
        if diffuse:
            pool[i] = torch.randn(C, H, W, generator=g, device=device) * 3
        else:
            low = torch.randn(1, C, 32, 64, generator=g, device=device) * 4
            pool[i] = F.interpolate(low, size=(H, W), mode='bilinear', align_corners=True)[0]
            pool[i] += torch.randn(C, H, W, generator=g, device=device) * 0.5
    return pool


def gpu_arm(args):
    import torch
    import torch.distributed as dist
    from types import SimpleNamespace
    from hiast_b200.ias_engine import IASEngine
    from hiast_b200.sharded import ShardedIAS, TokenRing

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (there is no CPU path); use --impl reference for the CPU arm')
    torch.cuda.set_device(local)
    device = torch.device('cuda', local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('TORCH_NCCL_SHOW_EAGER_INIT_P2P_SERIALIZATION_WARNING', 'false')
        dist.init_process_group('nccl', device_id=device)
    if args.gpus != world and rank == 0 and world > 1:
        print('warning: --gpus %d but WORLD_SIZE %d' % (args.gpus, world), file=sys.stderr)

    K, Wm = args.steps, max(args.warmup, 3)
    pool = make_pool(device, args.dist, WINDOW)
    # three window slots: phase A of window j, the threshold chain of j-1 and phase C of j-2 are in flight together
    engine = IASEngine(C, H, W, GROUP, ALPHA, BETA, GAMMA, CP_GAMMA, 3 * WINDOW, device=device)
    # SMs phase A leaves free for the threshold chain and phase C of the window before (they then run CONCURRENTLY with phase A
    # on the chain stream instead of between two launches; hiast_b200/sharded.py, "concurrent" schedule).  0 = serial schedule.
    engine.reserve_sms = int(os.environ.get('HIAST_RESERVE_SMS', RESERVE_SMS))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    a_events = []
    launches = {'n': 0}

    class TimedEngine:
        """Forwards to the engine; brackets every phase-A launch with CUDA events on the launching stream and counts the
        kernels of this library that are launched."""
        def __getattr__(self, name):
            return getattr(engine, name)

        def phase_a(self, logits, first_image=0):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            engine.phase_a(logits, first_image)
            e1.record()
            a_events.append((e0, e1))
            launches['n'] += 1                      # k_softmax_hist_grs

        def phase_b(self, first_image, n_images, **kw):
            engine.phase_b(first_image, n_images, **kw)
            launches['n'] += 2                      # k_hist_prefix + k_threshold_scan (token hand-off inside at N > 1)

        def phase_c(self, first_image, n_images):
            engine.phase_c(first_image, n_images)
            launches['n'] += 1                      # k_select_private

        def mean_prob_from_groups(self, confsum, counts):
            engine.mean_prob_from_groups(confsum, counts)
            launches['n'] += 1                      # k_meanprob_scan

    def run_job(n_steps, timed, on_window=None):
        total_windows = n_steps * world
        drv = ShardedIAS(TimedEngine() if timed else engine, WINDOW, total_windows * WINDOW, rank, world)
        return drv.run(lambda w: pool, on_window)

    # warm-up (also creates the NCCL p2p communicators)
    run_job(Wm, timed=False)
    # Schedule calibration, still warm-up: the concurrent schedule squeezes phase C (issue-bound) onto the reserved SMs, where it
    # takes about as long as one phase A -- on a box whose GPUs run power-capped it can become the longer of the two.  Four
    # windows per rank with the reserved SMs and four without; every rank adopts the faster one (max over ranks decides).
    ring_on = world > 1 and TokenRing.get(device, rank, world) is not None      # every rank asks (cached since the warm-up job)
    calibration = None
    if 'HIAST_RESERVE_SMS' not in os.environ and engine.reserve_sms > 0:
        calibration = {}
        for cand in (engine.reserve_sms, 0):
            engine.reserve_sms = cand
            run_job(2, timed=False)
            barrier()
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record()
            run_job(4, timed=False)
            c1.record()
            torch.cuda.synchronize()
            t = torch.tensor([c0.elapsed_time(c1) / 4], dtype=torch.float64, device=device)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            calibration[cand] = float(t[0])
        engine.reserve_sms = min(calibration, key=calibration.get)
    ShardedIAS(engine, WINDOW, K * world * WINDOW, rank, world).warm_collective()   # the timed job's all-gather shape
    engine.thr_state.fill_(0.9)
    engine.mean_state.zero_()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0 and os.environ.get('HIAST_BENCH_SAMPLER', '1') != '0':
        sampler.start()
        sampler.wait_first()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    wall0 = time.perf_counter()
    t0.record()
    run_job(K, timed=True)
    t1.record()
    barrier()
    wall1 = time.perf_counter()
    ms = t0.elapsed_time(t1)
    if os.environ.get('HIAST_BENCH_VERBOSE'):
        print('rank %d: %.3f ms per step, phase A launches %s' % (rank, ms / K, ' '.join('%.2f' % e0.elapsed_time(e1) for e0, e1 in a_events)),
              file=sys.stderr, flush=True)
    certified = engine.check_errors()
    n_launches = launches['n']
    a_ms = sum(e0.elapsed_time(e1) for e0, e1 in a_events) / max(len(a_events), 1)
    # The timed region of a short run is a few tens of milliseconds: keep the SAME load running (untimed) until the clock
    # sampler has seen it for at least 0.3 s, so that the clocks line rests on enough samples.
    # Every rank must run the SAME number of these jobs (each one ends in a collective): rank 0's clock decides.
    t_end = time.perf_counter() + max(0.0, 0.3 - (wall1 - wall0))
    while True:
        go = torch.tensor([1 if time.perf_counter() < t_end else 0], dtype=torch.int32, device=device)
        if world > 1:
            dist.broadcast(go, src=0)
        if not int(go.item()):
            break
        run_job(max(K, 10), timed=False)
        torch.cuda.synchronize()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    if clocks is not None:
        clocks['sampled_over'] = 'the timed steps plus the same load repeated untimed for >= 0.3 s, nvidia-smi -lms 5'
    starts = [round(t0.elapsed_time(e0), 2) for e0, _ in a_events]           # when each step's phase A started (this rank)
    per_rank_ms = [ms / K]
    if world > 1:
        per_rank = [None] * world
        dist.all_gather_object(per_rank, (ms / K, starts))
        per_rank_ms = [p[0] for p in per_rank]
        starts = max(per_rank, key=lambda p: p[0])[1]                        # of the slowest rank
        t = torch.tensor([ms, a_ms], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, a_ms = float(t[0]), float(t[1])
    images = K * WINDOW * world
    value = images / (ms / 1e3)

    parity = sharded_parity(engine, pool, rank, world, device, run_job) if world > 1 else None

    # ---- e2e through the reference-facing API with host buffers (every rank its own replica of the call)
    e2e = None if args.no_e2e else measure_e2e(args, device, rank, world, barrier)

    line = None
    if rank == 0:
        peak, peak_src = peaks()
        achieved = ALG_BYTES_PER_IMAGE * WINDOW / (a_ms / 1e3) / 1e9
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': K, 'warmup': Wm,
            'ms_per_step': ms / K, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': 'GTA5->Cityscapes IAS pseudo-labelling, 19x1024x2048 logit maps, batch 2 (configs[1])',
                       'images_per_step_per_gpu': WINDOW, 'images_total': images, 'alpha': ALPHA, 'beta': BETA,
                       'gamma': GAMMA, 'distribution': args.dist, 'resident_pool_maps': WINDOW,
                       'l2': 'inputs exceed L2 (10.2 GB streamed per step)',
                       'reserved_sms': engine.reserve_sms,
                       'schedule_calibration_ms_per_step': calibration,
                       'token_ring': ('none (one rank)' if world == 1 else
                                      'peer memory (CUDA IPC mailboxes)' if ring_on else
                                      'NCCL send/recv'),
                       'parallelism': (('one rank: ' if world == 1 else 'windows striped over %d ranks; per rank ' % world) +
                                       ('phase A(j+1) on %d SMs while the threshold chain and phase C of window j run on the other '
                                        '%d (chain stream)' % (148 - engine.reserve_sms, engine.reserve_sms) if engine.reserve_sms
                                        else 'phase A(j+1), then the chain of window j beside phase C of window j-1 (serial schedule)') +
                                       ('; no collective' if world == 1 else
                                        '; 19-double threshold state handed over INSIDE the scan kernel (see token_ring; '
                                        'HIAST_RING=nccl selects NCCL send/recv); one NCCL all-gather at the end'))},
            'hbm_frac_of_peak': ALG_BYTES_PER_IMAGE * value / world / 1e9 / peak,
            'roofline': {'kernel': 'k_softmax_hist_grs (phase A)', 'bound': 'hbm', 'achieved': achieved, 'peak': peak,
                         'unit': 'GB/s', 'frac': achieved / peak, 'traffic': ncu_traffic(), 'peak_source': peak_src,
                         'algorithmic_bytes_per_launch': ALG_BYTES_PER_IMAGE * WINDOW, 'launch_ms': a_ms,
                         'share_of_step': a_ms / (ms / K)},
            'step_timeline': {'ms_per_step_per_rank': [round(x, 4) for x in per_rank_ms],
                              'phase_a_starts_ms_slowest_rank': starts},
            'e2e': e2e,
            'gpu_launches': n_launches,
            'clocks': clocks,
            'pow_rounding_certified': bool(certified),
        }
        if parity is not None:
            line['parity'] = parity
        if world == 1 and not args.no_cpu_baseline:
            v, secs, threads, kind = run_cpu_arm(args.cpu_images)
            line['cpu_baseline'] = {
                'value': v, 'unit': UNIT, 'cores': threads, 'kind': kind, 'host_cpus': os.cpu_count(), 'note': CPU_NOTE[kind],
                'sample': '%d maps of 19x1024x2048, batch 2 (%.1f s of host work); PNG write excluded' % (args.cpu_images, secs)}

    # The legs for the other BASELINE configs come last and must never cost the headline: a watchdog prints the line without them
    # (and ends the process) if they hang -- a rank that dropped out of a collective would otherwise block the others for good.
    if not args.no_extra:
        done = threading.Event()

        def watchdog():
            if not done.wait(args.extra_timeout):
                if rank == 0:
                    line['extra'] = {'error': 'the extra legs did not finish within %d s' % args.extra_timeout}
                    print(json.dumps(line), flush=True)
                os._exit(0)
        threading.Thread(target=watchdog, daemon=True).start()
        del pool
        torch.cuda.empty_cache()
        try:
            extra = measure_extra(device, rank, world, barrier)
        except Exception as exc:                          # noqa: BLE001
            extra = {'error': '%s: %s' % (type(exc).__name__, exc)}
            if world > 1:                                 # the other ranks may sit in a collective: end everybody, headline first
                if rank == 0:
                    line['extra'] = extra
                    print(json.dumps(line), flush=True)
                time.sleep(1.0)
                os._exit(0)
        done.set()
        if rank == 0:
            line['extra'] = extra
    if rank == 0:
        print(json.dumps(line), flush=True)
    if parity is not None and not (parity['thr_equal'] and parity['plbl_sha_equal']):
        raise SystemExit('sharded run differs from the single-rank replay: %r' % (parity,))
    if world > 1:
        dist.destroy_process_group()


def sharded_parity(engine, pool, rank, world, device, run_job, windows_per_rank=2):
    """After the timed job: the same global job, `windows_per_rank` windows per rank, once sharded (every rank, NCCL token) and
    once replayed on rank 0 alone; thresholds per group and SHA-256 of every window's pseudo-labels must agree."""
    import hashlib
    import torch
    import torch.distributed as dist
    from hiast_b200.sharded import ShardedIAS

    def job(drv):
        got = {}

        def on_window(w, plbl, counts, thr_groups):
            got[w] = (thr_groups.clone(), plbl.clone())
        engine.thr_state.fill_(0.9)
        engine.mean_state.zero_()
        thr, mean, statics = drv.run(lambda w: pool, on_window)
        torch.cuda.synchronize()
        out = {}
        for w, (tg, pl) in got.items():
            out[w] = (tg.cpu().numpy().tobytes(), hashlib.sha256(pl.cpu().numpy().tobytes()).hexdigest())
        return out, thr.cpu().numpy().tobytes(), mean.cpu().numpy().tobytes(), statics.cpu().numpy().tobytes()

    n_windows = windows_per_rank * world
    mine = job(ShardedIAS(engine, WINDOW, n_windows * WINDOW, rank, world))
    parts = [None] * world
    dist.all_gather_object(parts, mine)
    res = None
    if rank == 0:
        ref_windows, ref_thr, ref_mean, ref_statics = job(ShardedIAS(engine, WINDOW, n_windows * WINDOW, 0, 1))
        merged = {}
        for p in parts:
            merged.update(p[0])
        res = {'windows': n_windows, 'images': n_windows * WINDOW,
               'thr_equal': sorted(merged) == sorted(ref_windows) and all(merged[w][0] == ref_windows[w][0] for w in ref_windows)
               and all(p[1] == ref_thr for p in parts),
               'plbl_sha_equal': all(merged.get(w, (None, None))[1] == ref_windows[w][1] for w in ref_windows),
               'mean_probs_equal': all(p[2] == ref_mean for p in parts),
               'statics_equal': all(p[3] == ref_statics for p in parts),
               'how': 'same global job sharded over %d ranks vs replayed on rank 0 alone, after the timed region' % world}
    dist.barrier()
    engine.thr_state.fill_(0.9)
    engine.mean_state.zero_()
    return res


def h2d_ceiling(host, device, world, barrier, seconds=0.25):
    """Concurrent pinned host-to-device copy bandwidth of all ranks (GB/s, whole job): what the full-resolution e2e leg can
    reach at best on this host."""
    import torch
    import torch.distributed as dist
    dst = torch.empty_like(host, device=device)
    dst.copy_(host, non_blocking=True)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 0
    t_end = time.perf_counter() + seconds
    e0.record()
    while time.perf_counter() < t_end:
        dst.copy_(host, non_blocking=True)
        n += 1
        if n % 4 == 0:
            torch.cuda.current_stream().synchronize()
    e1.record()
    torch.cuda.synchronize()
    gbs = n * host.numel() * host.element_size() / (e0.elapsed_time(e1) / 1e3) / 1e9
    t = torch.tensor([gbs], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t)
    barrier()
    return float(t[0])


def measure_extra(device, rank, world, barrier):
    """The other BASELINE.json configs (tools/bench_extra.py).  Kernel legs on rank 0 at N = 1 only; the sharded jobs
    (configs[3], configs[4]) on every rank.  A leg that fails is reported as its error, the headline stands."""
    import traceback
    from tools import bench_extra as bx
    peak, _ = peaks()
    out = {}

    def leg(name, fn, *a, **k):
        try:
            out[name] = fn(*a, **k)
        except Exception as exc:                          # noqa: BLE001
            out[name] = {'error': '%s: %s' % (type(exc).__name__, exc), 'trace': traceback.format_exc()[-600:]}
            if world > 1:
                raise                                     # a rank that drops out of a collective would hang the others

    if world == 1:
        leg('configs2_loss_fwd_bwd', bx.loss_leg, device, peak)
        leg('copy_paste_kernel', bx.copy_paste_leg, device, peak)
        leg('confusion_kernel', bx.confusion_leg, device, peak)
        leg('configs1_by_distribution', bx.distributions_leg, device, peak, make_pool,
            int(os.environ.get('HIAST_RESERVE_SMS', RESERVE_SMS)))
    leg('configs3_synthia_ias_copy_paste', bx.synthia_leg, device, rank, world, barrier)
    leg('configs4_full_round', bx.full_round_leg, device, rank, world, barrier)
    return out


def measure_e2e(args, device, rank, world, barrier):
    """IASPseudoGenerator.run() on pinned host logits: H2D per batch + D2H per label map inside the timed region."""
    import torch
    from types import SimpleNamespace
    from hiast_b200.pseudo_label_generator import IASPseudoGenerator, ShardedIASPseudoGenerator
    import shutil
    import tempfile
    import torch.distributed as dist

    n_host = 8                                     # pinned host pool: 8 maps = 1.3 GB
    steps = max(1, args.e2e_steps)
    host = torch.empty((n_host, C, H, W), dtype=torch.float32).pin_memory()
    host.copy_(synth_logits_cpu(2).repeat(n_host // 2, 1, 1, 1))
    ceiling_gbs = h2d_ceiling(host, device, world, barrier)

    class Identity:
        def eval(self):
            return self

        def __call__(self, x):
            return {'logits': x}

    def loader(n_images):
        for i in range(0, n_images, GROUP):
            j = i % n_host
            yield {'images': host[j:j + GROUP], 'image_paths': ['img_%06d.png' % (i + k) for k in range(GROUP)]}

    cfg = SimpleNamespace(
        dataset=SimpleNamespace(num_classes=C),
        pseudo_policy=SimpleNamespace(type='IAS', batch_size=GROUP, ias=SimpleNamespace(alpha=ALPHA, beta=BETA, gamma=GAMMA)),
        preprocessor=SimpleNamespace(copy_paste=SimpleNamespace(gamma=CP_GAMMA)))

    # N > 1: the same call through PSEUDO_POLICY['IAS_SHARDED'] -- every rank feeds its own windows of one global job of
    # n_images * world images, the threshold state travels rank to rank (NCCL) exactly as in the device-resident run.
    Base = ShardedIASPseudoGenerator if world > 1 else IASPseudoGenerator

    def total(n_images):
        return n_images * world if world > 1 else None

    class Gen(Base):
        def save_pseudo_label(self, plbl, img_path):      # PNG encode/write excluded, as in the CPU baseline
            self.last = plbl

        def save_data(self):
            pass

    def run(n_images):
        gen = Gen(cfg, model=Identity(), loader=loader(n_images), dataset_len=total(n_images), save_dir=tempfile.mkdtemp(),
                  window_batches=WINDOW // GROUP, device=device)
        gen.run()
        return gen

    # PCIe throughput on these VMs varies from run to run on the same box (the same call measured 341 and 259 images/s
    # back to back): every e2e leg is timed REPS times, the fastest repetition is reported, all are listed.
    REPS = 2

    def timed(fn):
        """Seconds of fn() as the slowest rank sees them: max(wall clock, CUDA events), max over ranks."""
        barrier()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        secs = max(time.perf_counter() - t0, e0.elapsed_time(e1) / 1e3)
        if world > 1:
            t = torch.tensor([secs], dtype=torch.float64, device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            secs = float(t[0])
        return secs

    run(WINDOW)                                    # warm-up
    reps_full = [timed(lambda: run(steps * WINDOW)) for _ in range(REPS)]
    secs = min(reps_full)
    res = {'value': steps * WINDOW * world / secs, 'unit': UNIT, 'steps': steps,
           'repetitions_images_per_s': [steps * WINDOW * world / t for t in reps_full], 'reported': 'fastest of %d repetitions' % REPS,
           'h2d_ceiling_gbs': ceiling_gbs, 'h2d_achieved_gbs': steps * WINDOW * world * C * H * W * 4 / secs / 1e9,
           'h2d_frac_of_ceiling': steps * WINDOW * world * C * H * W * 4 / secs / 1e9 / ceiling_gbs,
           'h2d_bytes_per_step': WINDOW * C * H * W * 4,
           'd2h_bytes_per_step': WINDOW * (H * W + C * 8) + (WINDOW // GROUP) * C * 8,
           'api': ("PSEUDO_POLICY['IAS'](cfg).run() on pinned host logits, identity model, PNG write stubbed" if world == 1 else
                   "PSEUDO_POLICY['IAS_SHARDED'](cfg).run() on pinned host logits (one global job, windows striped over the "
                   "ranks, NCCL threshold hand-off), identity model, PNG write stubbed")}

    # Same call fed with what the network actually produces: stride-8 logits [19,129,257] (2.5 MB per image instead of
    # 159 MB); the bilinear up-sampling of self_training_segmentor.py:27 is fused into phase A.
    # Reported next to the headline e2e, not instead of it: the metric's input is the full-resolution logit map.
    h_lr, w_lr = H // 8 + 1, W // 8 + 1
    host_lr = torch.empty((n_host, C, h_lr, w_lr), dtype=torch.float32).pin_memory()
    # What a segmentation network emits at stride 8 is spatially coherent (objects span tens to hundreds of pixels): a random
    # field with a correlation length of 8 stride-8 cells (64 pixels) plus per-cell noise.  Independent noise per stride-8 cell
    # -- a label change every 8 pixels in both directions, files of 330 KB where Cityscapes pseudo-labels have 20-60 KB -- is kept
    # as the worst case and timed on a short sample beside it (`white_noise_worst_case`).
    gen_lr = torch.Generator().manual_seed(5)
    white_lr = torch.randn(n_host, C, h_lr, w_lr, generator=gen_lr) * 4
    coarse = torch.randn(n_host, C, h_lr // 8 + 1, w_lr // 8 + 1, generator=gen_lr) * 4
    smooth_lr = torch.nn.functional.interpolate(coarse, size=(h_lr, w_lr), mode='bilinear', align_corners=True)
    smooth_lr += torch.randn(n_host, C, h_lr, w_lr, generator=gen_lr) * 0.5
    host_lr.copy_(smooth_lr)
    LR_KIND = 'random field, correlation length 64 px (8 stride-8 cells), + N(0, 0.5) per cell'

    class LowRes(Identity):
        def __call__(self, x):
            return {'logits_lr': x, 'size': (H, W)}

    def loader_lr(n_images):
        for i in range(0, n_images, GROUP):
            j = i % n_host
            yield {'images': host_lr[j:j + GROUP], 'image_paths': ['img_%06d.png' % (i + k) for k in range(GROUP)]}

    def run_lr(n_images):
        gen = Gen(cfg, model=LowRes(), loader=loader_lr(n_images), dataset_len=total(n_images), save_dir=tempfile.mkdtemp(),
                  window_batches=WINDOW // GROUP, device=device)
        gen.run()
        return gen

    steps_lr = max(4 * steps, 46)             # 46 windows = 2944 images, the size of the Cityscapes train set (2975)
    run_lr(WINDOW)
    reps_lr = [timed(lambda: run_lr(steps_lr * WINDOW)) for _ in range(REPS)]
    secs = min(reps_lr)
    lr_gbs = steps_lr * WINDOW * world * C * h_lr * w_lr * 4 / secs / 1e9
    res['from_stride8_logits'] = {'value': steps_lr * WINDOW * world / secs, 'unit': UNIT, 'steps': steps_lr,
                                  'repetitions_images_per_s': [steps_lr * WINDOW * world / t for t in reps_lr],
                                  'h2d_achieved_gbs': lr_gbs, 'h2d_frac_of_ceiling': lr_gbs / ceiling_gbs,
                                  'h2d_bytes_per_step': WINDOW * C * h_lr * w_lr * 4,
                                  'd2h_bytes_per_step': res['d2h_bytes_per_step'],
                                  'logits_lr': LR_KIND,
                                  'api': "same call, model returns {'logits_lr': [B,19,129,257]}: fused up-sampling + IAS"}

    # ... and with the on-disk output of the reference (pseudo_label_generator.py:43-46) INCLUDED: the label maps are
    # encoded as PNG files on the device (hiast_png_encode), only the files cross PCIe, a thread pool writes them.
    class GenPng(Base):
        def save_data(self):
            pass

    class GenPngHost(IASPseudoGenerator):               # rank 0 alone, for the reference writer beside it
        def save_data(self):
            pass

    # Where the files go.  The container's /tmp is an ext4 image on a throttled virtual disk (tools/fs_probe.py on the B200 boxes:
    # 5-10 k files/s from one process, erratic; 3.8 k images/s end to end once the page cache is dirty) -- that would time the VM's
    # disk, not this pipeline.  tmpfs (/dev/shm: 60 k files/s from one process) stands in for a local NVMe scratch directory; the
    # same call on the container disk is reported beside it on a short sample.
    def files_root(kind):
        if kind == 'fast' and os.path.isdir('/dev/shm') and os.access('/dev/shm', os.W_OK):
            try:
                st = os.statvfs('/dev/shm')
                if st.f_bavail * st.f_frsize > 4 << 30:
                    return '/dev/shm'
            except OSError:
                pass
        return tempfile.gettempdir()

    last_trace = {}

    def run_png(n_images, mode, where='fast'):
        """Seconds of construct + run() (files on disk when it returns), number of files, bytes on disk."""
        d = tempfile.mkdtemp(dir=files_root(where))
        try:
            cls = GenPng if mode == 'device' else GenPngHost
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            gen = cls(cfg, model=LowRes(), loader=loader_lr(n_images), save_dir=os.path.join(d, 'pl'),
                      dataset_len=total(n_images) if mode == 'device' else None,
                      window_batches=WINDOW // GROUP, device=device, png=mode)
            gen.run()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            last_trace.clear()
            last_trace.update(getattr(gen, 'pipeline_trace', {}))
            files = os.listdir(os.path.join(d, 'pl'))
            return dt, len(files), sum(os.path.getsize(os.path.join(d, 'pl', f)) for f in files)
        finally:
            shutil.rmtree(d, ignore_errors=True)

    run_png(WINDOW, 'device')
    reps_png = []
    for _ in range(REPS):
        barrier()
        secs_k, n_files, n_bytes = run_png(steps_lr * WINDOW, 'device')
        if world > 1:
            t = torch.tensor([secs_k], dtype=torch.float64, device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            secs_k = float(t[0])
        reps_png.append(secs_k)
    secs = min(reps_png)
    assert n_files == steps_lr * WINDOW
    tr = dict(last_trace)
    host_trace = {'rank0_seconds': {k: round(v, 4) for k, v in tr.items() if k in ('wait_completion', 'close', 'total', 'collect')},
                  'note': 'host time of rank 0 inside run(): blocked on a window\'s completion / closing windows (launches, token '
                          'calls, emit) / total between the first close and the end'}
    png = {'value': steps_lr * WINDOW * world / secs, 'unit': UNIT, 'steps': steps_lr, 'files_written': n_files,
           'repetitions_images_per_s': [steps_lr * WINDOW * world / t for t in reps_png],
           'h2d_frac_of_ceiling': steps_lr * WINDOW * world * C * h_lr * w_lr * 4 / secs / 1e9 / ceiling_gbs,
           'mean_file_bytes': n_bytes / max(1, n_files), 'd2h_bytes_per_step': n_bytes / steps_lr + WINDOW * C * 8,
           'files_dir': files_root('fast'), 'host_trace': host_trace,
           'api': 'same call with the PNG files written (device encoder, native writer pool with %d POSIX writers per rank, completion deferred by three windows)' % IASPseudoGenerator._default_workers()}
    secs_disk = run_png(8 * WINDOW, 'device', where='disk')[0]
    if world > 1:
        t = torch.tensor([secs_disk], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        secs_disk = float(t[0])
    png['on_container_disk'] = {'value': 8 * WINDOW * world / secs_disk, 'steps': 8, 'files_dir': files_root('disk'),
                                'note': 'ext4 image on the VM\'s virtual disk: file-system bound and erratic (tools/fs_probe.py)'}
    if rank == 0:                                   # the reference's writer (cv2.imwrite on host label maps) beside it
        n_host_png = WINDOW
        png['host_cv2_imwrite_images_per_s'] = n_host_png / run_png(n_host_png, 'host')[0]
    barrier()
    # worst case for the encoder, the device-to-host copies and the writers: independent noise per stride-8 cell
    host_lr.copy_(white_lr)
    steps_wn = 16
    run_png(WINDOW, 'device')
    barrier()
    secs_wn, n_files_wn, n_bytes_wn = run_png(steps_wn * WINDOW, 'device')
    if world > 1:
        t = torch.tensor([secs_wn], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        secs_wn = float(t[0])
    png['white_noise_worst_case'] = {'value': steps_wn * WINDOW * world / secs_wn, 'steps': steps_wn,
                                     'mean_file_bytes': n_bytes_wn / max(1, n_files_wn),
                                     'logits_lr': 'independent N(0, 16) per stride-8 cell: a label change every 8 pixels'}
    barrier()
    res['from_stride8_logits']['with_png_files'] = png
    return res


def main():
    args = parse_args()
    if args.impl == 'reference':
        reference_arm(args)
    else:
        gpu_arm(args)


if __name__ == '__main__':
    main()
