"""Multi-GPU IAS pseudo-labelling: windows striped over the ranks, threshold state handed rank to rank.

The reference pseudo-labels on one GPU in one process (no ``dist`` call in
``workflows/pseudo_label_generator.py``).  The path shards naturally: phase A
(softmax / arg-max / histograms) and phase C (mask / counts) are independent per group; the only
cross-group dependencies are the sequential ``class_threshold`` f64[C] (:207-209) and, after phase C,
``class_mean_probs`` f64[C] (:100-105).  With R ranks (one process per GPU) the pinned global order of
image groups is cut into windows of ``window_images`` images and window w belongs to rank w mod R:

    rank r, local window j  (global window w = j*R + r):
        phase A of window j+1            <- issued first, so the GPU has work while the token travels
        recv  thr f64[C] from rank r-1   (ncclRecv over NVLink; skipped for w = 0, which starts from 0.9)
        phase B of window j              (device scan over the window's groups)
        send  thr f64[C] to rank r+1     (ncclSend; skipped for the last window)
        phase C of window j

The 152-byte token goes round the ring once per round of R windows; as long as R x (scan + hop) is
shorter than one window of phase A + C the chain is hidden behind the bandwidth-bound work, which is
what makes image/s scale with R.  No other data-path collective exists.  At the end the per-group
(confidence sum, count)[C] of all windows are all-gathered (a few hundred KB) and every rank replays
the tiny ``class_mean_probs`` EMA in global order; the final thresholds are broadcast from the rank
that scanned the last window; ``statics_class`` falls out of the gathered counts.

Results are bit-identical to a single-rank run over the same global order (tests/test_sharded_gpu.py;
on CPU with gloo and a host stand-in engine in tests/test_sharded_gloo.py).  With R = 1 the same code
is the single-GPU windowed pipeline.
"""

from __future__ import annotations

import torch
import torch.distributed as dist


def window_owner(w, world_size):
    return w % world_size


def local_windows(n_windows_total, rank, world_size):
    """Global indices of the windows rank ``rank`` owns, in processing order."""
    return list(range(rank, n_windows_total, world_size))


def window_images(w, window_size, n_images_total):
    """(first global image, image count) of global window w."""
    i0 = w * window_size
    return i0, max(0, min(window_size, n_images_total - i0))


class ShardedIAS:
    """Drives one rank's engine (an ``IASEngine`` with room for two windows, or any object with the same phase
    methods and state tensors -- the CPU tests use a host stand-in)."""

    def __init__(self, engine, window_size, n_images_total, rank=None, world_size=None, process_group=None):
        if window_size % engine.B:
            raise ValueError('window_size must be a multiple of the group (batch) size')
        if engine.max_images < 2 * window_size:
            raise ValueError('the engine needs room for two windows (double buffering)')
        self.engine = engine
        self.pg = process_group
        use_dist = dist.is_available() and dist.is_initialized()
        self.rank = (dist.get_rank(process_group) if use_dist else 0) if rank is None else rank
        self.world = (dist.get_world_size(process_group) if use_dist else 1) if world_size is None else world_size
        self.window_size = int(window_size)
        self.n_total = int(n_images_total)
        self.n_windows_total = (self.n_total + self.window_size - 1) // self.window_size
        self.my_windows = local_windows(self.n_windows_total, self.rank, self.world)
        self._stash = []

    def _global(self, group_rank):
        return dist.get_global_rank(self.pg, group_rank) if self.pg is not None else group_rank

    def _slot(self, j):
        return (j % 2) * self.window_size

    def run(self, window_logits, on_window=None):
        """``window_logits(w) -> logits f32 [n_w,C,H,W]`` on the device for GLOBAL window index w (called once per
        owned window, one window ahead of its use).  ``on_window(w, plbl, counts, thr_groups)`` receives
        device views of window w's results right after phase C is enqueued; they stay valid until the window
        after next is started.  Returns (class_threshold, class_mean_probs, statics_class) device tensors."""
        e = self.engine
        wins = self.my_windows
        self._stash = []
        # One rank: nothing to wait for, so a window goes through the fused persistent kernel (A + B + C in one launch,
        # the conf / label spill never leaves L2).  With R > 1 the thresholds of a window arrive from another rank long
        # after its phase A has run, so the three phases stay separate kernels.
        fused = self.world == 1 and bool(getattr(e, 'fused', False)) and hasattr(e, 'process_fused')
        if wins and not fused:
            self._phase_a(wins[0], 0, window_logits)
        for j, w in enumerate(wins):
            _, n = window_images(w, self.window_size, self.n_total)
            slot = self._slot(j)
            if fused:
                logits = window_logits(w)
                if logits.shape[0] != n:
                    raise ValueError('window %d must hold %d images, got %d' % (w, n, logits.shape[0]))
                if not e.process_fused(logits, slot):
                    e.phase_a(logits, slot)
                    e.phase_b(slot, n)
                    e.phase_c(slot, n)
            else:
                if j + 1 < len(wins):
                    self._phase_a(wins[j + 1], j + 1, window_logits)
                if w > 0 and self.world > 1:
                    dist.recv(e.thr_state, src=self._global((self.rank - 1) % self.world), group=self.pg)
                e.phase_b(slot, n)
                if w < self.n_windows_total - 1 and self.world > 1:
                    dist.send(e.thr_state, dst=self._global((self.rank + 1) % self.world), group=self.pg)
                e.phase_c(slot, n)
            g0, g = slot // e.B, (n + e.B - 1) // e.B
            self._stash.append(torch.stack([e.confsum[g0:g0 + g], e.group_counts(slot, n)], dim=1).clone())
            if on_window is not None:
                on_window(w, e.plbl[slot:slot + n], e.counts[slot:slot + n], e.thr_groups[g0:g0 + g])
        return self.finish_state()

    def _phase_a(self, w, j, window_logits):
        logits = window_logits(w)
        _, n = window_images(w, self.window_size, self.n_total)
        if logits.shape[0] != n:
            raise ValueError('window %d must hold %d images, got %d' % (w, n, logits.shape[0]))
        self.engine.phase_a(logits, self._slot(j))

    def finish_state(self):
        """All-gather the per-group sums, replay the mean-prob EMA over all groups in global order, broadcast the
        final thresholds."""
        e = self.engine
        C, B = e.C, e.B
        gw = self.window_size // B
        kmax = (self.n_windows_total + self.world - 1) // self.world
        dev = e.thr_state.device
        packed = torch.zeros((max(kmax, 1), gw, 2, C), dtype=torch.int64, device=dev)
        for j, s in enumerate(self._stash):
            packed[j, :s.shape[0]] = s
        if self.world > 1:
            gathered = [torch.empty_like(packed) for _ in range(self.world)]
            dist.all_gather(gathered, packed, group=self.pg)
        else:
            gathered = [packed]
        rows = []
        for w in range(self.n_windows_total):
            _, n = window_images(w, self.window_size, self.n_total)
            rows.append(gathered[w % self.world][w // self.world, :(n + B - 1) // B])
        if rows:
            allg = torch.cat(rows)
            confsum, counts = allg[:, 0].contiguous(), allg[:, 1].contiguous()
            e.mean_prob_from_groups(confsum, counts)
            statics = counts.sum(dim=0)
        else:
            statics = torch.zeros(C, dtype=torch.int64, device=dev)
        if self.world > 1 and self.n_windows_total > 0:
            last_owner = window_owner(self.n_windows_total - 1, self.world)
            dist.broadcast(e.thr_state, src=self._global(last_owner), group=self.pg)
        return e.thr_state, e.mean_state, statics
