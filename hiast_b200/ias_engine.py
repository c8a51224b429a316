"""Device-resident IAS pipeline: buffers, state and the three kernel phases for a window of images.

This is the engine under ``IASPseudoGenerator`` (pseudo_label_generator.py in this package) and under
the sharded multi-GPU driver.  It restates the loop body of the reference's
``IASPseudoGenerator.run`` (``workflows/pseudo_label_generator.py:181-213``, /root/reference/code) as

    phase A  logits -> conf, label, key histograms         (parallel over images)
    phase B  histograms -> thresholds, group by group       (the only serial chain; 19 CTAs)
    phase C  conf, label, thresholds -> pseudo-labels, counts, confidence sums
    mean-prob EMA over groups

for a *window* of up to ``max_images`` images whose groups (the reference's DataLoader batches) are
processed in order.  All state lives on the device (``thr_state`` f64[C] starts at 0.9, :185;
``mean_state`` f64[C] at 0, :21); nothing here synchronises with the host.
"""

from __future__ import annotations

import torch

from . import ops


class IASEngine:
    def __init__(self, num_classes, height, width, group_size, alpha, beta, gamma, cp_gamma,
                 max_images, device='cuda', key_lo=None, hist_mode=0, fused=False):
        if max_images % group_size:
            raise ValueError('max_images must be a multiple of the group (batch) size')
        self.C, self.H, self.W, self.B = int(num_classes), int(height), int(width), int(group_size)
        self.alpha, self.beta, self.gamma, self.cp_gamma = float(alpha), float(beta), float(gamma), float(cp_gamma)
        self.device = torch.device(device)
        self.max_images = int(max_images)
        self.max_groups = self.max_images // self.B
        self.key_lo = ops.ias_key_lo(self.C) if key_lo is None else int(key_lo)
        self.hist_mode = int(hist_mode)
        dev = self.device
        n, g, c = self.max_images, self.max_groups, self.C
        self.conf = torch.empty((n, self.H, self.W), dtype=torch.float32, device=dev)
        self.label = torch.empty((n, self.H, self.W), dtype=torch.uint8, device=dev)
        self.plbl = torch.empty((n, self.H, self.W), dtype=torch.uint8, device=dev)
        self.hist = ops.ias_new_hist(g, c, self.key_lo, dev)
        self.thr_groups = torch.empty((g, c), dtype=torch.float64, device=dev)
        self.temp_groups = torch.empty((g, c), dtype=torch.float32, device=dev)
        self.counts = torch.zeros((n, c), dtype=torch.int64, device=dev)
        self.confsum = torch.zeros((g, c), dtype=torch.int64, device=dev)
        self.thr_state = torch.full((c,), 0.9, dtype=torch.float64, device=dev)      # :185
        self.mean_state = torch.zeros(c, dtype=torch.float64, device=dev)            # :21
        self.error_flag = torch.zeros(1, dtype=torch.int32, device=dev)
        # fused=True sends single-GPU windows through the persistent kernel hiast_ias_fused_window (A + B + C in one
        # launch, the conf / label spill stays in L2).  It is bit-identical to the three kernels but measured SLOWER
        # on B200 (2.9 ms against 2.1 ms per 64-image window: every hand-over between units costs a 4-5 us round trip
        # through a memory system that phase A keeps saturated), so it is off by default.
        self.fused = bool(fused)
        self.fused_ws = ops.ias_fused_workspace(n, self.B, dev)
        self.groups_in_flight = 0
        self.keep_spill = False
        # SMs phase A leaves free (its CTAs take whole SMs): the sharded driver runs the threshold chain and phase C of the
        # window before on them, CONCURRENTLY with phase A of the next window (sharded.ShardedIAS, "concurrent" schedule)
        self.reserve_sms = 0

    # ------------------------------------------------------------------ phases
    def _groups(self, n_images):
        return (n_images + self.B - 1) // self.B

    def phase_a(self, logits, first_image=0):
        """logits f32 [n,C,H,W] -> conf/label/hist slots [first_image, first_image+n)."""
        n = logits.shape[0]
        if first_image % self.B:
            raise ValueError('a window must start on a group boundary')
        if first_image + n > self.max_images:
            raise ValueError('window overflow: %d + %d > %d' % (first_image, n, self.max_images))
        g0 = first_image // self.B
        ops.ias_softmax_hist(logits, self.B, self.key_lo, self.conf[first_image:first_image + n],
                             self.label[first_image:first_image + n], self.hist[g0:g0 + self._groups(n)],
                             accumulate=False, hist_mode=self.hist_mode | (int(self.reserve_sms) << 8))

    def phase_a_lowres(self, logits_lr, first_image=0):
        """Phase A from the network's LOW-RESOLUTION logits f32 [n,C,h,w]: the bilinear up-sampling to (H,W) of
        self_training_segmentor.py:27 is fused into the kernel (bit-identical to interpolate -> softmax -> max)."""
        n = logits_lr.shape[0]
        if first_image % self.B or first_image + n > self.max_images:
            raise ValueError('bad window')
        g0 = first_image // self.B
        try:
            ops.ias_upsample_softmax_hist(logits_lr, (self.H, self.W), self.B, self.key_lo,
                                          self.conf[first_image:first_image + n], self.label[first_image:first_image + n],
                                          self.hist[g0:g0 + self._groups(n)], accumulate=False)
        except ops._lib.HiastError as err:
            if err.status != ops.UNSUPPORTED:
                raise
            # a shape the fused kernel does not cover (C not in {16, 19}, W % 4 != 0, down-sampling): the reference's own
            # F.interpolate (self_training_segmentor.py:27), group by group to bound the full-resolution scratch, then the
            # generic phase A -- same results, no fusion
            for k in range(0, n, self.B):
                full = torch.nn.functional.interpolate(logits_lr[k:k + self.B], size=(self.H, self.W), mode='bilinear',
                                                       align_corners=True)
                self.phase_a(full, first_image + k)

    def phase_a_from_conf(self, conf, label, first_image=0):
        """Same, from caller-provided conf f32 [n,H,W] / label (u8|i64) [n,H,W]; the engine's key range must
        cover the confidences (key_lo=0 covers every value in [0,1])."""
        n = conf.shape[0]
        if first_image % self.B or first_image + n > self.max_images:
            raise ValueError('bad window')
        g0 = first_image // self.B
        sl = slice(first_image, first_image + n)
        self.conf[sl].copy_(conf.reshape(n, self.H, self.W))
        ops.ias_conf_hist(self.conf[sl], label.reshape(n, self.H, self.W).contiguous(), self.C, self.B, self.key_lo,
                          self.hist[g0:g0 + self._groups(n)], accumulate=False, label_u8_out=self.label[sl])

    def phase_b(self, first_image, n_images, token=None):
        """``token``: (mailbox_in, in_seq, mailbox_out, out_seq) of the multi-GPU token ring (sharded.TokenRing), fused into
        the scan kernel; None on one GPU or when the state travels through torch.distributed send / recv."""
        g0, g = first_image // self.B, self._groups(n_images)
        ops.ias_threshold_scan(self.hist[g0:g0 + g], g, self.C, self.key_lo, self.alpha, self.beta, self.gamma,
                               self.thr_state, self.thr_groups[g0:g0 + g], self.temp_groups[g0:g0 + g],
                               self.error_flag, token=token)

    def phase_c(self, first_image, n_images):
        g0, g = first_image // self.B, self._groups(n_images)
        sl = slice(first_image, first_image + n_images)
        self.counts[sl].zero_()
        self.confsum[g0:g0 + g].zero_()
        ops.ias_select(self.conf[sl], self.label[sl], self.thr_groups[g0:g0 + g], self.C, self.B,
                       self.plbl[sl], self.counts[sl], self.confsum[g0:g0 + g])

    def mean_prob(self, first_image, n_images):
        g0, g = first_image // self.B, self._groups(n_images)
        sl = slice(first_image, first_image + n_images)
        ops.ias_meanprob_scan(self.confsum[g0:g0 + g], self.counts[sl], self.B, self.C, self.cp_gamma, self.mean_state)

    def group_counts(self, first_image, n_images):
        """Kept-pixel counts per group i64 [G,C] (sum of the per-image counts of each group) of a window."""
        g = self._groups(n_images)
        pad = g * self.B - n_images
        c = self.counts[first_image:first_image + n_images]
        if pad:
            c = torch.cat([c, torch.zeros((pad, self.C), dtype=c.dtype, device=c.device)])
        return c.view(g, self.B, self.C).sum(dim=1)

    def mean_prob_from_groups(self, confsum, group_counts):
        """EMA over an explicit list of groups (used by the sharded driver after its all-gather)."""
        ops.ias_meanprob_scan(confsum, group_counts, 1, self.C, self.cp_gamma, self.mean_state)

    # ----------------------------------------------------------------- windows
    def process(self, logits, first_image=0):
        """A, B, C and the mean-prob EMA for one window.  Returns views (plbl, counts, thr_groups)."""
        n = logits.shape[0]
        if not (self.fused and self.process_fused(logits, first_image)):
            self.phase_a(logits, first_image)
            self.phase_b(first_image, n)
            self.phase_c(first_image, n)
        self.mean_prob(first_image, n)
        g0 = first_image // self.B
        return (self.plbl[first_image:first_image + n], self.counts[first_image:first_image + n],
                self.thr_groups[g0:g0 + self._groups(n)])

    def process_fused(self, logits, first_image=0):
        """A + B + C of one window in a single persistent kernel.  False (nothing launched) if the shape is not
        covered.  self.conf / self.label hold no defined values afterwards."""
        n = logits.shape[0]
        if first_image % self.B or first_image + n > self.max_images:
            raise ValueError('bad window')
        g0, g = first_image // self.B, self._groups(n)
        sl = slice(first_image, first_image + n)
        return ops.ias_fused_window(logits, self.B, self.key_lo, self.alpha, self.beta, self.gamma, self.conf[sl],
                                    self.label[sl], self.hist[g0:g0 + g], self.thr_state, self.thr_groups[g0:g0 + g],
                                    self.temp_groups[g0:g0 + g], self.plbl[sl], self.counts[sl], self.confsum[g0:g0 + g],
                                    self.error_flag, self.fused_ws, keep_spill=self.keep_spill,
                                    groups_in_flight=self.groups_in_flight)

    def check_errors(self):
        """Host sync.  Mirrors numpy's ValueError for a quantile level outside [0,1] (np.quantile, :178).
        Returns True when every threshold is certified independent of the host libm's last-bit pow rounding."""
        flag = int(self.error_flag.item())
        if flag & 4:
            raise RuntimeError('hiast_ias_fused_window gave up waiting for a group to close (internal error)')
        if flag & 8:
            raise RuntimeError('the threshold token of the multi-GPU ring did not arrive within 10 s (a rank died or fell out of step)')
        if flag & 1:
            raise ValueError('Quantiles must be in the range [0, 1]')
        return not (flag & 2)
