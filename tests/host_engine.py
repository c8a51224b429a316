"""A CPU stand-in for hiast_b200.ias_engine.IASEngine, for tests of the HOST-side orchestration only
(world_size > 1 with gloo, where no GPU exists).  It lives under tests/ on purpose: the product has no
CPU path.  Arithmetic: torch CPU softmax (oracle.ias.softmax_max), the oracle's key histogram, and the
library's host test hook for the threshold step (the same header the device scan compiles)."""

import ctypes as C

import numpy as np
import torch

from oracle import ias as oias


class HostEngine:
    def __init__(self, num_classes, height, width, group_size, alpha, beta, gamma, cp_gamma, max_images):
        from hiast_b200 import _lib, build
        build.build()
        self.L = _lib.lib()
        self.C, self.H, self.W, self.B = num_classes, height, width, group_size
        self.alpha, self.beta, self.gamma, self.cp_gamma = alpha, beta, gamma, cp_gamma
        self.max_images = max_images
        g = max_images // group_size
        self.key_lo = 0
        self.conf = np.zeros((max_images, height, width), dtype=np.float32)
        self.label = np.zeros((max_images, height, width), dtype=np.int64)
        self.plbl = np.zeros((max_images, height, width), dtype=np.uint8)
        self.hist = [None] * g
        self.thr_groups = torch.zeros((g, num_classes), dtype=torch.float64)
        self.counts = torch.zeros((max_images, num_classes), dtype=torch.int64)
        self.confsum = torch.zeros((g, num_classes), dtype=torch.int64)
        self.thr_state = torch.full((num_classes,), 0.9, dtype=torch.float64)
        self.mean_state = torch.zeros(num_classes, dtype=torch.float64)

    def _groups(self, n):
        return (n + self.B - 1) // self.B

    def phase_a(self, logits, first_image=0):
        n = logits.shape[0]
        conf, label = oias.softmax_max(logits)
        self.conf[first_image:first_image + n] = conf
        self.label[first_image:first_image + n] = label
        for k in range(self._groups(n)):
            sl = slice(first_image + k * self.B, min(first_image + n, first_image + (k + 1) * self.B))
            self.hist[first_image // self.B + k] = oias.class_key_histogram(self.conf[sl], self.label[sl], self.C, 0)

    def phase_b(self, first_image, n_images):
        g0 = first_image // self.B
        for g in range(g0, g0 + self._groups(n_images)):
            for c in range(self.C):
                prefix = np.cumsum(self.hist[g][c]).astype(np.uint32)
                temp, err = C.c_float(), C.c_int()
                new = self.L.hiast_testhook_threshold_step(prefix.ctypes.data_as(C.c_void_p), 0,
                                                           float(self.thr_state[c]), self.alpha, self.beta, self.gamma,
                                                           C.byref(temp), C.byref(err))
                self.thr_state[c] = new
            self.thr_groups[g] = self.thr_state

    def phase_c(self, first_image, n_images):
        g0 = first_image // self.B
        self.counts[first_image:first_image + n_images] = 0
        self.confsum[g0:g0 + self._groups(n_images)] = 0
        for i in range(first_image, first_image + n_images):
            g = i // self.B
            thr = self.thr_groups[g].numpy()
            plbl = oias.select_confident(self.conf[i], self.label[i], thr)
            self.plbl[i] = plbl.astype(np.uint8)
            for c in range(self.C):
                kept = plbl == c
                self.counts[i, c] = int(kept.sum())
                fx = (self.conf[i][kept].astype(np.float64) * 4294967296.0).astype(np.int64)
                self.confsum[g, c] += int(fx.sum())

    def group_counts(self, first_image, n_images):
        g = self._groups(n_images)
        out = torch.zeros((g, self.C), dtype=torch.int64)
        for i in range(n_images):
            out[i // self.B] += self.counts[first_image + i]
        return out

    def mean_prob_from_groups(self, confsum, group_counts):
        omg = np.float32(1.0 - self.cp_gamma)
        for g in range(confsum.shape[0]):
            for c in range(self.C):
                n = int(group_counts[g, c])
                if n == 0:
                    continue
                m = np.float32(float(confsum[g, c]) * 2.0 ** -32 / n)
                cur = float(self.mean_state[c])
                self.mean_state[c] = float(m) if cur == 0.0 else cur * self.cp_gamma + float(np.float32(m * omg))

    def mean_prob(self, first_image, n_images):
        g0 = first_image // self.B
        g = self._groups(n_images)
        self.mean_prob_from_groups(self.confsum[g0:g0 + g], self.group_counts(first_image, n_images))
