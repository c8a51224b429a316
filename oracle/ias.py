"""Oracle: instance-adaptive selector (IAS) pseudo-labelling.

Restates ``workflows/pseudo_label_generator.py`` (reference, /root/reference/code):
``IASPseudoGenerator.run`` :181-213, ``get_ias_threshold`` :171-179 and
``BasePseudoGenerator.select_and_save_confident_label`` :67-106.

Two execution styles with identical results (checked in tests/test_oracle_golden.py):

* ``faithful=True``  -- the same operation sequence as the reference (Python lists
  of fp16 scalars fed to ``np.quantile``, a per-row Python lookup for the
  threshold map, built-in ``sum`` for the class counts).  This is the one timed as
  the CPU baseline in ``bench.py`` because its cost profile is the reference's.
* ``faithful=False`` -- numpy-vectorised forms of the same arithmetic, used by the
  tests so that parity checks finish in seconds.

A third form, ``threshold_from_hist``, is the histogram restatement that the CUDA
scan kernel implements (fp16 bit patterns of positive numbers order like their
values, so a per-class histogram over the keys + the closed-form lerp reproduces
``np.quantile`` exactly).
"""

from __future__ import annotations

import warnings

import numpy as np
import torch
from torch.nn import functional as F

IGNORE = 255
KEY_ONE = 0x3C00  # fp16 bit pattern of 1.0


# --------------------------------------------------------------------------- a1
def softmax_max(logits: torch.Tensor):
    """conf, label of a logits batch  (pseudo_label_generator.py:192-195).

    Runs on the device the tensor lives on (the reference does this on CUDA and
    copies to host); returns host numpy arrays: conf f32 [B,H,W], label int64.
    """
    probs = F.softmax(logits, dim=1)
    conf, lbl = probs.max(dim=1)
    return conf.cpu().numpy(), lbl.cpu().numpy()


# ------------------------------------------------------------------------ a2-a3
def ias_quantile_thresholds(conf, label, thr, num_classes, alpha, gamma, faithful=False):
    """temp_class_threshold f32[C] for one batch (:198-201 + get_ias_threshold :171-179).

    The sample list of class c is [thr_c] followed by the fp16-rounded confidences
    of every pixel of the WHOLE batch predicted as c; the quantile level is
    1 - alpha * thr_c ** gamma; the result is stored into a float32 array.
    """
    out = np.ones(num_classes, dtype=np.float32)
    for c in range(num_classes):
        q = 1 - alpha * thr[c] ** gamma
        if faithful:
            samples = [thr[c]]
            samples.extend(conf[label == c].astype(np.float16))
            out[c] = np.quantile(samples, q)
        else:
            vals = conf[label == c].astype(np.float16).astype(np.float64)
            out[c] = np.quantile(np.concatenate(([thr[c]], vals)), q)
    return out


# --------------------------------------------------------------------------- a4
def ias_ema_update(thr, temp, beta):
    """class_threshold update + clamp (:207-209).  thr f64[C], temp f32[C]."""
    new = beta * thr + (1 - beta) * temp  # f64*f64 + (f32 product) -> f64, numpy promotion rules
    new[new >= 1] = 0.999
    return new


# --------------------------------------------------------------------------- a5
def select_confident(conf_img, label_img, thr, faithful=False):
    """plbl int64 [H,W]: label where conf >= thr[label] else 255 (:74-78)."""
    if faithful:
        thr_map = np.apply_along_axis(lambda row: [thr[e] for e in row], 1, label_img)
    else:
        thr_map = thr[label_img]
    plbl = label_img.copy()
    plbl[conf_img < thr_map] = IGNORE
    return plbl


# ------------------------------------------------------------------ hist form
def fp16_keys(conf):
    """uint16 bit patterns of fp16_rn(conf)."""
    return np.asarray(conf, dtype=np.float32).astype(np.float16).view(np.uint16)


def class_key_histogram(conf, label, num_classes, key_lo=0):
    """hist uint32 [C, KEY_ONE - key_lo + 1] over fp16 keys of one batch."""
    nb = KEY_ONE - key_lo + 1
    keys = fp16_keys(conf).astype(np.int64).ravel() - key_lo
    lab = np.asarray(label).astype(np.int64).ravel()
    ok = (lab >= 0) & (lab < num_classes)
    assert keys[ok].min(initial=0) >= 0 and keys[ok].max(initial=0) < nb
    flat = np.bincount(lab[ok] * nb + keys[ok], minlength=num_classes * nb)
    return flat.reshape(num_classes, nb).astype(np.uint32)


def _key_value(key):
    return float(np.array([key], dtype=np.uint16).view(np.float16)[0])


def threshold_from_hist(hist_c, key_lo, thr_c, alpha, gamma):
    """One class of get_ias_threshold computed from the key histogram (float32 result).

    Follows numpy 2.x ``_quantile`` / ``_get_indexes`` / ``_lerp`` for
    method='linear' on the n = 1 + m samples {keys..., thr_c}.
    """
    hist_c = np.asarray(hist_c, dtype=np.int64)
    m = int(hist_c.sum())
    n = m + 1
    q = 1 - alpha * thr_c ** gamma
    if not (0.0 <= q <= 1.0):
        raise ValueError('Quantiles must be in the range [0, 1]')
    vi = (n - 1) * q
    lo = np.floor(vi)
    g = vi - lo
    if vi >= n - 1:
        lo_i = hi_i = n - 1
    else:
        lo_i = int(lo)
        hi_i = lo_i + 1
    prefix = np.cumsum(hist_c)  # inclusive
    vals = np.arange(key_lo, key_lo + len(hist_c)).astype(np.uint16).view(np.float16).astype(np.float64)
    r = int(hist_c[vals < thr_c].sum())  # rank of thr_c inside the merged sorted list

    def order_stat(k):
        if k == r:
            return float(thr_c)
        j = k if k < r else k - 1
        b = int(np.searchsorted(prefix, j, side='right'))
        return float(vals[b])

    a = np.float64(order_stat(lo_i))
    b = np.float64(order_stat(hi_i))
    d = b - a
    t = a + d * g
    if g >= 0.5:
        t = b - d * (1 - g)
    return np.float32(t)


# ------------------------------------------------------------------ full state
class IASOracle:
    """State + loop body of IASPseudoGenerator (:14-23, :181-213)."""

    def __init__(self, num_classes=19, alpha=0.2, beta=0.9, gamma=8.0, cp_gamma=0.99,
                 faithful=False, keep_labels=True):
        self.C = num_classes
        self.alpha, self.beta, self.gamma, self.cp_gamma = alpha, beta, gamma, cp_gamma
        self.faithful = faithful
        self.keep_labels = keep_labels
        self.class_threshold = 0.9 * np.ones(num_classes)                 # :185
        self.statics_class = np.array([0] * num_classes)                  # :18
        self.sample_stats = []                                            # :19
        self.samples_class = {i: [] for i in range(num_classes)}          # :20
        self.class_mean_probs = np.zeros(num_classes)                     # :21
        self.threshold_trace = []   # thr after each batch (f64[C])
        self.temp_trace = []        # quantile results per batch (f32[C])
        self.labels = []            # captured pseudo-labels (uint8 [H,W]) in order

    # :67-106
    def select_and_save(self, conf, label, paths):
        kept = []
        for conf_img, label_img, path in zip(conf, label, paths):
            plbl = select_confident(conf_img, label_img, self.class_threshold, self.faithful)
            stats = {}
            for i in range(self.C):
                if self.faithful:
                    n_i = int(sum(sum((plbl == i))))
                else:
                    n_i = int(np.count_nonzero(plbl == i))
                if n_i != 0:
                    stats[i] = n_i
                    self.samples_class[i].append([path, n_i])
                    self.statics_class[i] += n_i
            stats['file'] = path
            self.sample_stats.append(stats)
            if self.keep_labels:
                self.labels.append(plbl.astype(np.uint8))                  # :46 (png payload)
            kept.append(plbl[None])
        kept = np.concatenate(kept)
        with np.errstate(all='ignore'), warnings.catch_warnings():
            warnings.simplefilter('ignore')   # np.mean of an empty gather warns and gives nan
            for c in range(self.C):                                        # :97-105
                mean_value = np.mean(conf[kept == c])
                if not np.isnan(mean_value) and not np.isinf(mean_value):
                    if self.class_mean_probs[c] == 0:
                        self.class_mean_probs[c] = mean_value
                    else:
                        self.class_mean_probs[c] = self.class_mean_probs[c] * self.cp_gamma + \
                            mean_value * (1 - self.cp_gamma)
        return kept

    # :190-211
    def step_conf(self, conf, label, paths):
        temp = ias_quantile_thresholds(conf, label, self.class_threshold, self.C,
                                       self.alpha, self.gamma, self.faithful)
        self.class_threshold = ias_ema_update(self.class_threshold, temp, self.beta)
        self.temp_trace.append(temp.copy())
        self.threshold_trace.append(self.class_threshold.copy())
        return self.select_and_save(conf, label, paths)

    def step_logits(self, logits, paths):
        conf, label = softmax_max(logits)
        return self.step_conf(conf, label, paths)

    def run(self, batches):
        """batches: iterable of (logits tensor [B,C,H,W], [paths])."""
        with torch.no_grad():
            for logits, paths in batches:
                self.step_logits(logits, paths)
        return self


# ------------------------------------------------------------------------ CBST
def cbst_thresholds(batches_conf_label, num_classes, sample_interval, p, f64_quantile=True):
    """CBSTPseudoGenerator.get_constant_threshold (:145-165): every sample_interval-th fp16 confidence per class and
    batch (raster order), then the (1 - p) quantile per class.  float64[C].

    ``f64_quantile=True`` (default) evaluates the quantile in float64, which is what the reference's pinned numpy
    1.19.2 does and what the CUDA path implements.  With numpy >= 2 the reference's call ``np.quantile(list of
    np.float16, python float)`` casts q AND the virtual index (n-1)*q to float16 (NEP 50): the index is rounded
    to ~3 digits and overflows to inf beyond 65 504 samples, so every threshold degenerates to the class's
    largest sample on real data.  ``f64_quantile=False`` reproduces that installed-version behaviour (used only
    to pin this restatement against the fixture generated with numpy 2.3.5)."""
    samples = {c: [] for c in range(num_classes)}
    for conf, label in batches_conf_label:
        for c in range(num_classes):
            vals = conf[label == c].astype(np.float16)
            samples[c].extend(vals[0:len(vals):sample_interval])
    thr = np.ones(num_classes)
    for c in range(num_classes):
        if f64_quantile:
            thr[c] = np.quantile(np.asarray(samples[c], dtype=np.float64), 1 - p)
        else:
            thr[c] = np.quantile(samples[c], 1 - p)
    return thr
