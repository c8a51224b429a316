"""Pseudo-label policies on the CUDA IAS pipeline, behind the reference's ``PSEUDO_POLICY`` interface.

Mirrors ``workflows/pseudo_label_generator.py`` (reference, /root/reference/code):
``BasePseudoGenerator`` :14-106, ``ConstantThresholdPseudoGenerator`` ('CT') :109-132,
``NoThresholdPseudoGenerator`` ('NT') :135-139, ``IASPseudoGenerator`` ('IAS') :168-213.  Same
constructor (``PSEUDO_POLICY[type](cfg)``), same attributes (``class_threshold``, ``class_mean_probs``,
``statics_class``, ``sample_stats``, ``samples_class``), same methods and the same files written by
``save_pseudo_label`` / ``save_data``.

What changes is where the work happens.  The reference copies conf (f32) + label (int64) of every
batch to the host (12 B/px) and runs numpy / Python loops there; here logits never leave the GPU:
each batch goes through phase A as it arrives, a window of batches is then scanned (phase B) and
masked (phase C) on the device and ENCODED AS PNG FILES on the device (``hiast_png_encode``); only the finished
files (typically 2-5 % of the label bytes), the per-image class counts and the thresholds come back, and a small
thread pool writes the files.  ``png='host'`` (or an overridden ``save_pseudo_label`` hook) keeps the reference's
``cv2.imwrite`` on uint8 label maps copied back at 1 B/px.

The backbone and the datasets are outside this package: ``initialize`` takes an
injected ``model`` / ``loader`` (any callable returning ``{'logits': [B,C,H,W]}``; any iterable of
``{'images', 'image_paths'}``).  Inside a reference checkout, bind the reference's own
``initialize`` instead.

``CBSTPseudoGenerator`` ('CBST') :142-165 is implemented on the device as well (two passes over the loader).
"""

from __future__ import annotations

import json
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

from . import ops
from .ias_engine import IASEngine
from .registry import PSEUDO_POLICY


def _cfg_get(node, path, default=None):
    for part in path.split('.'):
        if node is None or not hasattr(node, part):
            return default
        node = getattr(node, part)
    return node


_COPY_STREAMS = {}


def _copy_stream(device):
    """One host-to-device copy stream per device for the life of the process: the caching allocator keeps a block pool per
    stream, so a fresh stream per run() would cudaMalloc every batch again (measured: 260 us per batch instead of 23)."""
    key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
    st = _COPY_STREAMS.get(key)
    if st is None:
        st = _COPY_STREAMS[key] = torch.cuda.Stream(device)
    return st


class BasePseudoGenerator:

    def __init__(self, cfg, model=None, loader=None, dataset_len=None, save_dir=None, window_batches=8,
                 device='cuda', png_workers=8, png='device', prefetch=None, defer_sync=True):
        self.cfg = cfg
        self.statics_class = np.array([0] * self.cfg.dataset.num_classes)                # :18
        self.sample_stats = []                                                           # :19
        self.samples_class = {i: [] for i in range(self.cfg.dataset.num_classes)}        # :20
        self.class_mean_probs = np.zeros(self.cfg.dataset.num_classes)                   # :21
        self.class_threshold = None
        self.device = torch.device(device)
        self.window_batches = int(window_batches)
        self.defer_sync = bool(defer_sync)            # complete a window's outputs while the next one is computed
        self.prefetch = None if prefetch is None else int(prefetch)   # H2D copies issued this many batches ahead (None: auto, 0: inline)
        self._model_arg, self._loader_arg, self._len_arg, self._save_dir_arg = model, loader, dataset_len, save_dir
        self._png_pool = ThreadPoolExecutor(max_workers=png_workers) if png_workers > 0 else None
        self._png_workers = max(1, int(png_workers))
        self._png_jobs = []
        self._png_slot_jobs = [[], []]              # device PNG path: writers of the two pinned blob buffers
        self._png_slot = 0
        if png not in ('device', 'host'):
            raise ValueError("png must be 'device' or 'host'")
        self.png = png
        self._png_encoder = None
        self._engine = None
        self.pow_rounding_certified = True
        self.initialize()

    # ------------------------------------------------------------------ set-up
    def initialize(self):
        """:25-41.  The reference builds the segmentation model and the target DataLoader from cfg; both are
        outside this package, so they are injected.  The save-dir contract (:38-41) is kept."""
        if self._model_arg is None or self._loader_arg is None:
            raise RuntimeError('hiast_b200 pseudo-label generators need model= and loader= (the DeepLabv2 backbone and '
                               'the dataset loaders are not part of this package; see INTEGRATION.md)')
        self.model = self._model_arg
        self.t_loader = self._loader_arg
        n = self._len_arg
        if n is None:
            ds = getattr(self.t_loader, 'dataset', None)
            n = len(ds) if ds is not None else None
        self.t_dataset = range(n) if n is not None else None
        self.pseudo_label_save_dir = self._save_dir_arg or _cfg_get(self.cfg, 'pseudo_policy.save_dir')
        assert self.pseudo_label_save_dir is not None and \
            (not os.path.exists(self.pseudo_label_save_dir) or len(os.listdir(self.pseudo_label_save_dir)) == 0)
        os.makedirs(self.pseudo_label_save_dir, exist_ok=True)

    # ------------------------------------------------------------------ outputs
    def save_pseudo_label(self, plbl, img_path):
        """:43-46  uint8 gray PNG '{stem}_pseudo_label.png'."""
        import cv2
        img_name = os.path.splitext(os.path.basename(img_path))[0]
        plbl_save_path = os.path.join(self.pseudo_label_save_dir, '{}_pseudo_label.png'.format(img_name))
        cv2.imwrite(plbl_save_path, plbl.astype(np.uint8))

    def _pseudo_label_path(self, img_path):
        """:44-45"""
        img_name = os.path.splitext(os.path.basename(img_path))[0]
        return os.path.join(self.pseudo_label_save_dir, '{}_pseudo_label.png'.format(img_name))

    def save_pseudo_label_file(self, png_bytes, img_path):
        """:43-46 with the file already encoded on the device: same name, same decoded pixels.  (Hook: when a subclass
        overrides it, files are handed over one by one; otherwise a window is written by one hiast_write_files call.)"""
        plbl_save_path = self._pseudo_label_path(img_path)
        with open(plbl_save_path, 'wb') as f:
            f.write(png_bytes)

    def _device_png(self):
        """The device writer is used unless the caller asked for the host one or hooked save_pseudo_label."""
        return self.png == 'device' and type(self).save_pseudo_label is BasePseudoGenerator.save_pseudo_label

    def _save_file_async(self, png_bytes, img_path, slot):
        if self._png_pool is None:
            self.save_pseudo_label_file(png_bytes, img_path)
        else:
            self._png_slot_jobs[slot].append(self._png_pool.submit(self.save_pseudo_label_file, png_bytes, img_path))

    def _wait_png_slot(self, slot):
        for job in self._png_slot_jobs[slot]:
            job.result()
        self._png_slot_jobs[slot] = []

    def _save_async(self, plbl, img_path):
        if self._png_pool is None or type(self).save_pseudo_label is not BasePseudoGenerator.save_pseudo_label:
            self.save_pseudo_label(plbl, img_path)       # overridden hooks are called inline, in order
        else:
            self._png_jobs.append(self._png_pool.submit(self.save_pseudo_label, plbl, img_path))

    def _finish_pending_emit(self):
        """Complete the window whose outputs were queued by ``_emit_window(..., defer=True)``: wait for ITS event (the GPU is
        one window further by now), do the per-image bookkeeping, hand the finished files to the native writer."""
        p = getattr(self, '_pending_emit', None)
        if p is None:
            return
        self._pending_emit = None
        blob_host, offsets = p['enc'].finish(p['handle'])
        counts_h = p['counts'].numpy()[:p['n']].copy()
        for i in range(p['n']):
            self._record_image(counts_h[i], p['paths'][i])
        targets = [self._pseudo_label_path(q) for q in p['paths']]
        self._png_slot_jobs[p['slot']].append(self._png_pool.submit(ops.write_files, targets, blob_host, offsets,
                                                                    self._png_workers))
        if p.get('after') is not None:
            p['after']()

    def _wait_png(self):
        self._finish_pending_emit()
        for job in self._png_jobs:
            job.result()
        self._png_jobs = []
        self._wait_png_slot(0)
        self._wait_png_slot(1)

    def save_data(self):
        """:48-62  same file names, formats and locations (one level above the PNG directory)."""
        root = os.path.join(self.pseudo_label_save_dir, '..')
        if self.class_threshold is not None:
            print('class threshold: {}'.format(self.class_threshold))
            np.save(os.path.join(root, 'class_threshold.npy'), self.class_threshold)
        print('class statics number: {}'.format(self.statics_class))
        np.save(os.path.join(root, 'statics_class.npy'), self.statics_class)
        print('class mean probabilities: {}'.format(self.class_mean_probs))
        np.save(os.path.join(root, 'class_mean_probabilities.npy'), self.class_mean_probs)
        with open(os.path.join(root, 'sample_class_stats.json'), 'a') as f:
            f.write(json.dumps(self.sample_stats))
        with open(os.path.join(root, 'samples_with_class.json'), 'a') as f:
            f.write(json.dumps(self.samples_class))

    def run(self):
        raise NotImplementedError

    # ------------------------------------------------------- host bookkeeping
    def _record_image(self, counts_row, img_path):
        """:82-89 from the per-image class counts computed on the device."""
        current_stats = {}
        for i in range(self.cfg.dataset.num_classes):
            pix_num = int(counts_row[i])
            if pix_num != 0:
                current_stats[i] = pix_num
                self.samples_class[i].append([img_path, pix_num])
                self.statics_class[i] += pix_num
        current_stats['file'] = img_path
        self.sample_stats.append(current_stats)

    def _cp_gamma(self):
        return float(_cfg_get(self.cfg, 'preprocessor.copy_paste.gamma', 0.99))

    def select_and_save_confident_label(self, probs_pred, lbls_pred, img_paths):
        """:67-106 for host (numpy) or device conf [B,H,W] / labels [B,H,W]; uses ``self.class_threshold``
        (None = keep everything), records the per-image statistics, saves the PNGs and updates
        ``class_mean_probs``.  Returns the last image's pseudo-label like the reference (:106)."""
        C = self.cfg.dataset.num_classes
        conf = torch.as_tensor(probs_pred).to(self.device, torch.float32).contiguous()
        label = torch.as_tensor(lbls_pred).to(self.device)
        label = label.contiguous() if label.dtype == torch.uint8 else label.to(torch.uint8).contiguous()
        b = conf.shape[0]
        thr = np.zeros(C) if self.class_threshold is None else np.asarray(self.class_threshold, dtype=np.float64)
        thr_groups = torch.from_numpy(thr.reshape(1, C).copy()).to(self.device)
        plbl, counts, confsum = ops.ias_select(conf, label, thr_groups, C, b)
        mean_state = torch.from_numpy(np.asarray(self.class_mean_probs, dtype=np.float64).copy()).to(self.device)
        ops.ias_meanprob_scan(confsum, counts, b, C, self._cp_gamma(), mean_state)
        plbl_h, counts_h = plbl.cpu().numpy(), counts.cpu().numpy()
        self.class_mean_probs = mean_state.cpu().numpy()
        for i, img_path in enumerate(img_paths):
            self._record_image(counts_h[i], img_path)
            self._save_async(plbl_h[i], img_path)
        self._wait_png()
        return plbl_h[-1].astype(np.int64)

    # ------------------------------------------------------------ device loop
    def _make_engine(self, logits, alpha=0.0, beta=0.0, gamma=1.0):
        b, c, h, w = logits.shape       # a LowResLogits reports the up-sampled (image) size
        group = int(_cfg_get(self.cfg, 'pseudo_policy.batch_size', b) or b)
        group = max(group, b)
        return IASEngine(c, h, w, group, alpha, beta, gamma, self._cp_gamma(), group * self.window_batches,
                         device=self.device)

    def _device_batches(self):
        """(images on the device, image_paths) for every loader batch.  With ``prefetch > 0`` a producer thread walks the loader
        and issues the host-to-device copies on its own stream, up to ``prefetch`` batches ahead, so the copies of the next window
        run while the main thread sits in the stream sync at the end of the current one (the consumer waits on each batch's event)."""
        depth = getattr(self, 'prefetch', 0)
        if (depth is not None and depth <= 0) or self.device.type != 'cuda':
            for data in self.t_loader:
                yield data['images'].to(self.device, non_blocking=True), list(data['image_paths'])
            return
        import queue
        import threading
        q = queue.Queue()
        copy_stream = _copy_stream(self.device)
        stop = threading.Event()
        slots = []                                # semaphore, created by the producer once the batch size is known

        def producer():
            try:
                with torch.cuda.stream(copy_stream):
                    for data in self.t_loader:
                        if not slots:
                            nbytes = max(1, data['images'].numel() * data['images'].element_size())
                            # None: about 1.5 GB of batches in flight, at most one window (full-resolution logits: 4 batches;
                            # images or stride-8 logits: the whole next window)
                            n = depth if depth is not None else max(2, min(self.window_batches, int(1.5e9 // nbytes)))
                            slots.append(threading.Semaphore(n))
                        while not slots[0].acquire(timeout=0.05):
                            if stop.is_set():
                                return
                        if stop.is_set():
                            return
                        imgs = data['images'].to(self.device, non_blocking=True)
                        ev = torch.cuda.Event()
                        ev.record(copy_stream)
                        q.put((imgs, ev, list(data['image_paths'])))
                q.put(None)
            except BaseException as exc:          # surfaces in the consumer
                q.put(exc)

        t = threading.Thread(target=producer, name='hiast-h2d', daemon=True)
        t.start()
        try:
            while True:
                item = q.get()
                if item is None:
                    break
                if isinstance(item, BaseException):
                    raise item
                imgs, ev, paths = item
                slots[0].release()
                main = torch.cuda.current_stream(self.device)
                main.wait_event(ev)
                imgs.record_stream(main)
                yield imgs, paths
        finally:
            stop.set()
            t.join(timeout=5.0)

    def _iterate_logits(self):
        """Yields (logits [B,C,H,W] on the device, image_paths) exactly as :189-192 produces them."""
        self.model.eval() if hasattr(self.model, 'eval') else None
        with torch.no_grad():
            for imgs, image_paths in self._device_batches():
                out = self.model(imgs)
                if 'logits' not in out and 'logits_lr' in out:
                    # the network's own (stride-8) output: the bilinear up-sampling of
                    # self_training_segmentor.py:27 is fused into phase A
                    lr = out['logits_lr']
                    lr = lr.float() if lr.dtype != torch.float32 else lr
                    yield LowResLogits(lr.contiguous(), tuple(out.get('size', imgs.shape[2:]))), image_paths
                    continue
                logits = out['logits']
                if logits.dtype != torch.float32:
                    logits = logits.float()
                yield logits.contiguous(), image_paths

    def _already_done(self):
        return self.t_dataset is not None and len(os.listdir(self.pseudo_label_save_dir)) >= len(self.t_dataset)


class LowResLogits:
    """Stride-8 logits [B,C,h,w] together with the size they are to be up-sampled to (bilinear, align_corners=True)."""

    def __init__(self, logits_lr, size):
        self.logits_lr, self.size = logits_lr, (int(size[0]), int(size[1]))

    @property
    def shape(self):
        b, c = self.logits_lr.shape[:2]
        return (b, c, self.size[0], self.size[1])


def _phase_a(engine, logits, first_image):
    """Phase A of one batch.  Full-resolution logits are consumed at once (they are 159 MB per image and the caller may
    reuse the buffer).  Stride-8 logits (2.5 MB per image) are only queued: consecutive batches are concatenated and go through
    ONE launch per window (`_flush_lowres`) -- a 2-image launch of the fused up-sampling kernel costs 106 us, its share of a
    64-image launch 67 us."""
    if isinstance(logits, LowResLogits) and hasattr(engine, 'phase_a_lowres'):
        pend = engine.__dict__.setdefault('_pending_lr', [])
        if pend:
            last, last_first = pend[-1]
            if (last_first + last.logits_lr.shape[0] != first_image or last.size != logits.size
                    or last.logits_lr.shape[1:] != logits.logits_lr.shape[1:]):
                _flush_lowres(engine)
                pend = engine.__dict__.setdefault('_pending_lr', [])
        pend.append((logits, first_image))
    elif isinstance(logits, LowResLogits):
        engine.phase_a_lowres(logits.logits_lr, first_image)
    else:
        engine.phase_a(logits, first_image)


def _flush_lowres(engine):
    """Launch phase A for the queued stride-8 batches (one launch for the contiguous run of images)."""
    pend = engine.__dict__.pop('_pending_lr', None)
    if not pend:
        return
    lr = pend[0][0].logits_lr if len(pend) == 1 else torch.cat([p[0].logits_lr for p in pend])
    engine.phase_a_lowres(lr, pend[0][1])


def _flush_window(gen, engine, paths, n_images, scan, before_emit=None):
    """Phases B/C for the images in the window, then the outputs (``_emit_window``, deferred when possible).
    ``before_emit()`` may queue more copies on the stream and returns the completion callback."""
    if n_images == 0:
        return
    _flush_lowres(engine)
    if scan:
        engine.phase_b(0, n_images)
    engine.phase_c(0, n_images)
    engine.mean_prob(0, n_images)
    after = before_emit() if before_emit is not None else None
    _emit_window(gen, engine, paths, 0, n_images, defer=True, after=after)


def _can_defer(gen, engine):
    """Deferred completion needs the device PNG writer with the native file writer and a CUDA engine."""
    return (getattr(gen, 'defer_sync', True) and gen._device_png() and gen._png_pool is not None
            and type(gen).save_pseudo_label_file is BasePseudoGenerator.save_pseudo_label_file
            and torch.is_tensor(engine.plbl) and engine.plbl.is_cuda and engine.plbl.shape[2] <= 128 * 256)


def _emit_window(gen, engine, paths, first, n_images, defer=False, after=None):
    """Outputs of the masked window engine.plbl[first:first+n] (PNG files + per-image statistics).  ``defer=True`` (single-GPU
    generators): everything is only QUEUED on the stream -- encoder, copies of the offset table / predicted blob bytes / counts,
    an event -- and completed when the next window is emitted (``_finish_pending_emit``), so the host never waits for the GPU
    at the end of a window and phase A of the next window is issued behind this one without a gap.  ``after`` runs on
    completion (the IAS generator reads its threshold copies there)."""
    if defer and _can_defer(gen, engine):
        plbl, counts = engine.plbl[first:first + n_images], engine.counts[first:first + n_images]
        gen._finish_pending_emit()                    # the previous window (its event is long past)
        enc = gen._png_encoder
        if enc is None or (enc.H, enc.W) != tuple(plbl.shape[1:]) or enc.max_images < n_images:
            enc = gen._png_encoder = ops.PngEncoder(plbl.shape[1], plbl.shape[2], engine.max_images, device=engine.device)
        slot = gen._png_slot = 1 - gen._png_slot
        gen._wait_png_slot(slot)                      # the writers of two windows ago still read this pinned blob
        cpins = gen.__dict__.setdefault('_pinned_counts2', {})
        cpin = cpins.get(slot)
        if cpin is None or cpin.shape[0] < n_images or cpin.shape[1:] != counts.shape[1:]:
            cpin = cpins[slot] = torch.empty(engine.counts.shape, dtype=torch.int64).pin_memory()
        cpin[:n_images].copy_(counts, non_blocking=True)
        handle = enc.encode_async(plbl, slot)
        gen._pending_emit = dict(enc=enc, handle=handle, slot=slot, paths=list(paths[:n_images]), n=n_images, counts=cpin,
                                 after=after)
        return
    gen._finish_pending_emit()
    _emit_window_sync(gen, engine, paths, first, n_images)
    if after is not None:
        after()


def _emit_window_sync(gen, engine, paths, first, n_images):
    """Outputs of the masked window engine.plbl[first:first+n]: PNG files + per-image statistics.

    Device path (default): the label maps are encoded as PNG files on the device, one D2H copy moves the finished files to a
    pinned blob (two blobs alternate) and ONE native call writes them while the next window is computed.  Host path
    (``png='host'`` or a hooked ``save_pseudo_label``): uint8 label maps come back at 1 B/px through a pinned buffer (one
    sync per window) and go to the hook / ``cv2.imwrite``."""
    plbl, counts = engine.plbl[first:first + n_images], engine.counts[first:first + n_images]
    if not torch.is_tensor(plbl):                     # a host stand-in engine (CPU tests of the orchestration)
        plbl_h, counts_h = np.asarray(plbl), torch.as_tensor(counts).numpy()
        for i in range(n_images):
            gen._record_image(counts_h[i], paths[i])
            gen._save_async(plbl_h[i].copy(), paths[i])
        gen._wait_png()
        return
    if gen._device_png() and plbl.shape[2] > 128 * 256:
        gen.png = 'host'                              # wider than the device writer's 32768-pixel rows: the reference's writer
    native = (type(gen).save_pseudo_label_file is BasePseudoGenerator.save_pseudo_label_file and gen._png_pool is not None)
    if gen._device_png():
        cpin = getattr(gen, '_pinned_counts', None)
        if cpin is None or cpin.shape[0] < n_images or cpin.shape[1:] != counts.shape[1:]:
            cpin = gen._pinned_counts = torch.empty(engine.counts.shape, dtype=torch.int64).pin_memory()
        cpin[:n_images].copy_(counts, non_blocking=True)
        enc = gen._png_encoder
        if enc is None or (enc.H, enc.W) != tuple(plbl.shape[1:]) or enc.max_images < n_images:
            enc = gen._png_encoder = ops.PngEncoder(plbl.shape[1], plbl.shape[2], engine.max_images, device=engine.device)
        slot = gen._png_slot = 1 - gen._png_slot
        gen._wait_png_slot(slot)                      # the writers of two windows ago still read this pinned blob
        files = enc.encode_to_host(plbl, slot)        # finished PNG files; syncs the stream once
        counts_h = cpin.numpy()[:n_images].copy()
        for i in range(n_images):
            gen._record_image(counts_h[i], paths[i])
        if native:
            # one native call writes the whole window (POSIX writer threads, no interpreter lock)
            blob_host, offsets = enc.host_blob(slot)
            targets = [gen._pseudo_label_path(p) for p in paths[:n_images]]
            gen._png_slot_jobs[slot].append(gen._png_pool.submit(ops.write_files, targets, blob_host, offsets,
                                                                 gen._png_workers))
        else:
            for i in range(n_images):
                gen._save_file_async(files[i], paths[i], slot)
        return                                        # the files are written while the next window is computed
    pins = getattr(gen, '_pinned', None)
    if pins is None or pins[0].shape[0] < n_images or pins[0].shape[1:] != plbl.shape[1:]:
        pins = gen._pinned = (torch.empty(engine.plbl.shape, dtype=torch.uint8).pin_memory(),
                              torch.empty(engine.counts.shape, dtype=torch.int64).pin_memory())
    pins[0][:n_images].copy_(plbl, non_blocking=True)   # host PNG path: 1 B/px back
    pins[1][:n_images].copy_(counts, non_blocking=True)
    torch.cuda.current_stream(engine.device).synchronize()
    gen._wait_png()                                   # the previous window's encoders still read the pinned buffer
    plbl_h, counts_h = pins[0].numpy(), pins[1].numpy()
    for i in range(n_images):
        gen._record_image(counts_h[i], paths[i])
        gen._save_async(plbl_h[i], paths[i])
    gen._wait_png()


@PSEUDO_POLICY.register('CT')
class ConstantThresholdPseudoGenerator(BasePseudoGenerator):

    def get_constant_threshold(self):
        """:112-113"""
        return self.cfg.pseudo_policy.ct.threshold * np.ones(self.cfg.dataset.num_classes)

    def run(self):
        """:115-132"""
        if self._already_done():
            print('%% pseudo labels have existed')
            return
        self.class_threshold = self.get_constant_threshold()
        C = self.cfg.dataset.num_classes
        thr = np.zeros(C) if self.class_threshold is None else np.asarray(self.class_threshold, dtype=np.float64)
        engine, paths, n = None, [], 0
        for logits, img_paths in self._iterate_logits():
            if engine is None:
                engine = self._engine = self._make_engine(logits)
                engine.thr_groups.copy_(torch.from_numpy(np.tile(thr, (engine.max_groups, 1))))
            b = logits.shape[0]
            _phase_a(engine, logits, n)
            paths += img_paths
            n += b
            if n + engine.B > engine.max_images or b != engine.B:
                _flush_window(self, engine, paths, n, scan=False)
                paths, n = [], 0
        if engine is not None:
            _flush_window(self, engine, paths, n, scan=False)
            self.class_mean_probs = engine.mean_state.cpu().numpy()
        self._wait_png()
        self.save_data()


@PSEUDO_POLICY.register('NT')
class NoThresholdPseudoGenerator(ConstantThresholdPseudoGenerator):

    def get_constant_threshold(self):
        """:138-139  no threshold: every arg-max label is kept."""
        return None


@PSEUDO_POLICY.register('CBST')
class CBSTPseudoGenerator(ConstantThresholdPseudoGenerator):

    def get_constant_threshold(self):
        """:145-165.  First pass over the target set: per batch and class, every ``cbst.sample_interval``-th
        confidence (raster order, fp16) goes into one per-class histogram on the device; the thresholds are the
        (1 - cbst.p) quantiles.  ``run`` then makes the second pass with these constant thresholds (:115-132)."""
        C = self.cfg.dataset.num_classes
        cbst = self.cfg.pseudo_policy.cbst
        key_lo = ops.ias_key_lo(C)
        hist = None
        for logits, _paths in self._iterate_logits():
            b = logits.shape[0]
            if isinstance(logits, LowResLogits):                        # phase A (its own histogram is not used here)
                conf, label, _ = ops.ias_upsample_softmax_hist(logits.logits_lr, logits.size, b)
            else:
                conf, label, _ = ops.ias_softmax_hist(logits, b)
            if hist is None:
                hist = torch.zeros((C, ops.ias_row_stride(key_lo)), dtype=torch.int32, device=conf.device)
            ops.cbst_sample_hist(conf, label, C, b, int(cbst.sample_interval), key_lo, hist)
        if hist is None:
            raise IndexError('index -1 is out of bounds for axis 0 with size 0')     # np.quantile([]) on an empty set
        flag = torch.zeros(1, dtype=torch.int32, device=hist.device)
        thr = ops.cbst_quantile(hist, C, key_lo, 1 - cbst.p, flag)
        code = int(flag.item())
        if code & 4:
            raise IndexError('index -1 is out of bounds for axis 0 with size 0')     # a class was never predicted
        if code & 1:
            raise ValueError('Quantiles must be in the range [0, 1]')
        return thr.cpu().numpy()


@PSEUDO_POLICY.register('IAS')
class IASPseudoGenerator(BasePseudoGenerator):

    def get_ias_threshold(self, class_probs_dict, num_classes, alpha, old_thresholds=None, gamma=1.0):
        """:171-179 on the device.  ``class_probs_dict[c]`` is the reference's sample list for class c: the
        previous threshold followed by the fp16 confidences of the batch (:198-201); None skips the class.
        Returns float32[num_classes] like the reference."""
        if old_thresholds is None:
            old_thresholds = np.ones(num_classes)
        out = np.ones(num_classes, dtype=np.float32)
        for c in range(num_classes):
            samples = class_probs_dict[c]
            if samples is None:
                continue
            head = np.float64(samples[0])
            if head != np.float64(old_thresholds[c]):
                raise ValueError('class_probs_dict[c][0] must be old_thresholds[c] (the reference prepends the '
                                 'previous threshold to the sample list, pseudo_label_generator.py:198)')
            vals = np.asarray(samples[1:], dtype=np.float32)
            if not np.array_equal(vals.astype(np.float16).astype(np.float32), vals):
                raise ValueError('samples must be fp16-representable confidences (pseudo_label_generator.py:201)')
            conf = torch.from_numpy(vals.reshape(1, 1, -1)).to(self.device) if vals.size else \
                torch.zeros((1, 1, 1), device=self.device)
            label = torch.zeros(conf.shape, dtype=torch.uint8, device=self.device)
            if vals.size == 0:
                label.fill_(255)
            hist, _ = ops.ias_conf_hist(conf, label, 1, 1, key_lo=0)
            state = torch.tensor([float(old_thresholds[c])], dtype=torch.float64, device=self.device)
            flag = torch.zeros(1, dtype=torch.int32, device=self.device)
            _, temp = ops.ias_threshold_scan(hist, 1, 1, 0, alpha, 0.0, gamma, state, error_flag=flag)
            if int(flag.item()) & 1:
                raise ValueError('Quantiles must be in the range [0, 1]')
            out[c] = temp.item()
        return out

    def run(self):
        """:181-213"""
        if self._already_done():
            print('%% pseudo labels have existed')
            return
        C = self.cfg.dataset.num_classes
        ias = self.cfg.pseudo_policy.ias
        self.class_threshold = 0.9 * np.ones(C)                                            # :185
        self.threshold_trace = []
        engine, paths, n = None, [], 0
        for logits, img_paths in self._iterate_logits():
            if engine is None:
                engine = self._engine = self._make_engine(logits, ias.alpha, ias.beta, ias.gamma)
                engine.thr_state.copy_(torch.from_numpy(self.class_threshold))
                engine.mean_state.copy_(torch.from_numpy(self.class_mean_probs))
            b = logits.shape[0]
            _phase_a(engine, logits, n)                   # asynchronous; overlaps the next forward pass
            paths += img_paths
            n += b
            if n + engine.B > engine.max_images or b != engine.B:
                self._finish_window(engine, paths, n)
                paths, n = [], 0
        if engine is not None:
            self._finish_window(engine, paths, n)
            self.pow_rounding_certified = engine.check_errors()
        self._wait_png()
        self.save_data()

    def _finish_window(self, engine, paths, n):
        if n == 0:
            return
        g = engine._groups(n)

        def before_emit():
            # per-window host copies of the thresholds: queued into pinned buffers (one set per PNG slot), read on completion
            if not _can_defer(self, engine):
                def read_now():
                    self.threshold_trace.append(engine.thr_groups[:g].cpu().numpy().copy())
                    self.class_threshold = engine.thr_state.cpu().numpy()
                    self.class_mean_probs = engine.mean_state.cpu().numpy()
                return read_now
            pins = self.__dict__.setdefault('_thr_pins', {})
            slot = 1 - self._png_slot                  # the slot _emit_window is about to take
            buf = pins.get(slot)
            if buf is None or buf[0].shape != engine.thr_groups.shape:
                buf = pins[slot] = (torch.empty(engine.thr_groups.shape, dtype=torch.float64).pin_memory(),
                                    torch.empty(engine.thr_state.shape, dtype=torch.float64).pin_memory(),
                                    torch.empty(engine.mean_state.shape, dtype=torch.float64).pin_memory())
            buf[0][:g].copy_(engine.thr_groups[:g], non_blocking=True)
            buf[1].copy_(engine.thr_state, non_blocking=True)
            buf[2].copy_(engine.mean_state, non_blocking=True)

            def read_later():
                self.threshold_trace.append(buf[0][:g].numpy().copy())
                self.class_threshold = buf[1].numpy().copy()
                self.class_mean_probs = buf[2].numpy().copy()
            return read_later

        _flush_window(self, engine, paths, n, scan=True, before_emit=before_emit)


def striped_batch_order(n_images_total, window_images, batch_size, rank, world_size):
    """Dataset indices, batch by batch, that rank ``rank`` processes (windows w = rank, rank + R, ... in the pinned global
    order; each window cut into batches of ``batch_size``).  Feed it to a ``DataLoader(batch_sampler=...)``."""
    from .sharded import local_windows, window_images as _win
    n_windows = (n_images_total + window_images - 1) // window_images
    batches = []
    for w in local_windows(n_windows, rank, world_size):
        i0, n = _win(w, window_images, n_images_total)
        for k in range(0, n, batch_size):
            batches.append(list(range(i0 + k, i0 + min(n, k + batch_size))))
    return batches


@PSEUDO_POLICY.register('IAS_SHARDED')
class ShardedIASPseudoGenerator(IASPseudoGenerator):
    """IAS pseudo-labelling over R ranks (one process per GPU, ``torch.distributed`` initialised), behind the same
    ``PSEUDO_POLICY[...](cfg).run()`` call.  New: the reference runs this stage in one process.

    The pinned global order of the target set is cut into windows of ``window_batches`` batches; window w belongs to
    rank w mod R and **the injected loader yields only this rank's batches, in order** (``striped_batch_order`` gives the
    batch sampler).  Per owned window: phase A of the NEXT window is issued batch by batch as the model produces logits,
    then the 19-double threshold state is received from rank r-1 (NCCL / gloo recv), the window is scanned, the state is
    sent on to rank r+1, and the window is masked, PNG-encoded on the device and written.  At the end the per-group sums
    are all-gathered and every rank replays the mean-probability EMA, the statistics lists are gathered in global order,
    and rank 0 writes ``save_data``'s files.  Results are bit-identical to the single-process generator
    (tests/test_sharded_gloo.py with a host stand-in engine, tests/test_sharded_gpu.py with NCCL)."""

    def __init__(self, cfg, *args, engine_factory=None, process_group=None, **kw):
        import torch.distributed as dist
        self._engine_factory = engine_factory
        self._pg = process_group
        super().__init__(cfg, *args, **kw)
        if dist.is_available() and dist.is_initialized():
            dist.barrier(group=process_group)          # every rank has seen the empty save dir before anyone writes

    def _make_engine(self, logits, alpha=0.0, beta=0.0, gamma=1.0):
        b, c, h, w = logits.shape
        group = max(int(_cfg_get(self.cfg, 'pseudo_policy.batch_size', b) or b), b)
        n = 2 * group * self.window_batches                       # two windows: the next one's phase A runs ahead
        if self._engine_factory is not None:
            return self._engine_factory(c, h, w, group, alpha, beta, gamma, self._cp_gamma(), n)
        return IASEngine(c, h, w, group, alpha, beta, gamma, self._cp_gamma(), n, device=self.device)

    def run(self):
        import torch.distributed as dist
        from .sharded import ShardedIAS, window_images
        if self._already_done():
            print('%% pseudo labels have existed')
            return
        if self.t_dataset is None:
            raise RuntimeError('the sharded generator needs dataset_len (the size of the whole target set)')
        C = self.cfg.dataset.num_classes
        ias = self.cfg.pseudo_policy.ias
        self.class_threshold = 0.9 * np.ones(C)                                            # :185
        n_total = len(self.t_dataset)
        it = self._iterate_logits()
        engine = drv = None
        pending = None
        per_window = {}                                   # global window -> (sample_stats rows, thr_groups)
        first = next(it, None)
        if first is not None:
            engine = self._engine = self._make_engine(first[0], ias.alpha, ias.beta, ias.gamma)
            engine.thr_state.copy_(torch.from_numpy(self.class_threshold))
            window = engine.max_images // 2
            drv = ShardedIAS(engine, window, n_total, process_group=self._pg)
            drv._stash = []
            for j, w in enumerate(drv.my_windows):
                _, n = window_images(w, window, n_total)
                slot, filled, paths = drv._slot(j), 0, []
                while filled < n:                          # phase A of window j, batch by batch
                    logits, p = first if first is not None else next(it)
                    first = None
                    _phase_a(engine, logits, slot + filled)
                    filled += logits.shape[0]
                    paths += p
                if filled != n:
                    raise ValueError('window %d must hold %d images, the loader delivered %d' % (w, n, filled))
                if pending is not None:                    # ... while the previous window waits for its thresholds
                    self._finish_sharded_window(drv, per_window, *pending)
                pending = (w, j, paths, n)
            if pending is not None:
                self._finish_sharded_window(drv, per_window, *pending)
        self._wait_png()
        self._gather_state(drv, per_window, n_total)
        rank = dist.get_rank(self._pg) if dist.is_available() and dist.is_initialized() else 0
        if rank == 0:
            self.save_data()

    def _finish_sharded_window(self, drv, per_window, w, j, paths, n):
        import torch.distributed as dist
        e = drv.engine
        slot = drv._slot(j)
        _flush_lowres(e)                                 # phase A of this and of the next window, before the token wait
        if w > 0 and drv.world > 1:
            dist.recv(e.thr_state, src=drv._global((drv.rank - 1) % drv.world), group=drv.pg)
        e.phase_b(slot, n)
        if w < drv.n_windows_total - 1 and drv.world > 1:
            dist.send(e.thr_state, dst=drv._global((drv.rank + 1) % drv.world), group=drv.pg)
        e.phase_c(slot, n)
        g0, g = slot // e.B, (n + e.B - 1) // e.B
        drv._stash.append(torch.stack([torch.as_tensor(e.confsum[g0:g0 + g]), e.group_counts(slot, n)], dim=1).clone())
        rows_before = len(self.sample_stats)
        _emit_window(self, e, paths, slot, n)
        thr_groups = torch.as_tensor(e.thr_groups[g0:g0 + g]).cpu().numpy().copy()
        per_window[w] = (self.sample_stats[rows_before:], thr_groups)

    def _gather_state(self, drv, per_window, n_total):
        """Global results on every rank: thresholds / mean probabilities / class totals from the driver, the per-image
        statistics lists re-assembled in the global image order."""
        import torch.distributed as dist
        C = self.cfg.dataset.num_classes
        if drv is None:
            return
        thr, mean, statics = drv.finish_state()
        self.class_threshold = thr.cpu().numpy().copy()
        self.class_mean_probs = mean.cpu().numpy().copy()
        self.pow_rounding_certified = drv.engine.check_errors() if hasattr(drv.engine, 'check_errors') else True
        parts = [per_window]
        if drv.world > 1:
            parts = [None] * drv.world
            dist.all_gather_object(parts, per_window, group=drv.pg)
        merged = {}
        for p in parts:
            merged.update(p)
        self.sample_stats, self.threshold_trace = [], []
        self.samples_class = {i: [] for i in range(C)}
        self.statics_class = np.array([0] * C)
        for w in sorted(merged):
            rows, thr_groups = merged[w]
            self.threshold_trace.append(thr_groups)
            for row in rows:
                self.sample_stats.append(row)
                for k, v in row.items():
                    if k != 'file':
                        self.samples_class[k].append([row['file'], v])
                        self.statics_class[k] += v
        assert np.array_equal(self.statics_class, statics.cpu().numpy()), 'gathered statistics disagree with the device sums'
