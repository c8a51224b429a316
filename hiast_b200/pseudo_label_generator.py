"""Pseudo-label policies on the CUDA IAS pipeline, behind the reference's ``PSEUDO_POLICY`` interface.

Mirrors ``workflows/pseudo_label_generator.py`` (reference, /root/reference/code):
``BasePseudoGenerator`` :14-106, ``ConstantThresholdPseudoGenerator`` ('CT') :109-132,
``NoThresholdPseudoGenerator`` ('NT') :135-139, ``CBSTPseudoGenerator`` ('CBST') :142-165,
``IASPseudoGenerator`` ('IAS') :168-213.  Same constructor (``PSEUDO_POLICY[type](cfg)``), same attributes
(``class_threshold``, ``class_mean_probs``, ``statics_class``, ``sample_stats``, ``samples_class``), same methods and the
same files written by ``save_pseudo_label`` / ``save_data``.

What changes is where the work happens.  The reference copies conf (f32) + label (int64) of every batch to the host
(12 B/px) and runs numpy / Python loops there.  Here the logits never leave the GPU and the interpreter never waits for it:

    host loop (one thread)        per batch: one foreign call stages the batch into a device ring over a copy stream
                                  (``hiast_stager_push``), the model runs, phase A is launched (full-resolution logits) or the
                                  stride-8 logits are queued for ONE fused up-sampling launch per window;
    window j closes               phase A(j) is complete on the main stream; the threshold chain of window j-1 (token receive,
                                  scan, token send) is queued on a high-priority side stream behind A(j), so that it runs in the
                                  shadow of phase C / the PNG encoder of window j-2, which are queued on the main stream next
                                  (``hiast_ias_emit_window``: one call for mask, counts, encoder and every device-to-host copy);
    completion                    a native writer pool sleeps on the window's event and puts the files on disk
                                  (``hiast_writer_*``); the interpreter picks up counts / thresholds from pinned buffers when it
                                  is about to reuse them, three windows later.

``png='host'`` (or an overridden ``save_pseudo_label`` hook) keeps the reference's per-image ``cv2.imwrite`` on uint8 label maps
copied back at 1 B/px, through the same deferred completion.

``initialize`` takes an injected ``model`` / ``loader`` (any callable returning ``{'logits': [B,C,H,W]}`` or the network's own
``{'logits_lr': stride-8 logits, 'size': (H, W)}``; any iterable of ``{'images', 'image_paths'}``).  Without them it does what the
reference's ``initialize`` (:25-41) does through the registries: ``MODEL[cfg.model.type](cfg)`` + checkpoint from
``cfg.pseudo_policy.resume_from``, ``DATASET[cfg.dataset.target.type](...)`` in a shuffled ``DataLoader`` -- the backbone and the
datasets themselves are outside this package and have to be registered by the host project.

``IAS_SHARDED`` runs the same pipeline over R ranks (one process per GPU): windows are striped over the ranks and the 19-double
threshold state is handed rank to rank (NCCL send / recv) inside the side-stream chain.
"""

from __future__ import annotations

import json
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

from . import ops
from ._lib import HiastError, stream_ptr
from .ias_engine import IASEngine
from .registry import DATASET, MODEL, PSEUDO_POLICY


def _cfg_get(node, path, default=None):
    for part in path.split('.'):
        if node is None or not hasattr(node, part):
            return default
        node = getattr(node, part)
    return node


def _stream(device, kind):
    from .sharded import device_stream
    return device_stream(device, kind)


class LowResLogits:
    """Stride-8 logits [B,C,h,w] together with the size they are to be up-sampled to (bilinear, align_corners=True)."""

    def __init__(self, logits_lr, size):
        self.logits_lr, self.size = logits_lr, (int(size[0]), int(size[1]))

    @property
    def shape(self):
        b, c = self.logits_lr.shape[:2]
        return (b, c, self.size[0], self.size[1])


class BasePseudoGenerator:

    def __init__(self, cfg, model=None, loader=None, dataset_len=None, save_dir=None, window_batches=8,
                 device='cuda', png_workers=None, png='device', prefetch=None, defer_sync=True, stage_bytes=3 << 30):
        self.cfg = cfg
        self.statics_class = np.array([0] * self.cfg.dataset.num_classes)                # :18
        self.sample_stats = []                                                           # :19
        self.samples_class = {i: [] for i in range(self.cfg.dataset.num_classes)}        # :20
        self.class_mean_probs = np.zeros(self.cfg.dataset.num_classes)                   # :21
        self.class_threshold = None
        self.device = torch.device(device)
        self.window_batches = int(window_batches)
        self.defer_sync = bool(defer_sync)            # False: complete every window before the next one starts (debugging)
        self.prefetch = None if prefetch is None else int(prefetch)   # 0: plain .to(device) on the main stream, no staging ring
        self.stage_bytes = int(stage_bytes)           # device memory of the host-to-device staging ring
        self._model_arg, self._loader_arg, self._len_arg, self._save_dir_arg = model, loader, dataset_len, save_dir
        if png_workers is None:
            png_workers = self._default_workers()
        self._png_pool = ThreadPoolExecutor(max_workers=png_workers) if png_workers > 0 else None
        self._png_workers = max(1, int(png_workers))
        self._png_jobs = []
        if png not in ('device', 'host'):
            raise ValueError("png must be 'device' or 'host'")
        self.png = png
        self._engine = None
        self._stager = None
        self._writer = None
        self.pow_rounding_certified = True
        self.initialize()

    # ``sample_stats`` (:19) and ``samples_class`` (:20) are rebuilt from the [n_images, C] count matrix of a run the first
    # time somebody looks at them (``save_data`` on rank 0, a test, the training code): for 12 k images that is ~40 ms of list
    # building that the other ranks of a sharded run never need.
    @property
    def sample_stats(self):
        self._materialize_stats()
        return self._sample_stats

    @sample_stats.setter
    def sample_stats(self, value):
        self._sample_stats = value

    @property
    def samples_class(self):
        self._materialize_stats()
        return self._samples_class

    @samples_class.setter
    def samples_class(self, value):
        self._samples_class = value

    def _materialize_stats(self):
        pend = self.__dict__.pop('_stats_pending', None)
        if pend is not None:
            self._record_images(*pend)

    @staticmethod
    def _default_workers():
        """File-writer threads: the host cores this rank may use (cores / ranks on the node) minus the interpreter's own,
        at most 8 -- eight ranks with eight writers each on a 32-core host only fight each other."""
        local = int(os.environ.get('LOCAL_WORLD_SIZE', os.environ.get('WORLD_SIZE', '1')) or 1)
        try:
            cores = len(os.sched_getaffinity(0))
        except (AttributeError, OSError):
            cores = os.cpu_count() or 1
        return max(1, min(8, cores // max(1, local) - 1))

    # ------------------------------------------------------------------ set-up
    def initialize(self):
        """:25-41.  Injected ``model`` / ``loader`` are used as they are; otherwise both are built from cfg through the
        registries exactly like the reference (model: ``utils.load_model``, utils/utils.py:68-89; loader: :29-36)."""
        cfg = self.cfg
        if self._model_arg is not None:
            self.model = self._model_arg
        else:
            self.model = self._load_model(_cfg_get(cfg, 'pseudo_policy.resume_from'))
        if self._loader_arg is not None:
            self.t_loader = self._loader_arg
            n = self._len_arg
            if n is None:
                ds = getattr(self.t_loader, 'dataset', None)
                n = len(ds) if ds is not None else None
            self.t_dataset = range(n) if n is not None else None
        else:
            self.t_dataset, self.t_loader = self._build_target_loader()
        self.pseudo_label_save_dir = self._save_dir_arg or _cfg_get(cfg, 'pseudo_policy.save_dir')
        assert self.pseudo_label_save_dir is not None and \
            (not os.path.exists(self.pseudo_label_save_dir) or len(os.listdir(self.pseudo_label_save_dir)) == 0)
        os.makedirs(self.pseudo_label_save_dir, exist_ok=True)

    def _load_model(self, resume_from):
        """utils/utils.py:68-89 ``load_model(cfg, resume_from)`` + ``.cuda()`` (:27)."""
        mtype = _cfg_get(self.cfg, 'model.type')
        if mtype not in MODEL:
            raise RuntimeError("PSEUDO_POLICY[...](cfg) without model=: MODEL[%r] is not registered.  The backbone is outside "
                               "hiast_b200: register the host project's segmentor (INTEGRATION.md section 1) or pass model=" % (mtype,))
        model = MODEL[mtype](self.cfg)
        if resume_from is not None:
            own = model.state_dict()
            loaded = torch.load(resume_from, map_location='cpu')
            strip = 7 if 'module' in list(loaded.keys())[0] else 0          # saved from DistributedDataParallel (:78-79)
            own.update({k[strip:]: v for k, v in loaded.items() if k[strip:] in own})
            model.load_state_dict(own)
            print('%% load model from {}'.format(resume_from))
        else:
            import warnings
            warnings.warn('not load model')
        return model.to(self.device)

    def _build_target_dataset(self):
        """:29-35  target dataset at ``pseudo_policy.resize_size``."""
        cfg = self.cfg
        ttype = _cfg_get(cfg, 'dataset.target.type')
        if ttype not in DATASET:
            raise RuntimeError("PSEUDO_POLICY[...](cfg) without loader=: DATASET[%r] is not registered.  The datasets are outside "
                               "hiast_b200: register the host project's dataset class or pass loader=" % (ttype,))
        rs = cfg.pseudo_policy.resize_size
        aug_type = ['PRS-{}-{}'.format(rs[0], rs[1])]
        return DATASET[ttype](cfg, cfg.dataset.target.json_path, cfg.dataset.target.image_dir, aug_type=aug_type,
                              num_classes=cfg.dataset.num_classes)

    def _build_target_loader(self):
        """:36  a shuffled DataLoader over the target set."""
        from torch.utils.data import DataLoader
        cfg = self.cfg
        ds = self._build_target_dataset()
        loader = DataLoader(ds, cfg.pseudo_policy.batch_size, shuffle=True, num_workers=cfg.dataset.num_workers, pin_memory=True)
        return ds, loader

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
        overrides it, files are handed over one by one; otherwise the native writer pool puts a window on disk.)"""
        plbl_save_path = self._pseudo_label_path(img_path)
        with open(plbl_save_path, 'wb') as f:
            f.write(png_bytes)

    def _device_png(self):
        """The device writer is used unless the caller asked for the host one or hooked save_pseudo_label."""
        return self.png == 'device' and type(self).save_pseudo_label is BasePseudoGenerator.save_pseudo_label

    def _native_files(self):
        return type(self).save_pseudo_label_file is BasePseudoGenerator.save_pseudo_label_file

    def _save_async(self, plbl, img_path):
        """One label map to ``save_pseudo_label``: overridden hooks inline and in order, cv2.imwrite on the thread pool."""
        if self._png_pool is None or type(self).save_pseudo_label is not BasePseudoGenerator.save_pseudo_label:
            self.save_pseudo_label(plbl, img_path)
            return None
        job = self._png_pool.submit(self.save_pseudo_label, plbl, img_path)
        self._png_jobs.append(job)
        return job

    def _wait_png(self):
        for job in self._png_jobs:
            job.result()
        self._png_jobs = []

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
        for i in np.flatnonzero(counts_row):
            pix_num = int(counts_row[i])
            i = int(i)
            current_stats[i] = pix_num
            self.samples_class[i].append([img_path, pix_num])
            self.statics_class[i] += pix_num
        current_stats['file'] = img_path
        self.sample_stats.append(current_stats)

    def _record_images(self, counts, paths):
        """:82-89 for a whole run at once: the same three structures ``_record_image`` appends to, rebuilt from the
        [n_images, C] count matrix in image order."""
        C = self.cfg.dataset.num_classes
        n = counts.shape[0]
        rows, cols = np.nonzero(counts)
        vals = counts[rows, cols].tolist()
        cols_l = cols.tolist()
        cut = np.searchsorted(rows, np.arange(n + 1)).tolist()
        stats = []
        for i in range(n):
            a, b = cut[i], cut[i + 1]
            d = dict(zip(cols_l[a:b], vals[a:b]))
            d['file'] = paths[i]
            stats.append(d)
        self._sample_stats = stats
        self._samples_class = {}
        for c in range(C):
            idx = np.flatnonzero(counts[:, c])
            self._samples_class[c] = [[paths[i], v] for i, v in zip(idx.tolist(), counts[idx, c].tolist())]

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
    def _group_size(self, b):
        group = int(_cfg_get(self.cfg, 'pseudo_policy.batch_size', b) or b)
        return max(group, b)

    def _make_engine(self, shape, alpha=0.0, beta=0.0, gamma=1.0):
        """Engine with room for THREE windows: phase A of window j, the threshold chain of j-1 and the outputs of j-2 are
        in flight together (module docstring)."""
        b, c, h, w = shape                  # a LowResLogits reports the up-sampled (image) size
        group = self._group_size(b)
        n = 3 * group * self.window_batches
        factory = getattr(self, '_engine_factory', None)
        if factory is not None:
            return factory(c, h, w, group, alpha, beta, gamma, self._cp_gamma(), n)
        return IASEngine(c, h, w, group, alpha, beta, gamma, self._cp_gamma(), n, device=self.device)

    def _staged(self, images, pipe):
        """Host batch -> (device tensor, staging slot or None).  The copy runs on the copy stream into a ring slot; the main
        stream is made to wait for it, the host is not."""
        if self.device.type != 'cuda' or not torch.is_tensor(images) or images.is_cuda:
            return (images.to(self.device) if torch.is_tensor(images) and images.device != self.device else images), None
        if self.prefetch == 0:
            return images.to(self.device, non_blocking=True), None
        st = self._stager
        nbytes = images.numel() * images.element_size()
        if st is None or nbytes > st.slot_bytes:
            if st is not None:                                       # a larger batch than the ring was built for
                if pipe is not None:
                    pipe.flush_queued()
                torch.cuda.current_stream(self.device).synchronize()
                st.close()
            slot_bytes = max(nbytes, 1)
            n_slots = 3 * self.window_batches                        # three windows of batches, window-aligned
            if n_slots * slot_bytes > self.stage_bytes:             # big batches (full-resolution logits): a short ring
                n_slots = max(4, min(n_slots, self.stage_bytes // slot_bytes))
            st = self._stager = ops.Stager(n_slots, slot_bytes, self.device, _stream(self.device, 'copy'))
            self._stage_cursor = 0
        slot = self._stage_cursor % st.n_slots
        if st.busy(slot):
            if pipe is None:
                raise HiastError('staging ring exhausted')
            pipe.flush_queued()                                      # launches the queued phase A and releases its slots
        self._stage_cursor += 1
        view = st.push(slot, images, self._main_stream_ptr())
        if pipe is not None and pipe._ev_dir and 10 <= pipe.j <= 12:
            pipe._mark('h2d_copy', pipe.j, st.copy_stream)
        return view, slot

    def _main_stream_ptr(self):
        """The current stream of the device as a void* (looked up once per run: the lookup costs ~15 us in torch)."""
        p = getattr(self, '_main_ptr', None)
        if p is None:
            p = self._main_ptr = stream_ptr(self.device)
        return p

    def _release_staged(self, slots):
        st = self._stager
        if st is None or not slots:
            return
        main = self._main_stream_ptr()
        slots = sorted(slots)
        run0 = prev = slots[0]
        for s in slots[1:] + [None]:
            if s is None or s != prev + 1:
                st.release(run0, prev - run0 + 1, main)
                run0 = s
            prev = s

    def _already_done(self):
        return self.t_dataset is not None and len(os.listdir(self.pseudo_label_save_dir)) >= len(self.t_dataset)

    # --------------------------------------------------- one run of any policy
    def _run_windows(self, scan, thr_const=None, alpha=0.0, beta=0.0, gamma=1.0, rank=0, world=1, pg=None, n_total=None):
        """The loader through the window pipeline; fills every result attribute.  ``scan``: IAS thresholds (else the
        constant ``thr_const`` per class)."""
        C = self.cfg.dataset.num_classes
        pipe = None
        self._main_ptr = None
        it = self._iterate_logits(lambda: pipe)
        for logits, img_paths, slot in it:
            if pipe is None:
                engine = self._engine = self._make_engine(logits.shape, alpha, beta, gamma)
                pipe = _WindowPipeline(self, engine, scan, rank, world, pg, n_total)
                if scan:
                    engine.thr_state.copy_(torch.from_numpy(np.asarray(self.class_threshold, dtype=np.float64)))
                else:
                    engine.thr_groups.copy_(torch.from_numpy(np.tile(np.asarray(thr_const, dtype=np.float64), (engine.max_groups, 1))))
            pipe.add(logits, img_paths, slot)
        if pipe is None:                                  # this rank owns no window (or the loader is empty)
            engine = self._engine = self._make_engine((self._group_size(1), C, 4, 4), alpha, beta, gamma)
            pipe = _WindowPipeline(self, engine, scan, rank, world, pg, n_total)
            if scan:
                engine.thr_state.copy_(torch.from_numpy(np.asarray(self.class_threshold, dtype=np.float64)))
        pipe.finish()
        self._wait_png()
        import time
        t0 = time.perf_counter()
        self._collect(pipe, scan, rank, world, pg)
        self.pipeline_trace['collect'] = time.perf_counter() - t0

    def _iterate_logits(self, get_pipe=lambda: None):
        """Yields (logits [B,C,H,W] on the device or LowResLogits, image_paths, staging slot) as :189-192 produces them."""
        self.model.eval() if hasattr(self.model, 'eval') else None
        with torch.no_grad():
            for data in self.t_loader:
                imgs, slot = self._staged(data['images'], get_pipe())
                image_paths = list(data['image_paths'])
                out = self.model(imgs)
                if 'logits' not in out and 'logits_lr' in out:
                    # the network's own (stride-8) output: the bilinear up-sampling of
                    # self_training_segmentor.py:27 is fused into phase A
                    lr = out['logits_lr']
                    lr = lr.float() if lr.dtype != torch.float32 else lr
                    yield LowResLogits(lr.contiguous(), tuple(out.get('size', imgs.shape[2:]))), image_paths, slot
                    continue
                logits = out['logits']
                if logits.dtype != torch.float32:
                    logits = logits.float()
                yield logits.contiguous(), image_paths, slot

    def _collect(self, pipe, scan, rank, world, pg):
        """Global results on every rank: the windows' thresholds, class counts and confidence sums are merged in the global
        window order (ONE all_gather_object when R > 1), the mean-probability EMA (:95-105) is replayed over all groups, the
        per-image statistics lists (:82-89) are rebuilt in image order."""
        import torch.distributed as dist
        e = pipe.e
        C = self.cfg.dataset.num_classes
        self.pow_rounding_certified = e.check_errors() if hasattr(e, 'check_errors') else True      # host sync
        mine = {'windows': {r['w']: (r['paths'], r['counts'], r['confsum'], r['thr']) for r in pipe.results},
                'thr_state': torch.as_tensor(e.thr_state).cpu().numpy().copy() if scan else None}
        parts = [mine]
        if world > 1:
            parts = [None] * world
            dist.all_gather_object(parts, mine, group=pg)
        merged = {}
        for p in parts:
            merged.update(p['windows'])
        order = sorted(merged)
        if scan and order:                                # the final thresholds live where the last window was scanned
            last_owner = max(range(len(parts)), key=lambda r: max(parts[r]['windows'], default=-1))
            self.class_threshold = parts[last_owner]['thr_state']
        B = e.B
        self.threshold_trace = [merged[w][3] for w in order]
        group_counts, confsums, all_counts, all_paths = [], [], [], []
        for w in order:
            paths, counts, confsum, _ = merged[w]
            n = counts.shape[0]
            g = (n + B - 1) // B
            padded = np.zeros((g * B, C), dtype=np.int64)
            padded[:n] = counts
            group_counts.append(padded.reshape(g, B, C).sum(axis=1))
            confsums.append(confsum[:g])
            all_counts.append(counts)
            all_paths += paths
        counts_all = np.concatenate(all_counts) if all_counts else np.zeros((0, C), dtype=np.int64)
        self.statics_class = np.array([0] * C) + (counts_all.sum(axis=0) if len(all_paths) else 0)
        self._stats_pending = (counts_all, all_paths)            # sample_stats / samples_class: built on first access
        if order:
            dev = e.thr_state.device
            e.mean_state.copy_(torch.from_numpy(np.asarray(self.class_mean_probs, dtype=np.float64)))
            e.mean_prob_from_groups(torch.from_numpy(np.concatenate(confsums)).to(dev),
                                    torch.from_numpy(np.concatenate(group_counts)).to(dev))
            self.class_mean_probs = torch.as_tensor(e.mean_state).cpu().numpy().copy()


class _WindowPipeline:
    """Windows of ``engine.max_images // 3`` images through phase A -> threshold chain -> outputs, three in flight (module
    docstring).  ``add`` takes the batches of the rank's windows in order; ``finish`` drains.  With a host stand-in engine
    (CPU tests of the orchestration) the same schedule runs eagerly."""

    N_SLOTS = 3

    def __init__(self, gen, engine, scan, rank=0, world=1, pg=None, n_total=None):
        from .sharded import window_images
        self.gen, self.e, self.scan = gen, engine, bool(scan)
        self.rank, self.world, self.pg, self.n_total = rank, world, pg, n_total
        if world > 1 and n_total is None:
            raise RuntimeError('the sharded generator needs dataset_len (the size of the whole target set)')
        self.window = engine.max_images // self.N_SLOTS
        self._window_images = window_images
        self.n_windows_total = None if n_total is None else (n_total + self.window - 1) // self.window
        self.cuda = torch.is_tensor(engine.plbl) and engine.plbl.is_cuda
        self.ring = None
        if self.cuda and world > 1 and self.scan:
            from .sharded import TokenRing
            self.ring = TokenRing.get(engine.device, rank, world, pg)      # None: torch.distributed send / recv
        self.j = 0                    # local window being filled
        self.filled, self.paths, self.staged = 0, [], []
        self.queued = []              # stride-8 batches of the current window awaiting their single phase-A launch
        self.closed = []              # (n_images, paths) per closed local window
        self.b_done = self.c_done = 0
        self.results = []
        self.pending = {}
        self._host_jobs = {}
        # where the host's time goes (seconds): waiting for a window's completion, closing windows, everything else is the loader loop
        self.trace = {'wait_completion': 0.0, 'close': 0.0, 'chain_host': 0.0, 'flush_host': 0.0, 'windows': 0, 't0': None, 'total': 0.0}
        # HIAST_PIPE_EVENTS=<dir>: device timeline of every window (timing events on the three streams), dumped per rank
        self._ev_dir = os.environ.get('HIAST_PIPE_EVENTS')
        self._evs = []
        self.emitter = self.writer = None
        if self.cuda:
            dev = engine.device
            self.main = torch.cuda.current_stream(dev)
            self.side = _stream(dev, 'chain')
            self.ev_a = [torch.cuda.Event() for _ in range(self.N_SLOTS)]
            self.ev_b = [torch.cuda.Event() for _ in range(self.N_SLOTS)]
            self.out = _stream(dev, 'out')             # device-to-host copies of a window's results
            self.out_ptr = ops.C.c_void_p(self.out.cuda_stream)
            self.ev_out = [None] * self.N_SLOTS
            if gen._device_png() and engine.W > 128 * 256:
                gen.png = 'host'                       # wider than the device writer's 32768-pixel rows: the reference's writer
            self.mode = 'host' if not gen._device_png() else ('files' if gen._native_files() else 'blobs')
            self.emitter = ops.WindowEmitter.get(engine, self.window, self.N_SLOTS, png=self.mode != 'host')
            self.main_ptr = stream_ptr(dev)
            if self.mode == 'files':
                if gen._writer is None:
                    gen._writer = ops.FileWriter(gen._png_workers, dev)
                self.writer = gen._writer

    # ---------------------------------------------------------------- input side
    def _global(self, j):
        return j * self.world + self.rank

    def _expected(self):
        if self.n_total is None:
            return self.window
        return self._window_images(self._global(self.j), self.window, self.n_total)[1]

    def add(self, logits, paths, staged_slot=None):
        e = self.e
        b = logits.shape[0]
        want = self._expected()
        if self.filled + b > want:
            raise ValueError('window %d must hold %d images, the loader delivered more' % (self._global(self.j), want))
        first = (self.j % self.N_SLOTS) * self.window + self.filled
        if isinstance(logits, LowResLogits) and hasattr(e, 'phase_a_lowres'):
            q = self.queued
            if q and (q[-1][0].size != logits.size or q[-1][0].logits_lr.shape[1:] != logits.logits_lr.shape[1:]):
                self.flush_queued()
            self.queued.append((logits, first))
            if staged_slot is not None:
                self.staged.append(staged_slot)
        else:
            if isinstance(logits, LowResLogits):                    # an engine without the fused up-sampling (host stand-in)
                logits = torch.nn.functional.interpolate(logits.logits_lr, size=logits.size, mode='bilinear', align_corners=True)
            e.phase_a(logits, first)
            if staged_slot is not None:
                self.gen._release_staged([staged_slot])             # consumed by the launch just queued
        if self._ev_dir and self.cuda and self.filled == 0 and self.gen._stager is not None:
            self._mark('h2d_first_done', self.j, self.gen._stager.copy_stream)
        self.filled += b
        self.paths += paths
        if self.filled == want or (self.n_total is None and b != e.B):
            self._close()

    def flush_queued(self):
        """ONE launch of the fused up-sampling kernel for the queued stride-8 batches (a 2-image launch costs 106 us, its share
        of a 64-image launch 67 us).  Batches staged into consecutive ring slots are passed as one view, without a copy."""
        q = self.queued
        if q:
            self.queued = []
            first = q[0][1]
            ts = [x[0].logits_lr for x in q]
            lr = ts[0]
            if len(ts) > 1:
                lr = None
                st = self.gen._stager
                step = ts[0].numel() * 4
                if st is not None and all(t.data_ptr() == ts[0].data_ptr() + k * step and t.shape == ts[0].shape
                                          for k, t in enumerate(ts)):
                    lr = st.view_of(ts[0].data_ptr(), step * len(ts), torch.float32, (len(ts) * ts[0].shape[0],) + tuple(ts[0].shape[1:]))
                if lr is None:
                    lr = torch.cat(ts)
            self.e.phase_a_lowres(lr, first)
        if self.staged:
            self.gen._release_staged(self.staged)
            self.staged = []

    def _close(self):
        import time
        t_close = time.perf_counter()
        if self.trace['t0'] is None:
            self.trace['t0'] = t_close
        self.trace['windows'] += 1
        self._close_inner()
        self.trace['close'] += time.perf_counter() - t_close

    def _mark(self, name, j, stream):
        if self._ev_dir and self.cuda:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record(stream)
            self._evs.append((name, j, ev))

    def _close_inner(self):
        if self._ev_dir and self.cuda:
            if self.gen._stager is not None:
                self._mark('h2d_last_done', self.j, self.gen._stager.copy_stream)
            self._mark('A_begin', self.j, self.main)
        import time
        t_f = time.perf_counter()
        self.flush_queued()
        self.trace['flush_host'] += time.perf_counter() - t_f
        if self._ev_dir and self.cuda:
            self._mark('A_end', self.j, self.main)
        if self.cuda:
            self.ev_a[self.j % self.N_SLOTS].record(self.main)
        self.closed.append((self.filled, self.paths))
        self.j += 1
        self.filled, self.paths = 0, []
        self._advance(final=False)
        if self.cuda and self.scan and self.b_done:
            # The next phase A starts only when the chain just queued has finished.  On one rank that costs nothing (the scan
            # is shorter than the outputs it runs beside).  On R ranks the token reaches rank r about r hops after the ranks'
            # phase A's end together; without this wait the next phase A would take every SM while the receive kernel still
            # spins on one of them, and the scan behind it would crawl on that single SM with phase A's last CTA queued behind
            # it (measured at R = 4: phase A 2.47 ms instead of 1.77).  With it rank r falls r hops behind ONCE and from then
            # on every token is already there when its window's outputs end: the ranks run staggered and nobody waits.
            self.main.wait_event(self.ev_b[(self.b_done - 1) % self.N_SLOTS])

    def finish(self):
        if self.filled:
            self._close()
        if self.n_total is not None:
            from .sharded import local_windows
            mine = len(local_windows(self.n_windows_total, self.rank, self.world))
            if len(self.closed) != mine:
                raise ValueError('this rank owns %d windows, the loader filled %d' % (mine, len(self.closed)))
        self._advance(final=True)
        for es in sorted(self.pending, key=lambda k: self.pending[k]['j']):
            self._complete(es)
        for es in list(self._host_jobs):
            self._wait_host_jobs(es)
        if self.emitter is not None:
            self.gen._wait_png()
            self.emitter.done()
        if self.ring is not None:
            self.ring.advance(self.n_windows_total)
        import time
        if self.trace['t0'] is not None:
            self.trace['total'] = time.perf_counter() - self.trace['t0']
        self.gen.pipeline_trace = dict(self.trace)
        if self._ev_dir and self._evs:
            torch.cuda.synchronize()
            base = self._evs[0][2]
            rows = [(name, j, round(base.elapsed_time(ev), 3)) for name, j, ev in self._evs]
            os.makedirs(self._ev_dir, exist_ok=True)
            with open(os.path.join(self._ev_dir, 'pipe_events_rank%d_%d.json' % (self.rank, len(rows))), 'w') as f:
                json.dump(rows, f)

    # ---------------------------------------------------------------- the schedule
    def _advance(self, final):
        n_closed = len(self.closed)
        b_target = n_closed if (final or not self.scan) else n_closed - 1
        while self.b_done < b_target:
            if self.scan:
                import time
                t_c = time.perf_counter()
                self._chain(self.b_done)
                self.trace['chain_host'] += time.perf_counter() - t_c
            self.b_done += 1
        c_target = self.b_done if (final or not self.scan) else self.b_done - 1
        while self.c_done < c_target:
            self._outputs(self.c_done)
            self.c_done += 1

    def _chain(self, j):
        """Token receive -> threshold scan of window j -> token send, on the chain stream behind the LATEST closed phase A
        (so that it does not compete with a running phase A for SMs but runs beside the outputs of the window before)."""
        import torch.distributed as dist
        e = self.e
        n = self.closed[j][0]
        slot = (j % self.N_SLOTS) * self.window
        w = self._global(j)
        ring = self.world > 1

        def body():
            if self.ring is not None:                 # hand-off fused into the scan kernel, over peer memory
                e.phase_b(slot, n, token=self.ring.token(w, self.n_windows_total))
                return
            if ring and w > 0:
                dist.recv(e.thr_state, src=self._peer(-1), group=self.pg)
            e.phase_b(slot, n)
            if ring and w < self.n_windows_total - 1:
                dist.send(e.thr_state, dst=self._peer(+1), group=self.pg)

        if not self.cuda:
            body()
            return
        self.side.wait_event(self.ev_a[(len(self.closed) - 1) % self.N_SLOTS])
        if self.ev_out[j % self.N_SLOTS] is not None:        # window j-3's thresholds have been copied out of this slot
            self.side.wait_event(self.ev_out[j % self.N_SLOTS])
        with torch.cuda.stream(self.side):
            self._mark('chain_begin', j, self.side)
            body()
            self._mark('chain_end', j, self.side)
            self.ev_b[j % self.N_SLOTS].record(self.side)

    def _peer(self, step):
        import torch.distributed as dist
        r = (self.rank + step) % self.world
        return dist.get_global_rank(self.pg, r) if self.pg is not None else r

    def _outputs(self, j):
        """Phase C, the PNG encoder and the device-to-host copies of window j (one foreign call), then the hand-over to the
        writer pool; nothing here waits for the GPU except the reuse of an emit slot three windows later."""
        e, gen = self.e, self.gen
        n, paths = self.closed[j]
        first = (j % self.N_SLOTS) * self.window
        w = self._global(j)
        if not self.cuda:                                   # host stand-in engine: eager
            e.phase_c(first, n)
            g0, g = first // e.B, (n + e.B - 1) // e.B
            plbl = np.asarray(e.plbl[first:first + n])
            for i in range(n):
                gen._save_async(plbl[i].copy(), paths[i])
            self.results.append(dict(w=w, paths=paths, counts=torch.as_tensor(e.counts[first:first + n]).numpy().copy(),
                                     confsum=torch.as_tensor(e.confsum[g0:g0 + g]).numpy().copy(),
                                     thr=torch.as_tensor(e.thr_groups[g0:g0 + g]).numpy().copy()))
            return
        es = j % self.N_SLOTS
        if es in self.pending:
            self._complete(es)
        self._wait_host_jobs(es)
        if self.scan:
            self.main.wait_event(self.ev_b[es])
        if self.ev_out[es] is not None:                      # the copies of window j-3 have left this slot's device buffers
            self.main.wait_event(self.ev_out[es])
        self._mark('emit_begin', j, self.main)
        copied = self.emitter.emit(es, first, n, stream=self.main_ptr, copy_stream=self.out_ptr)
        self._mark('emit_end', j, self.main)
        self._mark('copy_end', j, self.out)
        s = self.emitter.slots[es]
        rec = dict(j=j, w=w, n=n, paths=paths, copied=copied)
        if self.mode == 'files':
            targets = [gen._pseudo_label_path(p) for p in paths]
            rec['ticket'] = self.writer.submit(targets, s['blob_host'], s['offsets_host'], copied, s['blob_dev'], self.out_ptr)
        else:
            ev = rec['event'] = torch.cuda.Event(blocking=True)
            ev.record(self.out)
        if self.ev_out[es] is None:
            self.ev_out[es] = torch.cuda.Event()
        self.ev_out[es].record(self.out)
        self.pending[es] = rec
        if not gen.defer_sync:
            self._complete(es)

    def _complete(self, es):
        import time
        t_wait = time.perf_counter()
        try:
            self._complete_inner(es)
        finally:
            self.trace['wait_completion'] += time.perf_counter() - t_wait

    def _complete_inner(self, es):
        rec = self.pending.pop(es)
        e, gen, s = self.e, self.gen, self.emitter.slots[es]
        n, paths = rec['n'], rec['paths']
        g = (n + e.B - 1) // e.B
        if 'ticket' in rec:
            self.writer.wait(rec['ticket'])                 # foreign call: sleeps until the window's files are on disk
            self.emitter.learn(int(s['offsets_host'][n]))
        else:
            rec['event'].synchronize()
            if self.mode == 'blobs':                        # save_pseudo_label_file hook: finished files one by one
                o = s['offsets_host'][:n + 1].tolist()
                if o[-1] > rec['copied']:
                    s['blob_host'][rec['copied']:o[-1]].copy_(s['blob_dev'][rec['copied']:o[-1]])
                self.emitter.learn(o[-1])
                blob = s['blob_host'].numpy()
                for i in range(n):
                    gen.save_pseudo_label_file(blob[o[i]:o[i + 1]], paths[i])
            else:                                           # save_pseudo_label hook / cv2.imwrite on uint8 label maps
                plbl = s['plbl_host'].numpy()
                jobs = [gen._save_async(plbl[i], paths[i]) for i in range(n)]
                self._host_jobs[es] = [job for job in jobs if job is not None]
        self.results.append(dict(w=rec['w'], paths=paths, counts=s['counts_host'][:n].numpy().copy(),
                                 confsum=s['confsum_host'][:g].numpy().copy(), thr=s['thr_host'][:g].numpy().copy()))

    def _wait_host_jobs(self, es):
        for job in self._host_jobs.pop(es, []):             # cv2 encoders still reading the slot's pinned label maps
            job.result()


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
        self._run_windows(scan=False, thr_const=thr)
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
        for logits, _paths, slot in self._iterate_logits():
            b = logits.shape[0]
            if isinstance(logits, LowResLogits):                        # phase A (its own histogram is not used here)
                conf, label, _ = ops.ias_upsample_softmax_hist(logits.logits_lr, logits.size, b)
            else:
                conf, label, _ = ops.ias_softmax_hist(logits, b)
            if slot is not None:
                self._release_staged([slot])
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

    def _ranks(self):
        return 0, 1, None

    def run(self):
        """:181-213"""
        if self._already_done():
            print('%% pseudo labels have existed')
            return
        rank, world, pg = self._ranks()
        C = self.cfg.dataset.num_classes
        ias = self.cfg.pseudo_policy.ias
        self.class_threshold = 0.9 * np.ones(C)                                            # :185
        n_total = None                                   # one rank: windows simply close when they are full
        if world > 1:
            if self.t_dataset is None:
                raise RuntimeError('the sharded generator needs dataset_len (the size of the whole target set)')
            n_total = len(self.t_dataset)
        self._run_windows(scan=True, alpha=ias.alpha, beta=ias.beta, gamma=ias.gamma, rank=rank, world=world, pg=pg,
                          n_total=n_total)
        if rank == 0:
            self.save_data()


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
    batch sampler).  Every rank runs the three-windows-in-flight pipeline of this module; the 19-double threshold state is
    received from rank r-1 in front of a window's scan and sent on to rank r+1 behind it (NCCL / gloo), on the chain
    stream.  At the end ONE all_gather_object merges the windows' thresholds, counts and confidence sums, every rank replays
    the mean-probability EMA and rebuilds the statistics lists in the global image order, and rank 0 writes ``save_data``'s
    files.  A rank that owns no window (more ranks than windows) takes part in the collectives with an empty share.
    Results are bit-identical to the single-process generator (tests/test_sharded_gloo.py with a host stand-in engine,
    tests/test_sharded_gpu.py with NCCL)."""

    def __init__(self, cfg, *args, engine_factory=None, process_group=None, **kw):
        import torch.distributed as dist
        self._engine_factory = engine_factory
        self._pg = process_group
        super().__init__(cfg, *args, **kw)
        if dist.is_available() and dist.is_initialized():
            dist.barrier(group=process_group)          # every rank has seen the empty save dir before anyone writes

    def _ranks(self):
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(self._pg), dist.get_world_size(self._pg), self._pg
        return 0, 1, None

    def _build_target_loader(self):
        """The target set in its pinned order, this rank's windows only (``striped_batch_order`` as the batch sampler)."""
        from torch.utils.data import DataLoader
        cfg = self.cfg
        ds = self._build_target_dataset()
        rank, world, _ = self._ranks()
        b = cfg.pseudo_policy.batch_size
        order = striped_batch_order(len(ds), b * self.window_batches, b, rank, world)
        return ds, DataLoader(ds, batch_sampler=order, num_workers=cfg.dataset.num_workers, pin_memory=True)
