"""The reference's pseudo-label script on the B200 path: same flags, same configuration merge, same one-line run.

Mirrors ``generate_pseudo_labels.py`` (reference, /root/reference/code): ``parse_args`` :8-18, ``update_cfg`` :21-40, the run
:43-48.  ``python generate_pseudo_labels.py --config_file configs/sl_1.yaml --pseudo_resume_from ckpt.pth --pseudo_save_dir out/``
builds the configuration tree (defaults < ``--config_file`` < ``--setting_file`` < flags), freezes it and calls
``PSEUDO_POLICY[cfg.pseudo_policy.type](cfg).run()``.  The generator resolves the model through ``MODEL[cfg.model.type]`` and the
target set through ``DATASET[cfg.dataset.target.type]`` exactly as the reference's ``initialize`` does; the backbone and the
dataset classes are the host project's (``--register some.module`` imports a module that registers them, the one extra flag).

Under ``torchrun`` (WORLD_SIZE > 1) the process group is initialised here and ``IAS`` runs as ``IAS_SHARDED``: one rank per GPU,
windows of the pinned dataset order striped over the ranks (no shuffling: the reference's unseeded ``shuffle=True`` makes its
thresholds order dependent).

Reference quirk kept: ``--batch_size`` raises AttributeError, because :30 reads ``cfg.batch_size``, which does not exist.
"""

from __future__ import annotations

import argparse
import importlib
import os
import sys

from .config import default_cfg
from .registry import PSEUDO_POLICY, SEG_MODEL, register_all


def parse_args(argv=None):
    """:8-18 (+ ``--register``)."""
    parser = argparse.ArgumentParser()
    parser.add_argument('--config_file', required=True)
    parser.add_argument('--setting_file')
    parser.add_argument('--pseudo_resume_from')
    parser.add_argument('--pseudo_save_dir')
    parser.add_argument('--batch_size', type=int)
    parser.add_argument('--seg_model', choices=list(SEG_MODEL.keys()) or None)
    parser.add_argument('--register', action='append', default=[],
                        help='module to import before the run: registers the host project\'s MODEL / SEG_MODEL / DATASET entries')
    return parser.parse_args(argv)


def update_cfg(cfg, args):
    """:21-40, statement for statement (the order is the precedence)."""
    cfg.merge_from_file(args.config_file)
    if args.setting_file:
        cfg.merge_from_file(args.setting_file)
    if args.pseudo_resume_from:
        cfg.pseudo_policy.resume_from = args.pseudo_resume_from
    if args.batch_size:
        cfg.pseudo_policy.batch_size = cfg.batch_size          # :30 -- cfg.batch_size does not exist: AttributeError, as upstream
    if args.pseudo_save_dir:
        cfg.pseudo_policy.save_dir = args.pseudo_save_dir
    if args.seg_model:
        cfg.model.seg_model.type = args.seg_model
    cfg.freeze()
    return cfg


def main(argv=None):
    """:43-48"""
    register_all()
    argv_list = list(argv) if argv is not None else None
    # host-project modules first, so that --seg_model's choices see their SEG_MODEL entries
    scan = argv_list if argv_list is not None else sys.argv[1:]
    for i, tok in enumerate(scan):
        if tok == '--register' and i + 1 < len(scan):
            importlib.import_module(scan[i + 1])
        elif tok.startswith('--register='):
            importlib.import_module(tok.split('=', 1)[1])
    args = parse_args(argv_list)
    cfg = update_cfg(default_cfg(), args)
    policy = cfg.pseudo_policy.type
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if world > 1:
        import torch
        import torch.distributed as dist
        local = int(os.environ.get('LOCAL_RANK', '0'))
        torch.cuda.set_device(local)
        if not dist.is_initialized():
            os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
            dist.init_process_group('nccl', device_id=torch.device('cuda', local))
        if policy == 'IAS':
            policy = 'IAS_SHARDED'
    pseudo_generator = PSEUDO_POLICY[policy](cfg)
    pseudo_generator.run()
    return pseudo_generator
