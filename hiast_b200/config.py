"""Attribute-style configuration tree with the reference's defaults and merge rules.

The reference builds ``cfg`` with yacs (``utils/default_config.py:1-190``, /root/reference/code): a tree of defaults that yaml
files are merged into (``cfg.merge_from_file``) and that is frozen before use (``generate_pseudo_labels.py:21-40``).  yacs is
not a dependency of this package, so the three behaviours the call surface relies on are restated here on PyYAML:

* merging accepts only keys that already exist in the defaults (yacs: ``KeyError: Non-existent config key``);
* a string read from yaml is decoded as a Python literal when it is one (yacs' ``_decode_cfg_value``: PyYAML reads
  ``lr: 3e-6`` of ``configs/sl_1.yaml`` as the string '3e-6'), and a merged value must keep the type of its default
  unless the default is ``None`` (int -> float and list <-> tuple are coerced, as yacs does);
* a frozen tree rejects assignment.

``default_cfg()`` returns a fresh tree holding every key and default value of ``utils/default_config.py`` -- the key names and
defaults ARE the interface of the reference's yaml files (``configs/sl_1.yaml`` etc. merge without change).
"""

from __future__ import annotations

import ast
import copy

import yaml


class CfgNode(dict):
    _IMMUTABLE = '__immutable__'

    def __init__(self, init=None):
        super().__init__()
        self.__dict__[CfgNode._IMMUTABLE] = False
        for k, v in (init or {}).items():
            self[k] = CfgNode(v) if isinstance(v, dict) and not isinstance(v, CfgNode) else v

    # ------------------------------------------------------------ attribute access
    def __getattr__(self, name):
        if name in self:
            return self[name]
        raise AttributeError(name)

    def __setattr__(self, name, value):
        if self.is_frozen():
            raise AttributeError('Attempted to set {} to {}, but CfgNode is immutable'.format(name, value))
        self[name] = value

    # ------------------------------------------------------------ freezing
    def is_frozen(self):
        return self.__dict__[CfgNode._IMMUTABLE]

    def _set_immutable(self, flag):
        self.__dict__[CfgNode._IMMUTABLE] = flag
        for v in self.values():
            if isinstance(v, CfgNode):
                v._set_immutable(flag)

    def freeze(self):
        self._set_immutable(True)

    def defrost(self):
        self._set_immutable(False)

    def clone(self):
        return copy.deepcopy(self)

    def __deepcopy__(self, memo):
        out = CfgNode()
        for k, v in self.items():
            dict.__setitem__(out, k, copy.deepcopy(v, memo))
        out.__dict__[CfgNode._IMMUTABLE] = self.is_frozen()
        return out

    # ------------------------------------------------------------ merging
    def merge_from_file(self, path):
        with open(path, 'r') as f:
            loaded = yaml.safe_load(f) or {}
        self.merge_from_dict(loaded)

    def merge_from_dict(self, other, _trail=()):
        if self.is_frozen():
            raise AttributeError('CfgNode is immutable')
        for k, v in other.items():
            full = '.'.join(_trail + (str(k),))
            if k not in self:
                raise KeyError('Non-existent config key: {}'.format(full))
            cur = self[k]
            if isinstance(cur, CfgNode):
                if not isinstance(v, dict):
                    raise ValueError('Type mismatch for config key {}: a section was replaced by {!r}'.format(full, v))
                cur.merge_from_dict(v, _trail + (str(k),))
            else:
                dict.__setitem__(self, k, _coerce(_decode(v), cur, full))


def _decode(v):
    """yacs' ``_decode_cfg_value``: strings that are Python literals ('3e-6', '[1, 2]', 'None') become those values."""
    if not isinstance(v, str):
        return v
    try:
        return ast.literal_eval(v)
    except (ValueError, SyntaxError):
        return v


def _coerce(new, old, key):
    """yacs' ``_check_and_coerce_cfg_value_type``: same type, or a default of None, or one of the tolerated casts."""
    if old is None or new is None or type(new) is type(old):
        return new
    if isinstance(old, float) and isinstance(new, int) and not isinstance(new, bool):
        return float(new)
    if isinstance(old, tuple) and isinstance(new, list):
        return tuple(new)
    if isinstance(old, list) and isinstance(new, tuple):
        return list(new)
    raise ValueError('Type mismatch ({} vs. {}) with values ({} vs. {}) for config key: {}'.format(
        type(old), type(new), old, new, key))


# Every key and default of utils/default_config.py (line numbers of the reference file in the comments).
_DEFAULTS = {
    'trainer': None,                                            # :4
    'work_dir': './',                                           # :5
    'model': {                                                  # :9-45
        'type': None,
        'is_freeze_bn': True,
        'seg_model': {'type': 'DeepLab_V2', 'output_dim': 256},
        'predictor': {
            'seg_loss': {'type': 'CE', 'source_weight': 1.0, 'target_pseudo_weight': 1.0},
            'kld_loss': {'weight': 0.1},
            'ent_loss': {'weight': 3.0},
        },
        'discriminator': {
            'is_enabled': False, 'is_entropy_input': False, 'lr': 1e-4,
            'D_loss': {'type': 'MSE', 'weight': 1.0, 'adv_weight': 0.05},
        },
    },
    'dataset': {                                                # :50-76
        'num_classes': 19,
        'num_workers': 2,
        'source': {'type': None, 'json_path': None, 'image_dir': None, 'aug_type': []},
        'target': {'type': None, 'json_path': None, 'image_dir': None, 'pseudo_dir': None, 'aug_type': []},
        'val': {'type': None, 'json_path': None, 'image_dir': None, 'resize_size': None},
    },
    'pseudo_policy': {                                          # :81-101
        'resume_from': None,
        'batch_size': 2,
        'resize_size': None,
        'save_dir': None,
        'type': None,
        'ias': {'alpha': 0.2, 'beta': 0.9, 'gamma': 8.0},
        'cbst': {'p': 0.2, 'sample_interval': 4},
        'ct': {'threshold': 0.9},
    },
    'train': {                                                  # :106-130
        'batch_size': 4, 'lr': 1e-4, 'optimizer': 'Adam', 'resume_from': None, 'apex_opt': 'O1', 'gpu_num': 2,
        'random_seed': 888, 'port': 6789, 'is_save_all': False, 'is_debug': False,
        'total_iter': 10000, 'iter_report': 100, 'iter_val': 400,
        'lr_scheduler': {'type': 'Cosine', 'poly': {'power': 0.9}},
    },
    'validate': {                                               # :135-140
        'resume_from': None, 'resize_sizes': [], 'is_flip': False, 'batch_size': 2, 'color_mask_dir_path': None,
    },
    'cst_training': {                                           # :145-157
        'is_enabled': False,
        'ema_model': {'iter_update': 1, 'gamma': 0.999},
        'cst_loss': {'type': 'SoftCE', 'weight': 1.0, 'region': 'ignored'},
    },
    'mut_training': {                                           # :162-170
        'is_enabled': False, 'resume_from': None, 'is_strong_input': False,
        'mut_loss': {'weight': 0.1, 'region': 'ignored'},
    },
    'preprocessor': {                                           # :175-185
        'type': None,
        'copy_paste': {'mode': 'original', 'name': 'normal', 'selected_num_classes': 14, 'gamma': 0.99},
    },
}


def default_cfg():
    """A fresh, unfrozen tree with the reference's defaults (``from utils.default_config import cfg``)."""
    return CfgNode(copy.deepcopy(_DEFAULTS))
