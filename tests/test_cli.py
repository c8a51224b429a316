"""The script-level call surface (generate_pseudo_labels.py:8-48 of the reference): flags, yaml merge order, the frozen tree,
the one-argument ``PSEUDO_POLICY[type](cfg)`` path through the MODEL / DATASET registries.

``tests/golden/cli_config.json`` was written by the reference's OWN ``utils/default_config.py`` and
``generate_pseudo_labels.update_cfg`` running unmodified on ``hiast_b200.config.CfgNode`` (standing in for yacs, which is not
installed) over the reference's shipped yaml files (tests/golden/make_golden.py ``cli_config``)."""

import json
import os

import numpy as np
import pytest
import torch

from hiast_b200 import cli
from hiast_b200.config import CfgNode, default_cfg

GOLD = os.path.join(os.path.dirname(__file__), 'golden')
SL1 = """
trainer: 'SelfTrainingTrainer'
work_dir: '../log/x'
model:
  type: 'ToySegmentor'
dataset:
  num_classes: 7
  num_workers: 0
  target:
    type: 'ToyTarget'
    json_path: 'toy.json'
    image_dir: 'toy'
pseudo_policy:
  batch_size: 2
  resize_size: [ 24, 40 ]
  type: 'IAS'
  ias:
    alpha: 0.5
    beta: 0.9
    gamma: 8.0
train:
  lr: 3e-6
"""
SETTING = """
pseudo_policy:
  ias:
    alpha: 0.4
preprocessor:
  copy_paste:
    gamma: 0.98
"""


def plain(node):
    return json.loads(json.dumps(node))


def test_default_tree_is_the_reference_default_config():
    gold = json.load(open(os.path.join(GOLD, 'cli_config.json')))
    assert plain(default_cfg()) == gold['defaults']
    assert gold['batch_size_flag'] == 'AttributeError'           # generate_pseudo_labels.py:30 on the reference itself


def write(tmp_path, name, text):
    p = tmp_path / name
    p.write_text(text)
    return str(p)


def test_merge_order_flags_and_freeze(tmp_path):
    a = write(tmp_path, 'sl.yaml', SL1)
    b = write(tmp_path, 'setting.yaml', SETTING)
    args = cli.parse_args(['--config_file', a, '--setting_file', b, '--pseudo_resume_from', 'ckpt.pth', '--pseudo_save_dir', 'out/pl'])
    cfg = cli.update_cfg(default_cfg(), args)
    assert cfg.pseudo_policy.ias.alpha == 0.4 and cfg.pseudo_policy.ias.gamma == 8.0        # setting file over config file
    assert cfg.preprocessor.copy_paste.gamma == 0.98 and cfg.preprocessor.copy_paste.selected_num_classes == 14
    assert cfg.pseudo_policy.resume_from == 'ckpt.pth' and cfg.pseudo_policy.save_dir == 'out/pl'   # flags over files
    assert cfg.train.lr == 3e-6 and isinstance(cfg.train.lr, float)                             # yacs literal decoding
    assert cfg.dataset.num_classes == 7 and cfg.model.seg_model.type == 'DeepLab_V2'
    assert cfg.is_frozen()
    with pytest.raises(AttributeError):
        cfg.pseudo_policy.save_dir = 'elsewhere'
    with pytest.raises(AttributeError):
        cfg.pseudo_policy.ias.alpha = 0.1


def test_reference_quirks_are_kept(tmp_path):
    a = write(tmp_path, 'sl.yaml', SL1)
    with pytest.raises(AttributeError):                          # :30 reads cfg.batch_size, which does not exist
        cli.update_cfg(default_cfg(), cli.parse_args(['--config_file', a, '--batch_size', '4']))
    bad = write(tmp_path, 'bad.yaml', 'pseudo_policy:\n  no_such_key: 1\n')
    with pytest.raises(KeyError, match='Non-existent config key: pseudo_policy.no_such_key'):
        cli.update_cfg(default_cfg(), cli.parse_args(['--config_file', bad]))
    typed = write(tmp_path, 'typed.yaml', 'pseudo_policy:\n  batch_size: "two"\n')
    with pytest.raises(ValueError, match='Type mismatch'):
        cli.update_cfg(default_cfg(), cli.parse_args(['--config_file', typed]))
    with pytest.raises(SystemExit):                              # --config_file is required (:10)
        cli.parse_args([])
    n = CfgNode({'a': {'b': 1.0}})
    n.merge_from_dict({'a': {'b': 2}})                           # int -> float is coerced like yacs
    assert n.a.b == 2.0 and isinstance(n.a.b, float)


def test_one_argument_constructor_resolves_model_and_dataset_through_the_registries(tmp_path, monkeypatch):
    """PSEUDO_POLICY[type](cfg) with ONE argument (generate_pseudo_labels.py:47): initialize() builds the model with
    MODEL[cfg.model.type](cfg) + the checkpoint of pseudo_policy.resume_from (utils/utils.py:68-89) and the target loader
    with DATASET[cfg.dataset.target.type](cfg, json, dir, aug_type=['PRS-h-w'], num_classes=C) (:29-36).  CPU: construction
    only (run() needs the GPU; tests/test_generator_gpu.py::test_cli_script_end_to_end runs it)."""
    import hiast_b200
    from hiast_b200.registry import DATASET, MODEL, PSEUDO_POLICY
    hiast_b200.register_all()
    seen = {}

    class ToySegmentor(torch.nn.Module):
        def __init__(self, cfg):
            super().__init__()
            self.head = torch.nn.Conv2d(3, cfg.dataset.num_classes, 1)

        def forward(self, x):
            return {'logits': self.head(x)}

    class ToyTarget(torch.utils.data.Dataset):
        def __init__(self, cfg, json_path, image_dir, aug_type=None, num_classes=None):
            seen.update(json_path=json_path, image_dir=image_dir, aug_type=aug_type, num_classes=num_classes)

        def __len__(self):
            return 5

        def __getitem__(self, i):
            return {'images': torch.zeros(3, 24, 40), 'image_paths': 'img_%d.png' % i}

    monkeypatch.setitem(MODEL, 'ToySegmentor', ToySegmentor)
    monkeypatch.setitem(DATASET, 'ToyTarget', ToyTarget)
    ref = ToySegmentor(default_cfg().clone() if False else CfgNode({'dataset': {'num_classes': 7}}))
    ckpt = str(tmp_path / 'ckpt.pth')
    torch.save({'module.' + k: v for k, v in ref.state_dict().items()}, ckpt)      # saved from DDP: 'module.' prefix (:78-79)
    a = write(tmp_path, 'sl.yaml', SL1)
    args = cli.parse_args(['--config_file', a, '--pseudo_resume_from', ckpt, '--pseudo_save_dir', str(tmp_path / 'run' / 'pl')])
    cfg = cli.update_cfg(default_cfg(), args)
    gen = PSEUDO_POLICY[cfg.pseudo_policy.type](cfg, device='cpu')
    assert seen == dict(json_path='toy.json', image_dir='toy', aug_type=['PRS-24-40'], num_classes=7)
    assert isinstance(gen.model, ToySegmentor) and torch.equal(gen.model.head.weight, ref.head.weight)
    assert len(gen.t_dataset) == 5 and gen.t_loader.batch_size == 2
    assert os.path.isdir(str(tmp_path / 'run' / 'pl'))           # :38-41
    with pytest.raises(RuntimeError, match="MODEL\\['Nope'\\] is not registered"):
        PSEUDO_POLICY['IAS'](CfgNode({'dataset': {'num_classes': 7}, 'model': {'type': 'Nope'},
                                      'pseudo_policy': {'resume_from': None, 'save_dir': str(tmp_path / 'x')}}), device='cpu')


def test_install_into_overriding_keys_keeps_the_reference_script_working():
    """registry.install_into(reference_registries, suffix='') overrides 'IAS', 'SelfTrainingSegmentor', ... in the
    reference's registries AND brings the reference's backbone / dataset entries into this package's, so that
    MODEL['SelfTrainingSegmentor'](cfg) builds its backbone through SEG_MODEL like build_seg_model(cfg) does."""
    from types import SimpleNamespace
    import hiast_b200
    from hiast_b200 import registry as ours

    class Backbone(torch.nn.Module):
        def __init__(self, num_classes, output_dim):
            super().__init__()
            self.args = (num_classes, output_dim)

    ref = SimpleNamespace(**{name: ours.Registry() for name in ours._ALL})
    ref.SEG_MODEL.register('ToyNet_for_install_test', Backbone)
    ref.MODEL.register('SelfTrainingSegmentor', object)           # the reference's own entry, to be overridden
    ref.PSEUDO_POLICY.register('IAS', object)
    try:
        ours.install_into(ref, suffix='')
        assert ref.PSEUDO_POLICY['IAS'] is ours.PSEUDO_POLICY['IAS'] and ref.PSEUDO_POLICY['IAS'] is not object
        assert ref.MODEL['SelfTrainingSegmentor'] is ours.MODEL['SelfTrainingSegmentor']
        assert ours.SEG_MODEL['ToyNet_for_install_test'] is Backbone
        cfg = default_cfg()
        cfg.model.seg_model.type = 'ToyNet_for_install_test'
        cfg.dataset.num_classes = 7
        seg = ref.MODEL['SelfTrainingSegmentor'](cfg)
        assert isinstance(seg.seg_model, Backbone) and seg.seg_model.args == (7, 256)
    finally:
        dict.pop(ours.SEG_MODEL, 'ToyNet_for_install_test', None)
