"""String-keyed registries: the reference's plugin API and therefore the drop-in boundary.

Mirrors ``utils/registry/registry.py:6-43`` / ``registries.py:3-9`` (reference,
/root/reference/code): a dict subclass whose ``register(name)`` works as a call or a decorator and
refuses duplicate keys.  The B200 implementations are registered under the reference's own key
names, so ``PSEUDO_POLICY[cfg.pseudo_policy.type](cfg)``, ``LOSS[cfg...type]``,
``MODEL[cfg.model.type](cfg)`` and ``PREPROCESSOR['CopyPaste'](...)`` resolve to them when a script
imports ``hiast_b200.registry`` in place of ``utils.registry.registries``
(``install_into(reference_registries)`` does the same inside a live reference checkout).
"""

from __future__ import annotations


class Registry(dict):
    def register(self, module_name, module=None):
        def _add(obj):
            assert module_name not in self, '%r is already registered' % (module_name,)
            self[module_name] = obj
            return obj

        if module is not None:
            _add(module)
            return None
        return _add


LOSS = Registry()
DATASET = Registry()
MODEL = Registry()
TRAINER = Registry()
PSEUDO_POLICY = Registry()
PREPROCESSOR = Registry()
SEG_MODEL = Registry()

_ALL = {'LOSS': LOSS, 'DATASET': DATASET, 'MODEL': MODEL, 'TRAINER': TRAINER, 'PSEUDO_POLICY': PSEUDO_POLICY,
        'PREPROCESSOR': PREPROCESSOR, 'SEG_MODEL': SEG_MODEL}


def register_all():
    """Import every module that registers something (the reference's utils/registry/register.py:3-9)."""
    from . import losses, metrics, preprocessor, pseudo_label_generator, segmentor  # noqa: F401


def install_into(reference_registries, suffix=''):
    """Put the B200 implementations into the reference's own registry module, and make the reference's backbone / dataset /
    trainer entries visible to this package's registries.

    ``suffix=''`` overrides the reference's entries (same keys); a non-empty suffix (e.g. ``'_B200'``) adds new keys next to
    them (``PSEUDO_POLICY['IAS_B200']``).  Either way ``SEG_MODEL``, ``DATASET``, ``TRAINER`` and the ``MODEL`` / ``LOSS`` /
    ``PREPROCESSOR`` entries this package does not provide are copied FROM the reference's registries, so that
    ``MODEL['SelfTrainingSegmentor'](cfg)`` finds its backbone through ``SEG_MODEL[cfg.model.seg_model.type]`` and a
    one-argument ``PSEUDO_POLICY[type](cfg)`` finds the target dataset class (``utils.load_model`` and
    ``generate_pseudo_labels.py`` keep working with the overridden keys).
    """
    register_all()
    for name, reg in _ALL.items():
        target = getattr(reference_registries, name)
        for key, obj in list(target.items()):              # the reference's entries -> this package (never overriding ours)
            if key not in reg:
                dict.__setitem__(reg, key, obj)
        for key, obj in list(reg.items()):
            if obj is not target.get(key):                 # ours -> the reference
                dict.__setitem__(target, key + suffix, obj)
