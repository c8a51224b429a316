"""hiast_b200 -- B200-native (sm_100a) post-logit self-training hot path of HIAST.

Layout: ``csrc/`` CUDA kernels + the C ABI (``include/hiast_b200.h``), ``_lib`` / ``ops`` the ctypes
binding, and the host-side mirrors of the reference interface for this path:

    pseudo_label_generator  PSEUDO_POLICY['IAS' | 'CT' | 'NT' | 'CBST'] (PNG files encoded on the device)
                                                                     workflows/pseudo_label_generator.py
    losses                  LOSS['CE' | 'SoftCE' | 'KLDIV' | 'MSE']  sseg/models/modules/losses.py
    segmentor               MODEL['SelfTrainingSegmentor']           sseg/models/segmentors/self_training_segmentor.py
    preprocessor            PREPROCESSOR['CopyPaste']                sseg/datasets/preprocessor.py
    metrics                 intersectionAndUnionGPU, ConfusionMeter  utils/metrics.py
    validator               Validator (multi-scale / flip predict)   workflows/validator.py
    pseudo_store            stat_samples_with_class, load_pseudo_labels   sseg/datasets/loader/base_dataset.py
    ema                     update_ema_model                         utils/utils.py
    sharded                 multi-GPU IAS with the NCCL threshold hand-off (new)

There is no CPU fallback: without the built library (``python -m hiast_b200.build``) and a CUDA
device every compute entry point raises.
"""

__version__ = '0.1.0'

from .registry import DATASET, LOSS, MODEL, PREPROCESSOR, PSEUDO_POLICY, SEG_MODEL, TRAINER, register_all  # noqa: F401
