"""B200-native (sm_100a) DeepFwFM / DeepLight forward hot path.

Drop-in for the inference side of ``model/DeepFMs.py`` of ShanningLiu/xsDeepFwFM_deprecated:
``from xsdeepfwfm_deprecated_b200.model.DeepFMs import DeepFMs``.  Python is host glue only; the
arithmetic runs in hand-written CUDA kernels behind the C ABI in ``include/deepfwfm_b200.h``.
"""
__version__ = "0.1.0"
