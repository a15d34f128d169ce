"""kaldi_hmm_gmm_b200 — B200-native (sm_100a) E-step of diagonal-GMM acoustic-model
training, behind the class/method names of csukuangfj/kaldi-hmm-gmm's Python API.

Layers (bottom up):
  libkhg_b200.so        hand-written CUDA kernels + C ABI (include/khg_b200.h)
  _cabi / device        ctypes view of the ABI; DeviceModel / DeviceStats handles
  _khg_b200 (pybind11)  DiagGmm, AmDiagGmm, AccumDiagGmm, AccumAmDiagGmm,
                        DecodableAmDiagGmm*, GmmUpdateFlags ... with the reference's
                        signatures (python/csrc/*.cc of the reference)
There is no CPU fallback anywhere in this package.
"""
from . import _cabi  # noqa: F401
from .device import DeviceModel, DeviceStats, GraphBatch, align_batch, align_utterance_host  # noqa: F401

try:  # the pybind11 mirror of the reference classes (built by __graft_entry__.build())
    from ._khg_b200 import *  # noqa: F401,F403
    from ._khg_b200 import __doc__ as _ext_doc  # noqa: F401
    from .scripts import (TrainingGraph, gmm_acc_stats_ali, gmm_align_compiled, gmm_align_compiled_batch, gmm_est,  # noqa: F401
                          make_decodable, make_decodables)
    HAVE_EXTENSION = True
except ImportError as _e:  # pragma: no cover - reported loudly on use
    HAVE_EXTENSION = False
    _EXT_ERROR = _e
