"""CPU-side checks of the drop-in boundary: the shared library loads and exports
every symbol include/khg_b200.h declares; without a GPU compute calls fail loudly
(no CPU fallback)."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "khg_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(khg_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    from kaldi_hmm_gmm_b200 import _cabi

    L = _cabi.lib()
    declared = _declared_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/khg_b200.h but not exported"
    assert sorted(n for n, _, _ in _cabi.SYMBOLS) == declared
    assert L.khg_abi_version() == 4


def test_flag_augmentation_matches_reference():
    # csrc/model-common.cc:72-84
    from kaldi_hmm_gmm_b200 import _cabi

    aug = _cabi.lib().khg_augment_flags
    assert aug(2) == 7 and aug(1) == 5 and aug(4) == 4 and aug(0) == 4 and aug(15) == 15


def test_no_cpu_fallback_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from kaldi_hmm_gmm_b200 import DeviceModel

    with pytest.raises(RuntimeError):
        DeviceModel(4, np.array([0, 2, 5], np.int32))
