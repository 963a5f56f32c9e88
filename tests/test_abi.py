"""The C-ABI library: it loads without a GPU, exports everything include/smolyax_b200.h declares, and fails loudly
(no CPU fallback) when there is no device."""
import ctypes
import re
from pathlib import Path

import numpy as np
import pytest
import torch

from smolyax_b200 import _build, _lib

HEADER = Path(__file__).resolve().parent.parent / "include" / "smolyax_b200.h"


def test_library_exports_every_declared_symbol():
    text = HEADER.read_text()
    declared = set(re.findall(r"^(?:int|int64_t|const char\*)\s+(smx_\w+)\s*\(", text, flags=re.M))
    assert len(declared) >= 16
    raw = ctypes.CDLL(str(_build.CUDA_LIB))
    for name in declared:
        assert hasattr(raw, name), f"{name} is declared in the header but not exported"
    assert declared == set(_lib.EXPORTS), "the ctypes table and the header disagree"


def test_version_arch_and_error_string():
    assert _lib.lib.smx_version() >= 1
    assert _lib.lib.smx_arch() == b"sm_100a"
    assert isinstance(_lib.lib.smx_last_error(), bytes)
    assert _lib.lib.smx_launch_count() >= 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_create_fails_loudly_without_a_gpu():
    layout = {"offset": np.zeros(1)}
    with pytest.raises(_lib.SmolyaxCudaError, match="no CPU fallback|no sm_100"):
        _lib.create(layout, 3, 1, 0, -1)


def test_invalid_arguments_are_reported():
    assert _lib.lib.smx_eval(None, None, 1, 1, None, None) == 1
    assert b"smx_eval" in _lib.lib.smx_last_error()
    with pytest.raises(AssertionError):
        _lib.check(_lib.lib.smx_gradient(None, None, 1, 1, None, None), "smx_gradient")
    assert _lib.lib.smx_eval(None, None, -1, 1, None, None) == 1
