"""The torch restatement of the reference's batched vmap/einsum evaluation (benchmarks/reference_gpu_standin.py: the
"reference algorithm on the GPU" number of bench.py) reproduces the outputs of the unmodified reference (tests/golden)."""
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

from helpers import LAYOUT_CASES, golden_layout, load

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


@pytest.mark.parametrize("case", LAYOUT_CASES)
def test_standin_matches_reference_outputs(case):
    from benchmarks import reference_gpu_standin as standin

    g = load(case)
    tables = standin.upload(golden_layout(g), "cpu")
    y_ref = g["y_ref"]
    scale = max(1.0, float(np.max(np.abs(y_ref))))
    for limit in (4.0, 1e-7):  # one batch per group / one summand per batch (interpolation.py:286-287)
        y = standin.evaluate(tables, torch.from_numpy(g["x"]), memory_limit=limit).numpy()
        assert y.shape == y_ref.shape
        assert np.max(np.abs(y - y_ref)) <= 1e-12 * scale
