"""Seeded randomised differential test: random shapes (aligned and not), per-axis mode lists, the filter
families of the path, CUDA result against the oracle.  (This sweep found the mode-sequence semantics of the
fused gradient magnitude: a sequence means one mode per DERIVATIVE axis, filters.py:1175-1201.)"""
import os
import sys

import pytest

pytestmark = pytest.mark.gpu

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tools"))


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_random_shapes_and_modes(seed):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import fuzz_fused
    assert fuzz_fused.run(seed, 70, verbose=False) == 0


@pytest.mark.parametrize("seed", [11, 12, 13])
def test_random_exact_paths_bit_exact(seed):
    """Integer / float64 arrays through the exact kernels (streaming, tiled, general, min / max, N-d
    correlate): random dtypes, shapes, modes, cvals, origins — bit for bit against the oracle."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import fuzz_fused
    assert fuzz_fused.run_exact(seed, 80, verbose=False) == 0
