"""GPU parity against the golden vectors recorded from the unmodified reference
(noise-replay mode, through the C ABI).  Tolerance: the kernel folds the affine maps into the
matrix and sums in a different order than torch's bmm, so results agree to fp32 rounding
accumulated over the run, not bit for bit: |diff| <= 2e-4 * max(1, |ref|)."""
import numpy as np
import pytest
import torch

from tests import _cases as C

pytestmark = pytest.mark.gpu

ATOL = 2e-4


def _close(got, exp, name):
    got, exp = np.asarray(got, dtype=np.float64), np.asarray(exp, dtype=np.float64)
    err = np.abs(got - exp) / np.maximum(1.0, np.abs(exp))
    assert np.isfinite(got).all(), name
    assert err.max() <= ATOL, f"{name}: max scaled err {err.max():.3e}"


@pytest.mark.parametrize("name", C.loop_fixtures())
def test_loop_matches_reference_golden(name):
    z = C.load(name)
    got = C.run_engine(name, z)
    for key, exp in C.expected_outputs(name, z).items():
        _close(got[key].numpy(), exp, f"{name}:{key}")


def test_mf_tensor_s():
    z = C.load("mf_tensorS")
    got = C.run_engine("mf_tensorS", z)
    for key, exp in C.expected_outputs("mf_tensorS", z).items():
        _close(got[key].numpy(), exp, key)
