"""TTS/R99 reporting (host numpy) against closed forms and, when the reference tree is present in
the build container, against the reference's own SampleTTSMetric."""
import os
import sys

import numpy as np
import pytest

from ccvm_b200 import tts


def test_r99_closed_forms():
    assert tts.calc_r99(0) == np.inf and tts.calc_r99(1) == 1.0
    assert tts.calc_r99(0.995) == 1.0
    assert tts.calc_r99(0.5) == pytest.approx(np.log(0.01) / np.log(0.5))
    with pytest.raises(ValueError):
        tts.calc_r99(0.5, confidence=1.0)


def test_tts_table_shape_and_monotonicity():
    md = [dict(problem_size=20, batch_size=1000, solve_time=1e-3, solution_performance={"optimal": p})
          for p in (0.9, 0.5, 0.7, 0.2)]
    md += [dict(problem_size=30, batch_size=1000, solve_time=2e-3, solution_performance={"optimal": p})
           for p in (0.1, 0.05, 0.2)]
    table = tts.tts_table(md)
    assert set(table) == {20, 30} and set(table[20]) == {25.0, 50.0, 75.0}
    assert table[20][25.0] <= table[20][50.0] <= table[20][75.0]
    assert table[30][50.0] > table[20][50.0]


@pytest.mark.skipif(not os.path.isdir("/root/reference/ccvm_simulators"), reason="reference tree not present")
def test_matches_reference_bootstrap():
    # load utils/sampleTTSmetric.py without running ccvmplotlib/__init__.py (it imports matplotlib)
    import importlib
    import types
    pkg = types.ModuleType("_ref_tts_utils")
    pkg.__path__ = ["/root/reference/ccvm_simulators/ccvmplotlib/utils"]
    sys.modules["_ref_tts_utils"] = pkg
    if "future" not in sys.modules:  # the reference imports future.utils.iteritems (py2 compat shim, absent here)
        fut, futu = types.ModuleType("future"), types.ModuleType("future.utils")
        futu.iteritems = lambda d: iter(d.items())
        fut.utils = futu
        sys.modules["future"], sys.modules["future.utils"] = fut, futu
    try:
        SampleTTSMetric = importlib.import_module("_ref_tts_utils.sampleTTSmetric").SampleTTSMetric
    except Exception as e:
        pytest.skip(f"reference TTS metric not importable: {e}")
    probs = [0.9, 0.35, 0.6, 0.05, 0.75]
    ref = SampleTTSMetric(tau_attribute="time", percentile=50.0, seed=1, num_bootstraps=100)
    exp = ref.calc_R99_distribution(probs, 1000)
    got = tts.r99_distribution(probs, 1000, percentile=50.0, num_bootstraps=100, seed=1)
    assert np.allclose(got, exp)
