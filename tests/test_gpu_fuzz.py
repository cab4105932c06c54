"""Randomised shape sweep of the production (Philox, in-loop noise) kernels against the plain variant
of the same tile kernel (CCVM_NO_PIPE: noise drawn after the contraction, run-time panel stride):
same Philox stream, same arithmetic up to prefetch order, so short runs must agree closely for every
(n, batch) -- column-group counts around every code-path boundary (PIPE eligibility at 4K+1 chunks,
TMEM -> hybrid at n = 128/129, odd / even chunk counts for the peeled tail, one or two groups,
partially filled groups and CTAs)."""
import numpy as np
import pytest
import torch

from ccvm_b200 import engine as E, _native as nat
from tests.test_gpu_parity import instance, HP

pytestmark = pytest.mark.gpu

RNG = np.random.RandomState(20260101)
SHAPES = sorted({(int(n), int(b)) for n, b in zip(
    list(RNG.randint(17, 257, 22)) + [17, 20, 33, 36, 37, 64, 65, 127, 128, 129, 132, 133, 255, 256],
    list(RNG.randint(1, 3000, 22)) + [1, 2, 29, 148, 149, 300, 4096, 1000, 57, 3, 1184, 1185, 8, 2049])})
KW = {
    "dl": (nat.SOLVER_DL, dict(s=1.0, pump=8.0, dt=0.001, noise_ratio=10.0, g=0.05)),
    "mf": (nat.SOLVER_MF, dict(s=20.0, pump=0.0, dt=0.0025, j=5.0, feedback_scale=4000.0, g=0.01)),
    "lv": (nat.SOLVER_LANGEVIN, dict(s=0.5, dt=0.002, sigma=0.5, feedback_scale=1.0)),
    "plv": (nat.SOLVER_PUMPED_LANGEVIN, dict(s=0.5, pump=2.0, dt=0.002, sigma=0.5, feedback_scale=1.0)),
}


@pytest.mark.parametrize("n,b", SHAPES)
def test_pipe_kernels_match_plain_variant(monkeypatch, n, b):
    monkeypatch.setenv("CCVM_MMA", "0")   # the tiled kernels (the tensor-core kernel has its own tests: test_gpu_mma.py)
    t = 24
    for k, (name, (sid, kw)) in enumerate(KW.items()):
        adam = (n + b + k) % 2 == 1
        kw = dict(kw)
        if name == "dl" and not adam:
            kw["feedback_scale"] = 100.0
        if adam:
            kw["hyperparameters"] = HP
        alg = nat.ALG_ADAM if adam else nat.ALG_ORIGINAL
        q, v, _ = instance(n, n + k, 0.2 if name == "dl" else 0.05)
        qg, vg = q.cuda(), v.cuda()
        monkeypatch.delenv("CCVM_NO_PIPE", raising=False)
        fast, _ = E.solve(sid, alg, qg, vg, b, t, seed=11, offset=4 * k, **kw)
        fast = [o.clone() for o in fast]
        monkeypatch.setenv("CCVM_NO_PIPE", "1")
        plain, _ = E.solve(sid, alg, qg, vg, b, t, seed=11, offset=4 * k, **kw)
        for a, c in zip(plain, fast):
            assert a.shape == c.shape == (b, n)
            assert torch.isfinite(c).all(), (name, adam)
            err = (a - c).abs().max().item()
            assert err <= 2e-4 * max(a.abs().max().item(), 1.0), (name, adam, err)
