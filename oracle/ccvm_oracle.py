"""CPU oracle for the CCVM SDE hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` leg may import this module.  Nothing under ``ccvm_b200/``
imports it: the product path is CUDA-only and fails loudly without its extension.

What it is: a torch-CPU fp32 restatement of the reference's Euler-Maruyama loops,
post-processors, energy and solution statistics, keeping the reference's operation
order so that, fed the same noise, it reproduces the reference bit for bit.

Parity status: PINNED.  ``tests/golden/make_golden.py`` runs the unmodified reference
(``/root/reference``) under a seeded torch generator, records inputs/noise/outputs as
small ``.npz`` fixtures, and ``tests/test_oracle_golden.py`` checks this file against
every one of them (exact equality), plus the reference's own known-answer values
(test_mf_solver.py:63-154, test_solution.py:140-173, test_problem_instance.py:137-187).

Noise convention: each reference draw
``Normal(zeros(B), ones(B)).sample((N,)).transpose(0, 1)`` equals ``randn(N, B).T`` from
the same generator.  A ``NoiseSource`` either draws like that, or replays a recorded
tensor laid out ``[T][K][N][B]`` (K = draws per iteration, 2 for DL, else 1).

Every function cites the reference file:line it follows (paths under
``/root/reference/ccvm_simulators``).
"""

from __future__ import annotations

import numpy as np
import torch

EPS_ADAM = 1e-8
GAP_THRESHOLDS = (0.1, 1, 2, 3, 4, 5, 10)
GAP_NAMES = (
    "optimal",
    "one_percent",
    "two_percent",
    "three_percent",
    "four_percent",
    "five_percent",
    "ten_percent",
)


# --------------------------------------------------------------------------- noise
class NoiseSource:
    """Hands out (B, N) standard-normal draws in the reference's call order."""

    def __init__(self, n, batch, replay=None, generator=None, dtype=torch.float32):
        self.n, self.batch, self.dtype = n, batch, dtype
        self.replay = replay  # tensor [T][K][N][B] or None
        self.generator = generator
        self._cursor = 0
        self.record = None  # set to [] to capture draws

    def draw(self):
        if self.replay is not None:
            flat = self.replay.reshape(-1, self.n, self.batch)
            w = flat[self._cursor]
            self._cursor += 1
        else:
            # solvers/dl_solver.py:512-519,539: Normal(0,1).sample((N,)) is an (N,B) draw
            w = torch.randn(self.n, self.batch, generator=self.generator, dtype=torch.float32)
            if self.record is not None:
                self.record.append(w.clone())
        return w.transpose(0, 1).to(self.dtype)


def make_replay_noise(seed, iterations, draws_per_iter, n, batch):
    """[T][K][N][B] tensor equal to what the reference would draw after manual_seed(seed)."""
    # One randn(N, B) call per draw: a single large randn call is a different stream on CPU.
    g = torch.Generator().manual_seed(seed)
    draws = [torch.randn(n, batch, generator=g) for _ in range(iterations * draws_per_iter)]
    return torch.stack(draws).reshape(iterations, draws_per_iter, n, batch)


# --------------------------------------------------------------- shared building blocks
def _rowvec_times_q(x, q):
    # "bi,ij->bj": row vector times matrix (NOT q @ x).  dl_solver.py:145-149
    return torch.einsum("bi,ij -> bj", x, q)


def change_variables(y, lower, upper, s):
    """dl_solver.py:219-235 (identical in mf/langevin/pumped_langevin)."""
    return 0.5 * y / s * (upper - lower) + 0.5 * (upper + lower)


def energy(x, q, v, scaled_by=1):
    """problem_classes/boxqp/problem_instance.py:226-241."""
    e_quad = torch.einsum("bi, ij, bj -> b", x, q, x) * scaled_by
    e_lin = torch.einsum("bi, i -> b", x, v) * scaled_by
    return 0.5 * e_quad + e_lin


def scaling_factor(q, multiplier):
    """solvers/ccvm_solver.py:134-150."""
    return torch.sqrt(torch.sum(torch.abs(q))) * multiplier


def scale_coefs(q, v, scaled_by, factor):
    """problem_instance.py:243-255."""
    return q / factor, v / factor, scaled_by * factor


class _Adam:
    """The in-loop Adam transform shared by all four _solve_adam variants
    (dl_solver.py:637-727, mf_solver.py:659-738, langevin_solver.py:482-540,
    pumped_langevin_solver.py:360-420)."""

    def __init__(self, hp, like):
        self.alpha, self.beta1, self.beta2 = hp["alpha"], hp["beta1"], hp["beta2"]
        self.add_assign = hp["add_assign"]
        self.m = torch.zeros_like(like)
        self.v = None if self.beta2 == 1.0 else torch.zeros_like(like)

    def step(self, grads, i):
        self.m = self.beta1 * self.m + (1.0 - self.beta1) * grads
        mhat = self.m / (1.0 - self.beta1 ** (i + 1))
        if self.v is not None:
            self.v = self.beta2 * self.v + (1.0 - self.beta2) * torch.pow(grads, 2)
            vhat = self.v / (1.0 - self.beta2 ** (i + 1))
            upd = self.alpha * torch.div(mhat, torch.sqrt(vhat) + EPS_ADAM)
        else:
            upd = self.alpha * mhat
        return grads + upd if self.add_assign else upd


# ------------------------------------------------------------------------------ DL
def _dl_feedback(y, q, v, lower, upper, s):
    """G(y) = 1/4 ((y a/S + b) Q) a/S  and  V a/(2S).  dl_solver.py:143-154,191-201."""
    a, b = upper - lower, upper + lower
    g1 = 0.25 * _rowvec_times_q(y * a / s + b, q) * a / s
    g3 = v * a / (2 * s)
    return g1, g3


def dl_solve(q, v, batch, iterations, pump, dt, noise_ratio, feedback_scale, noise,
             pump_rate_flag=True, g=0.05, s_clamp=1, bounds=(0.0, 1.0), dtype=torch.float32,
             snapshots=None):
    """DLSolver._solve, dl_solver.py:468-569 with _calculate_drift_boxqp 117-172.

    The drift is evaluated with S = sqrt(pump-1) when pump > 1 else 1 (it is never handed
    the constructor's S), while the final clamp uses the constructor's S (``s_clamp``)."""
    n = q.shape[0]
    lower, upper = bounds
    c = torch.zeros((batch, n), dtype=dtype)
    s = torch.zeros((batch, n), dtype=dtype)
    s_drift = np.sqrt(pump - 1) if pump > 1 else 1
    rate = 1
    for i in range(iterations):
        if pump_rate_flag:
            rate = (i + 1) / iterations
        ratio_i = (noise_ratio - 1) * np.exp(-(i + 1) / iterations * 3) + 1
        c2, s2 = torch.pow(c, 2), torch.pow(s, 2)
        gc1, gc3 = _dl_feedback(c, q, v, lower, upper, s_drift)
        gc2 = (-1 + (pump * rate) - c2 - s2) * c
        gs1, gs3 = _dl_feedback(s, q, v, lower, upper, s_drift)
        gs2 = (-1 - (pump * rate) - c2 - s2) * s
        fsd = feedback_scale * (0.5 + rate)
        c_drift = -fsd * (gc1 + gc3) + gc2
        s_drift_term = -fsd * (gs1 + gs3) + gs2
        w_c = noise.draw() * np.sqrt(dt) * ratio_i
        w_s = noise.draw() * np.sqrt(dt) / ratio_i
        diff = 2 * g * torch.sqrt(c**2 + s**2 + 0.5)
        c += dt * c_drift + diff * w_c
        s += dt * s_drift_term + diff * w_s
        if snapshots is not None:
            snapshots(i, c, s)
    return torch.clamp(c, -s_clamp, s_clamp), s


def dl_solve_adam(q, v, batch, iterations, pump, dt, noise_ratio, noise, hyper,
                  pump_rate_flag=True, g=0.05, s_param=1, bounds=(0.0, 1.0),
                  dtype=torch.float32, snapshots=None):
    """DLSolver._solve_adam, dl_solver.py:571-769 with _calculate_grads_boxqp 174-217.
    No feedback_scale here; S is replaced by sqrt(pump-1) when pump > 1 and that same S
    clamps the result."""
    n = q.shape[0]
    lower, upper = bounds
    c = torch.zeros((batch, n), dtype=dtype)
    s = torch.zeros((batch, n), dtype=dtype)
    big_s = np.sqrt(pump - 1) if pump > 1 else s_param
    adam_c, adam_s = _Adam(hyper, c), _Adam(hyper, s)
    for i in range(iterations):
        p_i = pump * (i + 1) / iterations if pump_rate_flag else pump
        ratio_i = (noise_ratio - 1) * np.exp(-(i + 1) / iterations * 3) + 1
        gc1, gc3 = _dl_feedback(c, q, v, lower, upper, big_s)
        gs1, gs3 = _dl_feedback(s, q, v, lower, upper, big_s)
        c_grads = adam_c.step(-gc1 - gc3, i)
        s_grads = adam_s.step(-gs1 - gs3, i)
        c2, s2 = torch.pow(c, 2), torch.pow(s, 2)
        c_drift = (-1 + p_i - c2 - s2) * c
        s_drift = (-1 - p_i - c2 - s2) * s
        w_c = noise.draw() * np.sqrt(dt) * ratio_i
        w_s = noise.draw() * np.sqrt(dt) / ratio_i
        c += dt * (c_drift + c_grads) + 2 * g * torch.sqrt(c2 + s2 + 0.5) * w_c
        s += dt * (s_drift + s_grads) + 2 * g * torch.sqrt(c2 + s2 + 0.5) * w_s
        if snapshots is not None:
            snapshots(i, c, s)
    return torch.clamp(c, -big_s, big_s), s


# ------------------------------------------------------------------------------ MF
def _mf_feedback(mu_tilde_c, q, v, lower, upper, s):
    """mf_solver.py:176-189 / 216-229."""
    a, b = upper - lower, upper + lower
    t1 = -(1 / 4) * _rowvec_times_q(mu_tilde_c * a / s + b, q) * a / s
    t2 = -v * a / (2 * s)
    return t1, t2


def mf_solve(q, v, batch, iterations, s, pump, dt, j, feedback_scale, noise,
             pump_rate_flag=True, g=0.01, bounds=(0.0, 1.0), dtype=torch.float32,
             snapshots=None):
    """MFSolver._solve, mf_solver.py:493-593 with _calculate_drift_boxqp 141-198.
    Returns (mu, clamp(last measured mu_tilde), sigma)."""
    n = q.shape[0]
    lower, upper = bounds
    mu = torch.zeros((batch, n), dtype=dtype)
    sigma = torch.ones((batch, n), dtype=dtype) * (1 / 2)
    rate = 1
    mu_tilde = None
    for i in range(iterations):
        j_i = j * np.exp(-(i + 1) / iterations * 3.0)
        w_inc = noise.draw() / np.sqrt(dt)
        mu_tilde = mu + np.sqrt(1 / (4 * j_i)) * w_inc
        mu_tilde_c = torch.clamp(mu_tilde, -s, s)
        if pump_rate_flag:
            rate = (i + 1) / iterations
        pump_i = pump * rate + 1 + j_i
        mu2 = torch.pow(mu, 2)
        t1, t2 = _mf_feedback(mu_tilde_c, q, v, lower, upper, s)
        drift_mu = (-(1 + j_i) + pump_i - g**2 * mu2) * mu + feedback_scale * (t1 + t2)
        drift_sigma = (
            2 * (-(1 + j_i) + pump_i - 3 * g**2 * mu2) * sigma
            + -2 * j_i * (sigma - 0.5).pow(2)
            + ((1 + j_i) + 2 * g**2 * mu2)
        )
        diffusion = np.sqrt(j_i) * (sigma - 0.5) * w_inc
        mu += dt * (drift_mu + diffusion)
        sigma += dt * drift_sigma
        if snapshots is not None:
            snapshots(i, mu, sigma)
    return mu, torch.clamp(mu_tilde, -s, s), sigma


def mf_solve_adam(q, v, batch, iterations, s, pump, dt, j, feedback_scale, noise, hyper,
                  pump_rate_flag=True, g=0.01, bounds=(0.0, 1.0), dtype=torch.float32,
                  snapshots=None):
    """MFSolver._solve_adam, mf_solver.py:595-764 with _calculate_grads_boxqp 200-233."""
    n = q.shape[0]
    lower, upper = bounds
    mu = torch.zeros((batch, n), dtype=dtype)
    sigma = torch.ones((batch, n), dtype=dtype) * (1 / 2)
    adam = _Adam(hyper, mu)
    mu_tilde = None
    for i in range(iterations):
        j_i = j * np.exp(-(i + 1) / iterations * 3.0)
        w_inc = noise.draw() / np.sqrt(dt)
        mu_tilde = mu + np.sqrt(1 / (4 * j_i)) * w_inc
        mu_tilde_c = torch.clamp(mu_tilde, -s, s)
        rate = (i + 1) / iterations if pump_rate_flag else 1.0
        pump_i = pump * rate + 1 + j_i
        t1, t2 = _mf_feedback(mu_tilde_c, q, v, lower, upper, s)
        grads = adam.step(feedback_scale * (t1 + t2), i)
        mu2 = torch.pow(mu, 2)
        mu_drift = (-(1 + j_i) + pump_i - g**2 * mu2) * mu
        mu_drift += np.sqrt(j_i) * (sigma - 0.5) * w_inc
        mu += dt * (grads + mu_drift)
        sigma_drift = 2 * (-(1 + j_i) + pump_i - 3 * g**2 * mu2) * sigma
        sigma_drift += -2 * j_i * (sigma - 0.5).pow(2)
        sigma_drift += (1 + j_i) + 2 * g**2 * mu2
        sigma += dt * sigma_drift
        if snapshots is not None:
            snapshots(i, mu, sigma)
    return mu, torch.clamp(mu_tilde, -s, s), sigma


# ------------------------------------------------------------------------ Langevin
def _langevin_grad(c, q, v, lower, upper, s):
    """langevin_solver.py:131-139 (drift) == 157-166 (grads)."""
    a, b = upper - lower, upper + lower
    t1 = _rowvec_times_q(c * a / (2 * s) + b / 2, q)
    return -(t1 + v) * a / (2 * s)


def langevin_solve(q, v, batch, iterations, s, dt, sigma, feedback_scale, noise,
                   bounds=(0.0, 1.0), dtype=torch.float32, snapshots=None):
    """LangevinSolver._solve, langevin_solver.py:368-435."""
    n = q.shape[0]
    lower, upper = bounds
    c = torch.zeros((batch, n), dtype=dtype)
    for i in range(iterations):
        drift = _langevin_grad(c, q, v, lower, upper, s)
        w = noise.draw() * np.sqrt(dt)
        c += dt * feedback_scale * drift + sigma * w
        c = torch.clamp(c, -s, s)
        if snapshots is not None:
            snapshots(i, c)
    return c


def langevin_solve_adam(q, v, batch, iterations, s, dt, sigma, feedback_scale, noise, hyper,
                        bounds=(0.0, 1.0), dtype=torch.float32, snapshots=None):
    """LangevinSolver._solve_adam, langevin_solver.py:437-561.  (The beta2==1 lambdas read
    mhat_c through their closure, 494/500 -- the same value they are passed.)"""
    n = q.shape[0]
    lower, upper = bounds
    c = torch.zeros((batch, n), dtype=dtype)
    adam = _Adam(hyper, c)
    for i in range(iterations):
        grads = adam.step(_langevin_grad(c, q, v, lower, upper, s), i)
        w = noise.draw() * np.sqrt(dt)
        c += dt * feedback_scale * grads + sigma * w
        c = torch.clamp(c, -s, s)
        if snapshots is not None:
            snapshots(i, c)
    return c


# ------------------------------------------------------------------ pumped Langevin
def _pl_grad(c, q, v, lower, upper, s):
    """pumped_langevin_solver.py:132-147."""
    a, b = upper - lower, upper + lower
    t1 = _rowvec_times_q(c * a / (2 * s) + b / 2, q) * a / (2 * s)
    t2 = v * a / (2 * s)
    return -t1 - t2


def pumped_langevin_solve(q, v, batch, iterations, s, pump, dt, sigma, feedback_scale, noise,
                          pump_rate_flag=True, bounds=(0.0, 1.0), dtype=torch.float32,
                          snapshots=None):
    """PumpedLangevinSolver._solve, pumped_langevin_solver.py:232-309 (+ drift 95-114)."""
    n = q.shape[0]
    lower, upper = bounds
    c = torch.zeros((batch, n), dtype=dtype)
    for i in range(iterations):
        p_i = pump * (i + 1) / iterations if pump_rate_flag else pump
        c2 = torch.pow(c, 2)
        drift = (-1 + p_i - c2) * c + feedback_scale * _pl_grad(c, q, v, lower, upper, s)
        w = noise.draw() * np.sqrt(dt)
        c += dt * drift + sigma * w
        c = torch.clamp(c, -s, s)
        if snapshots is not None:
            snapshots(i, c)
    return c


def pumped_langevin_solve_adam(q, v, batch, iterations, s, pump, dt, sigma, feedback_scale,
                               noise, hyper, pump_rate_flag=True, bounds=(0.0, 1.0),
                               dtype=torch.float32, snapshots=None):
    """PumpedLangevinSolver._solve_adam, pumped_langevin_solver.py:311-449."""
    n = q.shape[0]
    lower, upper = bounds
    c = torch.zeros((batch, n), dtype=dtype)
    adam = _Adam(hyper, c)
    for i in range(iterations):
        grads = adam.step(_pl_grad(c, q, v, lower, upper, s), i)
        w = noise.draw() * np.sqrt(dt)
        p_i = pump * (i + 1) / iterations if pump_rate_flag else pump
        c2 = torch.pow(c, 2)
        c_pump = (-1 + p_i - c2) * c
        c += dt * (c_pump + feedback_scale * grads) + sigma * w
        c = torch.clamp(c, -s, s)
        if snapshots is not None:
            snapshots(i, c)
    return c


# ----------------------------------------------------------------- post-processors
def pp_grad_descent(x, q, v, lower=0.0, upper=1.0, num_iter_main=1000, num_iter_pp=None,
                    step_size=0.1):
    """post_processor/grad_descent.py:58-64.  Works on a copy (the reference mutates its
    input on the first step, which callers here never rely on)."""
    x = x.clone()
    if num_iter_pp is None:
        num_iter_pp = int(num_iter_main * 0.01)
    for _ in range(num_iter_pp):
        grads = _rowvec_times_q(x, q) + v
        x += -step_size * grads
        x = torch.clamp(x, lower, upper)
    return x


def pp_adam(x, q, v, lower=0.0, upper=1.0, lr=0.01, beta1=0.9, beta2=0.99, eps=1e-8):
    """post_processor/adam.py:58-66 for its effective behaviour: ONE torch.optim.Adam step
    on 1/2 xQx + Vx (gradient 1/2 (xQ + x Q^T) + V) followed by the box clamp.  Further
    iterations in the reference do not move the returned tensor, so
    num_iter is not a parameter.  First Adam step: m = (1-b1) g, v = (1-b2) g^2,
    step = lr * (m/(1-b1)) / (sqrt(v)/sqrt(1-b2) + eps)."""
    grad = 0.5 * (_rowvec_times_q(x, q) + torch.einsum("bj,ij -> bi", x, q)) + v
    m = (1 - beta1) * grad
    vv = (1 - beta2) * grad * grad
    denom = vv.sqrt() / np.sqrt(1 - beta2) + eps
    step = (lr / (1 - beta1)) * (m / denom)
    return torch.clamp(x - step, lower, upper)


# ---------------------------------------------------------------------- statistics
def solution_stats(objective_values, optimal_value):
    """solution.py:82 and 87-146: best = max(-E); seven gap-threshold fractions."""
    neg = -objective_values
    best = torch.max(neg).item()
    gap = (optimal_value - neg) * 100 / torch.abs(neg)
    nb = neg.shape[0]
    perf = {}
    for name, thr in zip(GAP_NAMES, GAP_THRESHOLDS):
        cnt = torch.where(gap <= thr, torch.ones_like(neg), torch.zeros_like(neg)).sum().item()
        perf[name] = round(cnt / nb, 4)
    return best, perf


# ----------------------------------------------------------------- __call__ epilogues
def epilogue(solver, state, q, v, scaled_by, s, bounds=(0.0, 1.0), post_processor=None,
             num_iter_main=1000):
    """The tail of each solver's __call__:
    DL   dl_solver.py:936-959 (note the double change of variables with a post-processor),
    MF   mf_solver.py:927-948, Langevin langevin_solver.py:710-726,
    PL   pumped_langevin_solver.py:603-622.
    ``state`` is c (DL/Langevin/PL) or the clamped mu_tilde (MF).
    Returns (problem_variables, objective_values)."""
    lower, upper = bounds

    def run_pp(x):
        if post_processor is None:
            return x
        if post_processor == "grad-descent":
            return pp_grad_descent(x, q, v, num_iter_main=num_iter_main)
        if post_processor == "adam":
            return pp_adam(x, q, v)
        raise AssertionError(f"Method type is not valid. Provided: {post_processor}")

    if solver == "dl":
        if post_processor:
            pv = run_pp(change_variables(state, lower, upper, s))
        else:
            pv = state
        confs = change_variables(pv, lower, upper, s)
        return pv, energy(confs, q, v, scaled_by)
    if solver == "mf":
        pv = run_pp(change_variables(state, lower, upper, s))
        return pv, energy(pv, q, v, scaled_by)
    if solver in ("langevin", "pumped_langevin"):
        pv = run_pp((state + s) / (2 * s))
        return pv, energy(pv, q, v, scaled_by)
    raise ValueError(solver)


# ------------------------------------------------------------------- instance loader
def load_instance(path, delimiter="\t"):
    """problem_instance.py:116-224: header (8 fields), V line, N Q lines, optional solution
    line; Q and V are NEGATED on load (183, 188)."""
    with open(path, "r") as fh:
        lines = fh.readlines()
    head = lines[0].split("\n")[0].split(delimiter)
    n = int(head[0])
    info = {
        "problem_size": n,
        "optimal_sol": float(head[1]),
        "best_sol": float(head[2]),
        "optimality": head[3].lower() == "true",
        "sol_time_gb": float(head[4]),
        "sol_time_bfgs": float(head[5]),
        "num_frac_values": int(head[7]),
    }
    vrow = lines[1].split("\n")[0].split(delimiter)
    v = -torch.tensor([float(t) for t in vrow[:n]], dtype=torch.float32)
    qrows = [[float(t) for t in ln.split("\n")[0].split(delimiter)[:n]] for ln in lines[2 : n + 2]]
    q = -torch.tensor(qrows, dtype=torch.float32)
    sol = []
    if len(lines) > n + 2:
        sol = [float(t) for t in lines[n + 2].split("\n")[0].split(delimiter) if t != ""]
    info["solution_vector"] = sol
    return q, v, info


def synthetic_boxqp(n, seed):
    """Synthetic dense BoxQP fitted to the bundled instances: returns the
    reference-convention (negated) Q, V before scaling."""
    g = torch.Generator().manual_seed(1000 + seed)
    a = torch.randn(n, n, generator=g)
    q_file = (a + a.T) / np.sqrt(2.0) * (28.5 / np.sqrt(n))
    v_file = 20.0 * torch.randn(n, generator=g)
    return (-q_file).float(), (-v_file).float()
