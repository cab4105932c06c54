"""Shared helpers: turn a golden ``.npz`` fixture into an oracle call / an engine call."""
import glob
import os

import numpy as np
import torch

from oracle import ccvm_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

LOOP_PREFIXES = ("dl_", "dladam_", "mf_", "mfadam_", "lv_", "lvadam_", "plv_", "plvadam_")


def loop_fixtures():
    names = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz")))
    return [n for n in names if n.startswith(LOOP_PREFIXES) and not n.startswith("mf_tensorS")]


def load(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: z[k] for k in z.files}


def _f(z, k, default=None):
    return float(z[k]) if k in z else default


def hyper_of(z):
    return dict(alpha=float(z["alpha"]), beta1=float(z["beta1"]), beta2=float(z["beta2"]),
                add_assign=bool(z["add_assign"]))


def kind_of(name):
    return name.split("_")[0]


def expected_outputs(name, z):
    k = kind_of(name)
    if k in ("dl", "dladam"):
        return {"c": z["out_c"], "s": z["out_s"]}
    if k in ("mf", "mfadam"):
        return {"mu": z["out_mu"], "mu_tilde": z["out_mu_tilde"], "sigma": z["out_sigma"]}
    return {"c": z["out_c"]}


def run_oracle(name, z, dtype=torch.float32):
    """Run the oracle on fixture ``z`` replaying its recorded noise."""
    k = kind_of(name)
    q = torch.from_numpy(z["q"]).to(dtype)
    v = torch.from_numpy(z["v"]).to(dtype)
    b, t = int(z["batch"]), int(z["iterations"])
    n = q.shape[0]
    bounds = tuple(float(x) for x in z["bounds"]) if "bounds" in z else (0.0, 1.0)
    flag = bool(z["flag"]) if "flag" in z else True
    noise = O.NoiseSource(n, b, replay=torch.from_numpy(z["noise"]), dtype=dtype)
    if k == "dl":
        c, s = O.dl_solve(q, v, b, t, _f(z, "pump"), _f(z, "dt"), _f(z, "noise_ratio"),
                          _f(z, "feedback_scale"), noise, flag, _f(z, "g"), _f(z, "s_ctor"), bounds, dtype)
        return {"c": c, "s": s}
    if k == "dladam":
        c, s = O.dl_solve_adam(q, v, b, t, _f(z, "pump"), _f(z, "dt"), _f(z, "noise_ratio"), noise,
                               hyper_of(z), flag, _f(z, "g"), _f(z, "s_ctor"), bounds, dtype)
        return {"c": c, "s": s}
    if k in ("mf", "mfadam"):
        s_val = torch.from_numpy(z["s_vec"]).to(dtype) if "s_vec" in z else _f(z, "S")
        if k == "mf":
            mu, mt, sg = O.mf_solve(q, v, b, t, s_val, _f(z, "pump"), _f(z, "dt"), _f(z, "j"),
                                    _f(z, "feedback_scale"), noise, flag, _f(z, "g"), bounds, dtype)
        else:
            mu, mt, sg = O.mf_solve_adam(q, v, b, t, s_val, _f(z, "pump"), _f(z, "dt"), _f(z, "j"),
                                         _f(z, "feedback_scale"), noise, hyper_of(z), flag, _f(z, "g"),
                                         bounds, dtype)
        return {"mu": mu, "mu_tilde": mt, "sigma": sg}
    if k == "lv":
        return {"c": O.langevin_solve(q, v, b, t, _f(z, "S"), _f(z, "dt"), _f(z, "sigma"),
                                      _f(z, "feedback_scale"), noise, bounds, dtype)}
    if k == "lvadam":
        return {"c": O.langevin_solve_adam(q, v, b, t, _f(z, "S"), _f(z, "dt"), _f(z, "sigma"),
                                           _f(z, "feedback_scale"), noise, hyper_of(z), bounds, dtype)}
    if k == "plv":
        return {"c": O.pumped_langevin_solve(q, v, b, t, _f(z, "S"), _f(z, "pump"), _f(z, "dt"),
                                             _f(z, "sigma"), _f(z, "feedback_scale"), noise, flag, bounds,
                                             dtype)}
    if k == "plvadam":
        return {"c": O.pumped_langevin_solve_adam(q, v, b, t, _f(z, "S"), _f(z, "pump"), _f(z, "dt"),
                                                  _f(z, "sigma"), _f(z, "feedback_scale"), noise,
                                                  hyper_of(z), flag, bounds, dtype)}
    raise ValueError(name)


def run_engine(name, z, device="cuda"):
    """Run the CUDA engine (through the C ABI) on fixture ``z`` in noise-replay mode."""
    from ccvm_b200 import engine as E
    from ccvm_b200 import _native as nat
    k = kind_of(name)
    q = torch.from_numpy(z["q"]).to(device)
    v = torch.from_numpy(z["v"]).to(device)
    b, t = int(z["batch"]), int(z["iterations"])
    bounds = tuple(float(x) for x in z["bounds"]) if "bounds" in z else (0.0, 1.0)
    flag = bool(z["flag"]) if "flag" in z else True
    noise = torch.from_numpy(z["noise"]).to(device)
    adam = k.endswith("adam")
    kw = dict(lower=bounds[0], upper=bounds[1], pump_rate_flag=flag, noise=noise,
              hyperparameters=hyper_of(z) if adam else None)
    alg = nat.ALG_ADAM if adam else nat.ALG_ORIGINAL
    if k in ("dl", "dladam"):
        outs, _ = E.solve(nat.SOLVER_DL, alg, q, v, b, t, s=_f(z, "s_ctor"), pump=_f(z, "pump"), dt=_f(z, "dt"),
                          noise_ratio=_f(z, "noise_ratio"), feedback_scale=_f(z, "feedback_scale"), g=_f(z, "g"),
                          **kw)
        return {"c": outs[0].cpu(), "s": outs[1].cpu()}
    if k in ("mf", "mfadam"):
        s_vec = torch.from_numpy(z["s_vec"]) if "s_vec" in z else None
        outs, _ = E.solve(nat.SOLVER_MF, alg, q, v, b, t, s=_f(z, "S", 0.0), s_vec=s_vec, pump=_f(z, "pump"),
                          dt=_f(z, "dt"), j=_f(z, "j"), feedback_scale=_f(z, "feedback_scale"), g=_f(z, "g"), **kw)
        return {"mu": outs[0].cpu(), "mu_tilde": outs[1].cpu(), "sigma": outs[2].cpu()}
    if k in ("lv", "lvadam"):
        outs, _ = E.solve(nat.SOLVER_LANGEVIN, alg, q, v, b, t, s=_f(z, "S"), dt=_f(z, "dt"),
                          sigma=_f(z, "sigma"), feedback_scale=_f(z, "feedback_scale"), **kw)
        return {"c": outs[0].cpu()}
    if k in ("plv", "plvadam"):
        outs, _ = E.solve(nat.SOLVER_PUMPED_LANGEVIN, alg, q, v, b, t, s=_f(z, "S"), pump=_f(z, "pump"),
                          dt=_f(z, "dt"), sigma=_f(z, "sigma"), feedback_scale=_f(z, "feedback_scale"), **kw)
        return {"c": outs[0].cpu()}
    raise ValueError(name)
