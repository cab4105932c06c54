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
