"""Shared helpers: run the CPU oracle and the CUDA FDM on the same ICs / control sequences and compare."""
import numpy as np

from oracle.fdm import OracleFdm, comp_names, prop_names


def oracle_named_state(f: OracleFdm):
    d = f.snapshot_dict()
    d.update(f.props_dict())
    pid = f.pid_state()
    for k, (name, ctype) in enumerate(comp_names()):
        if ctype == 5:
            d[f"pid:{name}:prev"], d[f"pid:{name}:prev2"], d[f"pid:{name}:itot"] = pid[k]
    return d


def random_ics(rng, n):
    ic = np.zeros((n, 12))
    ic[:, 0] = rng.uniform(119.5, 120.5, n)      # lon
    ic[:, 1] = rng.uniform(59.5, 60.5, n)        # lat
    ic[:, 2] = rng.uniform(14000, 30000, n)      # h ft
    ic[:, 3] = rng.uniform(0, 360, n)            # psi
    ic[:, 4] = rng.uniform(400, 1200, n)         # u fps
    return ic


def random_controls(rng, n):
    u = np.zeros((n, 4))
    u[:, 0:3] = rng.integers(0, 41, (n, 3)) / 20.0 - 1.0
    u[:, 3] = rng.integers(0, 30, n) / 58.0 + 0.4
    return u


# scale used to turn absolute differences into the relative error reported per field group
def field_scale(name, val):
    if name.startswith("ri_"):
        return 2.1e7
    if name.startswith(("vi_", "dqv")):
        return 1.5e3
    if name.startswith(("q", "wi_", "pqridot")) and name[1:2].isdigit():
        return 1.0
    return max(1.0, abs(val))
