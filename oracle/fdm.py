"""ctypes binding of the CPU oracle FDM (oracle/f16_oracle.cpp).

TEST INFRASTRUCTURE ONLY -- see the header of f16_oracle.cpp.  Mirrors the slice of the
``jsbsim.FGFDMExec`` surface that the reference's ``AircraftSimulator`` touches
(reference envs/JSBSim/core/simulatior.py:165-188,223,261,295,313).
"""
from __future__ import annotations

import ctypes
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB = None


def build(force: bool = False) -> Path:
    so = _HERE / "_build" / "liboracle.so"
    srcs = [_HERE / "f16_oracle.cpp", _HERE / "gen" / "f16_ir.inc", _HERE / "env_oracle.cpp"]
    srcs = [s for s in srcs if s.exists()]
    if force or not so.exists() or any(s.stat().st_mtime > so.stat().st_mtime for s in srcs):
        subprocess.check_call(["make", "-C", str(_HERE), "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(str(build()))
        d, vp, i = ctypes.c_double, ctypes.c_void_p, ctypes.c_int
        L.orc_fdm_create.restype = vp
        L.orc_fdm_create.argtypes = [d, d]
        L.orc_fdm_destroy.argtypes = [vp]
        L.orc_fdm_reset.argtypes = [vp] + [d] * 12
        L.orc_fdm_set_controls.argtypes = [vp, d, d, d, d]
        L.orc_fdm_run.argtypes = [vp, i]
        L.orc_fdm_snapshot_name.restype = ctypes.c_char_p
        L.orc_fdm_snapshot_name.argtypes = [i]
        L.orc_fdm_snapshot.argtypes = [vp, ctypes.POINTER(d)]
        L.orc_fdm_prop_name.restype = ctypes.c_char_p
        L.orc_fdm_prop_name.argtypes = [i]
        L.orc_fdm_get_props.argtypes = [vp, ctypes.POINTER(d)]
        L.orc_fdm_set_props.argtypes = [vp, ctypes.POINTER(d)]
        L.orc_fdm_comp_name.restype = ctypes.c_char_p
        L.orc_fdm_comp_name.argtypes = [i]
        L.orc_fdm_comp_type.argtypes = [i]
        L.orc_fdm_get_pid.argtypes = [vp, ctypes.POINTER(d)]
        L.orc_atmosphere.argtypes = [d, ctypes.POINTER(d)]
        L.orc_atmosphere_biased.argtypes = [d, d, ctypes.POINTER(d)]
        L.orc_vcas_from_mach.restype = d
        L.orc_vcas_from_mach.argtypes = [d, d]
        L.orc_geodetic.argtypes = [d, d, d, ctypes.POINTER(d)]
        pd = ctypes.POINTER(d)
        L.orc_kinemat.restype = d
        L.orc_kinemat.argtypes = [pd, pd, i, i, d, d, d]
        L.orc_pid.restype = d
        L.orc_pid.argtypes = [pd, d, d, d, d, d, i, d]
        L.orc_gravity.argtypes = [d, d, d, pd]
        L.orc_fdm_mass.argtypes = [vp, pd]
        L.orc_fdm_set_tanks.argtypes = [vp, pd]
        _LIB = L
    return _LIB


def snapshot_names():
    L = lib()
    return [L.orc_fdm_snapshot_name(k).decode() for k in range(L.orc_fdm_n_snapshot())]


def prop_names():
    L = lib()
    return [L.orc_fdm_prop_name(k).decode() for k in range(L.orc_fdm_n_props())]


def comp_names():
    L = lib()
    return [(L.orc_fdm_comp_name(k).decode(), L.orc_fdm_comp_type(k)) for k in range(L.orc_fdm_n_comps())]


class OracleFdm:
    """One F-16.  ``dt`` = 1/sim_freq; ``fcs_dt`` = the FCS components' latched dt (1/120)."""

    def __init__(self, dt: float = 1.0 / 60.0, fcs_dt: float = 1.0 / 120.0):
        self._L = lib()
        self._h = self._L.orc_fdm_create(dt, fcs_dt)
        self._snap = np.zeros(self._L.orc_fdm_n_snapshot())
        self._names = snapshot_names()
        self._idx = {n: k for k, n in enumerate(self._names)}

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.orc_fdm_destroy(self._h)
            self._h = None

    def reset(self, lon_deg=120.0, lat_geod_deg=60.0, h_sl_ft=20000.0, psi_deg=0.0, u_fps=800.0, v_fps=0.0,
              w_fps=0.0, p=0.0, q=0.0, r=0.0, phi_deg=0.0, theta_deg=0.0):
        self._L.orc_fdm_reset(self._h, lon_deg, lat_geod_deg, h_sl_ft, psi_deg, u_fps, v_fps, w_fps, p, q, r,
                              phi_deg, theta_deg)

    def set_controls(self, aileron, elevator, rudder, throttle):
        self._L.orc_fdm_set_controls(self._h, aileron, elevator, rudder, throttle)

    def run(self, nframes: int = 1):
        self._L.orc_fdm_run(self._h, nframes)

    def snapshot(self) -> np.ndarray:
        self._L.orc_fdm_snapshot(self._h, self._snap.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
        return self._snap.copy()

    def snapshot_dict(self):
        return dict(zip(self._names, self.snapshot()))

    def get(self, name):
        return self.snapshot()[self._idx[name]]

    def props(self) -> np.ndarray:
        out = np.zeros(self._L.orc_fdm_n_props())
        self._L.orc_fdm_get_props(self._h, out.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
        return out

    def set_props(self, arr):
        arr = np.ascontiguousarray(arr, dtype=np.float64)
        self._L.orc_fdm_set_props(self._h, arr.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))

    def props_dict(self):
        return dict(zip(prop_names(), self.props()))

    def pid_state(self) -> np.ndarray:
        out = np.zeros((self._L.orc_fdm_n_comps(), 3))
        self._L.orc_fdm_get_pid(self._h, out.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
        return out


def atmosphere(h_ft: float, delta_T_R: float = None):
    """ISA-1976 at geometric altitude ``h_ft``; ``delta_T_R`` = the property atmosphere/delta-T (a temperature bias over the
    whole profile, J/models/atmosphere/FGStandardAtmosphere.cpp:343-353) -- used by the known-answer tests only."""
    out = (ctypes.c_double * 6)()
    if delta_T_R is None:
        lib().orc_atmosphere(h_ft, out)
    else:
        lib().orc_atmosphere_biased(h_ft, delta_T_R, out)
    return dict(zip(["T", "P", "rho", "a", "density_altitude", "pressure_altitude"], list(out)))


def geodetic(x, y, z):
    out = (ctypes.c_double * 6)()
    lib().orc_geodetic(x, y, z, out)
    return dict(zip(["lon", "lat_gc", "lat_geod", "geod_alt", "radius", "slr"], list(out)))


def kinemat(detents, times, input_, output, dt, noscale=False) -> float:
    """One FGKinemat::Run (oracle restatement of J/models/flight_control/FGKinemat.cpp:99-157)."""
    n = len(detents)
    D = (ctypes.c_double * n)(*detents)
    T = (ctypes.c_double * n)(*times)
    return lib().orc_kinemat(D, T, n, int(noscale), input_, output, dt)


class Pid:
    """One FGPID component (oracle restatement of J/models/flight_control/FGPID.cpp:154-214); int_type 1 rect, 2 trap, 3 ab2, 4 ab3."""

    def __init__(self, kp=0.0, ki=0.0, kd=0.0, int_type=4, dt=1.0 / 120.0):
        self.kp, self.ki, self.kd, self.int_type, self.dt = kp, ki, kd, int_type, dt
        self.state = (ctypes.c_double * 3)(0.0, 0.0, 0.0)

    def run(self, input_, trigger=0.0) -> float:
        return lib().orc_pid(self.state, input_, trigger, self.kp, self.ki, self.kd, self.int_type, self.dt)


def gravity(x, y, z):
    out = (ctypes.c_double * 3)()
    lib().orc_gravity(x, y, z, out)
    return np.array(list(out))


def mass_properties(f: "OracleFdm"):
    out = (ctypes.c_double * 23)()
    lib().orc_fdm_mass(f._h, out)
    o = np.array(list(out))
    return {"weight": o[0], "mass": o[1], "cg": o[2:5], "J": o[5:14].reshape(3, 3), "Jinv": o[14:23].reshape(3, 3)}


def set_tanks(f: "OracleFdm", contents):
    t = (ctypes.c_double * 4)(*contents)
    lib().orc_fdm_set_tanks(f._h, t)
