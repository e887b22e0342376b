"""ctypes binding of the C ABI in ``include/acs.h`` (``libacs.so``, built by ``__graft_entry__.build``).

This is the only way the package reaches the device code.  There is no CPU fallback: if the
shared library is missing or a CUDA device is unavailable, calls raise.
"""
from __future__ import annotations

import ctypes
from pathlib import Path

import torch

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libacs.so"
_LIB = None


class AcsConfig(ctypes.Structure):
    _fields_ = [("n_envs", ctypes.c_int32), ("n_agents", ctypes.c_int32), ("sim_dt", ctypes.c_double),
                ("fcs_dt", ctypes.c_double)]


class AcsError(RuntimeError):
    pass


def lib():
    global _LIB
    if _LIB is None:
        if not LIB_PATH.exists():
            raise AcsError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(the CUDA extension is required; there is no CPU fallback)")
        L = ctypes.CDLL(str(LIB_PATH))
        vp, i, cp = ctypes.c_void_p, ctypes.c_int, ctypes.c_char_p
        L.acs_last_error.restype = cp
        L.acs_create.argtypes = [ctypes.POINTER(AcsConfig), i, ctypes.POINTER(vp)]
        L.acs_destroy.argtypes = [vp]
        L.acs_n_rows.argtypes = [vp]
        L.acs_state_field_name.restype = cp
        L.acs_state_field_name.argtypes = [i]
        L.acs_output_field_name.restype = cp
        L.acs_output_field_name.argtypes = [i]
        L.acs_fdm_reset.argtypes = [vp, vp, vp, vp]
        L.acs_fdm_set_controls.argtypes = [vp, vp, vp]
        L.acs_fdm_run.argtypes = [vp, i, vp, vp]
        L.acs_get_state.argtypes = [vp, vp, vp]
        L.acs_set_state.argtypes = [vp, vp, vp]
        L.acs_get_outputs.argtypes = [vp, vp, vp]
        _LIB = L
    return _LIB


def _check(rc):
    if rc != 0:
        raise AcsError(lib().acs_last_error().decode())


def state_field_names():
    L = lib()
    return [L.acs_state_field_name(k).decode() for k in range(L.acs_n_state_fields())]


def output_field_names():
    L = lib()
    return [L.acs_output_field_name(k).decode() for k in range(L.acs_n_output_fields())]


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t, dtype):
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous() and t.dtype == dtype, (t.device, t.dtype, t.is_contiguous())
    return ctypes.c_void_p(t.data_ptr())


class FdmBatch:
    """A batch of F-16 flight-dynamics models on one GPU -- the device-side counterpart of N
    ``jsbsim.FGFDMExec`` instances (reference envs/JSBSim/core/simulatior.py:165-188)."""

    def __init__(self, n_envs: int, n_agents: int = 1, sim_freq: int = 60, fcs_dt: float = 1.0 / 120.0,
                 device: int = 0):
        if not torch.cuda.is_available():
            raise AcsError("CUDA is not available; the simulator has no CPU fallback")
        self.device = torch.device("cuda", device)
        self.n_envs, self.n_agents = n_envs, n_agents
        self.n_rows = n_envs * n_agents
        cfg = AcsConfig(n_envs, n_agents, 1.0 / sim_freq, fcs_dt)
        h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _check(lib().acs_create(ctypes.byref(cfg), device, ctypes.byref(h)))
        self._h = h
        self.state_names = state_field_names()
        self.output_names = output_field_names()

    def close(self):
        if getattr(self, "_h", None):
            lib().acs_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self, ic: torch.Tensor, mask: torch.Tensor | None = None):
        """ic: [n_rows, 12] float64 (see acs.h acs_fdm_reset); mask: [n_rows] uint8 or None."""
        assert ic.shape == (self.n_rows, 12)
        _check(lib().acs_fdm_reset(self._h, _ptr(mask, torch.uint8), _ptr(ic, torch.float64), _stream()))

    def set_controls(self, u: torch.Tensor):
        assert u.shape == (self.n_rows, 4)
        _check(lib().acs_fdm_set_controls(self._h, _ptr(u, torch.float64), _stream()))

    def run(self, n_frames: int = 1, alive: torch.Tensor | None = None):
        _check(lib().acs_fdm_run(self._h, n_frames, _ptr(alive, torch.uint8), _stream()))

    def get_state(self) -> torch.Tensor:
        out = torch.empty((len(self.state_names), self.n_rows), dtype=torch.float64, device=self.device)
        _check(lib().acs_get_state(self._h, _ptr(out, torch.float64), _stream()))
        return out

    def set_state(self, st: torch.Tensor):
        assert st.shape == (len(self.state_names), self.n_rows)
        _check(lib().acs_set_state(self._h, _ptr(st, torch.float64), _stream()))

    def get_outputs(self) -> torch.Tensor:
        out = torch.empty((len(self.output_names), self.n_rows), dtype=torch.float64, device=self.device)
        _check(lib().acs_get_outputs(self._h, _ptr(out, torch.float64), _stream()))
        return out
