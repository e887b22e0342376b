"""ctypes binding of the C ABI in ``include/acs.h`` (``libacs.so``, built by ``__graft_entry__.build``).

This is the only way the package reaches the device code.  There is no CPU fallback: if the
shared library is missing or a CUDA device is unavailable, calls raise.
"""
from __future__ import annotations

import ctypes
from pathlib import Path

import torch

_PKG = Path(__file__).resolve().parent
import os as _os
LIB_PATH = Path(_os.environ["ACS_LIB"]) if _os.environ.get("ACS_LIB") else _PKG / "libacs.so"   # ACS_LIB: tuning experiments only
_LIB = None


class AcsConfig(ctypes.Structure):
    _fields_ = [("n_envs", ctypes.c_int32), ("n_agents", ctypes.c_int32), ("sim_dt", ctypes.c_double),
                ("fcs_dt", ctypes.c_double)]


MAX_AGENTS, MAX_REWARDS, MAX_TERMS, INFO_DIM = 8, 12, 6, 4


class AcsRewardSpec(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int32), ("potential", ctypes.c_int32), ("scale", ctypes.c_double), ("p0", ctypes.c_double),
                ("p1", ctypes.c_double), ("p2", ctypes.c_double)]


class AcsTaskConfig(ctypes.Structure):
    _fields_ = [("n_envs", ctypes.c_int32), ("n_ego", ctypes.c_int32), ("n_enm", ctypes.c_int32), ("substeps", ctypes.c_int32),
                ("max_steps", ctypes.c_int32), ("sim_dt", ctypes.c_double), ("fcs_dt", ctypes.c_double),
                ("altitude_limit", ctypes.c_double), ("acc_limit", ctypes.c_double * 3), ("center", ctypes.c_double * 3),
                ("obs_kind", ctypes.c_int32), ("obs_dim", ctypes.c_int32), ("act_kind", ctypes.c_int32), ("shoot_dim", ctypes.c_int32),
                ("n_rewards", ctypes.c_int32), ("rewards", AcsRewardSpec * MAX_REWARDS), ("n_terms", ctypes.c_int32),
                ("terms", ctypes.c_int32 * MAX_TERMS), ("dones_before_rewards", ctypes.c_int32), ("team_mean", ctypes.c_int32),
                ("share_obs", ctypes.c_int32), ("reward_gate", ctypes.c_int32), ("launch_kind", ctypes.c_int32),
                ("use_artillery", ctypes.c_int32), ("use_baseline", ctypes.c_int32), ("max_attack_angle", ctypes.c_double),
                ("max_attack_distance", ctypes.c_double), ("min_attack_interval", ctypes.c_int32), ("lock_len", ctypes.c_int32),
                ("num_missiles", ctypes.c_int32 * MAX_AGENTS), ("n_missile_slots", ctypes.c_int32),
                ("init_state", (ctypes.c_double * 12) * MAX_AGENTS), ("heading_increments", ctypes.c_double * 3),
                ("check_interval", ctypes.c_double), ("seed", ctypes.c_uint64), ("env_offset", ctypes.c_int32),
                ("reserved", ctypes.c_int32), ("curriculum_rule", ctypes.c_int32), ("curriculum_window", ctypes.c_int32),
                ("curriculum_threshold", ctypes.c_double)]


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
        L.acs_env_create.argtypes = [ctypes.POINTER(AcsTaskConfig), i, ctypes.POINTER(vp)]
        L.acs_env_destroy.argtypes = [vp]
        L.acs_env_set_init_states.argtypes = [vp, vp]
        L.acs_env_set_seed.argtypes = [vp, ctypes.c_uint64, vp]
        L.acs_env_set_stage_init_states.argtypes = [vp, i, vp]
        L.acs_env_reset.argtypes = [vp, vp, vp, vp, vp]
        L.acs_env_step.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, i, vp]
        L.acs_env_arena_info.argtypes = [vp, i, ctypes.POINTER(i), ctypes.POINTER(i), ctypes.POINTER(i)]
        L.acs_env_arena_field_name.restype = cp
        L.acs_env_arena_field_name.argtypes = [i, i]
        L.acs_env_get_arena.argtypes = [vp, i, vp, vp]
        L.acs_env_set_arena.argtypes = [vp, i, vp, vp]
        L.acs_env_arena_ptr.argtypes = [vp, i, ctypes.POINTER(vp)]
        L.acs_env_fdm.restype = vp
        L.acs_env_fdm.argtypes = [vp]
        L.acs_env_set_option.argtypes = [vp, cp, i]
        L.acs_env_get_option.argtypes = [vp, cp, ctypes.POINTER(i)]
        L.acs_env_set_timing.argtypes = [vp, i]
        L.acs_env_get_timing.argtypes = [vp, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(i), i]
        L.acs_bench_fp64_peak.argtypes = [i, ctypes.POINTER(ctypes.c_double)]
        L.acs_debug_fmath.argtypes = [i, vp, vp, vp, vp, i, vp]
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


ARENAS = {"fdm": 0, "out": 1, "ac_d": 2, "ac_i": 3, "env_d": 4, "env_i": 5, "ms_d": 6, "ms_i": 7}


def task_config(spec, n_envs: int, seed: int = 0, env_offset: int = 0) -> AcsTaskConfig:
    """Pack a ``taskspec.TaskSpec`` into the C struct of include/acs.h."""
    c = AcsTaskConfig()
    c.n_envs, c.n_ego, c.n_enm, c.substeps, c.max_steps = n_envs, spec.n_ego, spec.n_enm, spec.substeps, spec.max_steps
    c.sim_dt, c.fcs_dt, c.altitude_limit = spec.dt, spec.fcs_dt, spec.altitude_limit
    for k in range(3):
        c.acc_limit[k], c.center[k], c.heading_increments[k] = spec.acc_limit[k], spec.center[k], spec.heading_increments[k]
    c.obs_kind, c.obs_dim, c.act_kind, c.shoot_dim = spec.obs_kind, spec.obs_dim, spec.act_kind, spec.shoot_dim
    if len(spec.rewards) > MAX_REWARDS or len(spec.terminations) > MAX_TERMS or spec.n_agents > MAX_AGENTS:
        raise AcsError("task exceeds ACS_MAX_REWARDS / ACS_MAX_TERMS / ACS_MAX_AGENTS")
    c.n_rewards = len(spec.rewards)
    for k, r in enumerate(spec.rewards):
        c.rewards[k].kind, c.rewards[k].potential, c.rewards[k].scale = r.kind, int(r.potential), r.scale
        c.rewards[k].p0, c.rewards[k].p1, c.rewards[k].p2 = r.p0, r.p1, r.p2
    c.n_terms = len(spec.terminations)
    for k, t in enumerate(spec.terminations):
        c.terms[k] = t
    c.dones_before_rewards, c.team_mean, c.share_obs = int(spec.dones_before_rewards), int(spec.team_mean), int(spec.share_obs)
    c.reward_gate, c.launch_kind, c.use_artillery, c.use_baseline = spec.reward_gate, spec.launch_kind, int(spec.use_artillery), int(spec.use_baseline)
    c.max_attack_angle = spec.max_attack_angle
    c.max_attack_distance = spec.max_attack_distance
    c.min_attack_interval, c.lock_len = spec.min_attack_interval, spec.lock_len
    for k in range(spec.n_agents):
        c.num_missiles[k] = spec.num_missiles[k] if spec.num_missiles else 0
        for j in range(12):
            c.init_state[k][j] = spec.init_states[k][j]
    c.n_missile_slots = spec.n_missile_slots
    c.check_interval = spec.check_interval
    c.seed, c.env_offset = seed, env_offset
    c.curriculum_rule, c.curriculum_window = int(getattr(spec, "curriculum_rule", 0)), int(getattr(spec, "curriculum_window", 0))
    c.curriculum_threshold = float(getattr(spec, "curriculum_threshold", 0.0))
    return c


class EnvBatch:
    """``n_envs`` environments of one task on one GPU -- the device-side counterpart of ``n_envs`` reference Env objects
    (reference envs/JSBSim/envs/env_base.py).  Tensors in, tensors out; nothing here synchronises the stream."""

    def __init__(self, spec, n_envs: int, seed: int = 0, device: int = 0, env_offset: int = 0, device_share_obs: bool = False):
        """``device_share_obs``: also materialise share_obs [B, A, A*D] in HBM (the kernel writes it).  The default leaves
        it as what it is -- every agent's row is the same concatenation of all observations (reference
        envs/JSBSim/envs/env_base.py:183-189) -- and exposes it as a stride-0 view of ``obs``."""
        if not torch.cuda.is_available():
            raise AcsError("CUDA is not available; the simulator has no CPU fallback")
        self.spec, self.n_envs, self.n_agents = spec, n_envs, spec.n_agents
        self.device = torch.device("cuda", device)
        self.cfg = task_config(spec, n_envs, seed, env_offset)
        h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _check(lib().acs_env_create(ctypes.byref(self.cfg), device, ctypes.byref(h)))
        self._h = h
        B, A, D = n_envs, self.n_agents, spec.obs_dim
        # every per-step output lives in ONE device buffer so the host-facing layer fetches a step with a single D2H copy
        layout = [("obs", torch.float64, (B, A, D))]
        self.device_share_obs = bool(device_share_obs and spec.share_obs)
        if self.device_share_obs:
            layout.append(("share_obs", torch.float64, (B, A, A * D)))
        layout += [("rewards", torch.float64, (B, A)), ("info", torch.int32, (B, A, INFO_DIM)), ("dones", torch.uint8, (B, A)),
                   ("env_done", torch.uint8, (B,))]
        self.out_layout, off = [], 0
        for name, dt, shape in layout:
            nbytes = int(torch.tensor([], dtype=dt).element_size()) * int(torch.Size(shape).numel())
            self.out_layout.append((name, dt, shape, off, nbytes))
            off += (nbytes + 15) // 16 * 16
        self.out_buf = torch.zeros(off, dtype=torch.uint8, device=self.device)
        self._share_dev = None
        for name, dt, shape, o, nbytes in self.out_layout:
            setattr(self, "_share_dev" if name == "share_obs" else name, self.out_buf[o:o + nbytes].view(dt).view(shape))
        self.act_dim = 4 + spec.shoot_dim

    def close(self):
        if getattr(self, "_h", None):
            lib().acs_env_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_init_states(self, init_states):
        import numpy as np
        arr = np.ascontiguousarray(init_states, dtype=np.float64)
        assert arr.shape == (self.n_agents, 12)
        _check(lib().acs_env_set_init_states(self._h, arr.ctypes.data_as(ctypes.c_void_p)))

    def set_stage_init_states(self, stage: int, init_states):
        """Initial conditions of curriculum stage ``stage`` (acs.h acs_env_set_stage_init_states); envs reset from the stage in
        their "stage" field of the ``env_i`` arena."""
        import numpy as np
        arr = np.ascontiguousarray(init_states, dtype=np.float64)
        assert arr.shape == (self.n_agents, 12)
        _check(lib().acs_env_set_stage_init_states(self._h, int(stage), arr.ctypes.data_as(ctypes.c_void_p)))

    @property
    def share_obs(self):
        if not self.spec.share_obs:
            return None
        if self._share_dev is not None:
            return self._share_dev
        B, A, D = self.obs.shape
        return self.obs.view(B, 1, A * D).expand(B, A, A * D)

    def set_seed(self, seed: int):
        _check(lib().acs_env_set_seed(self._h, ctypes.c_uint64(seed & 0xFFFFFFFFFFFFFFFF), _stream()))

    def host_views(self, host_buf: torch.Tensor):
        """numpy views of a host copy of ``out_buf`` (same layout), keyed by output name."""
        views = {}
        for name, dt, shape, o, nbytes in self.out_layout:
            views[name] = host_buf[o:o + nbytes].view(dt).view(shape).numpy()
        return views

    def reset(self, env_mask: torch.Tensor | None = None):
        _check(lib().acs_env_reset(self._h, _ptr(env_mask, torch.uint8), _ptr(self.obs, torch.float64),
                                   _ptr(self._share_dev, torch.float64), _stream()))
        return self.obs, self.share_obs

    def step(self, actions: torch.Tensor, auto_reset: bool = False):
        """actions: int32 [n_envs, n_agents, 4 + shoot_dim] low-level discrete actions."""
        assert actions.shape == (self.n_envs, self.n_agents, self.act_dim), actions.shape
        _check(lib().acs_env_step(self._h, _ptr(actions, torch.int32), _ptr(self.obs, torch.float64),
                                  _ptr(self._share_dev, torch.float64), _ptr(self.rewards, torch.float64),
                                  _ptr(self.dones, torch.uint8), _ptr(self.info, torch.int32), _ptr(self.env_done, torch.uint8),
                                  int(auto_reset), _stream()))
        return self.obs, self.share_obs, self.rewards, self.dones, self.info

    def set_option(self, name: str, value: int):
        """Tuning knobs of include/acs.h (``frame_split``: 0 one thread per aircraft, 1 two-warp frame, -1 auto)."""
        _check(lib().acs_env_set_option(self._h, name.encode(), int(value)))

    def get_option(self, name: str) -> int:
        out = ctypes.c_int()
        _check(lib().acs_env_get_option(self._h, name.encode(), ctypes.byref(out)))
        return out.value

    # ---- measurement
    def set_timing(self, on: bool):
        _check(lib().acs_env_set_timing(self._h, int(on)))

    def get_timing(self, reset: bool = True):
        """({'substeps': ms, 'post': ms, 'reset': ms, 'missiles': ms}, n_steps) of the recorded steps (synchronises)."""
        ms, n = (ctypes.c_double * 4)(), ctypes.c_int()
        _check(lib().acs_env_get_timing(self._h, ms, ctypes.byref(n), int(reset)))
        return {"substeps": ms[0], "post": ms[1], "reset": ms[2], "missiles": ms[3]}, n.value

    # ---- introspection (parity tests, rendering)
    def arena(self, name: str):
        """Returns (field names, tensor [n_fields, n_per_field]) -- a copy."""
        which = ARENAS[name]
        nf, per, ii = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        _check(lib().acs_env_arena_info(self._h, which, ctypes.byref(nf), ctypes.byref(per), ctypes.byref(ii)))
        dt = torch.int32 if ii.value else torch.float64
        t = torch.empty((nf.value, per.value), dtype=dt, device=self.device)
        _check(lib().acs_env_get_arena(self._h, which, ctypes.c_void_p(t.data_ptr()), _stream()))
        names = [lib().acs_env_arena_field_name(which, k).decode() for k in range(nf.value)]
        return names, t

    def arena_view(self, name: str):
        """(field names, tensor [n_fields, n_per_field]) aliasing the library's arena -- no copy; read-only by convention."""
        which = ARENAS[name]
        nf, per, ii = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        _check(lib().acs_env_arena_info(self._h, which, ctypes.byref(nf), ctypes.byref(per), ctypes.byref(ii)))
        ptr = ctypes.c_void_p()
        _check(lib().acs_env_arena_ptr(self._h, which, ctypes.byref(ptr)))

        class _Holder:       # CUDA array interface v3: lets torch wrap a raw device pointer it does not own
            __cuda_array_interface__ = {"shape": (nf.value, per.value), "typestr": "<i4" if ii.value else "<f8",
                                        "data": (ptr.value, False), "version": 3, "strides": None}
        t = torch.as_tensor(_Holder(), device=self.device)
        names = [lib().acs_env_arena_field_name(which, k).decode() for k in range(nf.value)]
        return names, t

    def set_arena(self, name: str, t: torch.Tensor):
        assert t.is_cuda and t.is_contiguous()
        _check(lib().acs_env_set_arena(self._h, ARENAS[name], ctypes.c_void_p(t.data_ptr()), _stream()))


def fp64_peak_flops(device: int = 0) -> float:
    out = ctypes.c_double()
    _check(lib().acs_bench_fp64_peak(device, ctypes.byref(out)))
    return out.value


FMATH_OPS = ("div", "rcp", "sqrt", "rsqrt", "sincos", "sin", "exp", "log", "atan2", "acos", "tanh", "pow_ratio", "angle_sc",
             "sincos_small")


def fmath_probe(op: str, a: torch.Tensor, b=None):
    """One function of csrc/fmath.cuh over device arrays (include/acs.h: acs_debug_fmath) -> (out, out2).  Test helper."""
    assert a.is_cuda and a.dtype == torch.float64 and a.is_contiguous()
    out, out2 = torch.empty_like(a), torch.empty_like(a)
    bp = None
    if b is not None:
        assert b.is_cuda and b.dtype == torch.float64 and b.is_contiguous() and b.numel() == a.numel()
        bp = ctypes.c_void_p(b.data_ptr())
    _check(lib().acs_debug_fmath(FMATH_OPS.index(op), ctypes.c_void_p(a.data_ptr()), bp, ctypes.c_void_p(out.data_ptr()),
                                 ctypes.c_void_p(out2.data_ptr()), a.numel(), _stream()))
    return out, out2
