"""The product's flight-dynamics SOURCE on the host, against the CPU oracle -- no GPU needed.

tests/native/fdm_host.cpp compiles csrc/fdm_core.cuh + csrc/gen/f16_gen.cuh + csrc/fmath.cuh (the files libacs.so is built
from) with g++ and drives them the way k_fdm_reset / k_set_controls / k_fdm_run (full frame) and k_env_substeps (lean frame +
fdm_refresh) do.  It is test infrastructure, not a CPU path of the simulator.  The comparison and its bounds are those of
tests/test_fdm_gpu.py: state and outputs <= 1e-9 relative after a reset, one frame and one 12-frame interaction step,
<= 1e-8 after 120 frames.  The device build differs from this host build only in rounding (hardware reciprocal seeds, FMA
contraction), so a regression in the stage functions, the generated FCS / aero code or the table handling shows up here, in
the CPU-only suite, before it reaches a GPU box."""
import ctypes
import subprocess
from pathlib import Path

import numpy as np
import pytest

from oracle.fdm import OracleFdm
from tests.fdm_parity import field_scale, oracle_named_state, random_controls, random_ics

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def host_lib(tmp_path_factory):
    so = tmp_path_factory.mktemp("fdm_host") / "fdm_host.so"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-Wno-unknown-pragmas",
                           "-I", str(ROOT / "aircombat_selfplay_b200" / "csrc"), "-x", "c++",
                           str(ROOT / "tests" / "native" / "fdm_host.cpp"), "-o", str(so)])
    L = ctypes.CDLL(str(so))
    d, vp, i, pd = ctypes.c_double, ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_double)
    L.fh_create.restype = vp
    L.fh_create.argtypes = [d, d]
    L.fh_destroy.argtypes = [vp]
    L.fh_reset.argtypes = [vp, pd]
    L.fh_set_controls.argtypes = [vp, pd]
    L.fh_run.argtypes = [vp, i, i]
    L.fh_get.argtypes = [vp, pd, pd]
    L.fh_set_state.argtypes = [vp, pd]
    L.fh_state_name.restype = ctypes.c_char_p
    L.fh_state_name.argtypes = [i]
    L.fh_out_name.restype = ctypes.c_char_p
    L.fh_out_name.argtypes = [i]
    return L


class HostFdm:
    """One aircraft of the host build, with the surface of capi.FdmBatch that the parity comparison uses."""

    def __init__(self, L, lean: bool):
        self.L, self.lean = L, int(lean)
        self.h = L.fh_create(1.0 / 60.0, 1.0 / 120.0)
        self.state_names = [L.fh_state_name(k).decode() for k in range(L.fh_n_state())]
        self.output_names = [L.fh_out_name(k).decode() for k in range(L.fh_n_out())]

    def __del__(self):
        self.L.fh_destroy(self.h)

    @staticmethod
    def _p(a):
        return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))

    def reset(self, ic):
        self.L.fh_reset(self.h, self._p(np.ascontiguousarray(ic, dtype=np.float64)))

    def set_controls(self, u):
        self.L.fh_set_controls(self.h, self._p(np.ascontiguousarray(u, dtype=np.float64)))

    def run(self, n):
        self.L.fh_run(self.h, n, self.lean)

    def set_state(self, st):
        self.L.fh_set_state(self.h, self._p(np.ascontiguousarray(st, dtype=np.float64)))

    def get(self):
        st, out = np.zeros(len(self.state_names)), np.zeros(len(self.output_names))
        self.L.fh_get(self.h, self._p(st), self._p(out))
        return st, out


def _worst(hosts, oracles):
    worst, where = 0.0, None
    for i, (h, f) in enumerate(zip(hosts, oracles)):
        st, out = h.get()
        d = oracle_named_state(f)
        for k, name in enumerate(h.state_names):
            if name in d:
                e = abs(st[k] - d[name]) / field_scale(name, d[name])
                if e > worst:
                    worst, where = e, (name, i, st[k], d[name])
        for k, name in enumerate(h.output_names):
            if name in d:
                e = abs(out[k] - d[name]) / max(1.0, abs(d[name]))
                if e > worst:
                    worst, where = e, ("out:" + name, i, out[k], d[name])
    return worst, where


@pytest.mark.parametrize("lean", [False, True], ids=["full_frame", "lean_frame"])
def test_product_fdm_source_matches_the_oracle(host_lib, lean):
    n = 12
    rng = np.random.default_rng(11)
    ic = random_ics(rng, n)
    hosts = [HostFdm(host_lib, lean) for _ in range(n)]
    oracles = [OracleFdm() for _ in range(n)]
    for h, f, c in zip(hosts, oracles, ic):
        h.reset(c)
        f.reset(*c)
    w, where = _worst(hosts, oracles)
    assert w < 1e-9, ("reset", where)
    names = set(hosts[0].state_names) | set(hosts[0].output_names)
    covered = names & set(oracle_named_state(oracles[0]))
    assert len(covered) >= 90, len(covered)         # core state, carried properties, PID states and outputs are all compared

    def advance(frames, u=None):
        for h, f, c in zip(hosts, oracles, u if u is not None else [None] * n):
            if c is not None:
                h.set_controls(c)
                f.set_controls(*c)
            h.run(frames)
            f.run(frames)

    advance(1, random_controls(rng, n))                 # one frame
    w, where = _worst(hosts, oracles)
    assert w < 1e-9, ("one frame", where)
    advance(11)                                          # the rest of a 12-frame interaction step
    w, where = _worst(hosts, oracles)
    assert w < 1e-9, ("one step", where)
    for _ in range(9):                                   # short-horizon drift: 120 frames, new commands every step
        advance(12, random_controls(rng, n))
    w, where = _worst(hosts, oracles)
    assert w < 1e-8, ("120 frames", where)


def test_lean_and_full_frames_agree(host_lib):
    """The throughput kernel's lean frame recomputes what the full frame keeps; same arithmetic, so the two host builds stay
    together far inside the parity bound."""
    rng = np.random.default_rng(5)
    ic = random_ics(rng, 4)
    a = [HostFdm(host_lib, False) for _ in ic]
    b = [HostFdm(host_lib, True) for _ in ic]
    for x, y, c in zip(a, b, ic):
        x.reset(c)
        y.reset(c)
    for _ in range(5):
        u = random_controls(rng, len(ic))
        for x, y, c in zip(a, b, u):
            x.set_controls(c)
            y.set_controls(c)
            x.run(12)
            y.run(12)
    for x, y in zip(a, b):
        (sa, oa), (sb, ob) = x.get(), y.get()
        np.testing.assert_allclose(sa, sb, rtol=1e-11, atol=1e-11)
        np.testing.assert_allclose(oa, ob, rtol=1e-11, atol=1e-11)


def test_single_step_deltas_from_identical_states_over_a_wide_envelope(host_lib):
    """north_star: single-step state deltas <= 1e-9 from identical states.  The oracle flies aggressive random commands from
    random attitudes, rates, speeds (sub- and supersonic: the Rayleigh branch of Vcas), altitudes and latitudes; before every
    12-frame step its state is injected into the host build of the product source (what load_state does), both advance one
    step, and every state word and output is compared.  Both frames (full / lean)."""
    rng = np.random.default_rng(2)
    full, lean = HostFdm(host_lib, False), HostFdm(host_lib, True)
    names = full.state_names
    worst, where = 0.0, None
    mach, alpha, alt = [], [], []
    for n in range(40):
        ic = np.zeros(12)
        ic[0], ic[1], ic[2], ic[3] = rng.uniform(100, 140), rng.uniform(-70, 70), rng.uniform(4000, 45000), rng.uniform(0, 360)
        ic[4], ic[5], ic[6] = rng.uniform(250, 1800), rng.uniform(-30, 30), rng.uniform(-60, 60)
        ic[7:10] = rng.uniform(-0.3, 0.3, 3)
        ic[10], ic[11] = rng.uniform(-180, 180), rng.uniform(-60, 60)
        f = OracleFdm()
        f.reset(*ic)
        for step in range(60):
            d0 = oracle_named_state(f)
            st = np.array([d0[nm] for nm in names])
            if rng.random() < 0.7:
                u = np.array([rng.integers(0, 41) / 20 - 1, rng.integers(0, 41) / 20 - 1, rng.integers(0, 41) / 20 - 1, rng.integers(0, 30) / 58 + 0.4])
            else:          # full deflections, idle or full throttle
                u = np.array([rng.choice([-1.0, 1.0]), rng.choice([-1.0, 1.0]), rng.choice([-1.0, 1.0]), rng.choice([0.0, 0.9])])
            f.set_controls(*u)
            f.run(12)
            d1 = oracle_named_state(f)
            mach.append(d1["mach"]); alpha.append(d1["alpha_rad"]); alt.append(d1["h_sl_ft"])
            for h in (full, lean):
                h.set_state(st)
                h.set_controls(u)
                h.run(12)
                a, o = h.get()
                for k, nm in enumerate(names):
                    e = abs(a[k] - d1[nm]) / field_scale(nm, d1[nm])
                    if e > worst:
                        worst, where = e, (nm, n, step, h.lean, a[k], d1[nm])
                for k, nm in enumerate(h.output_names):
                    if nm in d1:
                        e = abs(o[k] - d1[nm]) / max(1.0, abs(d1[nm]))
                        if e > worst:
                            worst, where = e, ("out:" + nm, n, step, h.lean, o[k], d1[nm])
            if d1["h_sl_ft"] < 500.0:
                break
    assert worst < 1e-9, where
    assert max(mach) > 1.4 and min(mach) < 0.4 and max(alpha) > 0.25 and min(alpha) < -0.15 and max(alt) > 35000 and min(alt) < 3000
