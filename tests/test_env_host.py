"""The product's env-layer building blocks on the host, against the CPU oracle -- no GPU needed.

tests/native/env_host.cpp compiles csrc/env_core.cuh + csrc/fmath.cuh (the files libacs.so is built from) with g++: the keyed
RNG, LLA2NEU / the altitude half of NEU2LLA, get_AO_TA_R, one proportional-navigation missile substep (guidance + state
transition) and the posture reward shaping.  They are compared with oracle/env_oracle.py -- which reproduces golden episodes
recorded from the reference's own Python (tests/test_oracle_env_golden.py) -- at the tolerances of the GPU parity tests:
1e-9 on geometry / missile state from identical inputs, 1e-6 on reward terms, the RNG bit for bit.  Test infrastructure only:
the simulator has no CPU path."""
import ctypes
import math
import subprocess
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import pytest

from oracle import env_oracle as eo

ROOT = Path(__file__).resolve().parents[1]
D = ctypes.c_double
PD = ctypes.POINTER(D)


@pytest.fixture(scope="module")
def L(tmp_path_factory):
    so = tmp_path_factory.mktemp("env_host") / "env_host.so"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-Wno-unknown-pragmas",
                           "-I", str(ROOT / "aircombat_selfplay_b200" / "csrc"), "-x", "c++",
                           str(ROOT / "tests" / "native" / "env_host.cpp"), "-o", str(so)])
    lib = ctypes.CDLL(str(so))
    lib.eh_u01.restype = D
    lib.eh_u01.argtypes = [ctypes.c_uint64] + [ctypes.c_int64] * 5
    lib.eh_lla2neu.argtypes = [PD, PD, PD]
    lib.eh_neu2alt.restype = D
    lib.eh_neu2alt.argtypes = [PD, PD]
    lib.eh_ao_ta_r.argtypes = [PD, PD, ctypes.c_int, PD]
    lib.eh_missile_step.argtypes = [ctypes.c_int, PD, PD, PD, D, PD]
    lib.eh_posture.argtypes = [ctypes.c_int, ctypes.c_int, D, D, D, D, PD]
    lib.eh_delta_heading.restype = D
    lib.eh_delta_heading.argtypes = [D, D]
    return lib


def _p(a):
    return a.ctypes.data_as(PD)


def test_keyed_rng_is_bit_identical(L):
    rng = np.random.default_rng(0)
    for _ in range(2000):
        seed = int(rng.integers(0, 2 ** 63))
        k = [int(x) for x in rng.integers(0, 2 ** 40, 5)]
        assert L.eh_u01(seed, *k) == eo.u01(seed, *k)
    assert 0.0 <= L.eh_u01(0, 0, 0, 0, 0, 0) < 1.0


def test_local_frame_transforms(L):
    rng = np.random.default_rng(1)
    worst_neu = worst_alt = 0.0
    for _ in range(3000):
        origin = np.array([rng.uniform(-180, 180), rng.uniform(-75, 75), rng.uniform(0, 1000)])
        lla = origin + np.array([rng.uniform(-1.5, 1.5), rng.uniform(-1.5, 1.5), rng.uniform(0, 15000)])
        neu = np.zeros(3)
        L.eh_lla2neu(_p(origin), _p(lla), _p(neu))
        ref = eo.LLA2NEU(lla[0], lla[1], lla[2], *origin)
        worst_neu = max(worst_neu, float(np.max(np.abs(neu - ref))))
        alt = L.eh_neu2alt(_p(origin), _p(np.ascontiguousarray(ref)))
        worst_alt = max(worst_alt, abs(alt - eo.NEU2alt(ref[0], ref[1], ref[2], *origin)), abs(alt - lla[2]))
    assert worst_neu < 2e-8          # metres, on coordinates up to 2e5 m from an ECEF difference of 6e6 m numbers (1e-9 relative)
    assert worst_alt < 2e-8          # Fukushima's closed form vs the oracle's iteration (You 2000) vs the altitude that went in


def test_ao_ta_r(L):
    rng = np.random.default_rng(2)
    for k in range(4000):
        ego = np.concatenate([rng.uniform(-5e4, 5e4, 2), rng.uniform(1e3, 1e4, 1), rng.uniform(-350, 350, 3)])
        enm = np.concatenate([rng.uniform(-5e4, 5e4, 2), rng.uniform(1e3, 1e4, 1), rng.uniform(-350, 350, 3)])
        two_d = bool(k & 1)
        out = np.zeros(4)
        L.eh_ao_ta_r(_p(ego), _p(enm), int(two_d), _p(out))
        AO, TA, R, side = eo.get_AO_TA_R(ego, enm, two_d)
        # acos amplifies one ulp of its argument by 1 / sin(angle): the bound the GPU tests state
        assert abs(out[0] - AO) <= 1e-9 + 4e-15 / max(math.sin(AO), 1e-12)
        assert abs(out[1] - TA) <= 1e-9 + 4e-15 / max(math.sin(TA), 1e-12)
        assert abs(out[2] - R) <= 1e-9 * R and out[3] == side
    # the reference's own yaml geometry: nose on the target (AO ~ 0): relative accuracy towards small angles
    ego, enm = np.array([0.0, 0.0, 6000.0, 240.0, 0.0, 0.0]), np.array([12000.0, 1e-3, 6000.0, -240.0, 0.0, 0.0])
    out = np.zeros(4)
    L.eh_ao_ta_r(_p(ego), _p(enm), 0, _p(out))
    AO, TA, R, side = eo.get_AO_TA_R(ego, enm)
    assert out[0] == pytest.approx(AO, abs=1e-9) and out[1] == pytest.approx(TA, abs=1e-9)


def _oracle_missile(kind, center, st, tg, dt):
    m = eo.Missile.__new__(eo.Missile)
    m.pr, m.kind, m.dt, m.center = eo.MISSILE_PARAMS[kind], kind, dt, tuple(center)
    m.position, m.velocity = np.array(st[0:3]), np.array(st[3:6])
    m.posture = np.array([0.0, st[6], st[7]])
    m.alt, m.t, m.m, m.dtheta, m.dphi = st[8], st[9], st[10], st[11], st[12]
    m.target = SimpleNamespace(position=np.array(tg[0:3]), velocity=np.array([tg[3], tg[4], tg[5]]))
    return m


@pytest.mark.parametrize("kind", [0, 1])
def test_missile_substeps_from_identical_states(L, kind):
    """One substep from identical states (north_star's single-step bound), along whole fly-outs: the oracle missile flies
    towards a manoeuvring target; before every substep its state is handed to the host build of the product code."""
    rng = np.random.default_rng(10 + kind)
    dt, worst = 1.0 / 60.0, 0.0
    center = np.array([120.0, 60.0, 0.0])
    pr = eo.MISSILE_PARAMS[kind]
    for shot in range(60):
        v0, psi, th = rng.uniform(200, 400), rng.uniform(-math.pi, math.pi), rng.uniform(-0.3, 0.3)
        st = np.array([rng.uniform(-2e4, 2e4), rng.uniform(-2e4, 2e4), rng.uniform(3000, 9000),
                       v0 * math.cos(th) * math.cos(psi), v0 * math.cos(th) * math.sin(psi), v0 * math.sin(th),
                       th, psi, 0.0, 0.0, pr["m0"], 0.0, 0.0])
        st[8] = eo.NEU2alt(st[0], st[1], st[2], *center)
        rng_t = rng.uniform(3000, 15000)
        brg = psi + rng.uniform(-0.5, 0.5)
        tpos = st[0:3] + np.array([rng_t * math.cos(brg), rng_t * math.sin(brg), rng.uniform(-1500, 1500)])
        tvel = rng.uniform(-300, 300, 3) * np.array([1, 1, 0.2])
        for k in range(int(rng.integers(60, 600))):
            tg = np.concatenate([tpos, tvel])
            m = _oracle_missile(kind, center, st, tg, dt)
            m.t += dt
            (ny, nz), dist = m._guidance()
            m._state_trans((ny, nz))
            got, out = st.copy(), np.zeros(3)
            L.eh_missile_step(kind, _p(center), _p(got), _p(tg), dt, _p(out))
            ref = np.concatenate([m.position, m.velocity, [m.posture[1], m.posture[2], m.alt, m.t, m.m, m.dtheta, m.dphi]])
            scale = np.array([1e4, 1e4, 1e4, 1e3, 1e3, 1e3, 1, 1, 1e4, 1, 1e2, 1, 1])
            err = float(np.max(np.abs(got - ref) / scale))
            err = max(err, abs(out[0] - ny) / 50.0, abs(out[1] - nz) / 50.0, abs(out[2] - dist) / 1e4)
            worst = max(worst, err)
            st = ref                                  # the oracle's state goes on
            tpos = tpos + dt * tvel
            if k % 30 == 0:
                tvel = tvel + rng.uniform(-40, 40, 3) * np.array([1, 1, 0.2])
            if dist < 30.0 or m.t > pr["t_max"] or np.linalg.norm(m.velocity) < pr["v_min"]:
                break
    print(f"missile kind {kind}: worst single-substep deviation {worst:.2e}")
    assert worst < 1e-9, worst


def test_posture_reward_shaping(L):
    rng = np.random.default_rng(3)
    out = np.zeros(2)
    for _ in range(5000):
        AO, TA = rng.uniform(1e-7, math.pi), rng.uniform(1e-7, math.pi)
        R, td = rng.uniform(0.05, 40.0), rng.choice([3.0, 5.0, 8.0])
        for ov in (0, 1, 2):
            for rv in (0, 1, 2, 3):
                L.eh_posture(ov, rv, AO, TA, R, td, _p(out))
                assert out[0] == pytest.approx(eo.posture_orientation(ov, AO, TA), abs=1e-9)
                assert out[1] == pytest.approx(eo.posture_range(rv, R, td), abs=1e-9)


def test_delta_heading_wraps_like_python_modulo(L):
    for tgt, psi in [(10.0, 350.0), (350.0, 10.0), (180.0, 0.0), (0.0, 180.0), (-170.0, 170.0), (725.0, 3.0), (90.0, 90.0)]:
        ang = (tgt - psi) % 360.0
        ang = ang - 360.0 if ang > 180.0 else ang
        assert L.eh_delta_heading(tgt, psi) == pytest.approx(ang, abs=1e-12)
