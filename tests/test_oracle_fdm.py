"""CPU tests of the oracle FDM against closed-form / known-answer material that IS in the reference tree
(SURVEY.md section 8c): JSBSim's own test formulas under envs/JSBSim/data/tests/*.py.  These pin the pieces of the
restatement that can be pinned without a runnable JSBSim."""
import math

import numpy as np
import pytest

from oracle import fdm as ofdm


def isa_reference(h_ft):
    """ISA-1976 from the formulas of reference envs/JSBSim/data/tests/TestStdAtmosphere.py:49-110 (SI, then converted)."""
    g0, Mair, Rstar, Re = 9.80665, 28.9645e-3, 8.31432, 6356766.0
    Hb = [0.0, 11000.0, 20000.0, 32000.0, 47000.0, 51000.0, 71000.0, 84852.0]
    Tb = [288.15, 216.65, 216.65, 228.65, 270.65, 270.65, 214.65, 186.946]
    h = h_ft * 0.3048
    H = Re * h / (Re + h)
    P = 101325.0
    for b in range(7):
        L = (Tb[b + 1] - Tb[b]) / (Hb[b + 1] - Hb[b])
        top = min(H, Hb[b + 1])
        dH = top - Hb[b]
        if abs(L) > 1e-12:
            P *= (Tb[b] / (Tb[b] + L * dH)) ** (g0 * Mair / (Rstar * L))
        else:
            P *= math.exp(-g0 * Mair * dH / (Rstar * Tb[b]))
        if H <= Hb[b + 1]:
            T = Tb[b] + L * dH
            break
    rho = P * Mair / (Rstar * T)
    a = math.sqrt(1.4 * Rstar * T / Mair)
    return T * 1.8, P / 47.880258, rho / 515.378818, a / 0.3048


@pytest.mark.parametrize("h_ft", [0.0, 1000.0, 8202.1, 20000.0, 30000.0, 36089.0, 40000.0, 65000.0, 80000.0])
def test_isa_atmosphere(h_ft):
    got = ofdm.atmosphere(h_ft)
    T, P, rho, a = isa_reference(h_ft)
    assert got["T"] == pytest.approx(T, rel=2e-6)
    assert got["P"] == pytest.approx(P, rel=2e-5)
    assert got["rho"] == pytest.approx(rho, rel=2e-5)
    assert got["a"] == pytest.approx(a, rel=2e-6)


def test_gas_constant_matches_jsbsim_documentation():
    """envs/JSBSim/data/doc/spreadsheets/Standard_Atmosphere_constants.ods lists the value JSBSim uses: Reng = 1716.5572
    ft lbf / slug / R (= R* / M in British units); the oracle's follows from its sea-level state, P = rho Reng T."""
    a = ofdm.atmosphere(0.0)
    assert a["P"] / (a["rho"] * a["T"]) == pytest.approx(1716.5572, abs=5e-5)


def test_density_altitude_is_identity_on_a_standard_day():
    # CalculateDensityAltitude inverts the density profile (FGStandardAtmosphere.cpp:464-492)
    for h in [0.0, 5000.0, 20000.0, 36000.0, 45000.0]:
        assert ofdm.atmosphere(h)["density_altitude"] == pytest.approx(h, abs=1e-6 * max(1.0, h))


def test_vcas_equals_tas_at_sea_level_and_is_monotone():
    # FGJSBBase.cpp:245-296: at sea-level pressure Vcas = M * a0
    a0 = ofdm.atmosphere(0.0)["a"]
    for m in [0.1, 0.5, 0.9, 1.2, 1.8]:
        # supersonic: 10 fixed-point iterations of the Rayleigh formula (FGJSBBase.cpp:283-293) converge to ~1e-8
        assert ofdm.lib().orc_vcas_from_mach(m, 2116.228) == pytest.approx(m * a0, rel=1e-9 if m < 1 else 1e-6)
    p20 = ofdm.atmosphere(20000.0)["P"]
    v = [ofdm.lib().orc_vcas_from_mach(m, p20) for m in np.linspace(0.1, 1.5, 15)]
    assert all(b > a for a, b in zip(v, v[1:]))


def test_geodetic_roundtrip():
    # FGLocation::SetPositionGeodetic -> ComputeDerived (FGLocation.cpp:247-258, 283-370)
    a, b = 20925646.32546, 20855486.5951
    e2 = 1 - (b / a) ** 2
    for lat_deg, lon_deg, h in [(60.0, 120.0, 20000.0), (0.0, 0.0, 0.0), (-33.3, 18.4, 35000.0), (89.0, -170.0, 1000.0)]:
        lat, lon = math.radians(lat_deg), math.radians(lon_deg)
        N = a / math.sqrt(1 - e2 * math.sin(lat) ** 2)
        x = (N + h) * math.cos(lat) * math.cos(lon)
        y = (N + h) * math.cos(lat) * math.sin(lon)
        z = ((1 - e2) * N + h) * math.sin(lat)
        g = ofdm.geodetic(x, y, z)
        assert g["lat_geod"] == pytest.approx(lat, abs=1e-11)
        assert g["lon"] == pytest.approx(lon, abs=1e-12)
        assert g["geod_alt"] == pytest.approx(h, abs=1e-6)


def test_reset_reproduces_the_initial_conditions():
    f = ofdm.OracleFdm()
    f.reset(lon_deg=120.3, lat_geod_deg=59.8, h_sl_ft=21000.0, psi_deg=135.0, u_fps=750.0)
    d = f.snapshot_dict()
    assert d["lon_deg"] == pytest.approx(120.3, abs=1e-9)
    assert d["lat_geod_deg"] == pytest.approx(59.8, abs=1e-9)
    assert d["h_sl_ft"] == pytest.approx(21000.0, abs=1e-6)
    assert math.degrees(d["heading_rad"]) == pytest.approx(135.0, abs=1e-9)
    assert d["u_fps"] == pytest.approx(750.0, abs=1e-9)
    assert abs(d["v_fps"]) < 1e-9 and abs(d["w_fps"]) < 1e-9
    assert d["roll_rad"] == pytest.approx(0.0, abs=1e-12) and d["pitch_rad"] == pytest.approx(0.0, abs=1e-12)
    assert d["sim_time"] == 0.0
    # engine.init_running + get_steady_state: idle, N2 = IdleN2 (FGTurbine.cpp:604-616)
    assert d["N2"] == pytest.approx(53.0) and d["N1"] == pytest.approx(40.0)
    # fuel 2 x 3000 lb and nothing burnt during trim (FGPropulsion.cpp:166-167)
    assert d["tank0"] == 3000.0 and d["tank1"] == 3000.0


def test_attitude_conventions_known_answer():
    """The quaternion / matrix arithmetic JSBSim keeps in headers (FGQuaternion, FGMatrix33, FGColumnVector3: not in the
    reference tree) is pinned by the textbook definitions it implements: for random Euler angles and body velocities the
    reset state gives the angles back, a unit quaternion, NED velocity = (3-2-1 direction cosine matrix)^T (u, v, w) and
    |v_ECI| = |v_NED + Omega x r| (FGPropagate::SetInitialState, J/models/FGPropagate.cpp:143-186)."""
    rng = np.random.default_rng(4)
    omega = 7.292115e-5
    for _ in range(50):
        phi, tht, psi = rng.uniform(-179, 179), rng.uniform(-85, 85), rng.uniform(0.5, 359.5)
        uvw = np.array([rng.uniform(300, 1200), rng.uniform(-80, 80), rng.uniform(-120, 120)])
        lat, pqr = rng.uniform(-70, 70), rng.uniform(-0.5, 0.5, 3)
        f = ofdm.OracleFdm()
        f.reset(lon_deg=rng.uniform(-179, 179), lat_geod_deg=lat, h_sl_ft=rng.uniform(5000, 40000), psi_deg=psi,
                u_fps=uvw[0], v_fps=uvw[1], w_fps=uvw[2], p=pqr[0], q=pqr[1], r=pqr[2], phi_deg=phi, theta_deg=tht)
        d = f.snapshot_dict()
        assert math.degrees(d["roll_rad"]) == pytest.approx(phi, abs=1e-9)
        assert math.degrees(d["pitch_rad"]) == pytest.approx(tht, abs=1e-9)
        assert math.degrees(d["heading_rad"]) == pytest.approx(psi, abs=1e-9)
        assert d["q0"] ** 2 + d["q1"] ** 2 + d["q2"] ** 2 + d["q3"] ** 2 == pytest.approx(1.0, abs=1e-14)
        assert [d["u_fps"], d["v_fps"], d["w_fps"]] == pytest.approx(list(uvw), abs=1e-9)
        assert [d["p_rad_sec"], d["q_rad_sec"], d["r_rad_sec"]] == pytest.approx(list(pqr), abs=1e-12)
        cf, sf, ct, st, cp, sp = (math.cos(math.radians(phi)), math.sin(math.radians(phi)), math.cos(math.radians(tht)),
                                  math.sin(math.radians(tht)), math.cos(math.radians(psi)), math.sin(math.radians(psi)))
        Tl2b = np.array([[ct * cp, ct * sp, -st],                                   # local (NED) -> body, 3-2-1 sequence
                         [sf * st * cp - cf * sp, sf * st * sp + cf * cp, sf * ct],
                         [cf * st * cp + sf * sp, cf * st * sp - sf * cp, cf * ct]])
        ned = Tl2b.T @ uvw
        assert [d["v_north_fps"], d["v_east_fps"], d["v_down_fps"]] == pytest.approx(list(ned), abs=1e-8)
        # inertial velocity: the earth's rotation adds Omega x r = Omega * r_xy towards the east
        a, b = 20925646.32546, 20855486.5951
        e2, la = 1 - (b / a) ** 2, math.radians(lat)
        N = a / math.sqrt(1 - e2 * math.sin(la) ** 2)
        r_xy = (N + d["geod_alt_ft"]) * math.cos(la)
        # east in NED is exactly the direction of Omega x r; north / down in the geodetic frame are perpendicular to it
        v_eci = math.sqrt(ned[0] ** 2 + (ned[1] + omega * r_xy) ** 2 + ned[2] ** 2)
        assert d["eci_velocity_mag_fps"] == pytest.approx(v_eci, rel=1e-10)


def test_air_data_and_weight_known_answer():
    """FGAuxiliary's air data from their definitions (J/models/FGAuxiliary.cpp:134-231: alpha = atan2(w, u), beta = atan2(v,
    sqrt(u^2 + w^2)), Vt = |uvw| without wind, qbar = rho Vt^2 / 2, Mach = Vt / a), the calibrated airspeed from the isentropic
    pitot relations (J/FGJSBBase.cpp:245-296: impact pressure at Mach and static pressure, read back at sea-level pressure),
    and the weight FGMassBalance sums (X/:69-90, 279-288: empty 17 400 lb + pilot 230 lb + 2 x 3000 lb of fuel)."""
    rng = np.random.default_rng(6)
    for _ in range(30):
        h = rng.uniform(3000, 45000)
        uvw = np.array([rng.uniform(300, 1700), rng.uniform(-60, 60), rng.uniform(-100, 100)])
        f = ofdm.OracleFdm()
        f.reset(h_sl_ft=h, u_fps=uvw[0], v_fps=uvw[1], w_fps=uvw[2])
        d, atm = f.snapshot_dict(), ofdm.atmosphere(h)
        vt = float(np.linalg.norm(uvw))
        assert d["alpha_rad"] == pytest.approx(math.atan2(uvw[2], uvw[0]), abs=1e-12)
        assert d["beta_rad"] == pytest.approx(math.atan2(uvw[1], math.hypot(uvw[0], uvw[2])), abs=1e-12)
        assert d["vt_fps"] == pytest.approx(vt, rel=1e-12)
        assert d["density"] == pytest.approx(atm["rho"], rel=1e-9) and d["pressure_psf"] == pytest.approx(atm["P"], rel=1e-9)
        assert d["qbar"] == pytest.approx(0.5 * atm["rho"] * vt * vt, rel=1e-9)
        mach = vt / atm["a"]
        assert d["mach"] == pytest.approx(mach, rel=1e-9)
        # calibrated airspeed: the same impact pressure measured at sea-level static pressure
        g, psl, asl = 1.4, 2116.228, ofdm.atmosphere(0.0)["a"]
        if mach < 1.0:
            qc = atm["P"] * ((1 + 0.2 * mach * mach) ** 3.5 - 1.0)
        else:      # Rayleigh pitot formula behind a normal shock
            qc = atm["P"] * (((g + 1) / 2 * mach * mach) ** (g / (g - 1)) * ((g + 1) / (2 * g * mach * mach - (g - 1))) ** (1 / (g - 1)) - 1.0)
        mc = math.sqrt(5.0 * ((qc / psl + 1.0) ** (1 / 3.5) - 1.0))
        if mc > 1.0:                                     # solve Rayleigh for the Mach number that gives qc at psl
            for _ in range(60):
                mc = 0.88128485 * math.sqrt((qc / psl + 1.0) * (1.0 - 1.0 / (7.0 * mc * mc)) ** 2.5)
        assert d["vc_fps"] == pytest.approx(mc * asl, rel=1e-6)
        assert d["mass_slugs"] == pytest.approx((17400.0 + 230.0 + 6000.0) / 32.174049, rel=1e-12)


def test_translational_integrators_known_answer():
    """FGPropagate integrates with the PREVIOUS frame's derivatives (J/models/FGPropagate.cpp:218-297): inertial position by
    Adams-Bashforth 3, inertial velocity by Adams-Bashforth 2 (:93-96, :371-470), the derivative history primed with the
    initial derivative (InitializeDerivatives).  Checked with the textbook coefficients over the first three frames."""
    dt = 1.0 / 60.0
    f = ofdm.OracleFdm()
    f.reset(lat_geod_deg=35.0, h_sl_ft=18000.0, psi_deg=77.0, u_fps=900.0, w_fps=30.0, theta_deg=4.0, phi_deg=20.0)
    f.set_controls(0.3, -0.4, 0.1, 0.8)
    S = [f.snapshot_dict()]
    for _ in range(3):
        f.run(1)
        S.append(f.snapshot_dict())
    r = [np.array([d["ri_x"], d["ri_y"], d["ri_z"]]) for d in S]
    v = [np.array([d["vi_x"], d["vi_y"], d["vi_z"]]) for d in S]
    a = [np.array([d["uvwidot_x"], d["uvwidot_y"], d["uvwidot_z"]]) for d in S]      # inertial acceleration a frame leaves behind
    # frame 1: the history holds the initial derivative three times -> both schemes reduce to rectangular Euler
    assert r[1] == pytest.approx(r[0] + dt * v[0], rel=1e-15, abs=1e-7)
    assert v[1] == pytest.approx(v[0] + dt * a[0], rel=1e-14, abs=1e-10)
    # frame 2: AB2 on (a1, a0); AB3 on (v1, v0, v0)
    assert v[2] == pytest.approx(v[1] + dt * (1.5 * a[1] - 0.5 * a[0]), rel=1e-14, abs=1e-10)
    assert r[2] == pytest.approx(r[1] + dt * (23.0 * v[1] - 16.0 * v[0] + 5.0 * v[0]) / 12.0, rel=1e-15, abs=1e-7)
    # frame 3: the full three-point formula
    assert v[3] == pytest.approx(v[2] + dt * (1.5 * a[2] - 0.5 * a[1]), rel=1e-14, abs=1e-10)
    assert r[3] == pytest.approx(r[2] + dt * (23.0 * v[2] - 16.0 * v[1] + 5.0 * v[0]) / 12.0, rel=1e-15, abs=1e-7)
    # inertial angular rate: rectangular Euler on the newest derivative (:93)
    w = [np.array([d["wi_x"], d["wi_y"], d["wi_z"]]) for d in S]
    wd = [np.array([d["pqridot_x"], d["pqridot_y"], d["pqridot_z"]]) for d in S]
    for k in range(3):
        assert w[k + 1] == pytest.approx(w[k] + dt * wd[k], rel=1e-14, abs=1e-16)
        assert sum(S[k + 1][q] ** 2 for q in ("q0", "q1", "q2", "q3")) == pytest.approx(1.0, abs=2.1e-10)   # FGQuaternion::Normalize leaves |q| alone within 1e-10 of 1
    # the earth rotation angle advances with the sidereal rate, the clock with dt
    assert S[3]["epa"] == pytest.approx(3 * dt * 7.292115e-5, rel=1e-12) and S[3]["sim_time"] == pytest.approx(3 * dt, abs=1e-15)


def test_newton_euler_known_answer():
    """FGAccelerations (J/models/FGAccelerations.cpp:109-207) from first principles, on the state a few frames of manoeuvring
    leave behind: body acceleration = F / m, inertial angular acceleration = J^-1 (M - w x J w), inertial acceleration =
    T_b->i F / m + gravity, with T from the textbook quaternion -> direction-cosine formula, the ECEF position from a rotation
    by the earth rotation angle, and gravity from the J2 field at that position."""
    rng = np.random.default_rng(8)
    for _ in range(10):
        f = ofdm.OracleFdm()
        f.reset(lat_geod_deg=rng.uniform(-60, 60), h_sl_ft=rng.uniform(8000, 35000), psi_deg=rng.uniform(0, 360),
                u_fps=rng.uniform(500, 1100), phi_deg=rng.uniform(-60, 60), theta_deg=rng.uniform(-20, 20))
        f.set_controls(*(list(rng.uniform(-1, 1, 3)) + [rng.uniform(0.2, 0.9)]))
        f.run(int(rng.integers(3, 40)))
        d, mp = f.snapshot_dict(), ofdm.mass_properties(f)
        F = np.array([d["fx"], d["fy"], d["fz"]]); M = np.array([d["mx"], d["my"], d["mz"]])
        w = np.array([d["wi_x"], d["wi_y"], d["wi_z"]])
        assert d["mass_slugs"] == pytest.approx(mp["mass"], rel=1e-14)
        assert [d["bodyaccel_x"], d["bodyaccel_y"], d["bodyaccel_z"]] == pytest.approx(list(F / mp["mass"]), rel=1e-12, abs=1e-12)
        assert mp["J"] @ mp["Jinv"] == pytest.approx(np.eye(3), abs=1e-12)
        wdot = mp["Jinv"] @ (M - np.cross(w, mp["J"] @ w))
        assert [d["pqridot_x"], d["pqridot_y"], d["pqridot_z"]] == pytest.approx(list(wdot), rel=1e-11, abs=1e-13)
        q0, q1, q2, q3 = d["q0"], d["q1"], d["q2"], d["q3"]
        Ti2b = np.array([[q0 * q0 + q1 * q1 - q2 * q2 - q3 * q3, 2 * (q1 * q2 + q0 * q3), 2 * (q1 * q3 - q0 * q2)],
                         [2 * (q1 * q2 - q0 * q3), q0 * q0 - q1 * q1 + q2 * q2 - q3 * q3, 2 * (q2 * q3 + q0 * q1)],
                         [2 * (q1 * q3 + q0 * q2), 2 * (q2 * q3 - q0 * q1), q0 * q0 - q1 * q1 - q2 * q2 + q3 * q3]])
        ce, se = math.cos(d["epa"]), math.sin(d["epa"])
        Ti2ec = np.array([[ce, se, 0.0], [-se, ce, 0.0], [0.0, 0.0, 1.0]])
        ri = np.array([d["ri_x"], d["ri_y"], d["ri_z"]])
        g_i = Ti2ec.T @ ofdm.gravity(*(Ti2ec @ ri))
        a_i = Ti2b.T @ (F / mp["mass"]) + g_i
        assert [d["uvwidot_x"], d["uvwidot_y"], d["uvwidot_z"]] == pytest.approx(list(a_i), rel=1e-9, abs=1e-9)


def test_pilot_load_factor_uses_last_frames_accelerations():
    """FGAuxiliary runs before FGAccelerations in a frame (J/FGFDMExec.cpp:222-236), so the pilot-station load factor the FCS and
    the Overload termination see is built from the PREVIOUS frame's linear and angular accelerations and this frame's rates:
    n_pilot = (a_body + wdot x r + w x (w x r)) / g0 with r from the centre of gravity to the eye point (X/:50-54),
    J/models/FGAuxiliary.cpp:205-222."""
    g0 = 9.80665 / 0.3048
    f = ofdm.OracleFdm()
    f.reset(h_sl_ft=15000.0, u_fps=950.0, phi_deg=35.0)
    f.set_controls(0.6, -0.7, 0.2, 0.7)
    f.run(5)
    for _ in range(20):
        prev = f.snapshot_dict()
        f.run(1)
        d, mp = f.snapshot_dict(), ofdm.mass_properties(f)
        cg = mp["cg"]
        r = np.array([(cg[0] - (-336.2)) / 12.0, (0.0 - cg[1]) / 12.0, (cg[2] - 29.5) / 12.0])     # structural (in) -> body (ft)
        a = np.array([prev["bodyaccel_x"], prev["bodyaccel_y"], prev["bodyaccel_z"]])
        wd = np.array([prev["pqridot_x"], prev["pqridot_y"], prev["pqridot_z"]])
        w = np.array([d["wi_x"], d["wi_y"], d["wi_z"]])
        n = (a + np.cross(wd, r) + np.cross(w, np.cross(w, r))) / g0
        assert [d["n_pilot_x"], d["n_pilot_y"], d["n_pilot_z"]] == pytest.approx(list(n), rel=1e-10, abs=1e-12)
    assert abs(d["n_pilot_z"]) > 1.5          # the manoeuvre loads the aircraft: the check is not about a quiescent state


def test_turbine_spools_up_with_the_published_rate():
    """reference envs/JSBSim/data/tests/TestTurbine.py:36-41,99-105: N2 seeks IdleN2 + throttle*N2_factor at
    delay/(1+3(1-n)^3+(1-sigma)) per second with delay = 90/(BPR+3)."""
    f = ofdm.OracleFdm()
    f.reset(h_sl_ft=20000.0, u_fps=800.0)
    f.set_controls(0.0, 0.0, 0.0, 0.45)   # throttle-pos = 2*cmd = 0.9 (f16.xml:869-873) -> no afterburner
    d0 = f.snapshot_dict()
    f.run(1)
    d1 = f.snapshot_dict()
    sigma = ofdm.atmosphere(d1["h_sl_ft"])["rho"] / ofdm.atmosphere(0.0)["rho"]
    n = min(1.0, d0["N2norm"] + 0.1)
    rate = (90.0 / (0.4 + 3.0)) / (1 + 3 * (1 - n) ** 3 + (1 - sigma))
    assert d1["N2"] - d0["N2"] == pytest.approx(rate / 60.0, rel=1e-6)


def test_turbine_spool_sequence_known_answer():
    """TestTurbine.testSpoolUp / runScript (envs/JSBSim/data/tests/TestTurbine.py:25-69,99-105, run there on the F-16 itself):
    every frame N1 and N2 `seek` their targets idle + ThrottlePos * (max - idle) with the default spool-up rate
    delay / (1 + 3 (1 - n)^3 + (1 - sigma)), n = min(1, N2norm + 0.1), delay = 90 dt / (BPR + 3); N1 spools DOWN 2.4 times and
    N2 3.0 times faster.  Throttle up to the stop, then back to idle, as the JSBSim test does."""
    idle_n1, max_n1, idle_n2, max_n2, bpr = 40.0, 100.0, 53.0, 100.0, 0.4     # data/engine/F100-PW-229.xml:17-23
    dt = 1.0 / 60.0
    delay = 90.0 * dt / (bpr + 3.0)
    rho0 = ofdm.atmosphere(0.0)["rho"]

    def seek(x, target, accel, decel):                                      # TestTurbine.py:25-33
        if x < target:
            return min(x + accel, target)
        if x > target:
            return max(x - decel, target)
        return x

    f = ofdm.OracleFdm()
    f.reset(h_sl_ft=15000.0, u_fps=700.0)
    went_up = went_down = 0
    for cmd, frames in ((0.5, 240), (0.0, 420)):      # throttle-pos-norm = 2 cmd (f16.xml:869-873): 1.0, then 0.0
        f.set_controls(0.0, 0.0, 0.0, cmd)
        pos = min(1.0, 2.0 * cmd)
        for _ in range(frames):
            d0 = f.snapshot_dict()
            f.run(1)
            d1 = f.snapshot_dict()
            sigma = d1["density"] / rho0
            n = min(1.0, d0["N2norm"] + 0.1)
            up = delay / (1 + 3 * (1 - n) ** 3 + (1 - sigma))
            n1 = seek(d0["N1"], idle_n1 + pos * (max_n1 - idle_n1), up, 2.4 * up)
            n2 = seek(d0["N2"], idle_n2 + pos * (max_n2 - idle_n2), up, 3.0 * up)
            assert d1["N1"] == pytest.approx(n1, abs=1e-7)          # assertAlmostEqual: 7 places
            assert d1["N2"] == pytest.approx(n2, abs=1e-7)
            went_up += d1["N2"] > d0["N2"]
            went_down += d1["N1"] < d0["N1"]
    assert went_up > 50 and went_down > 50                          # both branches of both seeks were exercised
    assert d1["N2"] == pytest.approx(idle_n2, abs=1e-9) and d1["N1"] == pytest.approx(idle_n1, abs=1e-9)


def test_level_flight_is_quiescent_and_time_advances():
    f = ofdm.OracleFdm()
    f.reset()
    f.set_controls(0.0, 0.0, 0.0, 0.5)
    f.run(60)
    d = f.snapshot_dict()
    assert d["sim_time"] == pytest.approx(1.0, abs=1e-12)
    assert abs(d["roll_rad"]) < 1e-3 and abs(d["heading_rad"]) < 1e-3 or abs(d["heading_rad"] - 2 * math.pi) < 1e-3
    assert 19000 < d["h_sl_ft"] < 21000
    assert abs(d["n_pilot_y"]) < 1e-3
    # quaternion stays normalised (FGPropagate::Integrate normalises every frame)
    q = [d[k] for k in ("q0", "q1", "q2", "q3")]
    assert sum(x * x for x in q) == pytest.approx(1.0, abs=1e-12)


def test_fuel_burn_matches_fuel_flow():
    f = ofdm.OracleFdm()
    f.reset()
    f.set_controls(0.0, 0.0, 0.0, 0.9)
    f.run(600)
    d = f.snapshot_dict()
    burnt = 6000.0 - (d["tank0"] + d["tank1"] + d["tank2"] + d["tank3"])
    assert burnt > 0
    assert d["tank0"] == pytest.approx(d["tank1"], rel=1e-12)  # equal split (FGPropulsion.cpp:164-260)
    assert burnt < d["FuelFlow_pph"] / 3600.0 * 10.0 * 1.5


def test_fuel_is_drawn_equally_from_the_tanks_every_frame():
    """FGPropulsion::ConsumeFuel (J/models/FGPropulsion.cpp:164-260) after FGTurbine::CalcFuelNeed (propulsion/FGTurbine.cpp:381-387):
    each frame the engine's fuel flow of THAT frame, pph / 3600 * dt, leaves the two tanks that hold fuel in equal parts --
    afterburner included (throttle command 0.9 -> throttle position 1.8: augmentation)."""
    dt = 1.0 / 60.0
    f = ofdm.OracleFdm()
    f.reset()
    f.set_controls(0.0, 0.0, 0.0, 0.9)
    f.run(150)
    for _ in range(40):
        a = f.snapshot_dict()
        f.run(1)
        b = f.snapshot_dict()
        need = b["FuelFlow_pph"] / 3600.0 * dt
        assert a["tank0"] - b["tank0"] == pytest.approx(need / 2, rel=1e-9)
        assert a["tank1"] - b["tank1"] == pytest.approx(need / 2, rel=1e-9)
        assert b["tank2"] == 0.0 and b["tank3"] == 0.0
    assert int(b["engflags"]) & 2                # the augmentation flag is up: the afterburner branch of FGTurbine::Run was timed


def test_same_inputs_same_trajectory():
    a, b = ofdm.OracleFdm(), ofdm.OracleFdm()
    rng = np.random.default_rng(1)
    for f in (a, b):
        f.reset(psi_deg=30.0)
    for _ in range(20):
        u = [rng.integers(0, 41) / 20 - 1, rng.integers(0, 41) / 20 - 1, rng.integers(0, 41) / 20 - 1, rng.integers(0, 30) / 58 + 0.4]
        for f in (a, b):
            f.set_controls(*u)
            f.run(12)
    assert np.array_equal(a.snapshot(), b.snapshot())


# ---------------------------------------------------------------------------------------------- JSBSim's own known answers
# Ported from the reference's copy of JSBSim's test-suite (envs/JSBSim/data/tests/*.py).  Those tests drive a full
# jsbsim.FGFDMExec with other aircraft (c172r flaps, the "tripod" test systems); what they PIN is component semantics
# with closed-form answers.  The oracle runs the same component code for the F-16's FCS (kinemat_step / pid_step are
# the functions fcs_run calls), so the known answers pin the restatement of FGKinemat / FGPID / FGInertial / FGMassBalance.

def test_kinematic_timing_known_answer():
    """TestKinematic.testKinematicTiming (envs/JSBSim/data/tests/TestKinematic.py:27-83): the c172r flap kinematic --
    detents 0 / 10 / 20 / 30 deg with traverse times 0 / 2 / 1 / 1 s, command scaled by the last detent and clamped."""
    det, tim, dt = [0.0, 10.0, 20.0, 30.0], [0.0, 2.0, 1.0, 1.0], 1.0 / 120.0
    pos, n = 0.0, 0

    def run_until(t_end, cmd, expect):
        nonlocal pos, n
        while n * dt < t_end - 1e-12:
            assert pos == pytest.approx(expect(n * dt), abs=1e-7), n * dt
            pos = ofdm.kinemat(det, tim, cmd, pos, dt)
            n += 1
    run_until(2.0, 1.5, lambda t: 5.0 * t)                 # command 1.5 is clamped to the last detent
    run_until(4.0, 1.5, lambda t: 10.0 * (t - 1.0))
    run_until(5.0, 1.5, lambda t: 30.0)
    run_until(7.0, 0.25, lambda t: 30.0 - 10.0 * (t - 5.0))   # up again, interrupted at 7.5 deg
    run_until(7.5, 0.25, lambda t: 10.0 - 5.0 * (t - 7.0))
    run_until(8.0, 0.25, lambda t: 7.5)
    run_until(9.5, -1.0, lambda t: 10.0 - 5.0 * (t - 7.5))    # command -1 is clamped to the first detent
    run_until(10.0, -1.0, lambda t: 0.0)


def test_kinematic_noscale_known_answer():
    """TestKinematic.testKinematicNoScale (:123-140): with <noscale/> the command is the position itself; 12 deg is reached
    after 2 s + 0.2 s."""
    det, tim, dt = [0.0, 10.0, 20.0, 30.0], [0.0, 2.0, 1.0, 1.0], 1.0 / 120.0
    pos = 0.0
    for _ in range(int(round(2.2 / dt)) + 1):
        pos = ofdm.kinemat(det, tim, 12.0, pos, dt, noscale=True)
    assert pos == pytest.approx(12.0, abs=1e-7)
    # a zero traverse time moves instantly (FGKinemat.cpp:131-134)
    assert ofdm.kinemat([0.0, 1.0], [0.0, 0.0], 0.7, 0.0, dt) == 0.7


def test_pid_integrators_known_answer():
    """TestIntegrators.test_integrators (envs/JSBSim/data/tests/TestIntegrators.py:38-105, integrators.xml): the four
    integration schemes of FGPID on sin(8 pi t) at dt = 5 ms against the analytic integral, a positive trigger freezes
    the integrator, a negative one resets it, zero restarts it."""
    dt, k = 0.005, 8 * math.pi
    pids = {name: ofdm.Pid(ki=1.0, int_type=it, dt=dt) for name, it in (("rect", 1), ("trap", 2), ("ab2", 3), ("ab3", 4))}
    out = {name: 0.0 for name in pids}
    t = 0.0
    for i in range(100):
        x = math.sin(k * t)
        if i > 1:                                   # AB3 is not initialised before the third frame
            assert out["ab3"] == pytest.approx((1.0 - math.cos(k * t)) / k, abs=1e-4)
        for name, p in pids.items():
            out[name] = p.run(x, 0.0)
        t += dt
    frozen = dict(out)
    for _ in range(49):
        for name, p in pids.items():
            out[name] = p.run(math.sin(k * t), 1.0)
        assert out == pytest.approx(frozen, abs=1e-15)
    for name, p in pids.items():
        out[name] = p.run(0.0, -1.0)
        assert out[name] == 0.0
    t0 = 0.0
    for i in range(50):
        x = math.sin(k * t0)
        if i > 1:
            assert out["ab3"] == pytest.approx((1.0 - math.cos(k * t0)) / k, abs=1e-4)
        for name, p in pids.items():
            out[name] = p.run(x, 0.0)
        t0 += dt


def test_pid_gains_known_answer():
    """TestIntegrators.test_pid (:107-145): kp alone = kp sin, ki alone (default AB2) = ki (1 - cos) / k, kd alone ~ kd k cos,
    and a pid with all three negated is minus their sum."""
    dt, k, kp, ki, kd = 0.005, 2 * math.pi, 2.0, 0.5, -1.5
    P, I, D = ofdm.Pid(kp=kp, dt=dt), ofdm.Pid(ki=ki, int_type=3, dt=dt), ofdm.Pid(kd=kd, dt=dt)
    N = ofdm.Pid(kp=-kp, ki=-ki, kd=-kd, int_type=3, dt=dt)
    oi = 0.0
    for i in range(100):
        t = i * dt
        x = math.sin(k * t)
        assert oi == pytest.approx(ki * (1.0 - math.cos(k * t)) / k, abs=1e-4)
        op, oi, od, on = P.run(x), I.run(x), D.run(x), N.run(x)
        assert -on == pytest.approx(op + oi + od, abs=1e-12)
        assert op == pytest.approx(kp * math.sin(k * t), abs=1e-12)
        if i > 1:
            assert od == pytest.approx(kd * k * math.cos(k * t), abs=0.15)


def test_gravity_known_answer():
    """TestPlanet.py checks gravity-ft_sec2 and the terrain radius at the equator and the pole of a planet file; for the
    earth model the reference loads (FGInertial.cpp:55-61: GM, J2, a, b) the J2 field gives the WGS-84 gravitation at the
    surface: 9.8142 m/s^2 at the equator (normal gravity 9.78033 + centrifugal 0.03392), 9.8322 at the pole."""
    a, b = 20925646.32546, 20855486.5951
    ge, gp = ofdm.gravity(a, 0.0, 0.0), ofdm.gravity(0.0, 0.0, b)
    assert ge[1] == 0.0 and ge[2] == 0.0 and gp[0] == 0.0 and gp[1] == 0.0
    assert -ge[0] * 0.3048 == pytest.approx(9.7803253 + 0.00007292115 ** 2 * 6378137.0, abs=2e-4)
    assert -gp[2] * 0.3048 == pytest.approx(9.8321849, abs=2e-4)
    # sea-level (terrain) radius = a at the equator, b at the pole (FGLocation::GetSeaLevelRadius)
    assert ofdm.geodetic(a, 0.0, 0.0)["slr"] == pytest.approx(a, rel=1e-15)
    assert ofdm.geodetic(0.0, 1e-3, b)["slr"] == pytest.approx(b, rel=1e-12)


def test_geocentric_vs_geodetic_latitude_known_answer():
    """TestInitialConditions.py compares ic/lat-gc with ic/lat-geod: tan(lat_gc) = ((1 - e^2) N + h) / (N + h) tan(lat_geod)."""
    a, b = 20925646.32546, 20855486.5951
    e2 = 1 - (b / a) ** 2
    for lat_deg, h in [(10.0, 0.0), (45.0, 20000.0), (60.0, 30000.0), (-75.0, 5000.0)]:
        lat = math.radians(lat_deg)
        N = a / math.sqrt(1 - e2 * math.sin(lat) ** 2)
        x, z = (N + h) * math.cos(lat), ((1 - e2) * N + h) * math.sin(lat)
        g = ofdm.geodetic(x, 0.0, z)
        assert math.tan(g["lat_gc"]) == pytest.approx(((1 - e2) * N + h) / (N + h) * math.tan(lat), rel=1e-12)
        assert g["radius"] == pytest.approx(math.hypot(x, z), rel=1e-15)


def test_inertia_matrix_known_answer():
    """TestPointMassInertia.testInertiaMatrix (envs/JSBSim/data/tests/TestPointMassInertia.py:96-128, run there on f16_test):
    Jinv is the inverse of J; with the tanks empty the weight is empty weight + point masses; the point masses enter
    through the parallel-axis tensor m (|r|^2 I - r r^T) about the cg (GetPointmassInertia, pinned here independently)."""
    f = ofdm.OracleFdm()
    f.reset()
    m = ofdm.mass_properties(f)
    np.testing.assert_allclose(m["J"] @ m["Jinv"], np.eye(3), atol=1e-12)
    assert m["weight"] == pytest.approx(17400.0 + 6000.0 + 230.0)          # empty + 2 x 3000 lb fuel + the 230 lb pilot
    ofdm.set_tanks(f, [0.0, 0.0, 0.0, 0.0])
    f.run(1)
    m0 = ofdm.mass_properties(f)
    assert m0["weight"] == pytest.approx(17400.0 + 230.0)
    # independent build-up about the new cg (structural frame inches -> body frame feet: x and z flip sign)
    lb2slug = 1.0 / 32.174049
    parts = [(17400.0, np.array([-193.0, 0.0, -5.1])), (230.0, np.array([-336.2, 0.0, 0.0]))]
    cg = sum(w * r for w, r in parts) / sum(w for w, _ in parts)
    np.testing.assert_allclose(m0["cg"], cg, rtol=1e-13)
    J = np.array([[9496.0, 0.0, -982.0], [0.0, 55814.0, 0.0], [-982.0, 0.0, 63100.0]])    # f16.xml ixx iyy izz ixz (negated cross products)
    for w, r in parts:
        d = (r - cg) / 12.0 * np.array([-1.0, 1.0, -1.0])
        J = J + w * lb2slug * (np.dot(d, d) * np.eye(3) - np.outer(d, d))
    np.testing.assert_allclose(m0["J"], J, rtol=1e-12, atol=1e-9)
    np.testing.assert_allclose(m0["J"] @ m0["Jinv"], np.eye(3), atol=1e-12)


# ---------------------------------------------------------------------------------------------- JSBSim's own golden numbers
# The two tables below are OUTPUTS OF THE REAL JSBSim, committed in the reference tree as the reference data of its own tests:
# (geometric altitude ft, atmosphere/delta-T deg R, expected altitude ft).  They pin the layer table, lapse rates, pressure
# break points (pow / exp per layer), the geopotential conversion and both inverse functions of the restated atmosphere --
# including the density altitude the F-16's engine tables are indexed with -- through all eight layers.
JSBSIM_DENSITY_ALTITUDE = [   # envs/JSBSim/data/tests/TestDensityAltitude.py:29-72
    (0, 0, 0), (0, -27, -1838.3210293), (0, 27, 1724.0715454),
    (10000, 0, 10000), (10000, -27, 8842.6417730), (10000, 27, 11117.881412),
    (20000, 0, 20000), (20000, -27, 19524.3027252), (20000, 27, 20511.1447732),
    (30000, 0, 30000), (30000, -27, 30206.6618955), (30000, 27, 29903.8616766),
    (40000, 0, 40000), (40000, -27, 40795.49296642545), (40000, 27, 39370.88017359472),
    (50000, 0, 50000), (50000, -27, 51540.55687445524), (50000, 27, 48722.498495051164),
    (60000, 0, 60000), (60000, -27, 62286.38628793139), (60000, 27, 58073.53700852329),
    (100000, 0, 100000), (100000, -27, 105340.83866126923), (100000, 27, 95441.10933931815),
    (150000, 0, 150000), (150000, -27, 159983.12837740956), (150000, 27, 141847.12260237316),
    (160000, 0, 160000), (160000, -27, 171305.317547756), (160000, 27, 150762.4482763878),
    (220000, 0, 220000), (220000, -27, 233395.46205079104), (220000, 27, 207735.4268030979),
    (260000, 0, 260000), (260000, -27, 274351.9265767301), (260000, 27, 246964.3481013492),
    (290000, 0, 290000), (290000, -27, 305321.815863847), (290000, 27, 276290.37419984984),
    (320000, 0, 320000), (320000, -27, 337990.28144264355), (320000, 27, 304417.9280936986),
]
JSBSIM_PRESSURE_ALTITUDE = [  # envs/JSBSim/data/tests/TestPressureAltitude.py:29-72
    (0, 0, 0), (0, -27, 0), (0, 27, 0),
    (10000, 0, 10000), (10000, -27, 10549.426597202142), (10000, 27, 9504.969939165301),
    (20000, 0, 20000), (20000, -27, 21099.40877940678), (20000, 27, 19009.488882465),
    (30000, 0, 30000), (30000, -27, 31649.946590500946), (30000, 27, 28513.556862000503),
    (40000, 0, 40000), (40000, -27, 42294.25242340247), (40000, 27, 37972.879013500584),
    (50000, 0, 50000), (50000, -27, 53040.858132036126), (50000, 27, 47323.24573196882),
    (60000, 0, 60000), (60000, -27, 63788.23024872676), (60000, 27, 56673.032160129085),
    (100000, 0, 100000), (100000, -27, 107018.51146890492), (100000, 27, 93910.6118895332),
    (150000, 0, 150000), (150000, -27, 161956.60354430682), (150000, 27, 139810.93668842476),
    (160000, 0, 160000), (160000, -27, 172582.32995327076), (160000, 27, 148992.4108097521),
    (220000, 0, 220000), (220000, -27, 233772.88181515134), (220000, 27, 207274.89794422916),
    (260000, 0, 260000), (260000, -27, 275000.20893894637), (260000, 27, 246263.63421221747),
    (290000, 0, 290000), (290000, -27, 306867.8082342206), (290000, 27, 275185.69941262825),
    (320000, 0, 320000), (320000, -27, 339541.05112835445), (320000, 27, 302991.642663158),
]


@pytest.mark.parametrize("h_ft,delta_T,expected", JSBSIM_DENSITY_ALTITUDE)
def test_density_altitude_matches_jsbsim_reference_data(h_ft, delta_T, expected):
    """TestDensityAltitude.test_densityaltitude (:75-97): `atmosphere/density-altitude` after ic/h-sl-ft + atmosphere/delta-T,
    asserted there to 7 places relative; then the standard-day density AT the density altitude equals the density it came from."""
    got = ofdm.atmosphere(float(h_ft), float(delta_T))
    if expected < 1e-9:
        assert got["density_altitude"] == pytest.approx(expected, abs=5e-8)
    else:
        assert got["density_altitude"] / expected == pytest.approx(1.0, abs=5e-8)
    assert ofdm.atmosphere(got["density_altitude"])["rho"] == pytest.approx(got["rho"], abs=5e-8, rel=1e-12)


@pytest.mark.parametrize("h_ft,delta_T,expected", JSBSIM_PRESSURE_ALTITUDE)
def test_pressure_altitude_matches_jsbsim_reference_data(h_ft, delta_T, expected):
    """TestPressureAltitude.test_pressurealtitude (:75-92): `atmosphere/pressure-altitude`, asserted there with delta = 1e-7 ft
    (1e-12 relative at 100 000 ft); then the standard-day pressure AT the pressure altitude equals the pressure it came from."""
    got = ofdm.atmosphere(float(h_ft), float(delta_T))
    assert got["pressure_altitude"] == pytest.approx(expected, abs=1e-7)
    assert ofdm.atmosphere(got["pressure_altitude"])["P"] == pytest.approx(got["P"], abs=5e-8, rel=1e-12)
