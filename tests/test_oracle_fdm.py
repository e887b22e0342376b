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
