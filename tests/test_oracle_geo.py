"""Known-answer pin of the pymap3d restatement in oracle/env_oracle.py (LLA2NEU / NEU2alt, reference envs/JSBSim/utils/utils.py:30-55).

pymap3d is a third-party dependency of the reference that is absent from /root/reference and from this image (pinned by the
reference's requirements as `pymap3d`; algorithm: WGS-84 closed-form geodetic -> ECEF, ECEF -> ENU rotation, You (2000)
ECEF -> geodetic).  The vectors below are the fixed points of pymap3d's OWN published test-suite (pymap3d/tests:
`lla0 = (42, -82, 200)`, `xyz0`, `aer0 = (33, 70, 1000)`, `enu0`, `lla1`), i.e. numbers produced by the real library, not by
this repository; they are reproduced here to the digits that suite publishes.  Until now the golden episodes had the
restatement on both sides (tools/make_golden.py shims pymap3d with it)."""
import math

import numpy as np
import pytest

from oracle import env_oracle as eo

LLA0 = (42.0, -82.0, 200.0)                                              # lat, lon, alt
XYZ0 = (660675.2518247, -4700948.68316, 4245737.66222)                   # geodetic2ecef(*LLA0)
AER0 = (33.0, 70.0, 1000.0)                                              # az, el, slant range from LLA0
ENU0 = (186.277521, 286.84222, 939.69262)                                # aer2enu(*AER0)
LLA1 = (42.002581974253744, -81.99775196006746, 1139.7018)               # aer2geodetic(*AER0, *LLA0) (alt to 1e-4 m)


def test_geodetic2ecef_matches_pymap3d_vector():
    x, y, z = eo.geodetic2ecef(LLA0[0], LLA0[1], LLA0[2])
    assert x == pytest.approx(XYZ0[0], abs=1e-6)
    assert y == pytest.approx(XYZ0[1], abs=1e-5)
    assert z == pytest.approx(XYZ0[2], abs=1e-5)


def test_aer2enu_is_what_the_vector_says():
    az, el, r = math.radians(AER0[0]), math.radians(AER0[1]), AER0[2]
    enu = (r * math.cos(el) * math.sin(az), r * math.cos(el) * math.cos(az), r * math.sin(el))
    assert enu == pytest.approx(ENU0, abs=1e-5)


def test_lla2neu_matches_pymap3d_vector():
    # geodetic2ned(lla1, lla0) = (north, east, -up) of ENU0; the published lat / lon carry ~1e-14 deg, the altitude 1e-4 m
    neu = eo.LLA2NEU(LLA1[1], LLA1[0], LLA1[2], LLA0[1], LLA0[0], LLA0[2])
    assert neu[0] == pytest.approx(ENU0[1], abs=1e-4)
    assert neu[1] == pytest.approx(ENU0[0], abs=1e-4)
    assert neu[2] == pytest.approx(ENU0[2], abs=1e-4)


def test_neu2alt_matches_pymap3d_vector():
    # ned2geodetic(north, east, down, lla0) -> altitude of lla1 (the only component the hot path consumes: missile air density)
    alt = eo.NEU2alt(ENU0[1], ENU0[0], ENU0[2], LLA0[1], LLA0[0], LLA0[2])
    assert alt == pytest.approx(LLA1[2], abs=1e-3)


def test_round_trip_at_the_battle_field_origin():
    # the reference's battlefield centre (envs/JSBSim/configs: 120 E, 60 N, 0 m): LLA -> NEU -> altitude closes to 1e-6 m
    rng = np.random.default_rng(0)
    for _ in range(50):
        lon, lat, alt = 120 + rng.uniform(-1, 1), 60 + rng.uniform(-1, 1), rng.uniform(0, 12000)
        n, e, u = eo.LLA2NEU(lon, lat, alt, 120.0, 60.0, 0.0)
        assert eo.NEU2alt(n, e, u, 120.0, 60.0, 0.0) == pytest.approx(alt, abs=1e-6)
