"""csrc/fmath.cuh on the HOST: the guard-free Newton sequences and the polynomial kernels the FDM frame and the missile
path use instead of `/`, sqrt() and libdevice, checked against libm / long double.

The header compiles for the host with its MUFU seeds emulated at a relative error of 2^-12 -- far worse than the
hardware's -- so this pins the algebra (iteration counts, residual corrections, reduction constants, coefficient tables);
the device build of the same source is covered by the GPU parity tests.  Bounds are in ulp of the exact result."""
import json
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def report(tmp_path_factory):
    exe = tmp_path_factory.mktemp("fmath") / "fmath_host"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-I", str(ROOT / "aircombat_selfplay_b200" / "csrc"),
                           "-x", "c++", str(ROOT / "tests" / "native" / "fmath_host.cpp"), "-o", str(exe)])
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.strip().splitlines()
    d = {}
    for line in out:
        d.update(json.loads(line))
    return d


def test_division_and_roots_are_correctly_rounded_from_a_poor_seed(report):
    assert report["div"] <= 0.5001 and report["rcp"] <= 0.5001 and report["sqrt"] <= 0.5001
    assert report["rsqrt"] <= 1.0


def test_trigonometric_kernels(report):
    assert report["sin"] <= 1.0 and report["cos"] <= 1.5                      # |x| <= pi/4
    assert report["sincos_s"] <= 2.0 and report["sincos_c"] <= 2.0            # |x| <= 2e8, Cody-Waite reduction
    assert report["abs_s"] <= 2.5e-16 and report["abs_c"] <= 2.5e-16
    assert report["abs_sin"] <= 4e-16                                          # odd polynomial on |r| <= pi/2
    assert report["angle"] <= 5.0                                              # atan2 from (sin, cos): inputs carry ~3 ulp
    assert report["atan2"] <= 4.0 and report["abs_atan2"] <= 6e-16 and "atan2_special" not in report   # all quadrants, axes


def test_exp_and_the_isa_power_law(report):
    assert report["exp"] <= 2.5
    assert report["pow_rel"] <= 4e-15      # |y ln(ratio)| <= 5: the reference's exp(y log x) is good to ~1e-15 here


def test_special_operands(report):
    # sqrt0(0) = 0, 0 / b = 0, angle(0, 1) = 0, ratio 1 -> 1, fallback branches (ratio far from 1, |x| > pi/4, |angle| > 45 deg)
    assert report["bad"] == 0


def test_reward_and_geometry_functions(report):
    # acos keeps RELATIVE accuracy towards small angles (AO / TA start near 1e-7 in the reference's own scenarios)
    assert report["acos"] <= 6.0 and report["abs_acos"] <= 1e-15
    assert report["log"] <= 4.0
    assert report["abs_tanh"] <= 5e-16 and report["abs_atanh"] <= 3e-15 and report["abs_cos"] <= 3e-16
    assert report["bad2"] == 0      # acos(+-1), log(0) = -inf, atanh(+-1) = +-inf, exp clamps instead of underflowing
