#!/usr/bin/env python
"""Records tests/golden/fdm_jsbsim_*.npz: trajectories of the REAL JSBSim F-16 through the reference's own
``AircraftSimulator`` -- the pin the FDM oracle is missing ("parity unpinned").

It cannot run in the build container or on the GPU box (``import jsbsim`` fails on both, the wheel is not in the
wheelhouse and the vendored C++ has no headers).  Run it on ANY machine with the reference checkout and
``pip install jsbsim==1.1.6 gymnasium pymap3d`` (the reference's pins):

    python tools/make_fdm_golden.py --reference /path/to/aircombat-selfplay

and commit the files it writes; ``tests/test_fdm_jsbsim_golden.py`` switches itself on when they exist and compares the
CPU oracle (and, with a GPU, the CUDA FDM) with them frame by frame.

What is recorded.  For every case: the twelve initial-condition values (the order of acs_fdm_reset), the control schedule
(one row of normalised commands per interaction step of 12 frames) and, after reset and after 1, 2, 12, 120 and 1200
frames, every property the env layer reads (reference envs/JSBSim/core/simulatior.py:238-257, catalog.py) plus the
engine / FCS / mass properties the restatement carries as state, read through ``jsbsim_exec.get_property_value``.  The
jsbsim version string and the sha256 of the f16.xml / F100-PW-229.xml the run loaded are stored with them.
"""
from __future__ import annotations

import argparse
import hashlib
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

# oracle snapshot name -> JSBSim property (what AircraftSimulator._update_properties and the termination conditions read)
PROPS = {
    "lon_deg": "position/long-gc-deg", "lat_geod_deg": "position/lat-geod-deg", "h_sl_ft": "position/h-sl-ft",
    "roll_rad": "attitude/roll-rad", "pitch_rad": "attitude/pitch-rad", "heading_rad": "attitude/heading-true-rad",
    "v_north_fps": "velocities/v-north-fps", "v_east_fps": "velocities/v-east-fps", "v_down_fps": "velocities/v-down-fps",
    "u_fps": "velocities/u-fps", "v_fps": "velocities/v-fps", "w_fps": "velocities/w-fps", "vc_fps": "velocities/vc-fps",
    "n_pilot_x": "accelerations/n-pilot-x-norm", "n_pilot_y": "accelerations/n-pilot-y-norm", "n_pilot_z": "accelerations/n-pilot-z-norm",
    "p_rad_sec": "velocities/p-rad_sec", "q_rad_sec": "velocities/q-rad_sec", "r_rad_sec": "velocities/r-rad_sec",
    "eci_velocity_mag_fps": "velocities/eci-velocity-mag-fps", "sim_time": "simulation/sim-time-sec",
    "alpha_rad": "aero/alpha-rad", "beta_rad": "aero/beta-rad", "mach": "velocities/mach", "qbar": "aero/qbar-psf",
    "vt_fps": "velocities/vt-fps", "thrust_lbs": "propulsion/engine/thrust-lbs", "N1": "propulsion/engine/n1", "N2": "propulsion/engine/n2",
    "FuelFlow_pph": "propulsion/engine/fuel-flow-rate-pps", "tank0": "propulsion/tank[0]/contents-lbs", "tank1": "propulsion/tank[1]/contents-lbs",
    "mass_slugs": "inertia/mass-slugs", "cg_x": "inertia/cg-x-in", "cg_z": "inertia/cg-z-in", "geod_alt_ft": "position/geod-alt-ft",
    "temperature_R": "atmosphere/T-R", "pressure_psf": "atmosphere/P-psf", "density": "atmosphere/rho-slugs_ft3",
    "density_altitude": "atmosphere/density-altitude",
    # flight-control outputs (properties of the modified f16.xml the reference ships)
    "fcs/elevator-pos-rad": "fcs/elevator-pos-rad", "fcs/left-aileron-pos-rad": "fcs/left-aileron-pos-rad",
    "fcs/rudder-pos-rad": "fcs/rudder-pos-rad", "fcs/throttle-pos-norm": "fcs/throttle-pos-norm", "fcs/lef-pos-rad": "fcs/lef-pos-rad",
    "fcs/tef-pos-rad": "fcs/tef-pos-rad", "fcs/speedbrake-pos-rad": "fcs/speedbrake-pos-rad",
}
CHECKPOINTS = (0, 1, 2, 12, 120, 1200)
IC_KEYS = ["ic_long_gc_deg", "ic_lat_geod_deg", "ic_h_sl_ft", "ic_psi_true_deg", "ic_u_fps", "ic_v_fps", "ic_w_fps", "ic_p_rad_sec",
           "ic_q_rad_sec", "ic_r_rad_sec", "ic_phi_deg", "ic_theta_deg"]


def cases(rng):
    """name -> (ic row [12], controls [n_steps, 4]); the ICs / control distribution of tests/fdm_parity.py."""
    from tests.fdm_parity import random_controls, random_ics
    out = {"level_idle": (np.array([120.0, 60.0, 20000.0, 0.0, 800.0] + [0.0] * 7), np.tile([0.0, 0.0, 0.0, 0.4], (100, 1))),
           "level_full_throttle": (np.array([120.0, 60.0, 20000.0, 180.0, 800.0] + [0.0] * 7), np.tile([0.0, -0.05, 0.0, 0.9], (100, 1)))}
    ics = random_ics(rng, 6)
    for k in range(6):
        out[f"random_{k}"] = (ics[k], random_controls(rng, 100))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference", help="checkout of junghoseong/aircombat-selfplay")
    ap.add_argument("--out", default=str(ROOT / "tests" / "golden"))
    args = ap.parse_args()
    try:
        import jsbsim
    except ImportError:
        raise SystemExit("make_fdm_golden.py needs the real JSBSim: pip install jsbsim==1.1.6 (plus gymnasium, pymap3d for the "
                         "reference's imports); it cannot run in the offline build container")
    ref = Path(args.reference)
    sys.path.insert(0, str(ref))
    from envs.JSBSim.core.catalog import Catalog as c
    from envs.JSBSim.core.simulatior import AircraftSimulator
    data = ref / "envs" / "JSBSim" / "data"
    sha = {n: hashlib.sha256((data / p).read_bytes()).hexdigest() for n, p in
           (("f16.xml", "aircraft/f16/f16.xml"), ("F100-PW-229.xml", "engine/F100-PW-229.xml"))}
    action_var = [c.fcs_aileron_cmd_norm, c.fcs_elevator_cmd_norm, c.fcs_rudder_cmd_norm, c.fcs_throttle_cmd_norm]
    names = list(PROPS)
    for name, (ic, controls) in cases(np.random.default_rng(7)).items():
        init_state = {k: float(v) for k, v in zip(IC_KEYS, ic)}
        sim = AircraftSimulator(uid="A0100", color="Blue", model="f16", init_state=init_state, origin=(120.0, 60.0, 0.0), sim_freq=60)
        fdm = sim.jsbsim_exec

        def snap():
            return np.array([fdm.get_property_value(PROPS[n]) for n in names])
        rec, frame = {0: snap()}, 0
        for u in controls:
            sim.set_property_values(action_var, list(u))      # catalog clip on set, as the env applies actions
            for _ in range(12):
                sim.run()
                frame += 1
                if frame in CHECKPOINTS:
                    rec[frame] = snap()
        out = Path(args.out) / f"fdm_jsbsim_{name}.npz"
        np.savez_compressed(out, ic=ic, controls=controls, names=np.array(names), frames=np.array(sorted(rec)),
                            values=np.stack([rec[k] for k in sorted(rec)]), jsbsim_version=str(getattr(jsbsim, "__version__", "?")),
                            sha256=np.array([f"{k}:{v}" for k, v in sha.items()]), dt=1.0 / 60.0)
        print(f"[fdm golden] {out.name}: frames {sorted(rec)} jsbsim {getattr(jsbsim, '__version__', '?')}")


if __name__ == "__main__":
    main()
