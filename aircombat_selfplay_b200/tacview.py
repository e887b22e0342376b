"""TacView ``.acmi`` text log of ONE environment of a device batch -- the reference's ``render(mode="txt")``
(reference envs/JSBSim/envs/env_base.py:207-250, core/simulatior.py:73-79,535-551).  Off the hot path: each call copies
the arenas of the batch to the host and formats the lines of env ``env_index``."""
from __future__ import annotations

import math

import numpy as np

A_WGS, F_WGS = 6378137.0, 1.0 / 298.257223563
B_WGS = A_WGS * (1.0 - F_WGS)


def _geodetic2ecef(lat, lon, alt):
    la, lo = math.radians(lat), math.radians(lon)
    n = A_WGS ** 2 / math.hypot(A_WGS * math.cos(la), B_WGS * math.sin(la))
    return (n + alt) * math.cos(la) * math.cos(lo), (n + alt) * math.cos(la) * math.sin(lo), (n * (B_WGS / A_WGS) ** 2 + alt) * math.sin(la)


def neu2lla(n, e, u, lon0, lat0, alt0):
    """NEU2LLA of the reference (envs/JSBSim/utils/utils.py:44-55): (lon, lat, alt) of a local north-east-up point."""
    la, lo = math.radians(lat0), math.radians(lon0)
    x0, y0, z0 = _geodetic2ecef(lat0, lon0, alt0)
    t = math.cos(la) * u - math.sin(la) * n
    w = math.sin(la) * u + math.cos(la) * n
    x, y, z = x0 + math.cos(lo) * t - math.sin(lo) * e, y0 + math.sin(lo) * t + math.cos(lo) * e, z0 + w
    p = math.hypot(x, y)
    lat = math.atan2(z, p * (1 - (1 - (B_WGS / A_WGS) ** 2)))
    for _ in range(5):      # fixed-point iteration on the geodetic latitude; sub-millimetre after 3 rounds
        nn = A_WGS / math.sqrt(1 - (1 - (B_WGS / A_WGS) ** 2) * math.sin(lat) ** 2)
        alt = p / math.cos(lat) - nn
        lat = math.atan2(z, p * (1 - (1 - (B_WGS / A_WGS) ** 2) * nn / (nn + alt)))
    return math.degrees(math.atan2(y, x)), math.degrees(lat), alt


class AcmiWriter:
    def __init__(self, core, env_index: int = 0):
        self.core, self.i = core, env_index
        self._created = False
        self._exploded = set()

    def write(self, filepath: str, timestamp: float):
        core, i = self.core, self.i
        b, spec = core.batch, core.spec
        A, S = spec.n_agents, max(spec.n_missile_slots, 1)
        arenas = {k: (n, t.cpu().numpy()) for k, (n, t) in ((k, b.arena(k)) for k in ("out", "ac_d", "ac_i", "ms_d", "ms_i"))}

        def field(arena, name, idx):
            names, t = arenas[arena]
            return t[names.index(name), idx]
        if not self._created:
            with open(filepath, mode="w", encoding="utf-8-sig") as f:
                f.write("FileType=text/acmi/tacview\nFileVersion=2.1\n0,ReferenceTime=2020-04-01T00:00:00Z\n")
            self._created = True
        uids = core.ego_ids + core.enm_ids
        lines = [f"#{timestamp:.2f}"]
        cfgs = core.config["aircraft_configs"]
        for a, uid in enumerate(uids):
            row = i * A + a
            lon, lat, alt = field("out", "lon_deg", row), field("out", "lat_geod_deg", row), field("ac_d", "h_sl_m", row)
            r, p, y = (math.degrees(field("out", k, row)) for k in ("roll_rad", "pitch_rad", "heading_rad"))
            lines.append(f"{uid},T={lon}|{lat}|{alt}|{r}|{p}|{y},Name={str(cfgs[uid].get('model', 'f16')).upper()},"
                         f"Color={cfgs[uid].get('color', 'Red')}")
            for s in range(int(field("ac_i", "n_launched", row))):
                mid = row * S + s
                muid = f"{uid}{int(field('ms_i', 'keyn', mid))}"
                status = int(field("ms_i", "status", mid))
                n, e, u = (field("ms_d", k, mid) for k in ("pos_n", "pos_e", "pos_u"))
                mlon, mlat, malt = neu2lla(n, e, u, *spec.center)
                pitch, yaw = math.degrees(field("ms_d", "theta", mid)), math.degrees(field("ms_d", "phi", mid))
                color = cfgs[uid].get("color", "Red")
                if status == 0:
                    lines.append(f"{muid},T={mlon}|{mlat}|{malt}|0.0|{pitch}|{yaw},Name=AIM-9L,Color={color}")
                elif (i, mid) not in self._exploded:
                    self._exploded.add((i, mid))
                    rc = 300 if int(field("ms_i", "kind", mid)) == 0 else 5
                    lines.append(f"-{muid}\n{muid}F,T={mlon}|{mlat}|{malt}|0.0|{pitch}|{yaw},Type=Misc+Explosion,Color={color},Radius={rc}")
                else:
                    lines.append(f"-{muid}")
        with open(filepath, mode="a", encoding="utf-8-sig") as f:
            f.write("\n".join(lines) + "\n")
