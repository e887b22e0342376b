"""IR -> straight-line CUDA device code (``csrc/gen/f16_gen.cuh``): model compiler back end.

Where JSBSim walks a property tree and virtual component objects every frame
(reference data/src/models/FGFCS.cpp:153-178, FGAerodynamics.cpp:132-300), the kernel runs code
*compiled* from the same XML:

* every property becomes a named member of ``struct Props`` (scalar-replaced into registers);
* properties nothing ever writes are folded to constants (gear down, fbw-override 0, ...);
* properties that are read before they are written within a frame are the frame-to-frame carried
  state (one-frame lags of alpha/Mach/Vc/n-pilot seen by the FCS, actuator positions, ...) -- the
  set is derived by dataflow over the frame's model order, not hand-listed;
* all 43 tables are packed into one fp64 blob staged in shared memory; table lookups that share
  an independent variable and a breakpoint vector (25 tables on the 12-point alpha grid) share one
  bracket search.
"""
from __future__ import annotations

from pathlib import Path

from .jsbxml import load_f16
from .gen_oracle import prop_list, EXTRA_PROPS

AXES = ["DRAG", "SIDE", "LIFT", "ROLL", "PITCH", "YAW"]

# run-time inputs (set by the env step) -- carried in state
INPUT_PROPS = ["fcs/aileron-cmd-norm", "fcs/elevator-cmd-norm", "fcs/rudder-cmd-norm", "fcs/throttle-cmd-norm"]
# published by the hand-written core models, in frame order (stage name -> props)
CORE_PUBLISH = [
    ("propagate", ["attitude/pitch-rad", "attitude/roll-rad", "attitude/cos-pitch-cos-roll", "velocities/u-fps",
                   "velocities/v-fps"]),
    ("atmosphere", ["atmosphere/density-altitude"]),
    ("fcs", None),
    ("auxiliary", ["aero/alpha-rad", "aero/alpha-deg", "aero/beta-rad", "aero/qbar-psf", "velocities/mach",
                   "velocities/vc-kts", "velocities/vg-fps", "velocities/p-aero-rad_sec",
                   "velocities/q-aero-rad_sec", "velocities/r-aero-rad_sec", "accelerations/n-pilot-y-norm",
                   "accelerations/n-pilot-z-norm", "aero/h_b-mac-ft"]),
    ("propulsion", None),
    ("aero", None),
]
# initial values of never-written properties (FGFCS ctor, J/models/FGFCS.cpp:72-96; metrics)
CONST_INIT = {"gear/gear-cmd-norm": 1.0}
# props with non-zero initial value that ARE written by components
STATE_INIT = {"gear/gear-pos-norm": 1.0}
# derived at load time from another carried prop instead of being stored
DERIVED = {"aero/alpha-deg": ("aero/alpha-rad", "RADTODEG")}


def cid(name):
    return name.replace("/", "_").replace("-", "_").replace("[", "").replace("]", "")


def lit(x):
    s = repr(float(x))
    if "e" in s or "." in s or "inf" in s or "nan" in s:
        return s
    return s + ".0"


class CudaGen:
    def __init__(self, ir):
        self.ir = ir
        self.props = prop_list(ir)
        self.tab = []          # packed doubles
        self.packed = {}       # tuple(values) -> offset
        self.keyarrays = {}    # tuple(keys) -> offset
        self.out = []
        self.analyse()

    # ------------------------------------------------------------------ table blob
    def pack(self, arr):
        """Append to the table blob; identical arrays are stored once (the split aero functions re-reference them)."""
        key = tuple(float(x) for x in arr)
        if key not in self.packed:
            self.packed[key] = len(self.tab)
            self.tab.extend(key)
        return self.packed[key]

    def keys_off(self, keys):
        """Breakpoint vector followed by its reciprocal spacings: k[n + r] = 1 / (k[r] - k[r-1]), k[n] unused."""
        k = tuple(float(x) for x in keys)
        if k not in self.keyarrays:
            inv = [0.0] + [1.0 / (k[r] - k[r - 1]) for r in range(1, len(k))]
            self.keyarrays[k] = self.pack(list(k) + inv)
        return self.keyarrays[k]

    # ------------------------------------------------------------------ dataflow
    def comp_reads(self, c):
        r = [p for _, p in c["inputs"]]
        if c["type"] == "switch":
            for v in [c["default"]] + [t["value"] for t in c["tests"]]:
                if v[0] == "prop":
                    r.append(v[1])
            for t in c["tests"]:
                for cd in t["conds"]:
                    r.append(cd["prop"])
                    if cd["rhs"][0] == "prop":
                        r.append(cd["rhs"][1])
        if c["type"] == "scheduled_gain":
            r.append(c["table"]["row_var"])
        if c["type"] == "pid" and c["trigger"]:
            r.append(c["trigger"][1])
        if c["type"] == "fcs_function":
            r += self.fac_reads(c["factors"])
        if c["type"] == "kinematic":
            r.append(c["outputs"][0])
        return r

    @staticmethod
    def fac_reads(fs):
        r = []
        for f in fs:
            if f[0] in ("prop", "cos", "sin"):
                r.append(f[1])
            elif f[0] == "table":
                r.append(f[1]["row_var"])
                if f[1]["dim"] == 2:
                    r.append(f[1]["col_var"])
        return r

    def analyse(self):
        ir = self.ir
        written, carried = set(), []
        ever_written = set(INPUT_PROPS)

        def read(p):
            if p not in written and p not in carried:
                carried.append(p)

        def write(p):
            written.add(p)
            ever_written.add(p)
            if p == "fcs/speedbrake-pos-deg":
                written.add("fcs/speedbrake-pos-rad"); ever_written.add("fcs/speedbrake-pos-rad")

        for stage, pubs in CORE_PUBLISH:
            if pubs is not None:
                for p in pubs:
                    write(p)
            elif stage == "fcs":
                read("fcs/throttle-cmd-norm"); write("fcs/throttle-pos-norm")
                for c in ir["fcs"]:
                    for p in self.comp_reads(c):
                        read(p)
                    for o in c["outputs"]:
                        write(o)
            elif stage == "propulsion":
                for p in ("velocities/mach", "atmosphere/density-altitude", "fcs/throttle-pos-norm"):
                    read(p)
            elif stage == "aero":
                for f in ir["aero_pre"]:
                    for p in self.fac_reads(f["factors"]):
                        read(p)
                    write(f["name"])
                write("aero/bi2vel"); write("aero/ci2vel")
                for ax in ir["aero_axes"]:
                    for f in ax["functions"]:
                        for p in self.fac_reads(f["factors"]):
                            read(p)
        self.consts = {}
        self.carried = []
        for p in carried:
            if p in INPUT_PROPS:
                continue
            if p not in ever_written:
                self.consts[p] = CONST_INIT.get(p, 0.0)
            elif p in DERIVED:
                pass
            else:
                self.carried.append(p)
        for p, (src, _) in DERIVED.items():
            if p in carried and src not in self.carried:
                self.carried.append(src)
        self.derived = {p: v for p, v in DERIVED.items() if p in carried}
        metrics = {"metrics/Sw-sqft": ir["metrics"]["Sw"], "metrics/bw-ft": ir["metrics"]["bw"],
                   "metrics/cbarw-ft": ir["metrics"]["cbarw"]}
        for k, v in metrics.items():
            self.consts[k] = v
        self.pids = [c["name"] for c in ir["fcs"] if c["type"] == "pid"]

    # ------------------------------------------------------------------ expression helpers
    def pexpr(self, name, sign=1.0):
        if name in self.consts:
            e = lit(self.consts[name])
        else:
            e = "p." + cid(name)
        return e if sign >= 0 else f"(-{e})"

    def vexpr(self, v):
        if v[0] == "value":
            return lit(v[1])
        return self.pexpr(v[1], v[2] if len(v) > 2 else 1.0)

    # ------------------------------------------------------------------ table lookups with shared brackets
    def bracket(self, var, keys, scope):
        off = self.keys_off(keys)
        key = (var, off)
        if key not in scope["brackets"]:
            name = f"bk{len(scope['brackets'])}"
            scope["brackets"][key] = name
            scope.setdefault("vars", set()).add(var)
            k = [float(x) for x in keys]
            # "uniform enough": every breakpoint within a quarter step of the straight line through the end points (the
            # XML rounds radians to 4 digits), so the multiply lands within one row of the exact search result, which
            # the fix-up step in f16_bracket_u then restores
            step = (k[-1] - k[0]) / (len(k) - 1)
            uniform = len(k) >= 12 and all(abs(k[i] - (k[0] + i * step)) <= 0.25 * step for i in range(len(k)))
            if uniform:
                call = f"f16_bracket_u(T, {off}, {len(k)}, {self.pexpr(var)}, {lit(k[0])}, {lit((len(k) - 1) / (k[-1] - k[0]))})"
            else:
                call = f"f16_bracket(T, {off}, {len(k)}, {self.pexpr(var)})"
            scope.setdefault("decls", []).append(f"  const Bracket {name} = {call};")
        return scope["brackets"][key]

    def table_expr(self, t, scope):
        rb = self.bracket(t["row_var"], t["row_keys"], scope)
        if t["dim"] == 1:
            voff = self.pack(t["values"])
            return f"f16_tab1(T + {voff}, {len(t['row_keys'])}, {rb})"
        cb = self.bracket(t["col_var"], t["col_keys"], scope)
        flat = [x for row in t["values"] for x in row]
        voff = self.pack(flat)
        return f"f16_tab2(T + {voff}, {len(t['col_keys'])}, {rb}, {cb})"

    def product(self, fs, scope):
        terms = []
        fs = list(fs)
        # cos(pitch) * cos(roll) is element (3,3) of the local-to-body matrix the core already holds: no trig, and the
        # Euler angles need not be extracted every frame
        cp, cr = ("cos", "attitude/pitch-rad"), ("cos", "attitude/roll-rad")
        if any(tuple(f[:2]) == cp for f in fs) and any(tuple(f[:2]) == cr for f in fs):
            fs = [f for f in fs if tuple(f[:2]) not in (cp, cr)] + [("prop", "attitude/cos-pitch-cos-roll")]
        for f in fs:
            if f[0] == "prop":
                terms.append(self.pexpr(f[1]))
            elif f[0] == "value":
                terms.append(lit(f[1]))
            elif f[0] == "table":
                terms.append(self.table_expr(f[1], scope))
            elif f[0] in ("cos", "sin"):
                terms.append(f"{f[0]}({self.pexpr(f[1])})")
        e = terms[0]
        for t in terms[1:]:
            e = f"({e} * {t})"
        return e

    # ------------------------------------------------------------------ FCS
    def gen_fcs(self):
        scope = {"brackets": {}, "code": []}
        code = scope["code"]
        code.append("  p.fcs_throttle_pos_norm = p.fcs_throttle_cmd_norm;  // FGFCS::Run: ThrottlePos = ThrottleCmd")
        for ci, c in enumerate(self.ir["fcs"]):
            t = c["type"]
            code.append(f"  // [{c['channel']}] {t} {c['name']}")
            code.append("  {")
            code.append("    double out;")
            ins = [self.pexpr(p, s) for s, p in c["inputs"]]
            if t == "switch":
                first = True
                for te in c["tests"]:
                    conds = []
                    for cd in te["conds"]:
                        rhs = lit(cd["rhs"][1]) if cd["rhs"][0] == "value" else self.pexpr(cd["rhs"][1])
                        conds.append(f"({self.pexpr(cd['prop'])} {cd['op']} {rhs})")
                    joined = (" && " if te["logic"] == "AND" else " || ").join(conds)
                    code.append(f"    {'if' if first else 'else if'} ({joined}) out = {self.vexpr(te['value'])};")
                    first = False
                code.append(f"    {'else ' if not first else ''}out = {self.vexpr(c['default'])};")
            elif t == "pure_gain":
                code.append(f"    out = {lit(c['gain'])} * {ins[0]};")
            elif t == "scheduled_gain":
                code.append(f"    out = {lit(c['gain'])} * {self.table_expr(c['table'], scope)} * {ins[0]};")
            elif t == "aerosurface_scale":
                code.append(f"    {{ const double in = {ins[0]};")
                if c["zero_centered"]:
                    def over(d):   # in / d; a domain limit other than 0, +-1 becomes a multiplication by its reciprocal
                        return f"(in / {lit(d)})" if d in (0.0, 1.0, -1.0) else f"(in * (1.0 / {lit(d)}))"
                    code.append(f"      if (in == 0.0) out = 0.0; else if (in > 0) out = {over(c['in_max'])} * {lit(c['out_max'])}; "
                                f"else out = {over(c['in_min'])} * {lit(c['out_min'])};")
                else:
                    code.append(f"      out = {lit(c['out_min'])} + ((in - {lit(c['in_min'])}) / ({lit(c['in_max'])} - {lit(c['in_min'])})) * "
                                f"({lit(c['out_max'])} - {lit(c['out_min'])});")
                code.append(f"      out *= {lit(c['gain'])}; }}")
            elif t == "summer":
                code.append("    out = 0.0;")
                for e in ins:
                    code.append(f"    out += {e};")
                code.append(f"    out += {lit(c['bias'])};")
            elif t == "pid":
                n = cid(c["name"])
                trig = self.pexpr(c["trigger"][1], c["trigger"][0]) if c["trigger"] else "0.0"
                itype = {"none": 0, "rect": 1, "trap": 2, "ab2": 3, "ab3": 4}[c["int_type"]]
                code.append(f"    out = f16_pid({ins[0]}, {trig}, {lit(c['kp'])}, {lit(c['ki'])}, {lit(c['kd'])}, {itype}, fcs_dt, "
                            f"s.pid_{n}_prev, s.pid_{n}_prev2, s.pid_{n}_itot);")
            elif t == "kinematic":
                scale = "" if c["noscale"] else f" * {lit(c['detents'][-1])}"
                if len(c["detents"]) == 2:
                    code.append(f"    out = f16_kinemat2({lit(c['detents'][0])}, {lit(c['detents'][1])}, {lit(c['times'][1])}, "
                                f"{ins[0]}{scale}, {self.pexpr(c['outputs'][0])}, fcs_dt);")
                else:
                    doff = self.pack(c["detents"])
                    toff = self.pack(c["times"])
                    code.append(f"    out = f16_kinemat(T + {doff}, T + {toff}, {len(c['detents'])}, {ins[0]}{scale}, "
                                f"{self.pexpr(c['outputs'][0])}, fcs_dt);")
            elif t == "fcs_function":
                code.append(f"    out = {self.product(c['factors'], scope)};")
                if ins:
                    code.append(f"    out *= {ins[0]};")
            if c["clip"]:
                code.append(f"    out = f16_constrain({lit(c['clip'][0])}, out, {lit(c['clip'][1])});")
            for o in c["outputs"]:
                code.append(f"    p.{cid(o)} = out;")
                if o == "fcs/speedbrake-pos-deg":
                    code.append("    p.fcs_speedbrake_pos_rad = out * DEGTORAD;")
            code.append("  }")
        written = {o for c in self.ir["fcs"] for o in c["outputs"]}
        assert not (scope.get("vars", set()) & written), "bracket variable written inside the FCS"
        return "\n".join(scope.get("decls", []) + code)

    def aero_reads(self, axes):
        """Properties the coefficient functions of the given axes read (pre-functions they use included)."""
        reads = set()
        for ax in self.ir["aero_axes"]:
            if ax["axis"] in axes:
                for f in ax["functions"]:
                    reads |= set(self.fac_reads(f["factors"]))
        for f in self.ir["aero_pre"]:
            if f["name"] in reads:
                reads |= set(self.fac_reads(f["factors"]))
        return reads

    def gen_aero(self):
        """Body of f16_aero<AXMASK>: the brackets and pre-functions come first (pure, so the ones an axis subset does not
        use are removed by the compiler), each axis build-up sits under its mask bit."""
        scope = {"brackets": {}, "code": []}
        code = scope["code"]
        for f in self.ir["aero_pre"]:
            code.append(f"  p.{cid(f['name'])} = {self.product(f['factors'], scope)};")
        code.append("  // FGAerodynamics::Run: bi2vel/ci2vel after the pre-functions (J/models/FGAerodynamics.cpp:152-158)")
        code.append("  if (twovel != 0) { const double r2v = fm_rcp(twovel); p.aero_bi2vel = K_bw * r2v; p.aero_ci2vel = K_cbarw * r2v; }")
        for ax in self.ir["aero_axes"]:
            i = AXES.index(ax["axis"])
            code.append(f"  // axis {ax['axis']}")
            code.append(f"  if (AXMASK & {1 << i}) {{ double acc = 0.0;")
            for f in ax["functions"]:
                code.append(f"    acc += {self.product(f['factors'], scope)};  // {f['name']}")
            code.append(f"    f[{i}] = acc; }}")
        written = {f["name"] for f in self.ir["aero_pre"]} | {"aero/bi2vel", "aero/ci2vel"}
        assert not (scope.get("vars", set()) & written), "bracket variable written inside aero"
        return "\n".join(scope.get("decls", []) + code)

    def gen_engine(self):
        scope = {"brackets": {}, "code": []}
        e = self.ir["engine"]
        lines = []
        exprs = {k: self.table_expr(e[k], scope) for k in ("IdleThrust", "MilThrust", "AugThrust")}
        lines += scope.get("decls", [])
        lines.append(f"  idle = {exprs['IdleThrust']};")
        lines.append(f"  mil = {exprs['MilThrust']};")
        lines.append(f"  aug = {exprs['AugThrust']};")
        return "\n".join(lines)

    # ------------------------------------------------------------------ file
    def run(self):
        ir = self.ir
        fcs = self.gen_fcs()
        aero = self.gen_aero()
        eng = self.gen_engine()
        o = []
        o.append("// GENERATED by aircombat_selfplay_b200/modelc/gen_cuda.py from the reference's")
        o.append("// envs/JSBSim/data/aircraft/f16/f16.xml + engine/F100-PW-229.xml -- do not edit.")
        o.append("#pragma once")
        m, mass, e = ir["metrics"], ir["mass"], ir["engine"]
        o.append(f"#define F16_NTAB {len(self.tab)}")
        o.append(f"static constexpr double K_Sw = {lit(m['Sw'])}, K_bw = {lit(m['bw'])}, K_cbarw = {lit(m['cbarw'])};")
        for nm in ("AERORP", "EYEPOINT", "VRP"):
            o.append(f"static constexpr double K_{nm}_X = {lit(m[nm][0])}, K_{nm}_Y = {lit(m[nm][1])}, K_{nm}_Z = {lit(m[nm][2])};")
        for k in ("ixx", "iyy", "izz", "ixy", "ixz", "iyz", "emptywt"):
            o.append(f"static constexpr double K_{k} = {lit(mass[k])};")
        o.append(f"static constexpr bool K_negated_crossproduct_inertia = {'true' if mass['negated_crossproduct_inertia'] else 'false'};")
        o.append(f"static constexpr double K_CG_X = {lit(mass['cg'][0])}, K_CG_Y = {lit(mass['cg'][1])}, K_CG_Z = {lit(mass['cg'][2])};")
        o.append(f"static constexpr int K_NPM = {len(mass['pointmasses'])};")
        for i, pm in enumerate(mass["pointmasses"]):
            o.append(f"static constexpr double K_PM{i}_W = {lit(pm['weight'])}, K_PM{i}_X = {lit(pm['loc'][0])}, "
                     f"K_PM{i}_Y = {lit(pm['loc'][1])}, K_PM{i}_Z = {lit(pm['loc'][2])};")
        o.append(f"static constexpr int K_NTANKS = {len(ir['tanks'])};")
        for i, t in enumerate(ir["tanks"]):
            o.append(f"static constexpr double K_TANK{i}_X = {lit(t['loc'][0])}, K_TANK{i}_Y = {lit(t['loc'][1])}, "
                     f"K_TANK{i}_Z = {lit(t['loc'][2])}, K_TANK{i}_CONTENTS = {lit(t['contents'])};")
        assert len(mass["pointmasses"]) == 2 and len(ir["tanks"]) == 4, "fdm_core.cuh is written for 2 point masses / 4 tanks"
        for k in ("milthrust", "maxthrust", "bypassratio", "tsfc", "atsfc", "idlen1", "idlen2", "maxn1", "maxn2"):
            o.append(f"static constexpr double K_ENG_{k} = {lit(e[k])};")
        o.append(f"static constexpr double K_ENG_idleff = {lit(e['milthrust'] ** 0.2 * 107.0)};  // pow(MilThrust, 0.2) * 107 (FGTurbine.cpp:231)")
        o.append(f"static constexpr int K_ENG_augmented = {e['augmented']}, K_ENG_augmethod = {e['augmethod']};")
        o.append(f"static constexpr double K_THRUSTER_X = {lit(e['thruster_loc'][0])}, K_THRUSTER_Y = {lit(e['thruster_loc'][1])}, "
                 f"K_THRUSTER_Z = {lit(e['thruster_loc'][2])};")
        # props struct: everything that is not a folded constant
        members = [p for p in self.props if p not in self.consts]
        for stage, pubs in CORE_PUBLISH:          # core-published helpers that are not JSBSim properties
            for p in pubs or []:
                if p not in members and p not in self.consts:
                    members.append(p)
        o.append("struct Props {")
        for p in members:
            o.append(f"  double {cid(p)};  // {p}")
        o.append("};")
        o.append("struct FcsState {")
        for n in self.pids:
            c = cid(n)
            o.append(f"  double pid_{c}_prev, pid_{c}_prev2, pid_{c}_itot;")
        o.append("};")
        # carried-state field table (X-macro): name, expression
        o.append("// frame-to-frame carried properties (derived by dataflow) + run-time inputs + PID states")
        o.append("#define F16_CARRIED_FIELDS(X) \\")
        rows = []
        for p in INPUT_PROPS + self.carried:
            rows.append(f'  X("{p}", p.{cid(p)})')
        for n in self.pids:
            c = cid(n)
            rows.append(f'  X("pid:{n}:prev", s.pid_{c}_prev)')
            rows.append(f'  X("pid:{n}:prev2", s.pid_{c}_prev2)')
            rows.append(f'  X("pid:{n}:itot", s.pid_{c}_itot)')
        o.append(" \\\n".join(rows))
        o.append(f"#define F16_N_CARRIED {len(rows)}")
        o.append("// initial values of carried props at model load")
        o.append("__device__ __forceinline__ void f16_props_init(Props& p, FcsState& s) {")
        for p in members:
            o.append(f"  p.{cid(p)} = {lit(STATE_INIT.get(p, 0.0))};")
        for n in self.pids:
            c = cid(n)
            o.append(f"  s.pid_{c}_prev = 0.0; s.pid_{c}_prev2 = 0.0; s.pid_{c}_itot = 0.0;")
        o.append("}")
        o.append("__device__ __forceinline__ void f16_props_derive(Props& p) {")
        for p, (src, k) in self.derived.items():
            o.append(f"  p.{cid(p)} = p.{cid(src)} * {k};")
        o.append("}")
        o.append("__device__ __forceinline__ void f16_fcs(Props& p, FcsState& s, const double* __restrict__ T, const double fcs_dt) {")
        o.append(fcs)
        o.append("}")
        o.append("// AXMASK selects the axes to build (bit i = DRAG, SIDE, LIFT, ROLL, PITCH, YAW); f[i] is written only for those")
        o.append("template <int AXMASK = 63>")
        o.append("__device__ __forceinline__ void f16_aero(Props& p, const double* __restrict__ T, const double twovel, double f[6]) {")
        o.append(aero)
        o.append("}")
        # ---- two-warp frame (csrc/env_kernels.cuh, k_env_substeps_split): role A = equations of motion, atmosphere,
        # auxiliary, engine and a subset of the aero axes; role B = flight controls and the other axes.  The exchange
        # and ownership lists come from the same dataflow as the carried-state list: reads the other role does not use
        # are dead code on its side.
        fcs_written = {x for c in ir["fcs"] for x in c["outputs"]} | {"fcs/speedbrake-pos-rad", "fcs/throttle-pos-norm"}
        core_pub = [x for st, pubs in CORE_PUBLISH if st in ("atmosphere", "auxiliary") for x in (pubs or [])]
        fcs_reads = set()
        for c in ir["fcs"]:
            fcs_reads |= set(self.comp_reads(c))
        need_b = self.aero_reads(AXES) | fcs_reads
        x_aux = [x for x in core_pub if x in need_b and x not in self.consts]
        x_surf = sorted(x for x in (self.aero_reads(AXES) | {"fcs/throttle-pos-norm"}) if x in fcs_written and x not in self.consts)
        # the FCS reads pitch / roll only as cos(pitch) * cos(roll) (see product()), which Propagate publishes directly
        x_early = ["attitude/cos-pitch-cos-roll"] + [x for x in CORE_PUBLISH[0][1] if x in fcs_reads and not x.startswith("attitude/")]
        air = ("aero/qbar-psf", "velocities/mach", "velocities/vc-kts")     # published by the air-data half of Auxiliary
        x_kin = [x for x in x_aux if x not in air]
        x_air = [x for x in x_aux if x in air]
        for nm, lst in (("EARLY", x_early), ("AUX", x_aux), ("KIN", x_kin), ("AIR", x_air), ("SURF", x_surf)):
            o.append(f"#define F16_X_{nm}(X) " + " ".join(f"X(p.{cid(x)})" for x in lst))
            o.append(f"#define F16_N_X_{nm} {len(lst)}")
        own_b, own_a = [], []
        for r in rows:
            name = r.split('"')[1]
            (own_b if (name in INPUT_PROPS or name in fcs_written or name.startswith("pid:")) else own_a).append(name)
        o.append("// carried fields stored by the flight-control role (B) of the two-warp frame; the rest belong to role A")
        o.append("#define F16_CARRIED_ROLE_B(name) (" + " || ".join(f'f16_streq(name, "{n}")' for n in own_b) + ")")
        o.append("__device__ __forceinline__ void f16_engine_tables(const Props& p, const double* __restrict__ T, double& idle, double& mil, double& aug) {")
        o.append(eng)
        o.append("}")
        # table blob last (all packs done)
        blob = ["static const double F16_TAB_HOST[F16_NTAB] = {"]
        for i in range(0, len(self.tab), 8):
            blob.append("  " + ", ".join(lit(x) for x in self.tab[i:i + 8]) + ",")
        blob.append("};")
        text = "\n".join(o) + "\n" + "\n".join(blob) + "\n"
        return text.replace("#define F16_NTAB __PENDING__", f"#define F16_NTAB {len(self.tab)}")


def generate(out_path: Path, ir=None):
    ir = ir or load_f16()
    g = CudaGen(ir)
    text = g.run()
    # F16_NTAB was emitted before all tables were packed; patch it
    import re
    text = re.sub(r"#define F16_NTAB \d+", f"#define F16_NTAB {len(g.tab)}", text, count=1)
    out_path.parent.mkdir(parents=True, exist_ok=True)
    if not out_path.exists() or out_path.read_text() != text:
        out_path.write_text(text)
    return g


if __name__ == "__main__":
    root = Path(__file__).resolve().parents[1]
    g = generate(root / "csrc" / "gen" / "f16_gen.cuh")
    print("wrote csrc/gen/f16_gen.cuh:", len(g.tab), "table doubles;", len(g.carried), "carried props;",
          len(g.consts), "folded constants")
    print("carried:", g.carried)
    print("consts:", g.consts)
