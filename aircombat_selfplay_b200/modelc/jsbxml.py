"""JSBSim aircraft/engine XML -> flat model IR ("model compiler", front end).

The reference re-parses ``aircraft/f16/f16.xml`` and ``engine/F100-PW-229.xml`` inside every
``AircraftSimulator.reload`` (reference ``envs/JSBSim/core/simulatior.py:165-168``).  Here the XML is
parsed ONCE, at build time, into a plain-dict IR that the two back ends consume:

* ``gen_oracle.py``  -> C initialisers interpreted at run time by the CPU oracle (``oracle/``)
* ``gen_cuda.py``    -> straight-line CUDA device code for the sm_100a kernels (``csrc/gen``)

What is parsed (reference ``envs/JSBSim/data/aircraft/f16/f16.xml``):
  metrics (:38-59), mass_balance (:61-88), propulsion/tanks (:262-308), flight_control channels
  (:317-992), aerodynamics functions (:994-1925); engine tables from ``engine/F100-PW-229.xml``.
The ``pushback`` and ``hook`` system files are not compiled: their only outputs are the external
reaction magnitudes / tail-hook animation, which are identically zero in flight.
"""
from __future__ import annotations

import re
import xml.etree.ElementTree as ET
from pathlib import Path

IN2FT = 1.0 / 12.0


def _f(el, tag, default=None):
    sub = el.find(tag)
    if sub is None:
        if default is None:
            raise KeyError(tag)
        return default
    return float(sub.text.strip())


def _loc(el):
    return [_f(el, "x"), _f(el, "y"), _f(el, "z")]


def _signed_prop(text):
    """'-fcs/foo' -> (-1.0, 'fcs/foo')  (FGPropertyValue sign handling)."""
    text = text.strip()
    if text.startswith("-"):
        return (-1.0, text[1:].strip())
    return (1.0, text)


def _is_number(s):
    try:
        float(s)
        return True
    except ValueError:
        return False


def parse_table(el):
    """<table> -> dict.  1-D: rows [[key, val]...]; 2-D: row keys, col keys, data[r][c]
    (FGTable layout, reference data/src/math/FGTable.cpp:443-517)."""
    ivars = el.findall("independentVar")
    data = [float(x) for x in el.find("tableData").text.split()]
    if len(ivars) == 1:
        assert len(data) % 2 == 0
        keys = data[0::2]
        vals = data[1::2]
        assert all(keys[i] < keys[i + 1] for i in range(len(keys) - 1)), "1-D keys must increase"
        return {"dim": 1, "row_var": ivars[0].text.strip(), "row_keys": keys, "values": vals}
    assert len(ivars) == 2
    row_var = col_var = None
    for iv in ivars:
        if iv.get("lookup") == "row":
            row_var = iv.text.strip()
        elif iv.get("lookup") == "column":
            col_var = iv.text.strip()
    assert row_var and col_var
    # first text line holds the column keys; count them from the raw text
    lines = [ln.split() for ln in el.find("tableData").text.strip().splitlines() if ln.strip()]
    col_keys = [float(x) for x in lines[0]]
    ncol = len(col_keys)
    row_keys, values = [], []
    for ln in lines[1:]:
        assert len(ln) == ncol + 1, ln
        row_keys.append(float(ln[0]))
        values.append([float(x) for x in ln[1:]])
    return {"dim": 2, "row_var": row_var, "col_var": col_var, "row_keys": row_keys,
            "col_keys": col_keys, "values": values}


def parse_function(el):
    """<function> holding one <product> (the only form the F-16 aero model uses) or a bare <table>.
    Returns a list of factors: ('prop', name) | ('value', x) | ('table', tabledict)."""
    kids = [c for c in el if c.tag not in ("description",)]
    assert len(kids) == 1, (el.get("name"), [k.tag for k in kids])
    root = kids[0]
    if root.tag == "table":
        return [("table", parse_table(root))]
    assert root.tag == "product", root.tag
    factors = []
    for c in root:
        if c.tag == "property":
            factors.append(("prop", c.text.strip()))
        elif c.tag == "value":
            factors.append(("value", float(c.text)))
        elif c.tag == "table":
            factors.append(("table", parse_table(c)))
        elif c.tag in ("cos", "sin"):
            (inner,) = list(c)
            assert inner.tag == "property"
            factors.append((c.tag, inner.text.strip()))
        else:
            raise NotImplementedError(c.tag)
    return factors


def _parse_condition_lines(test_el):
    conds = []
    text = test_el.text or ""
    for ln in text.strip().splitlines():
        ln = ln.strip()
        if not ln:
            continue
        a, op, b = ln.split()
        op = {"lt": "<", "le": "<=", "gt": ">", "ge": ">=", "eq": "==", "ne": "!=",
              "LT": "<", "LE": "<=", "GT": ">", "GE": ">=", "EQ": "==", "NE": "!="}.get(op, op)
        assert op in ("<", "<=", ">", ">=", "==", "!=")
        rhs = ("value", float(b)) if _is_number(b) else ("prop", b)
        conds.append({"prop": a, "op": op, "rhs": rhs})
    assert len(list(test_el)) == 0, "nested <test> groups are not used by the F-16 model"
    return conds


def _value_or_prop(s):
    s = s.strip()
    if _is_number(s):
        return ("value", float(s))
    sign, name = _signed_prop(s)
    return ("prop", name, sign)


def parse_component(el):
    """One FCS component (reference data/src/models/flight_control/*.cpp)."""
    name = el.get("name")
    c = {"type": el.tag, "name": name if "/" in name else "fcs/" + name.lower().replace(" ", "-")}
    c["inputs"] = [_signed_prop(i.text) for i in el.findall("input")]
    outs = [o.text.strip() for o in el.findall("output")]
    # FGFCSComponent ctor collects <output> nodes first, then bind() appends the name node
    c["outputs"] = outs + [c["name"]]
    clip = el.find("clipto")
    c["clip"] = None
    if clip is not None:
        c["clip"] = (float(clip.find("min").text), float(clip.find("max").text))
    t = el.tag
    if t == "switch":
        d = el.find("default")
        c["default"] = _value_or_prop(d.get("value")) if d is not None else ("value", 0.0)
        c["tests"] = []
        for te in el.findall("test"):
            c["tests"].append({"logic": te.get("logic", "AND"),
                               "value": _value_or_prop(te.get("value")),
                               "conds": _parse_condition_lines(te)})
    elif t == "pure_gain":
        g = el.find("gain")
        c["gain"] = float(g.text) if g is not None else 1.0
    elif t == "scheduled_gain":
        g = el.find("gain")
        c["gain"] = float(g.text) if g is not None else 1.0
        c["table"] = parse_table(el.find("table"))
    elif t == "aerosurface_scale":
        g = el.find("gain")
        c["gain"] = float(g.text) if g is not None else 1.0
        dom = el.find("domain")
        c["in_min"], c["in_max"] = -1.0, 1.0
        if dom is not None and dom.find("min") is not None and dom.find("max") is not None:
            c["in_min"], c["in_max"] = _f(dom, "min"), _f(dom, "max")
        rng = el.find("range")
        c["out_min"], c["out_max"] = _f(rng, "min"), _f(rng, "max")
        zc = el.find("zero_centered")
        c["zero_centered"] = not (zc is not None and zc.text.strip() in ("0", "false"))
    elif t == "summer":
        b = el.find("bias")
        c["bias"] = float(b.text) if b is not None else 0.0
    elif t == "pid":
        assert el.get("type", "") != "standard"
        for k in ("kp", "ki", "kd"):
            e = el.find(k)
            c[k] = float(e.text) if e is not None else 0.0
        ki = el.find("ki")
        itype = ki.get("type", "") if ki is not None else None
        c["int_type"] = {None: "none", "rect": "rect", "trap": "trap", "ab2": "ab2", "ab3": "ab3"}.get(
            itype, "ab2")
        tr = el.find("trigger")
        c["trigger"] = _signed_prop(tr.text) if tr is not None else None
        assert el.find("pvdot") is None
    elif t == "kinematic":
        c["noscale"] = el.find("noscale") is not None
        c["detents"], c["times"] = [], []
        for s in el.find("traverse").findall("setting"):
            c["detents"].append(_f(s, "position"))
            c["times"].append(_f(s, "time"))
    elif t == "fcs_function":
        c["factors"] = parse_function(el.find("function"))
    else:
        raise NotImplementedError(t)
    return c


def parse_engine(path: Path):
    r = ET.parse(path).getroot()
    assert r.tag == "turbine_engine"
    e = {k: _f(r, k) for k in ("milthrust", "maxthrust", "bypassratio", "tsfc", "atsfc", "idlen1",
                               "idlen2", "maxn1", "maxn2")}
    e["augmented"] = int(_f(r, "augmented"))
    e["augmethod"] = int(_f(r, "augmethod"))
    e["injected"] = int(_f(r, "injected"))
    for fn in r.findall("function"):
        e[fn.get("name")] = parse_table(fn.find("table"))
    return e


def parse_aircraft(aircraft_xml: Path, engine_dir: Path):
    r = ET.parse(aircraft_xml).getroot()
    ir = {"name": r.get("name")}
    m = r.find("metrics")
    ir["metrics"] = {"Sw": _f(m, "wingarea"), "bw": _f(m, "wingspan"), "cbarw": _f(m, "chord")}
    for loc in m.findall("location"):
        ir["metrics"][loc.get("name")] = _loc(loc)
    mb = r.find("mass_balance")
    neg = mb.get("negated_crossproduct_inertia")
    ir["mass"] = {k: _f(mb, k, 0.0) for k in ("ixx", "iyy", "izz", "ixy", "ixz", "iyz")}
    ir["mass"]["negated_crossproduct_inertia"] = (neg != "false")
    ir["mass"]["emptywt"] = _f(mb, "emptywt")
    for loc in mb.findall("location"):
        if loc.get("name") == "CG":
            ir["mass"]["cg"] = _loc(loc)
    ir["mass"]["pointmasses"] = [{"name": pm.get("name"), "weight": _f(pm, "weight"),
                                  "loc": _loc(pm.find("location"))} for pm in mb.findall("pointmass")]
    pr = r.find("propulsion")
    engs = pr.findall("engine")
    assert len(engs) == 1
    eng = engs[0]
    ir["engine"] = parse_engine(engine_dir / (eng.get("file") + ".xml"))
    ir["engine"]["feeds"] = [int(f.text) for f in eng.findall("feed")]
    th = eng.find("thruster")
    ir["engine"]["thruster_loc"] = _loc(th.find("location"))
    ori = th.find("orient")
    ir["engine"]["thruster_orient"] = [_f(ori, "roll"), _f(ori, "pitch"), _f(ori, "yaw")]
    assert all(a == 0.0 for a in ir["engine"]["thruster_orient"])
    ir["tanks"] = [{"loc": _loc(t.find("location")), "capacity": _f(t, "capacity"),
                    "contents": _f(t, "contents")} for t in pr.findall("tank")]
    fc = r.find("flight_control")
    ir["fcs_declared"] = [p.text.strip() for p in fc.findall("property")]
    ir["fcs"] = []
    for ch in fc.findall("channel"):
        assert ch.get("execrate") is None
        for comp in ch:
            c = parse_component(comp)
            c["channel"] = ch.get("name")
            ir["fcs"].append(c)
    ae = r.find("aerodynamics")
    ir["aero_pre"] = [{"name": f.get("name"), "factors": parse_function(f)} for f in ae.findall("function")]
    ir["aero_axes"] = []
    for ax in ae.findall("axis"):
        ir["aero_axes"].append({"axis": ax.get("name"),
                                "functions": [{"name": f.get("name"), "factors": parse_function(f)}
                                              for f in ax.findall("function")]})
    assert [a["axis"] for a in ir["aero_axes"]] == ["DRAG", "SIDE", "LIFT", "ROLL", "PITCH", "YAW"]
    return ir


def referenced_properties(ir):
    """Every property name the compiled FCS/aero/engine code reads or writes."""
    reads, writes = set(), set()

    def tab(t):
        reads.add(t["row_var"])
        if t["dim"] == 2:
            reads.add(t["col_var"])

    def fac(fs):
        for f in fs:
            if f[0] in ("prop", "cos", "sin"):
                reads.add(f[1])
            elif f[0] == "table":
                tab(f[1])

    for c in ir["fcs"]:
        for _, p in c["inputs"]:
            reads.add(p)
        for o in c["outputs"]:
            writes.add(o)
        if c["type"] == "switch":
            vals = [c["default"]] + [t["value"] for t in c["tests"]]
            for v in vals:
                if v[0] == "prop":
                    reads.add(v[1])
            for t in c["tests"]:
                for cd in t["conds"]:
                    reads.add(cd["prop"])
                    if cd["rhs"][0] == "prop":
                        reads.add(cd["rhs"][1])
        if c["type"] == "scheduled_gain":
            tab(c["table"])
        if c["type"] == "pid" and c["trigger"]:
            reads.add(c["trigger"][1])
        if c["type"] == "fcs_function":
            fac(c["factors"])
        if c["type"] == "kinematic":
            reads.add(c["outputs"][0])
    for f in ir["aero_pre"]:
        fac(f["factors"])
        writes.add(f["name"])
    for ax in ir["aero_axes"]:
        for f in ax["functions"]:
            fac(f["factors"])
    for k in ("IdleThrust", "MilThrust", "AugThrust"):
        tab(ir["engine"][k])
    return reads, writes


REF_ROOT = Path("/root/reference/envs/JSBSim/data")


def load_f16(ref_root: Path = REF_ROOT):
    return parse_aircraft(ref_root / "aircraft" / "f16" / "f16.xml", ref_root / "engine")


if __name__ == "__main__":
    import json
    ir = load_f16()
    rd, wr = referenced_properties(ir)
    print(json.dumps({"n_fcs": len(ir["fcs"]), "reads": sorted(rd), "writes": sorted(wr)}, indent=1))
