#!/usr/bin/env python
"""Per-source-line / per-region instruction and stall-sample totals of one ncu capture taken with --import-source on.
usage: tools/ncu_lines.py <report.ncu-rep> [top-n]"""
import collections, csv, subprocess, sys
rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 50
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur, hdr, lines = "", None, []
for r in csv.reader(txt.splitlines()):
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif r and r[0] == "Line No":
        hdr = {c: i for i, c in enumerate(r)}
        i_s, i_n = r.index("# Samples"), r.index("Instructions Executed")
    elif hdr and r and r[0].strip().isdigit():
        try:
            lines.append((cur, int(r[0]), int(r[i_s] or 0), int(r[i_n] or 0), r[1].strip()))
        except ValueError:
            pass
tot_s = sum(x[2] for x in lines) or 1
tot_n = sum(x[3] for x in lines) or 1
REGIONS = [("fdm_core.cuh", 44, 80, "vector / matrix helpers"), ("fdm_core.cuh", 81, 122, "table lookups (bracket, tab1, tab2)"),
           ("fdm_core.cuh", 123, 183, "FCS primitives (pid, kinemat, pow)"), ("fdm_core.cuh", 186, 258, "ISA atmosphere + pitot"),
           ("fdm_core.cuh", 320, 400, "quaternion / Euler / location_derived"), ("fdm_core.cuh", 401, 475, "Propagate"),
           ("fdm_core.cuh", 476, 500, "gravity + atmosphere stage"), ("fdm_core.cuh", 558, 610, "Auxiliary"),
           ("fdm_core.cuh", 611, 690, "engine + fuel"), ("fdm_core.cuh", 770, 835, "mass / cg / inertia"),
           ("fdm_core.cuh", 836, 930, "lean frame body (accelerations)"), ("fdm_core.cuh", 931, 2000, "refresh / outputs / reset")]
reg = collections.OrderedDict()
by_file = collections.Counter(); by_file_s = collections.Counter()
for f, ln, s, n, t in lines:
    by_file[f] += n; by_file_s[f] += s
    key = None
    for rf, a, b, name in REGIONS:
        if f == rf and a <= ln <= b:
            key = name
    if key is None:
        key = f
    e = reg.setdefault(key, [0, 0]); e[0] += n; e[1] += s
print(f"warp instructions {tot_n}, samples {tot_s}")
print("\nby region: share of warp instructions, share of stall samples")
for k, (n, s) in sorted(reg.items(), key=lambda kv: -kv[1][0]):
    print(f"  {100 * n / tot_n:5.1f}%  {100 * s / tot_s:5.1f}%  {k}")
print(f"\ntop {topn} lines by samples: samples%, inst%, file:line, text")
for f, ln, s, n, t in sorted(lines, key=lambda x: -x[2])[:topn]:
    print(f"  {100 * s / tot_s:5.2f}% {100 * n / tot_n:5.2f}%  {f}:{ln}  {t[:140]}")
