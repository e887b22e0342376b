#!/usr/bin/env python
"""Summarise one `ncu --set full --import-source on` capture (.ncu-rep) as text: key metrics, SASS opcode mix, stall
reasons, fp64 FLOP count.  usage: tools/ncu_summary.py <report.ncu-rep> <out.txt> [units-per-launch]"""
import collections
import csv
import re
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
units = float(sys.argv[3]) if len(sys.argv) > 3 else None


def page(kind, extra=()):
    txt = subprocess.run(["ncu", "-i", rep, "--page", kind, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(txt.splitlines()))


raw = page("raw")
hdr, unit, val = raw[0], raw[1], raw[2]
m = {h: (v, u) for h, u, v in zip(hdr, unit, val)}
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.per_cycle_active",
        "sm__cycles_elapsed.max", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
        "smsp__issue_active.avg.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
        "smsp__sass_inst_executed_op_shared_ld.sum", "smsp__sass_inst_executed_op_global_ld.sum",
        "smsp__sass_inst_executed_op_global_st.sum"]
lines = [f"# ncu summary of {rep}", ""]
for k in keys:
    if k in m:
        lines.append(f"{k:75s} {m[k][0]} {m[k][1]}")


def num(k):
    try:
        return float(m[k][0].replace(",", ""))
    except Exception:
        return 0.0


if units:
    tr = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}
    dram = sum(num(k) * tr.get(m[k][1], 1.0) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum") if k in m)
    lines.append(f"DRAM traffic per launch: {dram:.6g} B = {dram / units:.1f} B per unit")

sass = page("source", ["--print-source", "sass"])
h = sass[1]
ix = {c: i for i, c in enumerate(h)}
op, samp, stall, thr = collections.Counter(), collections.Counter(), collections.Counter(), collections.Counter()
tot = tots = 0
scols = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
for r in sass[2:]:
    if len(r) < len(h):
        continue
    mm = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ix["Source"]].strip())
    o = (mm.group(2) if mm else r[ix["Source"]]).split(".")[0]
    n, s = int(r[ix["Instructions Executed"]] or 0), int(r[ix["# Samples"]] or 0)
    op[o] += n; samp[o] += s; tot += n; tots += s
    thr[o] += int(r[ix["Predicated-On Thread Instructions Executed"]] or 0)
    for c in scols:
        stall[c] += int(r[ix[c]] or 0)
flops = 2 * thr["DFMA"] + thr["DMUL"] + thr["DADD"]
lines += ["", f"fp64 FLOPs per launch (2*DFMA + DMUL + DADD, thread level, predicated on): {flops:.6g}"]
if units:
    lines.append(f"  per unit ({units:g} units per launch): {flops / units:.1f}")
lines += ["", f"static SASS instructions: {len(sass) - 2}  ({(len(sass) - 2) * 16 / 1024:.0f} KiB)", f"warp instructions executed: {tot}",
          "", "opcode            warp-inst    share   samples"]
for k, v in op.most_common(24):
    lines.append(f"{k:14s} {v:12d}  {100 * v / max(tot, 1):5.1f}%  {100 * samp[k] / max(tots, 1):5.1f}%")
lines += ["", "warp stall reasons (share of samples)"]
for k, v in stall.most_common(10):
    lines.append(f"{k:28s} {100 * v / max(tots, 1):5.1f}%")
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))

# ---- hottest CUDA source lines (per-line totals of the `cuda,sass` source page), appended to the summary
try:
    cur, hot, i_s, i_n = "", [], None, None
    for r in page("source", ["--print-source", "cuda,sass"]):
        if len(r) == 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif r and r[0] == "Line No":
            i_s, i_n = r.index("# Samples"), r.index("Instructions Executed")
        elif i_s is not None and r and r[0].strip().isdigit():
            try:
                hot.append((int(r[i_s] or 0), int(r[i_n] or 0), cur, int(r[0]), r[1].strip()[:130]))
            except ValueError:
                pass
    tot_s, tot_n = sum(x[0] for x in hot) or 1, sum(x[1] for x in hot) or 1
    by_file = collections.Counter()
    for x in hot:
        by_file[x[2]] += x[1]
    extra = ["", "warp instructions by source file: " + ", ".join(f"{k} {100 * v / tot_n:.1f}%" for k, v in by_file.most_common(6)),
             "", "hottest CUDA source lines (share of stall samples, share of warp instructions, file:line, text)"]
    for sm, n, f, ln, t in sorted(hot, reverse=True)[:40]:
        extra.append(f"{100 * sm / tot_s:5.2f}% {100 * n / tot_n:5.2f}%  {f}:{ln}  {t}")
    open(out, "a").write("\n".join(extra) + "\n")
except Exception as e:  # the summary above is the product; this section is best effort
    open(out, "a").write(f"\n(hot source lines unavailable: {e})\n")
