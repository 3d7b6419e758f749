#!/usr/bin/env python
"""Per-instruction view of an ncu report: opcode mix, stall mix, hottest SASS lines.
usage: python tools/ncu_src.py <rep> <kernel-regex> [top]"""
import collections, csv, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--kernel-name", "regex:" + rx], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines())); hdr = rows[0]
for r in rows[2:3]:
    def g(k): return r[hdr.index(k)] if k in hdr else "?"
    print("kernel", g("Kernel Name")[:60], "time", g("gpu__time_duration.sum"), "regs", g("launch__registers_per_thread"),
          "fp64%", g("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"), "issue%", g("smsp__issue_active.avg.pct_of_peak_sustained_active"),
          "inst", g("smsp__inst_executed.sum"), "lsu%", g("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
          "local ld", g("smsp__inst_executed_op_local_ld.sum"), "local st", g("smsp__inst_executed_op_local_st.sum"))
    items = [(h.replace("smsp__pcsamp_warps_issue_stalled_", ""), float(r[i].replace(",", ""))) for i, h in enumerate(hdr)
             if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("not_issued")]
    tot = sum(v for _, v in items)
    print("stalls: " + "  ".join(f"{h}:{100*v/tot:.1f}%" for h, v in sorted(items, key=lambda kv: -kv[1])[:9]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines())); hdr = rows[1]
isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
data = [(int(r[isamp]), int(r[iex]), r[isrc].strip(), i) for i, r in enumerate(rows[2:]) if len(r) > isamp and r[isamp].isdigit()]
half = len(data) // 2
if half and [d[2] for d in data[:half]] == [d[2] for d in data[half:]]: data = data[:half]   # same kernel listed per launch
tot = sum(d[0] for d in data); totex = sum(d[1] for d in data)
print("SASS instr", len(data), "executed", totex)
ops, samp = collections.Counter(), collections.Counter()
for s, e, text, i in data:
    t = text.split(); op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    ops[op] += e; samp[op] += s
print("  ".join(f"{op}:{100*c/totex:.1f}%/{100*samp[op]/tot:.1f}%" for op, c in ops.most_common(14)), "(exec/samples)")
for s, e, text, i in sorted(data, reverse=True)[:top]:
    print(f"{100*s/tot:5.2f}% idx={i:5d} exec={e:10d} {text[:90]}")
