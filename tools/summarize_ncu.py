#!/usr/bin/env python
"""Turn ncu outputs from a gpurun session into the tracked summaries under profiles/.

    python tools/summarize_ncu.py launches gpurun_out/<tag>/launches.csv  profiles/<name>.md
    python tools/summarize_ncu.py full     gpurun_out/<tag>/prof.ncu-rep  profiles/<name>.md

`launches`: per-kernel launch count, total/avg device time and SHARE of the step (the ncu times
are cold-cache and serialised: compare shares, not absolutes).
`full`: the metrics DESIGN.md / bench.py quote for the profiled launches (DRAM bytes, L1/L2/FP64
pipe utilisation, registers, occupancy).
"""
import collections
import csv
import subprocess
import sys

FULL_METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    # tensor / FP64 pipe evidence for the DMMA (fp64) and HMMA-TF32 (fp32) kernels
    "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active_realtime.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
    "smsp__warps_eligible.avg.per_cycle_active",
]
# every warp-stall reason ncu reports (warps per issue-active cycle), sorted by value in the summary
STALL_PREFIX = "smsp__average_warps_issue_stalled_"
STALL_SUFFIX = "_per_issue_active.ratio"


def launches(src, dst):
    rows = list(csv.reader(open(src)))
    hdr = None
    agg = collections.OrderedDict()
    for r in rows:
        if len(r) > 5 and r[0] == "ID":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            name = d["Kernel Name"].split("(")[0].replace("void ", "").replace("unnamed>::", "")
            v = float(d["Metric Value"].replace(",", ""))
            v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(d["Metric Unit"], 1.0)
            a = agg.setdefault(name, [0, 0.0])
            a[0] += 1
            a[1] += v
    tot = sum(v[1] for v in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list ({src})\n\n")
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none` over `bench.py`; times are "
                "cold-cache and serialised, compare SHARES.\n\n")
        f.write("| kernel | launches | total ms | avg ms | share |\n|---|---:|---:|---:|---:|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {v[0]} | {v[1]:.3f} | {v[1] / v[0]:.3f} | {100 * v[1] / tot:.1f}% |\n")
        f.write(f"\ntotal {tot:.3f} ms over {sum(v[0] for v in agg.values())} launches\n")


def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary ({src})\n\n")
        for r in rows[2:]:
            name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("unnamed>::", "")
            f.write(f"## `{name}`\n\n| metric | value | unit |\n|---|---:|---|\n")
            for mname in FULL_METRICS:
                if mname in idx:
                    f.write(f"| {mname} | {r[idx[mname]]} | {units[idx[mname]]} |\n")
            stalls = []
            for h, i in idx.items():
                if h.startswith(STALL_PREFIX) and h.endswith(STALL_SUFFIX) and "_not_issued" not in h:
                    try:
                        stalls.append((float(r[i].replace(",", "")), h[len(STALL_PREFIX):-len(STALL_SUFFIX)]))
                    except ValueError:
                        pass
            if stalls:
                f.write("\nwarp stall reasons (warps per issue-active cycle, largest first):\n\n| reason | value |\n|---|---:|\n")
                for v, nm in sorted(stalls, reverse=True)[:8]:
                    f.write(f"| {nm} | {v:.3f} |\n")
            if "dram__bytes_read.sum" in idx:
                def gb(name):
                    v = float(r[idx[name]].replace(",", ""))
                    return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(units[idx[name]], 1.0)
                f.write(f"\ntraffic (dram read + write) = {gb('dram__bytes_read.sum') + gb('dram__bytes_write.sum'):.0f} bytes\n\n")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
