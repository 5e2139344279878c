#!/usr/bin/env python
"""Digest of an ncu report (read here, no GPU): headline metrics + per-region SASS statistics.
usage: tools/ncu_digest.py gpurun_out/<tag>/prof.ncu-rep [evals_per_launch]"""
import csv, io, re, subprocess, sys

rep = sys.argv[1]
evals = int(sys.argv[2]) if len(sys.argv) > 2 else 37
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "launch__registers_per_thread", "sm__warps_active.avg.per_cycle_active", "smsp__inst_executed.sum",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]
for k in want:
    if k in hdr:
        i = hdr.index(k)
        print("%-70s %s %s" % (k, vals[i], units[i]))
for i, k in enumerate(hdr):
    if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio"):
        v = float(vals[i])
        if v > 0.05:
            print("  stall %-40s %.2f" % (k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")], v))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]
ix = {k: i for i, k in enumerate(h)}
data = [r for r in rows[2:] if r and r[0].startswith("0x")]
first = data[0][0]
ends = [i for i, r in enumerate(data) if r[0] == first]
data = data[:ends[1]] if len(ends) > 1 else data
tot_i = sum(int(r[ix["Instructions Executed"]]) for r in data)
tot_s = sum(int(r[ix["# Samples"]]) for r in data)
print("warp-instructions per evaluation: %.0f   samples %d" % (tot_i / evals, tot_s))
fp = re.compile(r"(@!?U?P\d+\s+)?(DFMA|DMUL|DADD|DSETP|F2F\.F64)")
with open(rep.replace(".ncu-rep", "_sass.txt"), "w") as f:
    for n, r in enumerate(data):
        f.write("%4d %9d %6d w=%s lsb=%s ssb=%s math=%s  %s\n" % (n, int(r[ix["Instructions Executed"]]), int(r[ix["# Samples"]]),
                r[ix["stall_wait"]], r[ix["stall_long_sb"]], r[ix["stall_short_sb"]], r[ix["stall_math"]], r[ix["Source"]].strip()))
reg, cur = [], None
for n, r in enumerate(data):
    c = int(r[ix["Instructions Executed"]]); s = int(r[ix["# Samples"]]); isfp = c if fp.match(r[ix["Source"]].strip()) else 0
    if cur and abs(c - cur["c"]) <= 0.15 * max(c, cur["c"]):
        cur["n"] += 1; cur["inst"] += c; cur["s"] += s; cur["fp"] += isfp; cur["end"] = n
    else:
        cur = {"start": n, "end": n, "c": c, "n": 1, "inst": c, "s": s, "fp": isfp}
        reg.append(cur)
print("fp64 share of instructions: %.1f%%" % (100.0 * sum(x["fp"] for x in reg) / tot_i))
for x in reg:
    if x["inst"] > 0.004 * tot_i or x["s"] > 0.004 * tot_s:
        print("%4d-%4d n=%3d cnt~%8d inst=%5.1f%% fp64=%5.1f%% samples=%5.1f%%" % (x["start"], x["end"], x["n"], x["c"],
              100 * x["inst"] / tot_i, 100 * x["fp"] / tot_i, 100 * x["s"] / tot_s))
