#!/usr/bin/env python
"""Registers / spills per kernel from `nvcc -Xptxas -v` output on stdin (c++filt-ed names, one line per kernel)."""
import re, subprocess, sys
txt = sys.stdin.read()
cur = None
rows = []
for line in txt.splitlines():
    m = re.search(r"Compiling entry function '(\S+)'", line)
    if m:
        cur = {"name": subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()}
        rows.append(cur)
        continue
    if cur is None:
        continue
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
    if m and "spill" not in cur:
        cur["spill"] = (int(m.group(2)), int(m.group(3)))
    m = re.search(r"Used (\d+) registers", line)
    if m:
        cur["regs"] = int(m.group(1))
        m2 = re.search(r"(\d+) bytes smem", line)
        cur["smem"] = int(m2.group(1)) if m2 else 0
for r in rows:
    n = re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", r["name"])
    n = re.sub(r"\(.*", "", n).replace("void ", "")
    print("%-60s regs %3d  spill st/ld %4d/%4d  smem %5d" % (n, r.get("regs", -1), *r.get("spill", (0, 0)), r.get("smem", 0)))
