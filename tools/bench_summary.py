#!/usr/bin/env python
"""Print the headline numbers of a bench.py JSON line read from stdin (developer convenience)."""
import json
import sys

for line in sys.stdin:
    line = line.strip()
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    if "roofline" not in d:
        print(line[:200])
        continue
    r = d["roofline"]
    print("%s us/eval(batched)=%.2f  us/eval(1 launch each)=%.2f  e2e us/step=%.2f  frac=%.3f  value=%.3g" % (
        " ".join(sys.argv[1:]), d["ms_per_step"] * 1e3, r.get("us_per_evaluation_one_launch_each", float("nan")),
        d["e2e"]["us_per_step"], r["frac"], d["value"]))
