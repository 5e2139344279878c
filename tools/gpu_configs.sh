#!/bin/bash
# The non-headline BASELINE configs on one GPU: bash tools/gpu_configs.sh [tag]
tag=${1:-configs}
out=gpurun_out/$tag
mkdir -p $out
for cfg in k4 hgdp200 batch64; do
  timeout 900 python bench.py --config $cfg --steps 10 --warmup 3 > $out/bench_$cfg.json 2> $out/bench_$cfg.err
  tail -2 $out/bench_$cfg.err
  python - <<PY
import json
try:
    d=json.load(open("$out/bench_$cfg.json"))
    c=d.get("cpu_baseline") or {}
    print("%-8s value %.3e  ms/step %.3f  frac %.3f  e2e %.3e (%.2f us/eval)  parity %s  cpu(all) %.3e" % ("$cfg", d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d["e2e"]["us_per_evaluation"], d["parity"], c.get("value", float("nan"))))
except Exception as e:
    print("$cfg failed:", e)
PY
done
