#!/bin/bash
# Quick GPU-box visit: parity tests + latency trace + the bench line (no profiler).  bash tools/gpu_quick.sh [tag]
tag=${1:-quick}
out=gpurun_out/$tag
mkdir -p $out
( time timeout 300 python -m pytest tests -m gpu -x -q ) > $out/pytest_gpu.log 2>&1
echo "pytest exit: $?" >> $out/pytest_gpu.log
tail -12 $out/pytest_gpu.log
timeout 100 python tools/trace_latency.py 2>&1 | grep -v "^rep [0-3]" | tee $out/trace.txt
timeout 200 python bench.py --no-cpu-baseline > $out/bench.json 2> $out/bench.err
python - <<PY
import json
d=json.load(open("$out/bench.json"))
print("us/eval %.3f" % d["roofline"]["us_per_evaluation"], "one-launch %.2f" % d["roofline"]["us_per_evaluation_one_launch_each"], "e2e %.2f" % d["e2e"]["us_per_step"], "frac %.3f" % d["roofline"]["frac"])
PY
tail -3 $out/bench.err
