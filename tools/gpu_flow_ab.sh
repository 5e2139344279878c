#!/bin/bash
# A/B of the two many-evaluations kernels on one box: llk_flow_kernel (default) against llk_stream_kernel
# (VB2_STREAM_KERNEL=queue), same library, the driver's bench command.   bash tools/gpu_flow_ab.sh [tag] [pytest -k expr]
tag=${1:-flow_ab}
out=gpurun_out/$tag
mkdir -p $out
if [ -n "$2" ]; then
  ( timeout 900 python -m pytest tests/test_llk_gpu.py -m gpu -x -q -k "$2" 2>&1 | tail -15 ) | tee $out/pytest.log
fi
for rep in 1 2; do
  for k in flow queue; do
    if [ $k = queue ]; then export VB2_STREAM_KERNEL=queue; else unset VB2_STREAM_KERNEL; fi
    timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $out/bench_$k.$rep.json 2> $out/bench_$k.$rep.err
    python - <<PY
import json
try:
    d = json.load(open("$out/bench_$k.$rep.json"))
    print("%-6s rep $rep: us/eval %.3f frac %.3f e2e %.2f parity %s" % ("$k", d["us_per_evaluation"], d["roofline"]["frac"], d["e2e"]["us_per_evaluation"], d["parity"]))
except Exception as e:
    print("$k failed", e); print(open("$out/bench_$k.$rep.err").read()[-800:])
PY
  done
done
unset VB2_STREAM_KERNEL
