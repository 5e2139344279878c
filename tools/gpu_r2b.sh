#!/bin/bash
# Round-2 visit B: parity tests, the bench line, the phase clock.   bash tools/gpu_r2b.sh [tag] [pytest -k expression]
tag=${1:-r2b}
out=gpurun_out/$tag
mkdir -p $out
( time timeout 900 python -m pytest tests -m gpu -x -q ${2:+-k "$2"} ) > $out/pytest_gpu.log 2>&1
echo "pytest exit: $?" >> $out/pytest_gpu.log
tail -5 $out/pytest_gpu.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > $out/bench.json 2> $out/bench.err
tail -5 $out/bench.err
python - <<PY
import json
d=json.load(open("$out/bench.json"))
print("us/eval %.3f" % d["us_per_evaluation"], "frac %.3f" % d["roofline"]["frac"], "one-launch %.2f" % d["roofline"]["us_per_evaluation_one_launch_each"],
      "e2e %.2f" % d["e2e"]["us_per_evaluation"], "batched-call %.2f" % d["e2e"]["us_per_evaluation_batched_public_call"], d["parity"])
PY
bash tools/gpu_phase.sh $tag
