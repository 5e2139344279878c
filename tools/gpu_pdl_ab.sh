#!/bin/bash
# A/B: the launches of a batch overlapped (programmatic stream serialization, default) or strictly serial (VB2_FLOW_PDL=0)
out=gpurun_out/${1:-pdl}
mkdir -p $out
( timeout 600 python -m pytest tests/test_llk_gpu.py -m gpu -x -q -k "many or batch or both or shards or bit_repro" 2>&1 | tail -3 )
for rep in 1 2; do
  for k in pdl serial; do
    if [ $k = serial ]; then export VB2_FLOW_PDL=0; else unset VB2_FLOW_PDL; fi
    timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $out/bench_$k.$rep.json 2> $out/bench_$k.$rep.err
    python - <<PY
import json
try:
    d = json.load(open("$out/bench_$k.$rep.json"))
    print("%-6s rep $rep: us/eval %.3f frac %.3f e2e batched call %.3f parity %s" % ("$k", d["us_per_evaluation"], d["roofline"]["frac"], d["e2e"]["us_per_evaluation_batched_public_call"], d["parity"]["batched_rel"]))
except Exception as e:
    print("$k failed", e); print(open("$out/bench_$k.$rep.err").read()[-800:])
PY
  done
done
