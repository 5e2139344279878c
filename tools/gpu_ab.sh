#!/bin/bash
# A/B of engine builds on one box: bash tools/gpu_ab.sh tag name1 name2 ...   (libvb2llk_<name>.so; "main" = libvb2llk.so)
tag=$1; shift
out=gpurun_out/$tag
mkdir -p $out
for rep in 1 2; do
for name in "$@"; do
  lib=$PWD/verifybamid_b200/libvb2llk_$name.so
  [ "$name" = main ] && lib=$PWD/verifybamid_b200/libvb2llk.so
  VB2_LLK_LIBRARY=$lib timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $out/bench_$name.$rep.json 2> $out/bench_$name.$rep.err
  python - <<PY
import json
try:
    d=json.load(open("$out/bench_$name.$rep.json"))
    print("%-12s rep $rep: us/eval %.3f frac %.3f one-launch %.2f e2e %.2f" % ("$name", d["us_per_evaluation"], d["roofline"]["frac"], d["roofline"]["us_per_evaluation_one_launch_each"], d["e2e"]["us_per_evaluation"]))
except Exception as e:
    print("$name failed", e); print(open("$out/bench_$name.$rep.err").read()[-500:])
PY
done
done
