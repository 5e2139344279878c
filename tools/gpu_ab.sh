#!/bin/bash
# A/B of library builds: bash tools/gpu_ab.sh tag lib1 lib2 ...   (paths relative to the repo root; "default" = in-tree build)
tag=$1; shift
out=gpurun_out/$tag; mkdir -p $out
for lib in "$@"; do
  if [ "$lib" = default ]; then unset VB2_LLK_LIBRARY; else export VB2_LLK_LIBRARY=$PWD/$lib; fi
  timeout 200 python bench.py --no-cpu-baseline --steps 1480 > $out/bench_$(basename $lib .so).json 2>> $out/err.log
  python - <<PY
import json
d=json.load(open("$out/bench_$(basename $lib .so).json"))
print("$lib", "us/eval %.3f" % d["roofline"]["us_per_evaluation"], "one-launch %.2f" % d["roofline"]["us_per_evaluation_one_launch_each"], "e2e %.2f" % d["e2e"]["us_per_step"])
PY
done
