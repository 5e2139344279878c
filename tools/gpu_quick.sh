#!/bin/bash
# Quick GPU-box visit: parity tests + the bench line (no profiler).  bash tools/gpu_quick.sh [tag] [extra env...]
tag=${1:-quick}
out=gpurun_out/$tag
mkdir -p $out
( time timeout 300 python -m pytest tests -m gpu -x -q ) > $out/pytest_gpu.log 2>&1
echo "pytest exit: $?" >> $out/pytest_gpu.log
tail -25 $out/pytest_gpu.log
timeout 200 python bench.py --no-cpu-baseline > $out/bench.json 2> $out/bench.err
cat $out/bench.json; tail -5 $out/bench.err
