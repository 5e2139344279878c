#!/bin/bash
# tests + bench + ncu full of one kernel.  bash tools/gpu_quickprof.sh tag kernel-regex
tag=${1:-qp}; rx=${2:-llk_stream_kernel}
out=gpurun_out/$tag
mkdir -p $out
( time timeout 300 python -m pytest tests -m gpu -x -q ) > $out/pytest_gpu.log 2>&1
echo "pytest exit: $?" >> $out/pytest_gpu.log
tail -12 $out/pytest_gpu.log
timeout 200 python bench.py --no-cpu-baseline > $out/bench.json 2> $out/bench.err
cut -c1-400 $out/bench.json; tail -5 $out/bench.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:$rx -s 4 -c 1 -f -o $out/prof \
    python bench.py --steps 370 --warmup 37 --no-cpu-baseline > $out/ncu_full.log 2>&1
tail -2 $out/ncu_full.log | cut -c1-200
