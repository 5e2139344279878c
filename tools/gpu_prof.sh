#!/bin/bash
# ncu --set full of one kernel.  bash tools/gpu_prof.sh tag kernel-regex [skip]
tag=${1:-prof}; rx=${2:-llk_stream_kernel}; skip=${3:-4}
out=gpurun_out/$tag
mkdir -p $out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -f -o $out/prof \
    python bench.py --steps 370 --warmup 37 --no-cpu-baseline > $out/ncu_full.log 2>&1
tail -2 $out/ncu_full.log | cut -c1-300
