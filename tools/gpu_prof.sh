#!/bin/bash
# ncu --set full of the batched LLK kernel only.  bash tools/gpu_prof.sh [tag]
tag=${1:-prof}
out=gpurun_out/$tag
mkdir -p $out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:llk_kernel -s 4 -c 1 -f -o $out/prof \
    python bench.py --steps 370 --warmup 37 --no-cpu-baseline > $out/ncu_full.log 2>&1
tail -3 $out/ncu_full.log
