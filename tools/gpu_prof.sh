#!/bin/bash
# ncu --set full of one launch of a kernel of the bench (a number printed under ncu is never a bench value).
#   bash tools/gpu_prof.sh tag [kernel-regex] [skip] [bench args...]
tag=${1:-prof}; rx=${2:-llk_stream_kernel}; skip=${3:-3}; shift 3
out=gpurun_out/$tag
mkdir -p $out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -f -o $out/prof \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-session "$@" > $out/ncu_full.log 2>&1
tail -2 $out/ncu_full.log | cut -c1-300
ls -la $out
