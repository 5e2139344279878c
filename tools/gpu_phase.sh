#!/bin/bash
# Per-phase cycle counters of the many-evaluations kernel (diagnostics build, csrc/Makefile target `phase`).
#   bash tools/gpu_phase.sh [tag]      (from the repo root, under gpurun)
tag=${1:-phase}
out=gpurun_out/$tag
mkdir -p $out
VB2_LLK_LIBRARY=$PWD/verifybamid_b200/libvb2llk_phase.so timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline \
    > $out/bench_phase.json 2> $out/phase.txt
grep "phase clock" $out/phase.txt | tail -3
