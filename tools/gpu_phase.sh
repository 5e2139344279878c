#!/bin/bash
# Per-phase cycle counters of the many-evaluations kernel (diagnostics build, csrc/Makefile target `phase`).
#   bash tools/gpu_phase.sh [tag]      (from the repo root, under gpurun)
# The counters are in llk_stream_kernel: VB2_STREAM_KERNEL=queue sends the bench's batches there.  Build the diagnostics
# library first (make -C verifybamid_b200/csrc phase); it is not part of `make all`.
tag=${1:-phase}
out=gpurun_out/$tag
mkdir -p $out
[ -f verifybamid_b200/libvb2llk_phase.so ] || make -s -C verifybamid_b200/csrc phase
VB2_STREAM_KERNEL=queue VB2_LLK_LIBRARY=$PWD/verifybamid_b200/libvb2llk_phase.so timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline \
    > $out/bench_phase.json 2> $out/phase.txt
grep "phase clock" $out/phase.txt | tail -3
