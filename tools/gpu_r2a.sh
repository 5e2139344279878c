#!/bin/bash
# Round-2 visit A: parity tests, the bench line (both arms, the driver's command), the phase clock.
tag=${1:-r2a}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/gpu.txt 2>&1
nproc >> $out/gpu.txt
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $out/pytest_gpu.log 2>&1
echo "pytest exit: $?" >> $out/pytest_gpu.log
tail -5 $out/pytest_gpu.log
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $out/bench_reference.json 2> $out/bench.err
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $out/bench.json 2>> $out/bench.err
tail -5 $out/bench.err
cut -c1-1500 $out/bench.json
bash tools/gpu_phase.sh $tag
