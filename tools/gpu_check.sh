#!/bin/bash
# One GPU-box visit: parity tests, the bench line, the ncu launch list and one full capture of the LLK kernel.
# Usage (from the repo root, under gpurun):  bash tools/gpu_check.sh [tag]
tag=${1:-run}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/gpu.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $out/pytest_gpu.log 2>&1
echo "pytest exit: $?" >> $out/pytest_gpu.log
timeout 600 python bench.py > $out/bench.json 2> $out/bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches.csv \
    python bench.py --steps 370 --warmup 37 --no-cpu-baseline > $out/bench_under_ncu.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:llk_kernel -s 4 -c 2 -f -o $out/prof \
    python bench.py --steps 370 --warmup 37 --no-cpu-baseline > $out/ncu_full.log 2>&1
tail -3 $out/pytest_gpu.log
cat $out/bench.json
