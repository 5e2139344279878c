#!/bin/bash
# Launch list + one ncu --set full capture of llk_flow_kernel from the bench (numbers under ncu are never bench values).
#   bash tools/gpu_flow_prof.sh [tag]
tag=${1:-flow_prof}
out=gpurun_out/$tag
mkdir -p $out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $out/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-session > $out/bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:llk_flow_kernel -s 30 -c 1 -f -o $out/prof_flow \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-session > $out/ncu_flow.log 2>&1
tail -2 $out/ncu_flow.log | cut -c1-300
ls -la $out
