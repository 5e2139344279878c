#!/bin/bash
# One GPU-box visit for the round's artefacts: parity tests, smoke, the bench line (both arms), stage traces, the ncu
# launch list and full captures of the two kernels.   bash tools/gpu_check.sh [tag]     (from the repo root, under gpurun)
tag=${1:-run}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/gpu.txt 2>&1
( time timeout 600 python -m pytest tests -m gpu -x -q ) > $out/pytest_gpu.log 2>&1
echo "pytest exit: $?" >> $out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > $out/smoke.log 2>&1
timeout 600 python bench.py > $out/bench.json 2> $out/bench.err
timeout 600 python bench.py --impl reference > $out/bench_reference.json 2>> $out/bench.err
timeout 100 python tools/trace_latency.py > $out/trace.txt 2>&1
# profiler passes: a number printed under ncu is never a bench value; --no-session because ncu serialises launches
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches.csv \
    python bench.py --steps 592 --warmup 296 --no-cpu-baseline --no-session > $out/bench_under_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:llk_stream_kernel -s 6 -c 1 -f -o $out/prof_stream \
    python bench.py --steps 592 --warmup 296 --no-cpu-baseline --no-session > $out/ncu_stream.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:llk_kernel -s 60 -c 1 -f -o $out/prof_latency \
    python bench.py --steps 592 --warmup 296 --no-cpu-baseline --no-session > $out/ncu_latency.log 2>&1
tail -3 $out/pytest_gpu.log; tail -2 $out/smoke.log
cut -c1-600 $out/bench.json; cut -c1-400 $out/bench_reference.json
