#!/bin/bash
# One GPU-box visit for a round's single-GPU artefacts: parity tests, smoke, the bench line (both arms, the driver's
# command), the other BASELINE configs, stage traces, ingest timings, the ncu launch list and full captures.
#   bash tools/gpu_check.sh [tag]     (from the repo root, under gpurun)
tag=${1:-run}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/gpu.txt 2>&1
nproc >> $out/gpu.txt
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $out/pytest_gpu.log 2>&1
echo "pytest exit: $?" >> $out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > $out/smoke.log 2>&1
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $out/bench_reference.json 2> $out/bench.err
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $out/bench.json 2>> $out/bench.err
for cfg in k4 hgdp200 batch64; do
  timeout 900 python bench.py --config $cfg --steps 10 --warmup 3 > $out/bench_$cfg.json 2>> $out/bench.err
done
timeout 200 python tools/trace_latency.py > $out/trace.txt 2>&1
timeout 200 python tools/ingest_time.py > $out/ingest_time.txt 2>/dev/null
VB2_CLI_TIMING=1 bash tools/gpu_cli_time.sh $tag/cli > $out/cli_time.txt 2>&1
# profiler passes: a number printed under ncu is never a bench value; --no-session because a resident kernel cannot be replayed
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $out/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-session > $out/bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:llk_flow_kernel -s 30 -c 1 -f -o $out/prof_flow \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-session > $out/ncu_flow.log 2>&1
VB2_STREAM_KERNEL=queue timeout 600 ncu --set full --clock-control none --import-source on -k regex:llk_stream_kernel -s 3 -c 1 -f -o $out/prof_stream \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-session > $out/ncu_stream.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:llk_kernel -s 40 -c 1 -f -o $out/prof_latency \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-session > $out/ncu_latency.log 2>&1
for mb in fp64 mix; do [ -x tools/microbench_$mb ] || nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench_$mb tools/microbench_$mb.cu; done
./tools/microbench_fp64 > $out/microbench_fp64.txt 2>&1
./tools/microbench_mix > $out/microbench_mix.txt 2>&1
VB2_STREAM_KERNEL=queue timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $out/bench_queue_kernel.json 2>> $out/bench.err
tail -3 $out/pytest_gpu.log; tail -2 $out/smoke.log; tail -3 $out/bench.err
cut -c1-400 $out/bench.json; echo; cut -c1-300 $out/bench_reference.json
