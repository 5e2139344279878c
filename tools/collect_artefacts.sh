#!/bin/bash
# Copy what a round wants judged from gpurun_out/<tag> (scratch) into profiles/ (tracked), named per round.
#   bash tools/collect_artefacts.sh <tag> <round-prefix, e.g. r02>
tag=$1; r=$2
src=gpurun_out/$tag
cp $src/bench.json profiles/${r}_bench_1gpu.json
cp $src/bench_reference.json profiles/${r}_bench_reference_arm.json
for cfg in k4 hgdp200 batch64; do cp $src/bench_$cfg.json profiles/${r}_bench_$cfg.json; done
cp $src/bench_queue_kernel.json profiles/${r}_bench_1gpu_queue_kernel.json
cp $src/pytest_gpu.log profiles/${r}_pytest_gpu.txt
cp $src/smoke.log profiles/${r}_smoke.log
cp $src/trace.txt profiles/${r}_trace_latency.txt
cp $src/ingest_time.txt profiles/${r}_ingest_time.txt
cp $src/launches.csv profiles/${r}_launches_bench.csv
cp $src/microbench_fp64.txt profiles/${r}_microbench_fp64.txt
cp $src/microbench_mix.txt profiles/${r}_microbench_mix.txt
grep -E "Finished phase|evaluations|process wall|^---" $src/cli_time.txt > profiles/${r}_cli_time.txt
for k in flow stream latency; do
  name=llk_${k}_kernel; [ $k = latency ] && name=llk_kernel
  n=120; [ $k = stream ] && n=2048; [ $k = latency ] && n=1
  python tools/ncu_digest.py $src/prof_$k.ncu-rep $n > profiles/${r}_${name}_digest.txt 2>&1
  ncu -i $src/prof_$k.ncu-rep --page details > profiles/${r}_${name}_ncu_details.txt 2>&1
done
for f in llk_flow_kernelILi2 llk_stream_kernelILi2ELb0 llk_kernelILb1ELb1ELi2ELb0 llk_session_kernelILi2 llk_reduce_kernel llk_gather_kernel; do
  short=$(echo $f | sed 's/IL.*//')
  bash tools/sass_of.sh $f > profiles/${r}_sass_$short.txt
done
ls -la profiles | grep ${r}_ | awk '{print $5, $9}'
