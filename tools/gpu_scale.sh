#!/bin/bash
# Scaling points of bench.py on this box under the driver's command (one rank per GPU, --steps 20 --warmup 5):
#   bash tools/gpu_scale.sh N [N ...]      (under gpurun --gpus N)
# per N: the headline workload with the NCCL all-reduce and with peer stores from the reduce kernel; at the largest N also
# the 64-sample cohort (no collective) and the collective A/B (a batched step and one dependent evaluation).
mkdir -p gpurun_out/scale
run() {  # n, tag, extra bench args...
  n=$1; tag=$2; shift 2
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $n --steps 20 --warmup 5 "$@" 2> gpurun_out/scale/err_${tag}_$n.txt | tail -1 > gpurun_out/scale/bench_${tag}_${n}gpu.json
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/scale/bench_${tag}_${n}gpu.json"))
    print("N=$n $tag: value %.3e us/eval %s e2e %.3e frac %.3f parity %s" % (d["value"], d.get("us_per_evaluation"), d["e2e"]["value"], d["roofline"]["frac"], {k: v for k, v in d["parity"].items() if k.endswith("rel")}))
except Exception as e:
    print("N=$n $tag failed:", e); print(open("gpurun_out/scale/err_${tag}_$n.txt").read()[-600:])
PY
}
last=1
for n in "$@"; do
  if [ $n = 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/scale/bench_nccl_1gpu.json 2> gpurun_out/scale/err_nccl_1.txt
  else
    run $n nccl
    if [ -z "$SCALE_QUICK" ]; then   # SCALE_QUICK=1: only the headline line per N
      run $n peer --collective peer
      VB2_STREAM_KERNEL=queue run $n queue
    fi
  fi
  last=$n
done
if [ $last -gt 1 ] && [ -z "$SCALE_QUICK" ]; then
  run $last batch64 --config batch64
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $last --master-addr 127.0.0.1 --master-port 29512 \
      tools/collective_ab.py > gpurun_out/scale/collective_ab_$last.txt 2>&1
  tail -6 gpurun_out/scale/collective_ab_$last.txt
fi
