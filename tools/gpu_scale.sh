#!/bin/bash
# strong-scaling points of bench.py on this box: bash tools/gpu_scale.sh N [N ...]
mkdir -p gpurun_out/scale
for n in "$@"; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $n --steps 4000 --warmup 50 2> gpurun_out/scale/err_$n.txt | tail -1 > gpurun_out/scale/bench_${n}gpu.json
  python - <<PY
import json
d=json.load(open("gpurun_out/scale/bench_${n}gpu.json"))
print("N=$n us/step %.3f value %.3e e2e us/step %.2f steps/launch %d" % (d["ms_per_step"]*1e3, d["value"], d["e2e"]["us_per_step"], d["config"]["steps_per_launch"]))
PY
done
