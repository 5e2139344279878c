mkdir -p gpurun_out/f4
for cfg in 100k30x k4 batch64 hgdp200; do
  timeout 600 python bench.py --config $cfg --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/f4/bench_$cfg.json 2> gpurun_out/f4/bench_$cfg.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/f4/bench_$cfg.json"))
    r = d["roofline"]
    print("$cfg", r["kernel"], "launches/step", r["kernel_launches_per_step"], "us/eval", r["us_per_evaluation"], "frac %.3f" % r["frac"], "e2e", d["e2e"]["us_per_evaluation"], "gpu_launches", d["gpu_launches"], d["parity"])
except Exception as e:
    print("$cfg failed", e); print(open("gpurun_out/f4/bench_$cfg.err").read()[-800:])
PY
done
