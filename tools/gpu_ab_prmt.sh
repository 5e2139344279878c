for rep in 1 2; do for name in main prmt; do
  lib=$PWD/verifybamid_b200/libvb2llk_$name.so; [ "$name" = main ] && lib=$PWD/verifybamid_b200/libvb2llk.so
  VB2_STREAM_KERNEL=queue VB2_LLK_LIBRARY=$lib timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > /tmp/b_$name.json 2>/tmp/b_$name.err
  python - <<PY
import json
d=json.load(open("/tmp/b_$name.json"))
print("$name rep $rep: queue-kernel us/eval %.3f  one-launch %.2f  e2e search %.2f  host-driven session %.2f" % (d["us_per_evaluation"], d["roofline"]["us_per_evaluation_one_launch_each"], d["e2e"]["us_per_evaluation"], d["e2e"]["us_per_evaluation_host_driven_session"]))
PY
done; done
