#!/bin/bash
# e2e + one-launch device time, several repetitions, for each library given (default = in-tree): host-side variance check
for lib in "$@"; do
if [ "$lib" = default ]; then unset VB2_LLK_LIBRARY; else export VB2_LLK_LIBRARY=$PWD/$lib; fi
echo "== $lib"
python - <<'PY'
import os, sys, numpy as np
sys.path.insert(0, os.getcwd())
import torch
import bench, verifybamid_b200 as vb
s = bench.make_workload()
stream = torch.cuda.Stream()
engines = [vb.LLKEngine(s.problem, stream=stream.cuda_stream) for _ in range(4)]
pc = np.full(2, 0.01)
for rep in range(3):
    secs, last = vb.time_host(engines, 50, 3000, pc, pc, 0.03)
    one = vb.time_device(engines, 50, 3000, pc, pc, 0.03) / 3000
    print("e2e %.2f us/step   one-launch %.2f us   llk %.6f" % (secs / 3000 * 1e6, one * 1e3, last))
PY
done
