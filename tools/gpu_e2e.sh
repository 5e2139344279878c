#!/bin/bash
# e2e (launch per evaluation / session) + one-launch device time, several repetitions: host-side variance check
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
    print("launch per evaluation: e2e %.2f us/step   one-launch %.2f us   llk %.6f" % (secs / 3000 * 1e6, one * 1e3, last))
eng = engines[:1]
eng[0].session_begin()
for rep in range(4):
    secs, last2 = vb.time_host(eng, 50, 3000, pc, pc, 0.03)
    print("session:               e2e %.2f us/step   llk %.6f" % (secs / 3000 * 1e6, last2))
eng[0].session_end()
secs, last3 = vb.time_host(eng, 50, 3000, pc, pc, 0.03)
print("after session_end:     e2e %.2f us/step   llk %.6f  same bits: %s" % (secs / 3000 * 1e6, last3, last2 == last3))
PY
