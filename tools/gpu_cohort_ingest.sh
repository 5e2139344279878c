#!/bin/bash
# Cohort throughput with the pileups read by the host vs by the device.  bash tools/gpu_cohort_ingest.sh [tag] [n_samples]
tag=${1:-cohort}; n=${2:-16}
out=gpurun_out/$tag
mkdir -p $out /tmp/cohort
python - <<PY
import sys, os
sys.path.insert(0, os.getcwd())
from verifybamid_b200 import panels, synth
panel = panels.load_bundled("1000g.phase3.100k.b37")
lines = []
for i in range($n):
    s = synth.make_sample(panel, n_pc=2, depth=30.0, alpha=0.02, seed=100 + i)
    if i == 0:
        panels.write_text_panel(s.panel, "/tmp/cohort/panel")
    s.write_pileup("/tmp/cohort/s%d.pileup" % i)
    lines.append("/tmp/cohort/s%d.pileup\t/tmp/cohort/o%d\n" % (i, i))
open("/tmp/cohort/list", "w").write("".join(lines))
PY
for mode in host device host device; do
  if [ $mode = host ]; then export VB2_HOST_INGEST=1; else unset VB2_HOST_INGEST; fi
  t0=$(date +%s.%N)
  ./verifybamid_b200/VerifyBamID --SVDPrefix /tmp/cohort/panel --PileupList /tmp/cohort/list --Reference x --NumPC 2 > $out/stdout.$mode.txt 2> $out/stderr.$mode.txt
  t1=$(date +%s.%N)
  python -c "print('$mode reader: process wall %.3f s;' % ($t1 - $t0), open('$out/stderr.$mode.txt').read().strip().splitlines()[-1])"
  md5sum /tmp/cohort/o3.selfSM /tmp/cohort/o3.Ancestry | awk '{printf "%s ", $1}'; echo
done
nproc
