#!/bin/bash
# Wall-clock of the product CLI on the headline workload, phase by phase.  bash tools/gpu_cli_time.sh [tag]
tag=${1:-cli}
out=gpurun_out/$tag
mkdir -p $out
python - <<PY
import sys, os
sys.path.insert(0, os.getcwd())
import bench
from verifybamid_b200 import panels
s = bench.make_workload("100k30x")
panels.write_text_panel(s.panel, "$out/panel")
s.write_pileup("$out/sample.pileup")
PY
ls -la $out/panel.* $out/sample.pileup | awk '{print $5, $9}'
nvidia-smi -L | wc -l
for rep in 1 2 3; do
  t0=$(date +%s.%N)
  ./verifybamid_b200/VerifyBamID --SVDPrefix $out/panel --PileupFile $out/sample.pileup --Reference x --NumPC 2 --Output $out/o $CLI_EXTRA > $out/stdout.$rep.txt 2> $out/stderr.$rep.txt
  t1=$(date +%s.%N)
  grep -E "Finished phase|evaluations|device" $out/stderr.$rep.txt
  python -c "print('process wall %.3f s' % ($t1 - $t0))"
  echo ---
done
rm -f $out/panel.* $out/sample.pileup
