#!/bin/bash
# Rebuild the library quietly; print errors/warnings and the per-kernel register/spill summary.
#   bash tools/build_lib.sh [grep-pattern] [extra NVFLAGS...]
pat=${1:-.}; shift
cd "$(dirname "$0")/../verifybamid_b200/csrc" || exit 1
touch llk_engine.cu
make all NVFLAGS_EXTRA="$*" > /tmp/vb2_build.log 2>&1; rc=$?
grep -E "error|warning" /tmp/vb2_build.log | head -20
python ../../tools/ptxas_summary.py < /tmp/vb2_build.log | grep -E "$pat"
exit $rc
