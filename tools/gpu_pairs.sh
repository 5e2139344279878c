#!/bin/bash
# the warp-specialised many-evaluations kernel (VB2_LLK_PAIRS=1) against the default one: parity tests, then the bench
export VB2_LLK_PAIRS=1
timeout 150 python -m pytest tests/test_llk_gpu.py -x -q 2>&1 | tail -3
timeout 100 python bench.py --no-cpu-baseline --no-session 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('PAIRS   us/eval %.3f frac %.3f' % (r['us_per_evaluation'], r['frac']))"
unset VB2_LLK_PAIRS
timeout 100 python bench.py --no-cpu-baseline --no-session 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('DEFAULT us/eval %.3f frac %.3f' % (r['us_per_evaluation'], r['frac']))"
