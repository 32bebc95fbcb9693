#!/bin/bash
mkdir -p gpurun_out
for v in 0 1 2 3 4; do
  XMP_SCORE_TMA=$v timeout 300 python bench.py --steps 20 --warmup 3 --kernel-only 2>&1 | tail -1 | python -c "
import sys, json, statistics
d = json.loads(sys.stdin.readline())
print('XMP_SCORE_TMA=$v ms_per_step %.4f median %.4f min %.4f checksum %d' % (d['ms_per_step'], statistics.median(d['ms_steps']), min(d['ms_steps']), d['checksum']))
"
done > gpurun_out/tma_vs_ldg2.log 2>&1
cat gpurun_out/tma_vs_ldg2.log
