#!/bin/bash
# Depth-aware walk stacks + CTAs-per-SM target: parity subset, then timings for both register targets.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu.py -m gpu -q -k "search or cli or shard or job or smoke or reserve" --timeout=180 -p no:cacheprovider -s > gpurun_out/pytest_search.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_search.log
SETS="mt-10 h1 h4 mh1 Metazoan mh3 rbc14x759 its36 53humans.28"
for wb in 6 4; do
  XMP_WALK_BLOCKS=$wb timeout 600 python scripts/time_search.py $SETS 2>&1 | grep rep1 > gpurun_out/time_search_wb$wb.log
done
grep -E "passed|failed|Error|error" gpurun_out/pytest_search.log | tail -10
for f in wb6 wb4; do echo "== $f"; cut -c1-16,110-250 gpurun_out/time_search_$f.log; done
