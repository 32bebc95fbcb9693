#!/bin/bash
# Search-side check after the arena / reserve / prefetch changes: parity subset, then wall-clock breakdowns.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu.py -m gpu -q -k "search or cli or shard or job or smoke or reserve" --timeout=180 -p no:cacheprovider -s > gpurun_out/pytest_search.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_search.log
SETS="mt-10 h1 h4 mh1 Metazoan mh3 rbc14x759 its36 53humans.28"
for pf in 1 2; do
  XMP_PREFETCH=$pf timeout 600 python scripts/time_search.py $SETS 2>&1 | grep rep1 > gpurun_out/time_search_pf$pf.log
done
XMP_NO_RESERVE=1 timeout 600 python scripts/time_search.py $SETS 2>&1 | grep rep1 > gpurun_out/time_search_noreserve.log
grep -E "s device|passed|failed|Error|error" gpurun_out/pytest_search.log | tail -30
for f in pf1 pf2 noreserve; do echo "== $f"; cut -c1-250 gpurun_out/time_search_$f.log; done
