#!/bin/bash
# Eager scoring: quick parity first (stop early on failure), then the search subset, then timings eager vs classic.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu.py -m gpu -q -x -k "random or scheduling or (matches_reference and (its10 or its14 or primate or h3 or mh6 or Eukaryotes or mt-10))" --timeout=120 -p no:cacheprovider > gpurun_out/pytest_eager_quick.log 2>&1
rc=$?
echo "quick exit $rc" >> gpurun_out/pytest_eager_quick.log
tail -25 gpurun_out/pytest_eager_quick.log
if [ $rc -ne 0 ]; then exit 0; fi
timeout 900 python -m pytest tests/test_gpu.py -m gpu -q -k "search or cli or shard or job or smoke or reserve" --timeout=180 -p no:cacheprovider > gpurun_out/pytest_search.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_search.log
tail -5 gpurun_out/pytest_search.log
SETS="mt-10 h1 h4 mh1 Metazoan mh3 rbc14x759 its36 53humans.28"
timeout 600 python scripts/time_search.py $SETS 2>&1 | grep rep1 > gpurun_out/time_search_eager.log
XMP_FLAGS=4 timeout 600 python scripts/time_search.py $SETS 2>&1 | grep rep1 > gpurun_out/time_search_classic.log
for f in eager classic; do echo "== $f"; cut -c1-16,110-250 gpurun_out/time_search_$f.log; done
