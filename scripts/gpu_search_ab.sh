#!/bin/bash
# A/B of the exact search on one B200: parity subset first, then per-dataset wall-clock breakdowns for each
# setting of an environment knob.  Usage (through gpurun):
#     bash scripts/gpu_search_ab.sh XMP_FLAGS 0 4        # eager (default) vs classic layout
#     bash scripts/gpu_search_ab.sh XMP_BEAM 128 512 2048
# Every measurement quoted in profiles/r01_search_tuning.md was taken this way.
KNOB=${1:-XMP_FLAGS}; shift
VALUES=${@:-0}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu.py -m gpu -q -x -k "search or cli or shard or job or smoke or reserve" --timeout=180 -p no:cacheprovider > gpurun_out/pytest_search.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_search.log
tail -4 gpurun_out/pytest_search.log
SETS="mt-10 h1 h4 mh1 Metazoan mh3 rbc14x759 its36 53humans.28"
for v in $VALUES; do
  env $KNOB=$v timeout 600 python scripts/time_search.py $SETS 2>&1 | grep rep1 > gpurun_out/time_search_${KNOB}_$v.log
  echo "== $KNOB=$v"; cut -c1-16,110-250 gpurun_out/time_search_${KNOB}_$v.log
done
