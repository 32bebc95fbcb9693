#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu.py -m gpu -q -x -k "search or cli or shard or job or smoke or reserve" --timeout=180 -p no:cacheprovider > gpurun_out/pytest_search.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_search.log
tail -4 gpurun_out/pytest_search.log
timeout 600 python scripts/time_search.py mt-10 h1 h4 mh1 Metazoan mh3 rbc14x759 its36 53humans.28 2>&1 | grep rep1 > gpurun_out/time_search_eager2.log
cut -c1-16,110-250 gpurun_out/time_search_eager2.log
