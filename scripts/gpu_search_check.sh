#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu.py -m gpu -q -k "search or cli or shard or job or smoke" --timeout=120 -p no:cacheprovider -s > gpurun_out/pytest_search.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_search.log
python scripts/time_search.py mt-10 h1 h4 mh1 Metazoan mh3 rbc14x759 EukaryotesrDNA 2>&1 | grep rep1 | cut -c1-16,95-260 > gpurun_out/time_search2.log
grep -E "s device|passed|failed|Error|error" gpurun_out/pytest_search.log | tail -45; cat gpurun_out/time_search2.log
