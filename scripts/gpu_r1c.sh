#!/bin/bash
# Tree-heavy searches after the emission / overflow / drain changes.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu.py -m gpu -q -k "deep or mh3 or its36 or cli or reserve or shard" --timeout=180 -p no:cacheprovider -s > gpurun_out/pytest_trees.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_trees.log
timeout 600 python scripts/time_search.py mh3 its36 53humans.24 53humans.26 53humans.28 h4 2>&1 | grep rep1 > gpurun_out/time_search_trees.log
grep -E "s device|passed|failed|Error|error" gpurun_out/pytest_trees.log | tail -30
cut -c1-250 gpurun_out/time_search_trees.log
