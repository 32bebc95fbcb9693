#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu.py -m gpu -q -k "build_midpoints or score_batch or pair or full_size or invalid" --timeout=600 -p no:cacheprovider > gpurun_out/pytest_quick.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_quick.log
timeout 900 python bench.py --no-bandb > gpurun_out/bench_quick.log 2> gpurun_out/bench_quick.err
echo "exit $?" >> gpurun_out/bench_quick.err
tail -5 gpurun_out/pytest_quick.log; python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench_quick.log') if l.startswith('{')][0])
print('value',d['value'],'ms',d['ms_per_step']); print('e2e',d['e2e']); print('cpu',d['cpu_baseline']); print(d['clocks'])
"; tail -3 gpurun_out/bench_quick.err
