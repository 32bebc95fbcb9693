#!/bin/bash
# Lean two-GPU check: the two-GPU tests (CLI threads + torchrun) and bench at N=2 with few steps.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu.py -m gpu -q -k "two_gpus" --timeout=200 -p no:cacheprovider > gpurun_out/pytest_multi.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_multi.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.log 2> gpurun_out/bench_n2.err
echo "exit $?" >> gpurun_out/bench_n2.err
tail -4 gpurun_out/pytest_multi.log; python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_n2.log").readline())
print({k: d[k] for k in ("value", "ms_per_step", "n_gpus")}, d["e2e"]["value"], d["roofline"]["frac"])
for b in d["bandb"]["datasets"]:
    print(b["dataset"], round(b["time_to_optimum_s"], 3), round(b["device_s_max_rank"], 3), b["matches_reference"])
PY
tail -3 gpurun_out/bench_n2.err
