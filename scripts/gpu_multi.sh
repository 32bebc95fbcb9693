#!/bin/bash
# Two-GPU run: sharded search (torchrun + CLI threads), bench at N=2.
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
timeout 900 python -m pytest tests/test_gpu.py -m gpu -q -k "two_gpus" --timeout=600 -p no:cacheprovider -s > gpurun_out/pytest_multi.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/multi_gpu_check.py h4 mh3 Metazoan > gpurun_out/multi_check.log 2>&1
echo "exit $?" >> gpurun_out/multi_check.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.log 2> gpurun_out/bench_n2.err
echo "exit $?" >> gpurun_out/bench_n2.err
tail -8 gpurun_out/pytest_multi.log; cat gpurun_out/multi_check.log; cat gpurun_out/bench_n2.log; tail -5 gpurun_out/bench_n2.err
