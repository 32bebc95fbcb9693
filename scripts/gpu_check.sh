#!/bin/bash
# GPU run: parity tests, smoke, bench, then the ncu launch list and one full capture of the top kernel.
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia-smi.txt 2>&1
timeout 2400 python -m pytest tests -m gpu -q --timeout=900 -p no:cacheprovider -s > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench exit $?" >> gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err
if [ "$1" = "ncu" ]; then
  timeout 300 python bench.py --steps 3 --warmup 3 --kernel-only > gpurun_out/plain_ko.log 2>&1 &&
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 3 --warmup 3 --kernel-only > gpurun_out/ncu_launches.log 2>&1
  timeout 300 python bench.py --steps 3 --warmup 3 --kernel-only > gpurun_out/plain_ko2.log 2>&1 &&
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_score_wide -s 3 -c 2 -f -o gpurun_out/prof_score_wide \
      python bench.py --steps 3 --warmup 3 --kernel-only > gpurun_out/ncu_full.log 2>&1
fi
tail -15 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log; cat gpurun_out/bench.log; tail -5 gpurun_out/bench.err; cat gpurun_out/bench_ref.log
