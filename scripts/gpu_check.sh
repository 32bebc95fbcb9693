#!/bin/bash
# Full GPU validation: parity tests, smoke, bench (ours + reference arm), then -- with "ncu" -- the launch
# list and full captures of the top kernels.  Everything lands in gpurun_out/.
mkdir -p gpurun_out /tmp/xw
nvidia-smi > gpurun_out/nvidia-smi.txt 2>&1
if [ "$2" != "nopytest" ]; then
timeout 2400 python -m pytest tests -m gpu -q --timeout=900 -p no:cacheprovider -s > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
fi
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench exit $?" >> gpurun_out/bench.err
if [ "$2" != "nopytest" ]; then
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err
fi
if [ "$1" = "ncu" ] && [ "$2" != "nopytest" ]; then
  timeout 300 python bench.py --steps 3 --warmup 3 --kernel-only > gpurun_out/plain_ko.log 2>&1 &&
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 3 --warmup 3 --kernel-only > gpurun_out/ncu_launches.log 2>&1
  timeout 300 python bench.py --steps 3 --warmup 3 --kernel-only > gpurun_out/plain_ko2.log 2>&1 &&
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_score_wide -s 3 -c 2 -f -o gpurun_out/prof_score_wide \
      python bench.py --steps 3 --warmup 3 --kernel-only > gpurun_out/ncu_full.log 2>&1
fi
if [ "$1" = "ncu" ]; then
  python - <<'PY'
import gzip, json
b = json.loads(gzip.open("tests/golden/alignments.json.gz").read())
open("/tmp/xw/h4.phy", "wb").write(b["h4.phy"].encode("latin-1"))
PY
  CMD="$GRAFT_REPO_ROOT/fastdnamp_b200 --tbr=N --seed=1 --displaynewbounds=n --displaynewtrees=n --updatesecs=0 -o /tmp/xw/out.nex /tmp/xw/h4.phy"
  (cd /tmp/xw && $CMD > $GRAFT_REPO_ROOT/gpurun_out/h4_plain.log 2>&1 &&
   timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $GRAFT_REPO_ROOT/gpurun_out/search_launches_h4.csv $CMD > $GRAFT_REPO_ROOT/gpurun_out/ncu_search_h4.log 2>&1)
  (cd /tmp/xw && $CMD > /dev/null 2>&1 &&
   timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_expand_last_fused -s 12 -c 1 -f -o $GRAFT_REPO_ROOT/gpurun_out/prof_fused_h4 $CMD > $GRAFT_REPO_ROOT/gpurun_out/ncu_fused_h4.log 2>&1)
  # mh3 (20 taxa x 22 patterns, one block per set): the narrow, deep case -- launch list and the two kernels that share its time
  python - <<'PY'
import gzip, json
b = json.loads(gzip.open("tests/golden/alignments.json.gz").read())
open("/tmp/xw/mh3.phy", "wb").write(b["mh3.phy"].encode("latin-1"))
PY
  CMD="$GRAFT_REPO_ROOT/fastdnamp_b200 --tbr=N --seed=1 --displaynewbounds=n --displaynewtrees=n --updatesecs=0 -o /tmp/xw/out3.nex /tmp/xw/mh3.phy"
  (cd /tmp/xw && $CMD > $GRAFT_REPO_ROOT/gpurun_out/mh3_plain.log 2>&1 &&
   timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $GRAFT_REPO_ROOT/gpurun_out/search_launches_mh3.csv $CMD > $GRAFT_REPO_ROOT/gpurun_out/ncu_search_mh3.log 2>&1)
  (cd /tmp/xw && timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_select -s 300 -c 1 -f -o $GRAFT_REPO_ROOT/gpurun_out/prof_select_mh3 $CMD > $GRAFT_REPO_ROOT/gpurun_out/ncu_select_mh3.log 2>&1)
  (cd /tmp/xw && timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_build_children -s 300 -c 1 -f -o $GRAFT_REPO_ROOT/gpurun_out/prof_children_mh3 $CMD > $GRAFT_REPO_ROOT/gpurun_out/ncu_children_mh3.log 2>&1)
fi
tail -6 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log; cat gpurun_out/bench.log | cut -c1-600; tail -3 gpurun_out/bench.err; cat gpurun_out/bench_ref.log | cut -c1-300
