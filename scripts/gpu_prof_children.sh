#!/bin/bash
# Launch list of the search of one dataset + one full capture of k_build_children: bash scripts/gpu_prof_children.sh mh3 [skip]
mkdir -p gpurun_out /tmp/xw
DS=${1:-mh3}
python - <<PY
import gzip, json
b = json.loads(gzip.open("tests/golden/alignments.json.gz").read())
open("/tmp/xw/$DS.phy", "wb").write(b["$DS.phy"].encode("latin-1"))
PY
cd /tmp/xw
OUT=$GRAFT_REPO_ROOT/gpurun_out
CMD="$GRAFT_REPO_ROOT/fastdnamp_b200 --tbr=N --seed=1 --displaynewbounds=n --displaynewtrees=n --updatesecs=0 -o /tmp/xw/out.nex /tmp/xw/$DS.phy"
$CMD > $OUT/${DS}_plain.log 2>&1 &&
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file $OUT/search_launches_$DS.csv $CMD > $OUT/ncu_search_$DS.log 2>&1
$CMD > /dev/null 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_build_children -s ${2:-150} -c 1 -f -o $OUT/prof_children_$DS $CMD > $OUT/ncu_children_$DS.log 2>&1
tail -2 $OUT/ncu_children_$DS.log
grep -E "launches|search:" $OUT/${DS}_plain.log
