#!/bin/bash
# ncu launch list of the search of one bundled dataset: bash scripts/gpu_launchlist.sh mh3 [max launches]
mkdir -p gpurun_out /tmp/xw
DS=${1:-mh3}
python - <<PY
import gzip, json
b = json.loads(gzip.open("tests/golden/alignments.json.gz").read())
open("/tmp/xw/$DS.phy", "wb").write(b["$DS.phy"].encode("latin-1"))
PY
cd /tmp/xw
CMD="$GRAFT_REPO_ROOT/fastdnamp_b200 --tbr=N --seed=1 --displaynewbounds=n --displaynewtrees=n --updatesecs=0 -o /tmp/xw/out.nex /tmp/xw/$DS.phy"
$CMD > $GRAFT_REPO_ROOT/gpurun_out/${DS}_plain.log 2>&1 &&
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c ${2:-6000} --csv --log-file $GRAFT_REPO_ROOT/gpurun_out/search_launches_$DS.csv $CMD > $GRAFT_REPO_ROOT/gpurun_out/ncu_search_$DS.log 2>&1
grep -E "launches|search:" $GRAFT_REPO_ROOT/gpurun_out/${DS}_plain.log
