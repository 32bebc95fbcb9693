#!/bin/bash
# Launch lists of the eager search on mh3 and its36.
mkdir -p gpurun_out /tmp/xw
python - <<'PY'
import gzip, json
b = json.loads(gzip.open("tests/golden/alignments.json.gz").read())
for n in ("mh3", "its36"):
    open("/tmp/xw/%s.phy" % n, "wb").write(b[n + ".phy"].encode("latin-1"))
PY
cd /tmp/xw
F="--tbr=N --seed=1 --displaynewbounds=n --displaynewtrees=n --updatesecs=0"
$GRAFT_REPO_ROOT/fastdnamp_b200 $F -o o1.nex mh3.phy > $GRAFT_REPO_ROOT/gpurun_out/mh3_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $GRAFT_REPO_ROOT/gpurun_out/search_launches_mh3.csv $GRAFT_REPO_ROOT/fastdnamp_b200 $F -o o1.nex mh3.phy > $GRAFT_REPO_ROOT/gpurun_out/ncu_search_mh3.log 2>&1
$GRAFT_REPO_ROOT/fastdnamp_b200 $F -Bp -b 233 -m 100000 -o o2.nex its36.phy > $GRAFT_REPO_ROOT/gpurun_out/its36_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file $GRAFT_REPO_ROOT/gpurun_out/search_launches_its36.csv $GRAFT_REPO_ROOT/fastdnamp_b200 $F -Bp -b 233 -m 100000 -o o2.nex its36.phy > $GRAFT_REPO_ROOT/gpurun_out/ncu_search_its36.log 2>&1
cd $GRAFT_REPO_ROOT
python scripts/launch_summary.py gpurun_out/search_launches_mh3.csv | head -6
python scripts/launch_summary.py gpurun_out/search_launches_its36.csv | head -6
grep "Time to perform" gpurun_out/mh3_plain.log gpurun_out/its36_plain.log
