#!/bin/bash
# ncu launch list + full captures of the search kernels on h1 (12 taxa x 64 patterns, W = 2).
mkdir -p gpurun_out /tmp/xw
python - <<'PY'
import gzip, json
b = json.loads(gzip.open("tests/golden/alignments.json.gz").read())
for n in ("h1.phy", "h4.phy", "mt-10.phy", "mh1.phy", "Metazoan.phy"):
    open("/tmp/xw/" + n, "wb").write(b[n].encode("latin-1"))
PY
cd /tmp/xw
DS=${1:-h1}
CMD="$GRAFT_REPO_ROOT/fastdnamp_b200 --tbr=N --seed=1 --displaynewbounds=n --displaynewtrees=n --updatesecs=0 -o /tmp/xw/out.nex /tmp/xw/$DS.phy"
$CMD > $GRAFT_REPO_ROOT/gpurun_out/${DS}_plain.log 2>&1 &&
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $GRAFT_REPO_ROOT/gpurun_out/search_launches_$DS.csv $CMD > $GRAFT_REPO_ROOT/gpurun_out/ncu_search.log 2>&1
$CMD > /dev/null 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_expand_last_fused" -s ${2:-10} -c 2 -f -o $GRAFT_REPO_ROOT/gpurun_out/prof_fused_$DS $CMD > $GRAFT_REPO_ROOT/gpurun_out/ncu_search_full.log 2>&1
tail -3 $GRAFT_REPO_ROOT/gpurun_out/ncu_search_full.log
