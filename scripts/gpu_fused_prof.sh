#!/bin/bash
mkdir -p gpurun_out /tmp/xw
python - <<'PY'
import gzip, json
b = json.loads(gzip.open("tests/golden/alignments.json.gz").read())
open("/tmp/xw/h4.phy", "wb").write(b["h4.phy"].encode("latin-1"))
PY
CMD="$GRAFT_REPO_ROOT/fastdnamp_b200 --tbr=N --seed=1 --displaynewbounds=n --displaynewtrees=n --updatesecs=0 -o /tmp/xw/out.nex /tmp/xw/h4.phy"
cd /tmp/xw && $CMD > /dev/null 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_expand_last_fused -s 12 -c 1 -f -o $GRAFT_REPO_ROOT/gpurun_out/prof_fused_h4 $CMD > $GRAFT_REPO_ROOT/gpurun_out/ncu_fused_h4.log 2>&1
tail -2 $GRAFT_REPO_ROOT/gpurun_out/ncu_fused_h4.log
