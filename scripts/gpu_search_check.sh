#!/bin/bash
mkdir -p gpurun_out /tmp/xw
timeout 1500 python -m pytest tests/test_gpu.py -m gpu -q -k "search or cli or shard or job or smoke" --timeout=900 -p no:cacheprovider -s > gpurun_out/pytest_search.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_search.log
python - <<'PY'
import gzip, json
b = json.loads(gzip.open("tests/golden/alignments.json.gz").read())
for n in ("h1.phy", "h4.phy", "mt-10.phy", "mh1.phy", "Metazoan.phy", "mh3.phy"):
    open("/tmp/xw/" + n, "wb").write(b[n].encode("latin-1"))
PY
cd /tmp/xw
for d in h1 h4 mh1 Metazoan mt-10 mh3 h1; do
  $GRAFT_REPO_ROOT/fastdnamp_b200 --tbr=N --seed=1 --displaynewbounds=n --displaynewtrees=n --updatesecs=0 -o /tmp/xw/out_$d.nex /tmp/xw/$d.phy 2>&1 | grep -E "executed|built|Site merges|launches|branch and bound search:" | sed "s/^/$d: /"
done > $GRAFT_REPO_ROOT/gpurun_out/cli_stats2.log
cd $GRAFT_REPO_ROOT
grep -E "s device|passed|failed|Error|error" gpurun_out/pytest_search.log | tail -45; grep -E "launches|search:" gpurun_out/cli_stats2.log
