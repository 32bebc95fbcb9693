#!/bin/bash
# Full ncu captures of the search kernels in mid-search (one launch each), plus launch lists, through the command line.
ROOT=${GRAFT_REPO_ROOT:-$(cd "$(dirname "$0")/.." && pwd)}
OUT=$ROOT/gpurun_out
mkdir -p $OUT /tmp/probe
cd /tmp/probe
python - <<PY
import gzip, json
b = json.loads(gzip.open("$ROOT/tests/golden/alignments.json.gz").read())
for k in ("its36.phy", "h4.phy", "mh3.phy", "Metazoan.phy"):
    open("/tmp/probe/" + k, "w", encoding="latin-1").write(b[k])
PY
FLAGS="--tbr=N --seed=1 --displaynewbounds=n --displaynewtrees=n --updatesecs=0 -q"
cap() {  # tag kernel-regex skip dataset
	ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -f -o $OUT/r2_prof_$1 $ROOT/fastdnamp_b200 $FLAGS -o /tmp/probe/o.nex /tmp/probe/$4.phy > $OUT/r2_ncu_$1.log 2>&1
	tail -1 $OUT/r2_ncu_$1.log
}
cap children_mh3 k_build_children_eager 150 mh3
cap select_mh3 k_select 300 mh3
cap fused_h4 k_expand_last_fused 12 h4
cap children_its36 k_build_children_eager 3000 its36
for ds in mh3 h4; do
	ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/r2_launches_$ds.csv $ROOT/fastdnamp_b200 $FLAGS -o /tmp/probe/o.nex /tmp/probe/$ds.phy > /dev/null 2>&1
done
ls -la $OUT/r2_prof_*.ncu-rep
