#!/bin/bash
# Probe for search workloads that last seconds on one B200 (candidates for bench.py's scaling leg): runs the
# command line on bundled alignments with a time limit each and reports the search's own statistics.
ROOT=${GRAFT_REPO_ROOT:-$(cd "$(dirname "$0")/.." && pwd)}
OUT=$ROOT/gpurun_out
mkdir -p $OUT /tmp/probe
cd /tmp/probe
python - <<PY
import gzip, json
b = json.loads(gzip.open("$ROOT/tests/golden/alignments.json.gz").read())
for k, v in b.items():
    if "/" in k: continue
    open("/tmp/probe/" + k, "w", encoding="latin-1").write(v)
rows = b["53humans.32.phy"].splitlines()
for n in (29, 30, 31):
    open("/tmp/probe/53humans.%d.phy" % n, "w").write("%d 202\n" % n + "\n".join(rows[1:1 + n]) + "\n")
PY
FLAGS="--tbr=N --seed=1 --displaynewbounds=y --displaynewtrees=n --updatesecs=0"
: > $OUT/probe.log
run() {   # limit name flags...
	local lim=$1 n=$2; shift 2
	local t0=$EPOCHREALTIME
	timeout $lim $ROOT/fastdnamp_b200 $FLAGS "$@" -o /tmp/probe/out.nex /tmp/probe/$n.phy > /dev/null 2> /tmp/probe/err.txt
	local rc=$?
	echo "== $n $* rc=$rc wall=$(awk "BEGIN{print $EPOCHREALTIME - $t0}") trees_sha=$(sha256sum /tmp/probe/out.nex | cut -c1-16)" >> $OUT/probe.log
	grep -E "New bound|Optimal trees|BranchAndBound\(\)|Site merges|launches|Time to perform|Time to determine lower|on trees of" /tmp/probe/err.txt | tail -${TAILN:-60} >> $OUT/probe.log
	rm -f /tmp/probe/out.nex
}
for job in "$@"; do
	IFS=: read lim name flags <<< "$job"
	run $lim $name $flags
done
cat $OUT/probe.log
