#!/bin/bash
# Scheduler knob sweep through the command line: minimum batch and frontier budget, on the launch-bound searches.
ROOT=${GRAFT_REPO_ROOT:-$(cd "$(dirname "$0")/.." && pwd)}
OUT=$ROOT/gpurun_out
mkdir -p $OUT /tmp/probe
cd /tmp/probe
python - <<PY
import gzip, json
b = json.loads(gzip.open("$ROOT/tests/golden/alignments.json.gz").read())
for k in ("its36.phy", "mh3.phy", "Metazoan.phy", "mh1.phy"):
    open("/tmp/probe/" + k, "w", encoding="latin-1").write(b[k])
PY
FLAGS="--tbr=N --seed=1 --displaynewbounds=n --displaynewtrees=n --updatesecs=0"
: > $OUT/knobs.log
run() {  # label dataset extra...
	local label=$1 ds=$2; shift 2
	$ROOT/fastdnamp_b200 $FLAGS "$@" -o /tmp/probe/o.nex /tmp/probe/$ds.phy > /dev/null 2> /tmp/probe/err.txt
	echo "$label $ds $* | $(grep -E 'Time to perform|launches' /tmp/probe/err.txt | tr -s ' ' | tr '\n' '|') sha=$(sha256sum /tmp/probe/o.nex | cut -c1-12)" >> $OUT/knobs.log
}
for ds in mh3 Metazoan its36; do
	EXTRA=""; [ $ds = its36 ] && EXTRA="-Bp -b 233 -m 100000"
	for mb in 8192 32768 131072; do
		for bud in 49152 98304; do
			XMP_MIN_BATCH=$mb run "minbatch=$mb budget=${bud}MiB" $ds --membudget=$bud $EXTRA
		done
	done
done
cat $OUT/knobs.log
