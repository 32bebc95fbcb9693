#!/bin/bash
# One short GPU call without Python: (1) the m-sweep of the synthetic microbenchmark (tests/sweep_m_check),
# (2) A/B of walk-kernel variants through the command line, each run checked against the reference's output
# file (sha256 from tests/golden/datasets.json, prepared in variants/data by scripts/build_variant.sh).
ROOT=${GRAFT_REPO_ROOT:-$(cd "$(dirname "$0")/.." && pwd)}
OUT=$ROOT/gpurun_out
mkdir -p $OUT /tmp/xw
cd /tmp/xw
T0=$EPOCHREALTIME
timeout 22 $ROOT/tests/sweep_m_check 6554.2 ${SWEEP_MS:-4 8 16 63} > $OUT/sweep_m.jsonl 2> $OUT/sweep_m.err
echo "exit $? after $(awk "BEGIN{print $EPOCHREALTIME - $T0}") s" >> $OUT/sweep_m.err
FLAGS="--tbr=N --seed=1 --displaynewbounds=n --displaynewtrees=n --updatesecs=0"
: > $OUT/ab.log
run() {   # variant name extra-flags...
	local v=$1 n=$2; shift 2
	local lib=$ROOT/xmp_b200
	[ "$v" != base ] && lib=$ROOT/variants/$v
	LD_LIBRARY_PATH=$lib timeout 12 $ROOT/fastdnamp_b200 $FLAGS "$@" -o /tmp/xw/out.nex $ROOT/variants/data/$n.phy > /dev/null 2> /tmp/xw/err.txt
	local rc=$?
	local sha=$(sha256sum /tmp/xw/out.nex | cut -d' ' -f1)
	local want=$(grep "^$n " $ROOT/variants/data/expected_sha256.txt | cut -d' ' -f2)
	local ok=MISMATCH; [ "$sha" = "$want" ] && ok=OK
	echo "$v $n rc=$rc $ok $(grep 'Site merges' /tmp/xw/err.txt | tr -s ' ') | $(grep 'BranchAndBound() executed' /tmp/xw/err.txt | tr -s ' ') | $(grep 'launches' /tmp/xw/err.txt | tr -s ' ') | t=$(awk "BEGIN{print $EPOCHREALTIME - $T0}")" >> $OUT/ab.log
	rm -f /tmp/xw/out.nex
}
V="base ${VARIANTS:-six ul ul6}"    # build them first: scripts/build_variant.sh six -DXMP_EAGER_MIN_CTAS=6 etc.
DEADLINE=${DEADLINE:-46}
late() { awk "BEGIN{exit !($EPOCHREALTIME - $T0 > $DEADLINE)}"; }
its36() { run $1 its36 -Bf $ROOT/variants/data/its36_restBound.txt -b 233 -m 100000; }
for v in $V; do late || run $v mh3; done
for v in $V; do late || its36 $v; done
for v in $V; do late || run $v Metazoan; done
for v in $V; do late || run $v mh3; done
for v in $V; do late || its36 $v; done
for v in $V; do late || run $v h4; done
for v in $V; do late || run $v mh1; done
cat $OUT/sweep_m.jsonl; tail -2 $OUT/sweep_m.err; cat $OUT/ab.log
