#!/bin/bash
# One 8-GPU call: parity at N = 8, the full bench line under torchrun (the driver's own command), the command line with
# --gpus=8 on the seconds-long search.
N=${1:-8}
ROOT=${GRAFT_REPO_ROOT:-$(cd "$(dirname "$0")/.." && pwd)}
OUT=$ROOT/gpurun_out
mkdir -p $OUT /tmp/probe
cd $ROOT
XMP_REQUIRE_PEER=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
	tests/multi_gpu_check.py mh6 h3 mh1 h4 > $OUT/r2_multi_check_n$N.log 2>&1
echo "multi_gpu_check rc=$? OK lines: $(grep -c ' OK ' $OUT/r2_multi_check_n$N.log) MISMATCH: $(grep -c MISMATCH $OUT/r2_multi_check_n$N.log)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 \
	bench.py --gpus $N --steps 10 --warmup 3 > $OUT/r2_bench_n$N.json 2> $OUT/r2_bench_n$N.err
echo "bench n$N rc=$?"; tail -3 $OUT/r2_bench_n$N.err
python - <<PY
import json
try:
    d = json.loads([l for l in open("$OUT/r2_bench_n$N.json") if l.startswith("{")][-1])
    print("value", d["value"], "e2e", d["e2e"]["value"], "ms", d["ms_per_step"])
    for r in d["bandb"]["datasets"]:
        print(d["n_gpus"], r["dataset"], "ok" if r["matches_reference"] else "MISMATCH", "%.4f s" % r["time_to_optimum_s"], "busy %.2f" % r["busy_fraction"],
              "nodes %d" % r["nodes"], r["transport"], "steals %d" % r["steals"], "eff %.2f infl %.3f solo %.4f" % (r.get("efficiency", 0), r.get("node_inflation", 0), r.get("one_gpu_same_run_s", 0)),
              r.get("rank0_host_seconds"))
except Exception as e:
    print("bench output unreadable", e)
PY
python - <<PY
import gzip, json
b = json.loads(gzip.open("$ROOT/tests/golden/alignments.json.gz").read())
for k in ("its36.phy", "h4.phy", "mh3.phy"):
    open("/tmp/probe/" + k, "w", encoding="latin-1").write(b[k])
PY
cd /tmp/probe
for ds in h4 mh3 its36; do
	timeout 120 $ROOT/fastdnamp_b200 --gpus=$N --tbr=N --seed=1 --displaynewbounds=n --displaynewtrees=n --updatesecs=0 -o /tmp/probe/$ds.nex /tmp/probe/$ds.phy 2> /tmp/probe/$ds.err > /dev/null
	echo "cli --gpus=$N $ds rc=$? sha=$(sha256sum /tmp/probe/$ds.nex | cut -c1-16) $(grep -E 'Time to perform|Steals|BranchAndBound' /tmp/probe/$ds.err | tr -s ' ' | tr '\n' '|')"
done
