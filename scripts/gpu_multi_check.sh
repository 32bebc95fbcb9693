#!/bin/bash
# One multi-GPU call: the two-GPU parity tests, the exact-search leg of bench.py under torchrun, and the command line
# with --gpus=N on the seconds-long search.  N = number of GPUs of the box (argument 1).
N=${1:-2}
ROOT=${GRAFT_REPO_ROOT:-$(cd "$(dirname "$0")/.." && pwd)}
OUT=$ROOT/gpurun_out
mkdir -p $OUT
cd $ROOT
nvidia-smi topo -m > $OUT/topo.txt 2>&1
timeout 600 python -m pytest tests/test_gpu.py -m gpu -x -q -k "two_gpus or deep_split or export_import" > $OUT/r2_pytest_multi.log 2>&1
tail -3 $OUT/r2_pytest_multi.log
XMP_REQUIRE_PEER=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
	tests/multi_gpu_check.py mh6 h3 mh1 h4 > $OUT/r2_multi_check_n$N.log 2>&1
echo "multi_gpu_check rc=$?"; grep -c " OK " $OUT/r2_multi_check_n$N.log; grep "rank 0" $OUT/r2_multi_check_n$N.log | tail -5; tail -3 $OUT/r2_multi_check_n$N.log
DS=${DATASETS:-h4,mh3,mh1,Metazoan,its36,its36.Bd}
timeout 600 python bench.py --bandb-only $DS > $OUT/r2_bandb_n1.json 2> $OUT/r2_bandb_n1.err
echo "bandb n1 rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 \
	bench.py --gpus $N --bandb-only $DS > $OUT/r2_bandb_n$N.json 2> $OUT/r2_bandb_n$N.err
echo "bandb n$N rc=$?"; tail -3 $OUT/r2_bandb_n$N.err
python - <<PY
import json
for f in ("$OUT/r2_bandb_n1.json", "$OUT/r2_bandb_n$N.json"):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    for r in d["datasets"]:
        print(d["n_gpus"], r["dataset"], "ok" if r["matches_reference"] else "MISMATCH", "%.3f s" % r["time_to_optimum_s"], "busy %.2f" % r["busy_fraction"],
              "nodes %d" % r["nodes"], r["transport"], "steals %d" % r["steals"], "eff %.2f infl %.3f solo %.3f" % (r.get("efficiency", 0), r.get("node_inflation", 0), r.get("one_gpu_same_run_s", 0)))
PY
