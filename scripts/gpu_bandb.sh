#!/bin/bash
# bench.py's exact-search leg alone under torchrun on N GPUs (N = argument 1), summarised.
N=${1:-8}
DS=${2:-h4,h5,mh1,mh3,rbc14x759,Metazoan,its36,its36.Bd}
TAG=${3:-x}
ROOT=${GRAFT_REPO_ROOT:-$(cd "$(dirname "$0")/.." && pwd)}
OUT=$ROOT/gpurun_out
mkdir -p $OUT
cd $ROOT
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 \
	bench.py --gpus $N --bandb-only $DS > $OUT/r2_bandb_n${N}_$TAG.json 2> $OUT/r2_bandb_n${N}_$TAG.err
echo "bandb n$N rc=$?"; tail -2 $OUT/r2_bandb_n${N}_$TAG.err
python - <<PY
import json
d = json.loads([l for l in open("$OUT/r2_bandb_n${N}_$TAG.json") if l.startswith("{")][-1])
for r in d["datasets"]:
    h = r.get("rank0_host_seconds") or {}
    print(d["n_gpus"], r["dataset"], "ok" if r["matches_reference"] else "MISMATCH", "%.4f s" % r["time_to_optimum_s"], "busy %.2f" % r["busy_fraction"],
          "nodes %d" % r["nodes"], r["transport"], "steals %d" % r["steals"], "eff %.2f infl %.3f solo %.4f" % (r.get("efficiency", 0), r.get("node_inflation", 0), r.get("one_gpu_same_run_s", 0)),
          " ".join("%s=%.1fms" % (k, v * 1e3) for k, v in h.items()))
    pr = r.get("per_rank")
    if pr:
        print("     run_s", pr["run_s"], "wait_s", pr["steal_wait_s"]); print("     Mnodes", [x // 1000000 for x in pr["nodes"]], "recv", pr["steals_received"], "served", pr["steals_served"])
PY
