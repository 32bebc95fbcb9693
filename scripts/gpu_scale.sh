#!/bin/bash
# Scaling run on one node: bench.py (ours and the reference arm) at N = 1, 2, 4, 8, launched the way the driver does.
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
rm -f gpurun_out/scale_status.txt
if [ $NG -ge 2 ]; then
  timeout 900 python -m pytest tests/test_gpu.py -m gpu -q -k "two_gpus" --timeout=600 -p no:cacheprovider > gpurun_out/pytest_multi.log 2>&1
  echo "two-GPU tests exit $?" >> gpurun_out/scale_status.txt
  tail -3 gpurun_out/pytest_multi.log
fi
for N in 1 2 4 8; do
  [ $N -gt $NG ] && break
  if [ $N -eq 1 ]; then
    timeout 900 python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) \
      bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
  fi
  echo "N=$N exit $?" >> gpurun_out/scale_status.txt
done
timeout 600 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/scale_ref.json 2> gpurun_out/scale_ref.err
cat gpurun_out/scale_status.txt
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/scale_n*.json")):
    for l in open(f):
        if l.startswith("{"):
            d = json.loads(l)
            print(f, "value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"], 3), d["clocks"])
            for b in d.get("bandb", {}).get("datasets", []):
                print("   ", b["dataset"], "t2o %.3f s" % b["time_to_optimum_s"], "dev %.3f" % b["device_s_max_rank"], "nodes/s %.0fM" % (b["nodes_per_s"] / 1e6), b["matches_reference"], "cpu %.1f s" % b["reference_cpu_bandb_s"])
            if "int_pipe_peak" in d: print("   int", {k: round(v) for k, v in d["int_pipe_peak"].items() if isinstance(v, float)})
PY
