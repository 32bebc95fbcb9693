#!/bin/bash
# Builds libxmp_b200.so with extra compiler flags into variants/<name>/ (git-ignored; travels with gpurun), for A/B
# runs through scripts/gpu_oneshot.sh (VARIANTS="<name> ...").  Also prepares variants/data (alignments, expected
# output hashes, the -Bp restBound file of its36) on first use.
#   scripts/build_variant.sh six -DXMP_EAGER_MIN_CTAS=6
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
name=$1; shift
mkdir -p $ROOT/variants/$name
cd $ROOT
/usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-Wall,-Wextra -Xptxas -v --use_fast_math "$@" \
	-shared -cudart static -o variants/$name/libxmp_b200.so xmp_b200/csrc/api.cu xmp_b200/csrc/kernels.cu xmp_b200/csrc/search.cu xmp_b200/csrc/microbench.cu \
	2> variants/$name/ptxas.log || { cat variants/$name/ptxas.log; exit 1; }
cuobjdump -res-usage variants/$name/libxmp_b200.so 2>/dev/null | awk '/Function/ {f=$2} /REG:/ && f ~ /k_build_children|k_expand_last_fusedILi1E|k_selectILi8/ {sub(/_ZN3xmp[0-9]*/, "", f); print substr(f, 1, 28), $1}' 
if [ ! -f variants/data/expected_sha256.txt ]; then
	mkdir -p variants/data
	python - <<'PY'
import gzip, json
b = json.loads(gzip.open("tests/golden/alignments.json.gz").read())
g = json.load(open("tests/golden/datasets.json"))
exp = []
for n in ("mh3", "its36", "Metazoan", "h4", "mh1", "rbc14x759", "mt-10", "h1", "53humans.28"):
    if n in g and "output_sha256" in g[n] and not any(f == "-m" for f in g[n]["flags"][:0]):
        open("variants/data/%s.phy" % n, "wb").write(b[g[n]["alignment"]].encode("latin-1"))
        exp.append("%s %s" % (n, g[n]["output_sha256"]))
open("variants/data/expected_sha256.txt", "w").write("\n".join(exp) + "\n")
PY
	(cd variants/data && ../../fastdnamp_b200 --tbr=N --seed=1 --updatesecs=0 --onlycomputelowerbound=Y -Bp -b 233 -m 100000 its36.phy > /dev/null 2>&1 && mv restBound.txt its36_restBound.txt)
fi
