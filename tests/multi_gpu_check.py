"""Run under torchrun on N GPUs: exact search of the given datasets sharded over the ranks through
xmp_b200.distributed (nccl), checked against the goldens.  Used by tests/test_gpu.py and by hand:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/multi_gpu_check.py mh6 h3 mh1
"""
import gzip
import hashlib
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from tests.newick import canonical_newick  # noqa: E402
from xmp_b200 import cuda as xc  # noqa: E402
from xmp_b200 import distributed as xd  # noqa: E402
from xmp_b200 import hostlib as xh  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    gold = os.path.join(ROOT, "tests", "golden")
    goldens = json.load(open(os.path.join(gold, "datasets.json")))
    bundle = json.loads(gzip.open(os.path.join(gold, "alignments.json.gz")).read())
    ok = True
    for name in sys.argv[1:]:
        g = goldens[name]
        with tempfile.NamedTemporaryFile("wb", suffix=".phy", delete=False) as f:
            f.write(bundle[g["alignment"]].encode("latin-1"))
        a = xh.Alignment(f.name)
        os.unlink(f.name)
        p = xh.Problem(a)
        ctx = xc.Context(local)
        ctx.load_problem(p)
        connected = xd.connect_exchange(ctx, dev)
        dist.barrier()
        t0 = time.perf_counter()
        score, total, paths, info = xd.sharded_search(ctx, p.num_taxa, dev, begin={"bound": p.bound}, connected=connected)
        wall = time.perf_counter() - t0
        allpaths = xd.gather_paths(paths, p.num_taxa, dev)
        st = ctx.search_stats()
        xd.disconnect_exchange(ctx)
        ctx.close()
        canon = sorted(canonical_newick(t) for t in p.newick_list(allpaths))
        blob = ("\n".join(canon) + "\n").encode() if canon else b""
        good = (score == g["score"] and total == g["num_trees"] == len(allpaths)
                and hashlib.sha256(blob).hexdigest() == g["canonical_sha256"])
        if os.environ.get("XMP_REQUIRE_PEER") and info["transport"] != "peer":
            good = False
        ok &= good
        print("rank %d %s: %s score %d trees %d (mine %d) %.3f s, transport %s, %d waits, %d nodes here, steals served %d received %d" %
              (rank, name, "OK" if good else "MISMATCH", score, total, len(paths), wall, info["transport"], info["rounds"], st["nodes"],
               st["steals_served"], st["steals_received"]), flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
