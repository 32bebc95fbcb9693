"""The multi-rank search protocol (xmp_b200/distributed.py) over gloo on CPU, world sizes 2 and 3:
bound exchange by all-reduce(min), collective work stealing with job descriptors, termination, and the
final gather.  The compute engine is the oracle-backed stand-in; the invariants are the ones the
reference established by model checking its MPI protocol (spin/fastdnamp.mpiprom): every job is
executed exactly once, the final bound reaches every rank before trees are collected, no rank hangs."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.conftest import GOLDEN, ROOT, golden_trees
from tests.newick import canonical_newick


def _worker(rank, world, port, name, alignment_path, out_dir, skew):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import json
    from tests.oracle_engine import OracleEngine
    from tests.test_host import prepare
    from xmp_b200 import distributed as xd
    goldens = json.load(open(os.path.join(GOLDEN, "datasets.json")))
    import pathlib
    a, p = prepare(pathlib.Path(alignment_path).parent, goldens, name)
    eng = OracleEngine(p)
    if skew:
        # all the work starts on rank 0: the others must obtain everything by stealing
        if rank == 0:
            eng.search_begin(p.bound, 0, 1, 6)
        else:
            eng.bound = p.bound
    else:
        eng.search_begin(p.bound, rank, world, 6)
    log = []
    score, total, paths, rounds = xd.sharded_search(eng, p.num_taxa, "cpu", slice_batches=3, steal_nodes=4,
                                                    on_round=lambda r, c, b: log.append((r, c, b)))
    allpaths = xd.gather_paths(paths, p.num_taxa, "cpu")
    trees = p.newick_list(allpaths)
    np.save(os.path.join(out_dir, "r%d.npy" % rank), np.array([score, total, len(paths), rounds, eng.nodes]))
    if rank == 0:
        open(os.path.join(out_dir, "trees.txt"), "w").write("\n".join(trees))
        json.dump(log, open(os.path.join(out_dir, "log.json"), "w"))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,name,skew", [(2, "mh6", False), (3, "53humans.16", False), (2, "53humans.16", True), (3, "quick2", True)])
def test_sharded_search_over_gloo(tmp_path, alignment_dir, goldens, world, name, skew):
    g = goldens[name]
    port = 29500 + (os.getpid() * 7 + world * 3 + skew) % 2000
    mp.spawn(_worker, args=(world, port, name, str(alignment_dir / g["alignment"]), str(tmp_path), skew), nprocs=world, join=True)
    stats = [np.load(tmp_path / ("r%d.npy" % r)) for r in range(world)]
    # every rank ends with the same optimal length and the same total tree count
    assert all(int(s[0]) == g["score"] for s in stats)
    assert all(int(s[1]) == g["num_trees"] for s in stats)
    # every optimal tree found exactly once across ranks
    assert sum(int(s[2]) for s in stats) == g["num_trees"]
    trees = open(tmp_path / "trees.txt").read().split()
    assert sorted(canonical_newick(t) for t in trees) == golden_trees(name)
    if skew:
        # the ranks that started empty did real work, i.e. jobs moved
        assert all(int(s[4]) > 0 for s in stats[1:]), [int(s[4]) for s in stats]
