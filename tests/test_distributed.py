"""The multi-rank search driver (xmp_b200/distributed.py) over gloo on CPU, world sizes 2 and 3.

Two transports:
  * "peer" -- the product path: the exchange protocol (xmp_b200/csrc/exchange_proto.h: bound words, steal requests and
    replies, the busy counter) between PROCESSES.  Here the words live in a /dev/shm file instead of peer-mapped device
    memory: the engine is xmp_b200.cuda.Context bound to the CPU test double (tests/stub), which compiles the
    product's protocol source over host atomics;
  * "collective" -- the lock-step fallback (all-gather rounds), with the Python oracle stand-in as the engine.
The invariants are the ones the reference established by model checking its MPI protocol (spin/fastdnamp.mpiprom):
every job is executed exactly once, the final bound reaches every rank before trees are collected, no rank hangs."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.conftest import GOLDEN, ROOT, golden_trees
from tests.newick import canonical_newick


def _worker(rank, world, port, name, alignment_path, out_dir, skew, stub, fail_connect=None):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import json
    import pathlib
    from tests.test_host import prepare
    from xmp_b200 import distributed as xd
    goldens = json.load(open(os.path.join(GOLDEN, "datasets.json")))
    a, p = prepare(pathlib.Path(alignment_path).parent, goldens, name)
    log = []
    if stub:
        # the product's Context class over the test double (never done by the product: tests swap the loaded library)
        import ctypes
        import xmp_b200.cuda as xc
        os.environ["XMP_STUB_SHM"] = os.path.join("/dev/shm", "xmp_stub_%d" % port)
        os.environ["XMP_STUB_DEVICES"] = str(world)
        if skew:
            os.environ["XMP_STUB_SKEW"] = "1"
        if fail_connect is not None:
            os.environ["XMP_STUB_FAIL_CONNECT"] = str(fail_connect)
        xc._LIB = xc.bind(ctypes.CDLL(os.path.join(stub, "libxmp_b200.so")), partial=True)
        eng = xc.Context(rank)
        eng.load_problem(p)
        score, total, paths, info = xd.sharded_search(eng, p.num_taxa, "cpu", begin={"bound": p.bound})
        nodes = eng.search_stats()["nodes"]
        assert info["transport"] == ("peer" if fail_connect is None else "collective")
    else:
        from tests.oracle_engine import OracleEngine
        eng = OracleEngine(p)
        begin = {"bound": p.bound, "split_level": 6}
        if skew:
            # all the work starts on rank 0: the others must obtain everything by stealing
            begin.update({"shard_index": 0, "shard_count": 1} if rank == 0 else {"shard_index": 10 ** 9, "shard_count": 10 ** 9 + 1})
        score, total, paths, info = xd.sharded_search(eng, p.num_taxa, "cpu", begin=begin, slice_batches=3, steal_nodes=4,
                                                      on_round=lambda r, c, b: log.append((r, c, b)))
        nodes = eng.nodes
        assert info["transport"] == "collective"
    allpaths = xd.gather_paths(paths, p.num_taxa, "cpu")
    trees = p.newick_list(allpaths)
    np.save(os.path.join(out_dir, "r%d.npy" % rank), np.array([score, total, len(paths), info["rounds"], nodes]))
    if rank == 0:
        open(os.path.join(out_dir, "trees.txt"), "w").write("\n".join(trees))
        json.dump(log, open(os.path.join(out_dir, "log.json"), "w"))
    xd.disconnect_exchange(eng)
    if stub:
        eng.close()
        if rank == 0 and os.path.exists(os.environ["XMP_STUB_SHM"]):
            os.unlink(os.environ["XMP_STUB_SHM"])
    dist.destroy_process_group()


@pytest.fixture(scope="module")
def stub_dir(tmp_path_factory):
    from tests.test_cli import build_stub
    return str(build_stub(tmp_path_factory.mktemp("stub_dist")))


@pytest.mark.parametrize("transport", ["peer", "collective"])
@pytest.mark.parametrize("world,name,skew", [(2, "mh6", False), (3, "53humans.16", False), (2, "53humans.16", True), (3, "quick2", True)])
def test_sharded_search_over_gloo(tmp_path, alignment_dir, goldens, stub_dir, world, name, skew, transport):
    g = goldens[name]
    port = 29500 + (os.getpid() * 7 + world * 3 + skew + 5 * (transport == "peer")) % 2000
    mp.spawn(_worker, args=(world, port, name, str(alignment_dir / g["alignment"]), str(tmp_path), skew, stub_dir if transport == "peer" else None),
             nprocs=world, join=True)
    stats = [np.load(tmp_path / ("r%d.npy" % r)) for r in range(world)]
    # every rank ends with the same optimal length and the same total tree count
    assert all(int(s[0]) == g["score"] for s in stats)
    assert all(int(s[1]) == g["num_trees"] for s in stats)
    # every optimal tree found exactly once across ranks
    assert sum(int(s[2]) for s in stats) == g["num_trees"]
    trees = open(tmp_path / "trees.txt").read().split()
    assert sorted(canonical_newick(t) for t in trees) == golden_trees(name)
    if skew and name != "quick2":
        # the ranks that started empty did real work, i.e. jobs moved (quick2 is over before anybody can ask)
        assert all(int(s[4]) > 0 for s in stats[1:]), [int(s[4]) for s in stats]


def test_exchange_protocol_invariants_under_random_timing(tmp_path):
    """tests/proto_check.cpp: the product's protocol source (xmp_b200/csrc/exchange_proto.h) with 2 to 8 ranks as threads over
    host atomics, synthetic job trees and random pauses -- every job executed exactly once, idle ranks never give work away, the
    busy counter ends at zero, every rank stops with the finished flag and without work, all bound words agree at the end; once
    more under ThreadSanitizer where its runtime is available."""
    import subprocess
    exe = str(tmp_path / "proto_check")
    src = os.path.join(ROOT, "tests", "proto_check.cpp")
    subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", "-o", exe, src], check=True)
    for ranks, rounds in ((2, 20), (3, 20), (5, 12), (8, 12)):
        r = subprocess.run([exe, str(ranks), str(rounds), str(ranks * 7)], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0 and r.stdout.startswith("OK "), r.stdout + r.stderr
    tsan = str(tmp_path / "proto_check_tsan")
    c = subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-pthread", "-fsanitize=thread", "-o", tsan, src], capture_output=True, text=True)
    if c.returncode == 0:
        r = subprocess.run([tsan, "4", "4", "11"], capture_output=True, text=True, timeout=600)
        if "unexpected memory mapping" not in r.stderr:
            assert r.returncode == 0 and "ThreadSanitizer" not in r.stderr and r.stdout.startswith("OK "), r.stdout + r.stderr[-2000:]


def test_driver_falls_back_together_when_one_rank_cannot_map_its_peers(tmp_path, alignment_dir, goldens, stub_dir):
    """One rank's xmp_cuda_exchange_connect fails (a node whose GPUs are not peers): every rank must end up disconnected and
    on the lock-step transport -- agreed through one all-reduce -- and the search still returns the reference's set."""
    name, world = "53humans.16", 3
    g = goldens[name]
    port = 29500 + (os.getpid() * 7 + 977) % 2000
    mp.spawn(_worker, args=(world, port, name, str(alignment_dir / g["alignment"]), str(tmp_path), False, stub_dir, 1), nprocs=world, join=True)
    stats = [np.load(tmp_path / ("r%d.npy" % r)) for r in range(world)]
    assert all(int(s[0]) == g["score"] and int(s[1]) == g["num_trees"] for s in stats)
    assert sum(int(s[2]) for s in stats) == g["num_trees"]
    trees = open(tmp_path / "trees.txt").read().split()
    assert sorted(canonical_newick(t) for t in trees) == golden_trees(name)
