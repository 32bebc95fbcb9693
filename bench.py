#!/usr/bin/env python3
"""bench.py -- BASELINE.json config 2: batched Fitch insertion scoring, all edges of 4096 partial trees
on a synthetic random 4-state alignment of 64 taxa x 2^20 sites, on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one pass of the hot path over the whole batch: taxon m is scored against every edge
MIDPOINT set of every partial tree (53,248 (tree, edge) pairs of 2^20 sites on each GPU), pruned against
the bound and the survivors ballot-compacted.  All weights are 1, so the kernel under test is the
weight-1 scorer (reference FitchScoreWeight1, src/fitchtemplate_b2_w8_fastc64.inc:43-80).

  value     Gsite-merges/s, whole job (all ranks), MIDPOINT arrays resident in HBM (26 GiB per GPU, far
            larger than the 126 MB L2, so every step streams from HBM)
  e2e       the same count of insertion-scoring site-merges through the host-buffer C ABI
            (xmp_cuda_load + xmp_cuda_score_insertions_host): alignment and topologies come from
            pinned host memory every step, state sets are built on the device, costs are read back
  roofline  algorithmic bytes (0.5 B per site-merge) / step time, against the measured HBM copy peak
  cpu_baseline / --impl reference
            the reference's own fastc64 Fitch code (oracle/_ref, compiled from its sources; the
            oracle port if that is absent) doing the same host-inputs-to-costs work on the host cores

N > 1 (torchrun, one rank per GPU): trees are independent, so every rank scores its own 4096 trees
(weak scaling, no data-path collective); time is the max over ranks.

  bandb     the second half of BASELINE's metric: exact branch and bound on the reference's measure/ datasets (bundled
            under tests/golden), time from search_begin to the proven optimum with the full MP set, nodes/s, and -- at
            N > 1, over the peer exchange of include/xmp_cuda.h (2d) -- speed-up, efficiency and node inflation against
            rank 0's own one-GPU search of the same dataset in the same run.  Every search is checked against the
            reference's score, tree count and canonical tree-set hash.  At N = 1 the reference's serial program
            (oracle/_ref/fastdnamp_ref) is timed on the same box for the datasets it finishes within ~35 s.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NUM_TAXA = 64
NUM_SITES = 1 << 20
N_TREES = 4096
M = 8                      # taxa on every partial tree; taxon M is the one inserted
METRIC = "fitch_insertion_scoring_site_merges_per_s"
UNIT = "Gsite-merges/s"


def make_rows(torch, device, seed=1):
    """64 x 2^20 i.i.d. uniform single-state sites, packed two per byte (even site = low nibble)."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    states = torch.randint(0, 4, (NUM_TAXA, NUM_SITES), generator=g, device=device, dtype=torch.uint8)
    nib = torch.bitwise_left_shift(torch.ones_like(states), states)
    return (nib[:, 0::2] | (nib[:, 1::2] << 4)).contiguous()


def make_paths(n_trees, m, seed=2):
    """Random stepwise-addition topologies on taxa 0..m-1 as insertion paths."""
    rng = np.random.default_rng(seed)
    paths = np.zeros((n_trees, NUM_TAXA), dtype=np.uint8)
    for t in range(3, m):
        paths[:, t] = rng.integers(0, 2 * t - 3, n_trees)
    return paths


def make_workload(torch, n_trees=N_TREES, m=M, device="cuda", seed_offset=0):
    d_rows = make_rows(torch, device, 1)
    paths = make_paths(n_trees, m, 2 + seed_offset)
    n_blocks = d_rows.shape[1] // 16
    E = 2 * m - 3
    d_mid = torch.empty((n_trees * E, n_blocks * 16), dtype=torch.uint8, device=device)
    return {"d_rows": d_rows, "paths": paths, "m": m, "E": E, "n_blocks": n_blocks, "d_mid": d_mid,
            "n_pairs": n_trees * E, "n_trees": n_trees}


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU every 5 ms through NVML while the timed region
    runs (nvidia-smi's own loop is too coarse for a region of ~100 ms)."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        import threading
        self.sm, self.reasons, self.max_mhz, self.err = [], set(), None, None
        self._stop = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as e:  # noqa: BLE001
            self.nv, self.err = None, repr(e)
            return
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception as e:  # noqa: BLE001
                self.err = repr(e)
                return
            time.sleep(0.005)

    def stop(self):
        if not self.nv:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: %s" % self.err]}
        self._stop.set()
        self.t.join()
        return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.sm)}


# ------------------------------------------------------------------------------- CPU reference legs
def cpu_cost_of_trees(paths, rows, m, threads, seconds_budget):
    """The reference's own Fitch code turning host inputs (alignment, topologies) into insertion costs:
    for each tree build every MIDPOINT set (FitchBases) and score taxon m on every edge
    (FitchScoreWeight1).  Returns (trees done, seconds, kind, score-only seconds for the same trees, their costs [done][2m-3])."""
    from concurrent.futures import ThreadPoolExecutor
    from tests import oracle_lib as ol
    nbytes = rows.shape[1]
    if ol.have_ref():
        R = ol.ref()
        kind = "reference"

        def bases(a, b, d):
            R.FitchBases_b2_w8_fastc64(a, b, d, nbytes // 8)

        def score1(a, b):
            return R.FitchScoreWeight1_b2_w8_fastc64(a, b, nbytes // 8)
    else:
        O = ol.oracle()
        kind = "port"

        def bases(a, b, d):
            O.xo_fitch_bases_swar64(a, b, d, nbytes // 8)

        def score1(a, b):
            return O.xo_fitch_score_weight1_swar64(a, b, nbytes // 8)

    E = 2 * m - 3
    rowp = [rows[t].ctypes.data for t in range(rows.shape[0])]

    def topo(path):
        par, kid = {0: 2, 1: 2, 2: None}, {2: (0, 1)}
        top = 2
        for t in range(3, m):
            v, a, b = int(path[t]), 2 * t - 3, 2 * t - 2
            pv = par[v]
            par[b] = pv
            if pv is None:
                top = b
            else:
                kid[pv] = tuple(b if k == v else k for k in kid[pv])
            kid[b] = (a, v)
            par[a] = par[v] = b
        pre, stack = [], [top]
        while stack:
            e = stack.pop()
            pre.append(e)
            if e in kid:
                stack.extend((kid[e][1], kid[e][0]))
        return par, kid, pre

    score_only = [0.0] * threads

    def one(args):
        slot, path = args
        buf = one.bufs[slot]
        up, down, mid = buf[0], buf[1], buf[2]
        par, kid, pre = topo(path)
        upp = {}
        for e in reversed(pre):
            if e in kid:
                bases(upp[kid[e][0]], upp[kid[e][1]], up[e].ctypes.data)
                upp[e] = up[e].ctypes.data
            else:
                upp[e] = rowp[(e + 3) // 2]
        dn = {}
        for e in pre:
            p = par[e]
            if p is None:
                dn[e] = rowp[0]
            else:
                sib = kid[p][1] if kid[p][0] == e else kid[p][0]
                bases(dn[p], upp[sib], down[e].ctypes.data)
                dn[e] = down[e].ctypes.data
            bases(upp[e], dn[e], mid[e].ctypes.data)
        t0 = time.perf_counter()
        costs = [score1(mid[e].ctypes.data, rowp[m]) for e in range(E)]
        score_only[slot] += time.perf_counter() - t0
        return costs

    one.bufs = [np.zeros((3, E, nbytes), np.uint8) for _ in range(threads)]
    done, t_start = 0, time.perf_counter()
    costs = np.zeros((len(paths), E), dtype=np.uint32)
    with ThreadPoolExecutor(threads) as ex:
        # waves of `threads` trees, one scratch slot per worker
        while done < len(paths) and time.perf_counter() - t_start < seconds_budget:
            wave = [(i, paths[done + i]) for i in range(min(threads, len(paths) - done))]
            for i, c in enumerate(ex.map(one, wave)):
                costs[done + i] = c
            done += len(wave)
    return done, time.perf_counter() - t_start, kind, max(score_only), costs[:done]


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch  # only for the same seeded alignment
    rows = make_rows(torch, "cpu", 1).numpy()
    paths = make_paths(N_TREES, M, 2)
    threads = os.cpu_count() or 1
    E = 2 * M - 3
    per_tree = E * NUM_SITES
    budget = 8.0
    vals, tot_trees, tot_s = [], 0, 0.0
    for step in range(args.warmup + args.steps):
        done, secs, kind, _, _ = cpu_cost_of_trees(paths, rows, M, threads, budget)
        if step >= args.warmup:
            vals.append(done * per_tree / secs / 1e9)
            tot_trees += done
            tot_s += secs
    value = tot_trees * per_tree / tot_s / 1e9
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * tot_s / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": "64 taxa x 2^20 sites, insertion scoring of all 13 edges of partial trees on 8 taxa; "
                               "host inputs (alignment + topologies) to costs on the host cores", "trees_per_step": tot_trees // max(args.steps, 1),
                   "taxa": NUM_TAXA, "sites": NUM_SITES, "m": M},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind,
                         "sample": "%d of %d trees per step (time-boxed to %.0f s), build MIDPOINT sets + score, %d threads" % (tot_trees // max(args.steps, 1), N_TREES, budget, threads)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ------------------------------------------------------------------- exact search (second half of the metric)
# name -> (golden record, options): its36.Bd is its36 under the default -Bd bound -- seconds on one GPU, the scaling
# workload; the reference did not finish it in 45 CPU-minutes, but the optimum and the MP set do not depend on the lower
# bound, so it must reproduce the golden of the author's -Bp run (233, 62,370 trees, the same canonical set).
BANDB_DATASETS = ["mt-10", "h1", "h4", "h5", "mh1", "mh3", "rbc14x759", "Metazoan", "its36", "its36.Bd"]
BANDB_GOLDEN = {"its36.Bd": "its36"}
REF_CPU_SAME_BOX = ["mt-10", "h3", "mh6", "h1", "mh1"]       # the reference's serial program finishes these within ~35 s


def _bandb_problem(xh, goldens, bundle, name):
    import tempfile
    g = goldens[BANDB_GOLDEN.get(name, name)]
    with tempfile.NamedTemporaryFile("wb", suffix=".phy", delete=False) as f:
        f.write(bundle[g["alignment"]].encode("latin-1"))
    a = xh.Alignment(f.name)
    os.unlink(f.name)
    if name not in BANDB_GOLDEN and "-Bp" in g["flags"]:
        # the author's its36 setting (-Bp -b 233 -m 100000); the partition bound is host precomputation
        # (host/partbound.c), outside the timed search like the reference's "Time to determine lower bounds"
        p = xh.Problem(a, options=xh.OPT_USEPARTBOUND | xh.OPT_IMPROVEINITUPPERBOUND, bound=int(g["flags"][g["flags"].index("-b") + 1]))
    else:
        p = xh.Problem(a)          # default options: -Bd bound, MST upper bound, like the reference run
    return g, p


def _canonical_hash(p, paths):
    import hashlib
    from tests.newick import canonical_newick
    canon = sorted(canonical_newick(t) for t in p.newick_list(paths))
    return hashlib.sha256((("\n".join(canon) + "\n").encode()) if canon else b"").hexdigest()


def run_bandb(torch, xc, dist, rank, world, local, names):
    """B&B nodes/s and time-to-optimum on the reference's measure/ datasets (bundled under tests/golden), sharded over
    the ranks.  At N > 1 every dataset is first searched by rank 0 ALONE (the other ranks wait), so that the speed-up and
    the node inflation are measured in the same run on the same box.  Returns a list of per-dataset dicts on rank 0."""
    import gzip
    from xmp_b200 import distributed as xd
    from xmp_b200 import hostlib as xh
    gold = os.path.join(ROOT, "tests", "golden")
    goldens = json.load(open(os.path.join(gold, "datasets.json")))
    bundle = json.loads(gzip.open(os.path.join(gold, "alignments.json.gz")).read())
    dev = torch.device("cuda", local)
    out = []
    # Untimed warm-up on every rank: the first search of a process loads the search kernels' code (lazy module
    # loading, several milliseconds) -- at N > 1 only rank 0 would have paid that in its solo pass.
    _, p = _bandb_problem(xh, goldens, bundle, "mt-10")
    ctx = xc.Context(local)
    ctx.load_problem(p)
    ctx.search_begin(bound=p.bound)
    while not ctx.search_run(0):
        pass
    ctx.close()
    for name in names:
        g, p = _bandb_problem(xh, goldens, bundle, name)
        ctx = xc.Context(local)
        ctx.load_problem(p)
        ctx.search_reserve(p.num_taxa, p.n_blocks)      # device memory set up before the clock, like the reference's arena
        solo = None
        if world > 1:
            if rank == 0 and not os.environ.get("XMP_BENCH_NO_SOLO"):      # A/B runs: skip the one-GPU pass
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                ctx.search_begin(bound=p.bound)
                while not ctx.search_run(0):
                    pass
                s1, n1, _ = ctx.search_result()
                torch.cuda.synchronize()
                solo = {"s": time.perf_counter() - t0, "nodes": ctx.search_stats()["nodes"], "ok": (s1, n1) == (g["score"], g["num_trees"])}
            connected = xd.connect_exchange(ctx, dev)    # set-up, like MPI_Init: 64-byte handles through one all-gather
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if world == 1:
            ctx.search_begin(bound=p.bound)
            while not ctx.search_run(0):
                pass
            score, n_mine, paths = ctx.search_result()
            total, info = n_mine, {"transport": "none", "rounds": 0}
        else:
            score, total, paths, info = xd.sharded_search(ctx, p.num_taxa, dev, begin={"bound": p.bound}, connected=connected)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        st = ctx.search_stats()
        t = torch.tensor([wall, st["seconds"], float(st["nodes"]), float(st["score_calls"]), float(st["site_merges"]),
                          float(st["steals_received"]), float(st["launches"])], dtype=torch.float64, device=dev)
        set_hash, per_rank = None, None
        if dist:
            hs = info.get("host_seconds") or {}
            mine = torch.tensor([hs.get("run", 0.0), hs.get("steal_wait", 0.0), float(st["nodes"]), float(st["steals_received"]),
                                 float(st["steals_served"]), float(st["launches"])], dtype=torch.float64, device=dev)
            allr = torch.zeros(world * mine.numel(), dtype=torch.float64, device=dev)
            dist.all_gather_into_tensor(allr, mine)
            per_rank = allr.cpu().numpy().reshape(world, -1)
            tmax = t.clone()
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            wall, dev_s = float(tmax[0]), float(tmax[1])
            allpaths = xd.gather_paths(paths, p.num_taxa, dev)       # the reference's MSG_TREES collection (outside the clock, as there)
            if rank == 0:
                set_hash = _canonical_hash(p, allpaths)
            xd.disconnect_exchange(ctx)
        else:
            wall, dev_s = float(t[0]), float(t[1])
            set_hash = _canonical_hash(p, paths)
        ctx.close()
        if rank == 0:
            ok = (score == g["score"] and total == g["num_trees"] and set_hash == g["canonical_sha256"])
            rec = {"dataset": name, "taxa": p.num_taxa, "patterns": p.num_sites, "score": score, "trees": total,
                   "matches_reference": ok, "canonical_set_sha256_ok": set_hash == g["canonical_sha256"],
                   "time_to_optimum_s": wall, "device_s_max_rank": dev_s,
                   "busy_fraction": float(t[1]) / (world * wall),      # sum over ranks of device-busy seconds / (N x wall)
                   "nodes": int(t[2]), "nodes_per_s": float(t[2]) / wall, "edge_evals_per_s": float(t[3]) / wall,
                   "gsite_merges_per_s": float(t[4]) / wall / 1e9, "launches": int(t[6]),
                   # the two kinds of site-operation apart (a cost = one insertion scored on one site, a merge = one state
                   # set recomputed on one site): the integer roofline weighs them by their own measured rates
                   "cost_site_ops": float(t[3]) * p.num_sites, "merge_site_ops": float(t[4]) - float(t[3]) * p.num_sites,
                   "transport": info["transport"], "steals": int(t[5]), "rank0_host_seconds": info.get("host_seconds"),
                   "rank0_begin_device_s": st["begin_seconds"],
                   "reference_cpu_bandb_s": g["ref_bandb_seconds"] if name not in BANDB_GOLDEN else None,
                   "reference_cpu_nodes": g["ref_bandb_nodes"] if name not in BANDB_GOLDEN else None,
                   "reference_cpu_source": "golden-time (measured when tests/golden/datasets.json was generated, another machine)"
                   if name not in BANDB_GOLDEN else "not available: the reference does not finish this search in 45 CPU-minutes"}
            if per_rank is not None:
                rec["per_rank"] = {"run_s": [round(float(x), 4) for x in per_rank[:, 0]], "steal_wait_s": [round(float(x), 4) for x in per_rank[:, 1]],
                                   "nodes": [int(x) for x in per_rank[:, 2]], "steals_received": [int(x) for x in per_rank[:, 3]],
                                   "steals_served": [int(x) for x in per_rank[:, 4]], "launches": [int(x) for x in per_rank[:, 5]]}
            if solo:
                rec.update({"one_gpu_same_run_s": solo["s"], "one_gpu_same_run_nodes": solo["nodes"], "one_gpu_same_run_ok": solo["ok"],
                            "speedup_vs_one_gpu": solo["s"] / wall, "efficiency": solo["s"] / wall / world,
                            "node_inflation": float(t[2]) / max(solo["nodes"], 1)})
            out.append(rec)
    return out if rank == 0 else None


def run_reference_bandb(names, budget_s=120.0):
    """The reference's own serial program (oracle/_ref/fastdnamp_ref, its unmodified sources compiled by oracle/Makefile)
    on this box's host cores, one process per dataset side by side; its "Time to perform branch and bound search" line
    (src/common.c:471-481).  Returns {name: seconds}.  Test infrastructure used as the reported CPU baseline only."""
    import gzip
    import re
    import tempfile
    exe = os.path.join(ROOT, "oracle", "_ref", "fastdnamp_ref")
    if not os.path.exists(exe):
        return {}
    gold = os.path.join(ROOT, "tests", "golden")
    goldens = json.load(open(os.path.join(gold, "datasets.json")))
    bundle = json.loads(gzip.open(os.path.join(gold, "alignments.json.gz")).read())
    procs = {}
    with tempfile.TemporaryDirectory(prefix="xmpref_") as top:
        for name in names:
            d = os.path.join(top, name)
            os.makedirs(d)
            open(os.path.join(d, "in.phy"), "wb").write(bundle[goldens[name]["alignment"]].encode("latin-1"))
            procs[name] = subprocess.Popen([exe, "--tbr=N", "--seed=1", "--displaynewbounds=n", "--displaynewtrees=n", "--updatesecs=0", "-o", "out.nex",
                                            "in.phy"], cwd=d, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True)
        out, deadline = {}, time.time() + budget_s
        for name, pr in procs.items():
            try:
                _, err = pr.communicate(timeout=max(1.0, deadline - time.time()))
            except subprocess.TimeoutExpired:
                pr.kill()
                continue
            m = re.search(r"Time to perform branch and bound search:\s+([\d.]+)s", err or "")
            ok = re.search(r"Each tree has score %d\b" % goldens[name]["score"], open(os.path.join(top, name, "out.nex")).read())
            if m and ok:
                out[name] = float(m.group(1))
    return out


# The reference's PARALLEL program on this box: its own boss / worker sources (src/mpiboss.c, src/mpiworker.c) compiled with
# TARGETMULTI against the one-node MPI of oracle/shim/mpi (there is no MPI installation in the image), one boss plus one worker
# per host core.  Ordered by weight in the review: the hard h / mh sets and its36 first.
REF_MPI_SAME_BOX = ["h4", "mh3", "its36", "h5", "Metazoan", "rbc14x759", "h1", "mh1", "mt-10"]


def run_reference_bandb_mpi(names, workers, budget_s=60.0):
    """`xmpirun -n workers + 1 fastdnamp_ref_mpi <the golden's flags>` per dataset, one after the other (each run uses every
    core), until the time budget is spent.  Returns {name: {"s": its "Time to perform branch and bound search", "workers": W}}
    for the runs that reproduced the golden score and tree count.  Test infrastructure used as a reported CPU baseline only."""
    import gzip
    import re
    import signal
    import tempfile
    exe = os.path.join(ROOT, "oracle", "_ref", "fastdnamp_ref_mpi")
    launcher = os.path.join(ROOT, "oracle", "_ref", "xmpirun")
    if not (os.path.exists(exe) and os.path.exists(launcher)):
        return {}
    gold = os.path.join(ROOT, "tests", "golden")
    goldens = json.load(open(os.path.join(gold, "datasets.json")))
    bundle = json.loads(gzip.open(os.path.join(gold, "alignments.json.gz")).read())
    out, deadline = {}, time.time() + budget_s
    with tempfile.TemporaryDirectory(prefix="xmprefmpi_") as top:
        for name in names:
            left = deadline - time.time()
            if left < 2.0:
                break
            g = goldens[name]
            d = os.path.join(top, name)
            os.makedirs(d)
            open(os.path.join(d, "in.phy"), "wb").write(bundle[g["alignment"]].encode("latin-1"))
            cmd = [launcher, "-n", str(workers + 1), exe, "--tbr=N", "--seed=1", "--displaynewbounds=n", "--displaynewtrees=n",
                   "--updatesecs=0"] + list(g["flags"]) + ["-o", "out.nex", "in.phy"]
            pr = subprocess.Popen(cmd, cwd=d, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True, start_new_session=True)
            try:
                _, err = pr.communicate(timeout=left)
            except subprocess.TimeoutExpired:
                try:
                    os.killpg(pr.pid, signal.SIGKILL)          # the launcher and its ranks: our own process group, nothing else
                except ProcessLookupError:
                    pass
                pr.wait()
                continue
            m = re.search(r"Time to perform branch and bound search:\s+([\d.]+)s", err or "")
            try:
                text = open(os.path.join(d, "out.nex")).read()
            except OSError:
                continue
            ok = re.search(r"Each tree has score %d\b" % g["score"], text) and text.count("TREE FASTDNAMP_") == g["num_trees"]
            if pr.returncode == 0 and m and ok:
                out[name] = {"s": float(m.group(1)), "workers": workers}
    return out


def run_planar_search_ab(names, timeout_s=150.0):
    """A/B of the search with every state set in planar form (XMP_SEARCH_PLANAR=1: an off-by-default switch of the library, DESIGN.md
    section 7) on a few datasets, in a SEPARATE process so that nothing it does can touch this run: `bench.py --bandb-only` with the
    switch in its environment.  Returns the records of that run (time to optimum, matches_reference, ...) or {"error": ...}."""
    env = dict(os.environ, XMP_SEARCH_PLANAR="1")
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT"):
        env.pop(k, None)
    try:
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--bandb-only", ",".join(names)], env=env, capture_output=True, text=True,
                           timeout=timeout_s)
        lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
        if r.returncode != 0 or not lines:
            return {"error": "exit code %d: %s" % (r.returncode, (r.stderr or "")[-400:])}
        return {"datasets": json.loads(lines[-1])["datasets"]}
    except Exception as e:  # noqa: BLE001
        return {"error": repr(e)}


# ------------------------------------------------------------------------------------------- ours
def run_ours(args):
    import torch
    import xmp_b200.cuda as xc

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL's own log lines (e.g. "NCCL version ...") belong on stderr: stdout carries the one JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: xmp_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    if args.bandb_only:
        # development / scaling runs: the exact-search leg alone (not the bench line)
        recs = run_bandb(torch, xc, dist, rank, world, local, args.bandb_only.split(","))
        if rank == 0:
            print(json.dumps({"bandb_only": True, "n_gpus": world, "datasets": recs}))
        if dist:
            dist.destroy_process_group()
        return

    n_trees = args.trees
    wl = make_workload(torch, n_trees, M, dev, seed_offset=rank)
    # Launch on a torch side stream so that torch.cuda.Event brackets exactly our kernels (the default
    # stream's handle is 0, which the C ABI reads as "use the context's own stream").
    torch.cuda.synchronize()         # the workload above was generated on torch's default stream
    side = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(side)
    ctx = xc.Context(local, stream=side.cuda_stream)
    ctx.load(wl["d_rows"])
    ctx.build_midpoints(wl["paths"], M, wl["d_mid"])
    n_pairs, E = wl["n_pairs"], wl["E"]
    d_cost = torch.zeros(n_pairs, dtype=torch.int32, device=dev)
    d_surv = torch.zeros(n_pairs, dtype=torch.int32, device=dev)
    d_n = torch.zeros(1, dtype=torch.int32, device=dev)
    # bound: the median insertion cost, so that about half of the pairs survive the pruning
    ctx.score_batch(wl["d_mid"], n_pairs, E, M, d_cost)
    ctx.synchronize()
    bound = int(d_cost.float().median().item())

    def step():
        ctx.score_batch(wl["d_mid"], n_pairs, E, M, d_cost, None, bound, d_surv, d_n)

    def barrier():
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    ev[0].record()
    for k in range(args.steps):
        step()
        ev[k + 1].record()
    barrier()
    ms_steps = [ev[k].elapsed_time(ev[k + 1]) for k in range(args.steps)]
    ms_total = ev[0].elapsed_time(ev[-1])
    clocks = sampler.stop() if sampler else None
    checksum = int(d_cost.to(torch.int64).sum().item())
    survivors = int(d_n.item())

    if args.kernel_only:
        print(json.dumps({"kernel_only": True, "ms_per_step": ms_total / args.steps, "ms_steps": ms_steps, "checksum": checksum}))
        return
    # ---- end to end through the host-buffer C ABI, same trees
    rows_pinned = wl["d_rows"].cpu().pin_memory()
    rows_np = rows_pinned.numpy()
    e2e_trees = n_trees
    cost_host = np.zeros((e2e_trees, E), dtype=np.uint32)
    ctx2 = xc.Context(local)
    h2d = rows_np.nbytes + wl["paths"].nbytes
    d2h = cost_host.nbytes

    def e2e_step():
        ctx2.load(rows_np)
        ctx2.score_insertions_host(wl["paths"][:e2e_trees], M, cost_host)

    del wl["d_mid"]
    torch.cuda.empty_cache()
    e2e_step()
    barrier()
    e2e_reps = max(2, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(e2e_reps):
        e2e_step()
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_reps
    assert int(cost_host.astype(np.int64).sum()) == checksum, "end-to-end costs differ from the resident-input costs"
    # The same call with the end-to-end kernel in its PACKED form (XMP_E2E_PACKED=1: the kernel every measurement up to the end of
    # round 2 was taken with; the shipped one keeps its sets in planar form): an A/B inside one run.  A reported
    # extra, so a failure here is recorded and does not cost the line.
    e2e_packed = None
    try:
        cost_packed = np.zeros_like(cost_host)
        os.environ["XMP_E2E_PACKED"] = "1"
        ctx2.load(rows_np)
        ctx2.score_insertions_host(wl["paths"][:e2e_trees], M, cost_packed)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(e2e_reps):
            ctx2.load(rows_np)
            ctx2.score_insertions_host(wl["paths"][:e2e_trees], M, cost_packed)
        torch.cuda.synchronize()
        e2e_packed = {"ms_per_step": (time.perf_counter() - t0) / e2e_reps * 1e3, "costs_equal_planar_form": bool((cost_packed == cost_host).all()),
                      "planar_form_ms_per_step_this_rank": e2e_s * 1e3}
    except Exception as e:  # noqa: BLE001
        e2e_packed = {"error": repr(e)}
    finally:
        os.environ.pop("XMP_E2E_PACKED", None)
    ctx2.close()

    t = torch.tensor([ms_total, e2e_s * 1e3], dtype=torch.float64, device=dev)
    if dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms = float(t[0].item()), float(t[1].item())
    int_peak = ctx.int_pipe_peak() if rank == 0 else None
    bandb, bandb_error = None, None
    if not args.no_bandb:
        if world == 1:
            # the exact-search leg is a secondary report: a failure there must not cost the headline line
            # (with several ranks an exception stays fatal: swallowing it on one rank would hang the others)
            try:
                bandb = run_bandb(torch, xc, dist, rank, world, local, BANDB_DATASETS)
            except Exception as e:  # noqa: BLE001
                print("bench.py: exact-search leg failed: %r" % (e,), file=sys.stderr)
                bandb_error = repr(e)
        else:
            bandb = run_bandb(torch, xc, dist, rank, world, local, BANDB_DATASETS)
    planar_ab = None
    if world == 1 and bandb is not None and not os.environ.get("XMP_SEARCH_PLANAR") and not os.environ.get("XMP_BENCH_NO_PLANAR_AB"):
        # the planar-form search variant on the three searches the review names, while the GPU is otherwise idle
        planar_ab = run_planar_search_ab(["h4", "mh3", "its36"])
    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return

    merges_per_step = n_pairs * NUM_SITES * world
    value = merges_per_step * args.steps / (ms_total * 1e-3) / 1e9
    e2e_value = e2e_trees * E * NUM_SITES * world / (e2e_ms * 1e-3) / 1e9
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    ms_kernel = statistics.median(ms_steps)
    achieved = n_pairs * NUM_SITES * 0.5 / (ms_kernel * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get("k_score_wide_dram_bytes_per_launch")

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32",
        "data": "synthetic",
        "config": {"workload": "synthetic 64 taxa x 2^20 sites (i.i.d. uniform over 4 states, seed 1), all weights 1; taxon 8 scored against "
                               "all 13 edges of %d random partial trees on 8 taxa per GPU (%d pairs, %.1f GiB of MIDPOINT state per GPU)" % (n_trees, n_pairs, n_pairs * NUM_SITES / 2 / 2**30),
                   "trees_per_gpu": n_trees, "taxa": NUM_TAXA, "sites": NUM_SITES, "m": M, "pairs_per_gpu": n_pairs,
                   "cache": "inputs larger than L2 (%.1f GiB per step vs 126 MB), no flush needed" % (n_pairs * NUM_SITES / 2 / 2**30),
                   "bound": bound, "survivors_last_step": survivors, "cost_checksum": checksum},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms,
                "what": "xmp_cuda_load(alignment from pinned host memory) + xmp_cuda_score_insertions_host(topologies) -> costs in host memory; "
                        "state sets are built on the device inside the timed region; counts insertion-scoring site-merges only",
                "packed_form": e2e_packed},
        "gpu_launches": 2 * args.steps,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "kernel": "k_score_wide<8>", "algorithmic_bytes_per_launch": n_pairs * NUM_SITES // 2,
                     "launch_ms": ms_kernel, "peak_source": peak_src,
                     "note": "launch_ms is the median CUDA-event time of one step = k_score_wide + k_prune + two 4-byte/213-KB memsets; "
                             "the ncu launch list under profiles/ gives k_score_wide's share"},
    }
    out["int_pipe_peak"] = {"unit": "Gsite-ops/s", "what": "register-resident microbenchmark of the Fitch instruction mixes on this GPU "
                            "(xmp_cuda_int_pipe_peak): the integer roofline for narrow state sets; the wide scoring kernel above "
                            "runs at %.1f %% of the weight-1 cost rate, i.e. it is HBM-bound, not ALU-bound" % (100.0 * value / world / int_peak["cost_w1"]),
                            **int_peak}
    if bandb is not None:
        for b in bandb:
            # Integer-roofline fraction of a search: the time the integer pipe needs for the merges and the costs the
            # search did, each at its own measured rate (xmp_cuda_int_pipe_peak), over the time the search took.
            floor_s = (b["merge_site_ops"] / int_peak["merge"] + b["cost_site_ops"] / int_peak["cost_w1"]) / 1e9 / world
            b["int_roof_floor_s"] = floor_s
            b["frac_of_int_roof"] = floor_s / b["time_to_optimum_s"]
            # Kept for comparison with round 1 and NOT a roofline fraction: it sets merges + costs, counted one by one,
            # against the rate of merge-and-cost PAIRS, which overstates the fraction (by 1.8 at this mix of 2.2 merges
            # per cost).  frac_of_int_roof above is the number to quote.
            b["frac_of_int_pipe_merge_cost_peak"] = b["gsite_merges_per_s"] / world / int_peak["merge_cost"]
        out["bandb"] = {"what": "exact branch and bound (default -Bd bound; its36 also in the author's -Bp setting), time from search_begin to the "
                                "proven optimum with the full MP set (wall clock around barriers, max over ranks), checked against the "
                                "reference's score, tree count and canonical tree-set hash.  At N > 1: bound by peer atomicMin and "
                                "asynchronous stealing over the exchange blocks (transport 'peer'), each dataset also searched by rank 0 "
                                "alone in the same run (one_gpu_same_run_s) for speed-up, efficiency and node inflation",
                        "n_gpus": world, "datasets": bandb}
    if bandb_error is not None:
        out["bandb"] = {"error": bandb_error}
    if planar_ab is not None and "bandb" in out:
        ab = {"what": "the same searches by a second process of this run with XMP_SEARCH_PLANAR=1: every state set of the search in planar form "
                      "(word k of a 128-bit block = base k of its 32 sites: a merge is 8 LOP3 instead of 28 instructions), an A/B switch that is "
                      "off by default; seconds to the proven optimum and whether score, tree count and canonical tree-set hash equal the "
                      "reference's"}
        if "error" in planar_ab:
            ab["error"] = planar_ab["error"]
        else:
            base = {b["dataset"]: b for b in (bandb or [])}
            ab["datasets"] = [{"dataset": d["dataset"], "time_to_optimum_s": d["time_to_optimum_s"], "matches_reference": d["matches_reference"],
                               "nodes": d["nodes"], "packed_form_time_to_optimum_s": base.get(d["dataset"], {}).get("time_to_optimum_s")}
                              for d in planar_ab["datasets"]]
        out["bandb"]["planar_search_variant"] = ab
    if world == 1 and not os.environ.get("XMP_BENCH_SKIP_CPU"):       # A/B runs: skip the 12-second CPU leg
        rows = rows_np
        threads = os.cpu_count() or 1
        done, secs, kind, score_s, cpu_costs = cpu_cost_of_trees(wl["paths"], rows, M, threads, 12.0)
        # full-size parity for free: every tree the CPU leg finished must cost what the GPU said, edge by edge
        # (cost_host = the end-to-end path's costs, whose checksum equals the resident path's)
        assert (cpu_costs == cost_host[:done]).all(), "GPU insertion costs differ from the reference's Fitch code at full size"
        out["cpu_baseline"] = {"value": done * E * NUM_SITES / secs / 1e9, "unit": UNIT, "cores": threads, "kind": kind,
                               "sample": "%d of %d trees, host inputs to costs (build MIDPOINT sets with FitchBases, then FitchScoreWeight1), "
                                         "time-boxed to 12 s" % (done, n_trees),
                               "score_only_value": done * E * NUM_SITES / max(score_s, 1e-9) / 1e9,
                               "costs_equal_gpu": "%d trees x %d edges compared, all equal" % (done, E)}
        if bandb is not None and not args.no_ref_bandb:
            # the reference's serial program on THIS box, same run (src/common.c:471-481 "Time to perform branch and bound search")
            ref_s = run_reference_bandb(REF_CPU_SAME_BOX)
            by_name = {b["dataset"]: b for b in bandb}
            for name, secs_ref in ref_s.items():
                if name in by_name:
                    by_name[name]["reference_cpu_bandb_s"] = secs_ref
                    by_name[name]["reference_cpu_source"] = "same box, same run (oracle/_ref/fastdnamp_ref, 1 core)"
            out["bandb"]["reference_cpu_same_box_s"] = ref_s
            # ... and its parallel (boss / worker) program on all the cores of this box, same run
            workers = 0
            try:
                workers = max(1, min(len(os.sched_getaffinity(0)) or (os.cpu_count() or 2), 32))     # one worker per core (at most 32), plus the boss
                mpi_s = run_reference_bandb_mpi([n for n in REF_MPI_SAME_BOX if n in by_name], workers)
            except Exception as e:  # noqa: BLE001  (a reported baseline must not cost the headline line)
                print("bench.py: the reference's parallel program failed: %r" % (e,), file=sys.stderr)
                mpi_s = {}
            for name, r in mpi_s.items():
                by_name[name]["reference_cpu_parallel_s"] = r["s"]
                by_name[name]["reference_cpu_parallel_workers"] = r["workers"]
            out["bandb"]["reference_cpu_parallel_same_box"] = {
                "what": "the reference's own MPI boss / worker program (src/mpiboss.c, src/mpiworker.c, unmodified, TARGETMULTI) on this "
                        "box's host cores in this run: one boss and one worker per core over the one-node MPI of oracle/shim/mpi; its "
                        "'Time to perform branch and bound search'; score and tree count checked against the goldens",
                "workers": workers if mpi_s else 0, "seconds": {k: v["s"] for k, v in mpi_s.items()}}
    print(json.dumps(out))
    if dist:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--trees", type=int, default=N_TREES, help="partial trees per GPU (default: the BASELINE workload)")
    ap.add_argument("--kernel-only", action="store_true", help="profiling runs: skip the end-to-end and CPU legs")
    ap.add_argument("--no-bandb", action="store_true", help="skip the exact-search leg (B&B nodes/s, time-to-optimum)")
    ap.add_argument("--no-ref-bandb", action="store_true", help="do not time the reference's serial program on this box's cores (~35 s)")
    ap.add_argument("--bandb-only", default=None, help="comma-separated datasets: run only the exact-search leg on them and print its JSON")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
