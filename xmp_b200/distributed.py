"""One-process-per-GPU exact search over torch.distributed: the replacement for the reference's MPI
boss/worker scheduler (reference src/mpiboss.c:30-429, src/mpiworker.c:214-380) when the host is
Python (bench.py under torchrun).  The C host (host/fastdnamp.c) does the same with threads.

Differences from the reference, by design:
  * no boss: every rank searches (the reference's rank 0 only relays, src/fastdnamp.c:112-118);
  * the upper bound is kept consistent by an all-reduce(min) of one word every round instead of
    worker -> boss -> all workers point-to-point relays (src/mpiboss.c:162-171, src/mpiworker.c:19-41);
  * work stealing is decided collectively from an all-gather of the open-node counts: ranks that ran
    dry receive job descriptors (insertion paths, a few hundred bytes) from the ranks holding the most
    open subtrees through one more all-gather (over NVLink with the nccl backend); the thief rebuilds the
    subtree state from the descriptor alone, as the reference's fast-forward does (src/yanbader.c:136-143);
  * termination: the all-gathered counts are all zero.

The engine is any object with the search_* methods of xmp_b200.cuda.Context; the CPU tests drive the
same protocol over gloo with a CPU stand-in engine (tests/test_distributed.py).
"""
import numpy as np
import torch
import torch.distributed as dist


def _t(x, device, dtype=torch.int64):
    return torch.tensor(x, dtype=dtype, device=device)


def sharded_search(engine, num_taxa, device="cpu", slice_batches=8, steal_nodes=256, group=None, on_round=None):
    """Runs engine (already search_begin()'ed with this rank's shard_index / shard_count) to completion
    together with the other ranks.  Returns (score, total number of optimal trees, this rank's optimal
    paths, number of rounds).  Every rank returns the same score and total.

    One collective per round (an all-gather of [bound, open subtrees] per rank) plus, only in rounds
    where some rank ran dry, one all-gather of job descriptors.  Point-to-point send/recv is avoided on
    purpose: with NCCL every new (donor, thief) pair pays a connection set-up of ~100 ms, more than a
    whole search of the hard datasets takes."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    rounds = 0
    done = False
    rec = num_taxa + 1
    state = torch.zeros(world * 2, dtype=torch.int64, device=device)          # flat: gloo wants 1-D outputs
    outbox = torch.zeros(steal_nodes * rec, dtype=torch.uint8, device=device)
    inbox = torch.zeros(world * steal_nodes * rec, dtype=torch.uint8, device=device)
    while True:
        if not done:
            done = engine.search_run(slice_batches)
        mine = torch.tensor([engine.search_get_bound(), int(np.asarray(engine.search_pending()).sum())], dtype=torch.int64, device=device)
        dist.all_gather_into_tensor(state, mine, group=group)
        st = state.cpu().numpy().reshape(world, 2)
        bound = int(st[:, 0].min())                       # upper bound: min over ranks
        counts = [int(c) for c in st[:, 1]]
        engine.search_set_bound(bound)
        rounds += 1
        if on_round:
            on_round(rounds, counts, bound)
        if sum(counts) == 0:
            break
        # collective steal plan, identical on every rank: the i-th idle rank takes from the i-th richest
        idle = [r for r in range(world) if counts[r] == 0]
        donors = sorted((r for r in range(world) if counts[r] > 1), key=lambda r: (-counts[r], r))
        if not idle or not donors:
            continue
        plan = [(donors[i % len(donors)], thief) for i, thief in enumerate(idle)]
        out = np.zeros((steal_nodes, num_taxa + 1), dtype=np.uint8)     # column 0: destination rank + 1 (0 = empty slot)
        mine_out = [thief for donor, thief in plan if donor == rank]
        used = 0
        for thief in mine_out:
            want = min((steal_nodes - used), max(1, counts[rank] // (2 * (len(mine_out) + 1))))
            if want <= 0:
                break
            jobs = engine.search_export_jobs(want)
            out[used:used + len(jobs), 0] = thief + 1
            out[used:used + len(jobs), 1:] = jobs
            used += len(jobs)
        outbox.copy_(torch.from_numpy(out.reshape(-1)))
        dist.all_gather_into_tensor(inbox, outbox, group=group)
        if rank in idle:
            got = inbox.cpu().numpy().reshape(-1, num_taxa + 1)
            jobs = got[got[:, 0] == rank + 1][:, 1:]
            if len(jobs):
                engine.search_import_jobs(np.ascontiguousarray(jobs))
                done = False
    if not done:
        engine.search_run(0)      # nothing is pending anywhere; lets an engine that had not noticed yet wrap up
    score, n, paths = engine.search_result()
    tot = _t([n], device)
    dist.all_reduce(tot, op=dist.ReduceOp.SUM, group=group)
    return score, int(tot.item()), paths, rounds


def gather_paths(paths, num_taxa, device="cpu", group=None):
    """All ranks' optimal paths on every rank (the reference's MSG_TREES collection, src/common.c:397-413)."""
    world = dist.get_world_size(group)
    n = _t([len(paths)], device)
    sizes = [_t([0], device) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    cap = max(max(sizes), 1)
    mine = torch.zeros((cap, num_taxa), dtype=torch.uint8, device=device)
    if len(paths):
        mine[:len(paths)] = torch.from_numpy(np.ascontiguousarray(paths)).to(device)
    bufs = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(bufs, mine, group=group)
    parts = [bufs[r][:sizes[r]].cpu().numpy() for r in range(world)]
    return np.concatenate(parts) if parts else np.zeros((0, num_taxa), np.uint8)
