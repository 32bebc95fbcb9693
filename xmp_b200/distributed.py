"""One-process-per-GPU exact search under torchrun: the replacement for the reference's MPI boss/worker scheduler
(reference src/mpiboss.c:30-429, src/mpiworker.c:214-380) when the host is Python (bench.py).  The C host
(host/fastdnamp.c) does the same with threads.

torch.distributed is plumbing here: it carries the 64-byte exchange handles once, provides the barrier before a
search and the sums / gathers after it.  The search itself exchanges nothing through it:

  * the upper bound lives in peer-mapped device memory; the kernels that find a better tree lower every rank's word
    with a system-scope atomicMin over NVLink (the reference relays MSG_NEWBOUND through the boss,
    src/mpiboss.c:162-171, src/mpiworker.c:19-41);
  * a rank whose frontier ran dry asks one peer for open subtrees and receives job descriptors (insertion paths) in
    its device mailbox; the thief rebuilds the subtree state from the descriptor alone, as the reference's
    fast-forward does (src/yanbader.c:136-143).  No collective and no lock step in the loop
    (include/xmp_cuda.h 2d, xmp_b200/csrc/exchange_proto.h);
  * termination: a counter of ranks holding work, kept in rank 0's block.

Fallback transport ("collective"): the round-1 scheme -- one all-gather of [bound, open subtrees] per round and a
second one for job descriptors -- for engines without an exchange (the CPU stand-in of tests/test_distributed.py)
and for nodes where the GPUs cannot map each other's memory.  It is lock-step and is reported as such.
"""
import numpy as np
import torch
import torch.distributed as dist


def _t(x, device, dtype=torch.int64):
    return torch.tensor(x, dtype=dtype, device=device)


def connect_exchange(engine, device="cpu", group=None):
    """Maps every rank's exchange block into this rank (once per engine / context).  Returns True when ALL ranks
    succeeded; otherwise every rank is left disconnected and the caller uses the collective transport."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if not hasattr(engine, "exchange_handle"):
        ok = 0
    else:
        try:
            mine = np.asarray(engine.exchange_handle(), dtype=np.uint8)
            ok = 1
        except Exception:  # noqa: BLE001
            mine, ok = np.zeros(64, np.uint8), 0
    flag = _t([ok], device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    if int(flag.item()) == 0:
        return False
    handles = torch.zeros(world * 64, dtype=torch.uint8, device=device)
    dist.all_gather_into_tensor(handles, torch.from_numpy(mine).to(device), group=group)
    try:
        engine.exchange_connect(rank, world, handles.cpu().numpy())
        ok = 1
    except Exception as e:  # noqa: BLE001
        engine._exchange_error = repr(e)
        ok = 0
    flag = _t([ok], device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    if int(flag.item()) == 0:
        if ok:
            engine.exchange_disconnect()
        return False
    return True


def disconnect_exchange(engine, group=None):
    """Unmaps the peers' blocks on every rank, then a barrier: only after it may a rank free its own block (close its
    context) -- memory that a peer still has mapped must not be freed."""
    if hasattr(engine, "exchange_disconnect"):
        engine.exchange_disconnect()
    dist.barrier(group=group)


def sharded_search(engine, num_taxa, device="cpu", group=None, begin=None, connected=None, slice_batches=8, steal_nodes=256,
                   on_round=None):
    """Runs one search over all ranks.  begin: keyword arguments for engine.search_begin (bound, max_trees, ...);
    shard_index / shard_count are filled in.  connected: the result of connect_exchange (None = call it here).
    Returns (score, total number of optimal trees, this rank's optimal paths, info) with info = {"transport": "peer" |
    "collective", "rounds": host rounds (collective) or steal waits (peer)}.  Every rank returns the same score and total."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    kw = dict(begin or {})
    kw.setdefault("shard_index", rank)
    kw.setdefault("shard_count", world)
    if connected is None:
        connected = world > 1 and connect_exchange(engine, device, group)
    import time
    t0 = time.perf_counter()
    if connected:
        engine.exchange_reset()
        dist.barrier(group=group)                 # nobody writes into a block that has not been reset yet
        t1 = time.perf_counter()
        engine.search_begin(**kw)
        t2 = time.perf_counter()
        waits, t_run, t_steal = 0, 0.0, 0.0
        while True:
            ta = time.perf_counter()
            while not engine.search_run(0):
                pass
            tb = time.perf_counter()
            waits += 1
            state = engine.search_steal()         # 1 = got work, 2 = finished everywhere
            t_run += tb - ta
            t_steal += time.perf_counter() - tb
            if state != 1:
                break
        info = {"transport": "peer", "rounds": waits,
                "host_seconds": {"reset_barrier": t1 - t0, "begin": t2 - t1, "run": t_run, "steal_wait": t_steal}}
    else:
        engine.search_begin(**kw)
        info = {"transport": "collective", "rounds": _collective_rounds(engine, num_taxa, device, group, slice_batches, steal_nodes, on_round)}
    # belt and braces: the bound every rank filters its trees against is the minimum over ranks
    b = _t([engine.search_get_bound()], device)
    dist.all_reduce(b, op=dist.ReduceOp.MIN, group=group)
    engine.search_set_bound(int(b.item()))
    t3 = time.perf_counter()
    score, n, paths = engine.search_result()
    tot = _t([n], device)
    dist.all_reduce(tot, op=dist.ReduceOp.SUM, group=group)
    total = int(tot.item())
    info.setdefault("host_seconds", {})["finish"] = time.perf_counter() - t3
    return score, total, paths, info


def _collective_rounds(engine, num_taxa, device, group, slice_batches, steal_nodes, on_round):
    """Lock-step fallback: one all-gather of [bound, open subtrees] per round plus, only in rounds where some rank ran
    dry, one all-gather of job descriptors.  (Point-to-point send/recv is avoided on purpose: with NCCL every new
    (donor, thief) pair pays a connection set-up of ~100 ms.)"""
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
    return rounds


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
