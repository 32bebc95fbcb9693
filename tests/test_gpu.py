"""Parity tests proper (run with -m gpu on a B200): every entry point of the C ABI against the CPU
oracle on seeded inputs, against the committed known-answer vectors, and against the goldens
generated from the reference program.  Integer work: everything is compared bit for bit."""
import gzip
import hashlib
import json
import os
import subprocess

import numpy as np
import pytest

from tests import oracle_lib as ol
from tests.conftest import GOLDEN, ROOT, golden_output, golden_trees
from tests.newick import canonical_newick
from tests.test_host import prepare

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def xc():
    import xmp_b200.cuda as xc
    return xc


@pytest.fixture(scope="module")
def torch():
    import torch
    assert torch.cuda.is_available()
    return torch


@pytest.fixture()
def ctx(xc):
    c = xc.Context(0)
    yield c
    c.close()


O = ol.oracle()


# ---------------------------------------------------------------------------------- per-pair family
def test_pair_family_known_answers(xc):
    cases = json.load(open(os.path.join(GOLDEN, "fitch_kat.json")))
    for c in cases:
        if c["nbytes"] % 16:
            continue
        s1, s2, prev = (np.frombuffer(bytes.fromhex(c[k]), np.uint8).copy() for k in ("s1", "s2", "prev"))
        w = np.array(c["weights"], np.uint32)
        assert xc.fitch_bases(s1, s2).tobytes().hex() == c["bases"]
        d, ch = xc.fitch_bases_and_compare(s1, s2, prev)
        assert d.tobytes().hex() == c["bases"] and int(ch != 0) == c["changed_vs_prev"]
        _, ch = xc.fitch_bases_and_compare(s1, s2, d)
        assert int(ch != 0) == c["changed_vs_self"]
        assert xc.fitch_score_weight1(s1, s2) == c["score_weight1"]
        assert xc.fitch_score(s1, s2, w) == c["score"]
        assert xc.fitch_score(s1, s2, w, fw1=True) == c["score_fw1"]
        # stop-early contract: exact when within the limit, otherwise anything above the limit
        for got, true, limit in ((xc.fitch_score(s1, s2, w, fw1=True, max_score=c["limit"]), c["score_fw1"], c["limit"]),
                                 (xc.fitch_score(s1, s2, w, max_score=c["limit"]), c["score"], c["limit"]),
                                 (xc.fitch_score_weight1(s1, s2, max_score=c["score_weight1"] // 2), c["score_weight1"], c["score_weight1"] // 2)):
            assert got == true if true <= limit else got > limit


def test_pair_family_random_vs_oracle(xc):
    rng = np.random.default_rng(5)
    for nbytes in (16, 48, 80, 1248, 4096 + 16):
        for amb in (0.0, 0.3):
            s1, s2, prev = (ol.random_sets(rng, nbytes, amb) for _ in range(3))
            w = ol.random_weights(rng, nbytes * 2, 0.4, 3000)
            want = np.zeros(nbytes, np.uint8)
            O.xo_fitch_bases(ol.p(s1), ol.p(s2), ol.p(want), nbytes)
            assert (xc.fitch_bases(s1, s2) == want).all()
            assert xc.fitch_score_weight1(s1, s2) == O.xo_fitch_score_weight1(ol.p(s1), ol.p(s2), nbytes)
            assert xc.fitch_score(s1, s2, w) == O.xo_fitch_score(ol.p(s1), ol.p(s2), ol.p(w), nbytes)
            # unsorted weights: the plain variant is exact, fw1 follows the reference's hand-over rule
            wu = rng.integers(1, 9, nbytes * 2).astype(np.uint32)
            assert xc.fitch_score(s1, s2, wu) == O.xo_fitch_score(ol.p(s1), ol.p(s2), ol.p(wu), nbytes)
            assert xc.fitch_score(s1, s2, wu, fw1=True) == O.xo_fitch_score_stopearly(ol.p(s1), ol.p(s2), ol.p(wu), nbytes, 16, 1, 0x7FFFFFFF)
    e = np.zeros(0, np.uint8)
    assert xc.fitch_score_weight1(e, e) == 0 and xc.fitch_bases(e, e).size == 0


# ------------------------------------------------------------------------------- batched scoring
def oracle_costs(rows, weights, sets, taxon):
    nbytes = rows.shape[1]
    w = None if weights is None else np.ascontiguousarray(weights, np.uint32)
    out = np.zeros(len(sets), np.uint32)
    for i, s in enumerate(sets):
        s = np.ascontiguousarray(s)
        out[i] = (O.xo_fitch_score(ol.p(s), ol.p(rows[taxon]), ol.p(w), nbytes) if w is not None
                  else O.xo_fitch_score_weight1(ol.p(s), ol.p(rows[taxon]), nbytes))
    return out


@pytest.mark.parametrize("n_blocks", [1, 2, 3, 5, 9, 17, 33, 64, 78, 130, 300, 513])
@pytest.mark.parametrize("weighted", [False, True])
def test_score_batch_matches_oracle(ctx, torch, n_blocks, weighted):
    rng = np.random.default_rng(n_blocks * 2 + weighted)
    num_taxa, n_pairs, ppt = 6, 1000 + n_blocks, 7
    nbytes = n_blocks * 16
    rows = np.stack([ol.random_sets(rng, nbytes, 0.1) for _ in range(num_taxa)])
    weights = ol.random_weights(rng, nbytes * 2, 0.35, 500) if weighted else None
    rest = rng.integers(0, 5, num_taxa).astype(np.uint32)
    sets = np.stack([ol.random_sets(rng, nbytes, 0.2) for _ in range(n_pairs)])
    base = rng.integers(0, 50, (n_pairs + ppt - 1) // ppt).astype(np.uint32)
    ctx.load(rows, weights, rest, 0)
    d_sets = torch.from_numpy(sets).cuda()
    d_base = torch.from_numpy(base.astype(np.int32)).cuda()
    d_cost = torch.zeros(n_pairs, dtype=torch.int32, device="cuda")
    d_surv = torch.full((n_pairs,), -1, dtype=torch.int32, device="cuda")
    d_n = torch.zeros(1, dtype=torch.int32, device="cuda")
    taxon = 4
    torch.cuda.synchronize()     # torch filled the buffers on its own stream; the library launches on another
    want = oracle_costs(rows, weights, sets, taxon)
    bound = int(np.median(want + base[np.arange(n_pairs) // ppt])) + int(rest[taxon])
    ctx.score_batch(d_sets, n_pairs, ppt, taxon, d_cost, d_base, bound, d_surv, d_n)
    ctx.synchronize()
    got = d_cost.cpu().numpy().astype(np.uint32)
    assert (got == want).all()
    keep = (want.astype(np.int64) <= bound - base[np.arange(n_pairs) // ppt].astype(np.int64) - int(rest[taxon]))
    n = int(d_n.item())
    assert n == int(keep.sum())
    assert sorted(d_surv[:n].cpu().numpy().tolist()) == np.nonzero(keep)[0].tolist()
    # no pruning requested
    ctx.score_batch(d_sets, n_pairs, ppt, taxon, d_cost)
    ctx.synchronize()
    assert (d_cost.cpu().numpy().astype(np.uint32) == want).all()


def test_tma_staged_scorer_gives_the_same_costs():
    """XMP_SCORE_TMA selects the TMA bulk-copy version of the wide scorer (read once per process, hence a
    child process): same parity tests, both ring configurations."""
    import sys
    for variant in ("1", "2"):
        env = dict(os.environ, XMP_SCORE_TMA=variant)
        r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu.py"), "-q", "-m", "gpu", "-p", "no:cacheprovider",
                            "-k", "score_batch_matches_oracle and (300 or 513)"], env=env, capture_output=True, text=True, cwd=ROOT)
        assert r.returncode == 0 and "4 passed" in r.stdout, r.stdout[-2000:]


def test_score_batch_empty_and_single(ctx, torch):
    rng = np.random.default_rng(1)
    rows = np.stack([ol.random_sets(rng, 32) for _ in range(4)])
    ctx.load(rows)
    d_n = torch.ones(1, dtype=torch.int32, device="cuda")
    d_cost = torch.zeros(1, dtype=torch.int32, device="cuda")
    d_surv = torch.zeros(1, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    ctx.score_batch(None, 0, 1, 3, d_cost, None, 5, d_surv, d_n)
    ctx.synchronize()
    assert int(d_n.item()) == 0
    one = torch.from_numpy(rows[1].copy()).cuda()
    torch.cuda.synchronize()
    ctx.score_batch(one, 1, 1, 3, d_cost)
    ctx.synchronize()
    assert int(d_cost.item()) == O.xo_fitch_score_weight1(ol.p(rows[1]), ol.p(rows[3]), 32)


@pytest.mark.parametrize("m,n_blocks,weighted", [(3, 2, True), (4, 1, False), (8, 5, True), (13, 40, True), (20, 3, False), (40, 2, True), (64, 2, False),
                                                  (8, 300, True), (13, 513, False), (20, 256, True), (40, 257, False)])
def test_build_midpoints_and_insertion_costs(ctx, torch, m, n_blocks, weighted):
    rng = np.random.default_rng(m * 31 + n_blocks)
    num_taxa, n_trees = m + 1, 37
    nbytes = n_blocks * 16
    rows = np.stack([ol.random_sets(rng, nbytes, 0.1) for _ in range(num_taxa)])
    weights = ol.random_weights(rng, nbytes * 2, 0.3, 77) if weighted else None
    ctx.load(rows, weights, None, 11)
    paths = ol.random_paths(rng, n_trees, m, num_taxa)
    E = 2 * m - 3
    d_mid = torch.zeros((n_trees, E, nbytes), dtype=torch.uint8, device="cuda")
    d_len = torch.zeros(n_trees, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    ctx.build_midpoints(paths, m, d_mid, d_len)
    ctx.synchronize()
    got_mid = d_mid.cpu().numpy()
    got_len = d_len.cpu().numpy()
    costs = ctx.score_insertions_host(paths, m)
    for i in range(n_trees):
        mid, _, _, length = ol.tree_sets(rows, m, paths[i], weights)
        assert (got_mid[i] == mid).all(), i
        assert got_len[i] == length + 11
        assert (costs[i] == oracle_costs(rows, weights, mid, m)).all()


def test_invalid_arguments_are_reported(ctx, xc):
    with pytest.raises(xc.XmpCudaError, match="no problem loaded"):
        ctx.score_insertions_host(np.zeros((1, 8), np.uint8), 5)
    rng = np.random.default_rng(2)
    ctx.load(np.stack([ol.random_sets(rng, 16) for _ in range(6)]))
    bad = np.zeros((1, 6), np.uint8)
    bad[0, 4] = 9   # a 4-taxon tree has edges 0..4
    with pytest.raises(xc.XmpCudaError, match="does not exist"):
        ctx.score_insertions_host(bad, 5)
    with pytest.raises(xc.XmpCudaError, match="no search in progress"):
        ctx.search_result()


# --------------------------------------------------------------------------------------- the search
def run_search(xc, p, **kw):
    score, n, paths, stats = xc.search(p, **kw)
    return score, n, p.newick_list(paths), stats


QUICK = ["its10", "its12", "its14", "primate", "quick2", "53humans.12", "53humans.16", "mt-10", "EukaryotesrDNA",
         "h3", "mh6", "mt-10.Bm", "h3.Bm", "h3.Bdm", "h3.b359"]
HARD = ["h1", "h2", "mh1", "mh2", "mh4", "rbc14x759", "Metazoan", "h4", "h5"]


def check_against_golden(name, g, p, score, n, trees):
    assert (score, n) == (g["score"], g["num_trees"]), name
    canon = sorted(canonical_newick(t) for t in trees)
    blob = ("\n".join(canon) + "\n").encode() if canon else b""
    assert hashlib.sha256(blob).hexdigest() == g["canonical_sha256"]
    if n <= 2000:
        assert canon == golden_trees(name)
        # sorted into the reference's depth-first order the text is the reference's tree block
        body = "".join("\tTREE FASTDNAMP_%d = [&U] %s\n" % (i + 1, t) for i, t in enumerate(trees))
        assert body.encode() in golden_output(name)


@pytest.mark.parametrize("name", QUICK + HARD)
def test_search_matches_reference(xc, alignment_dir, goldens, name):
    g = goldens[name]
    a, p = prepare(alignment_dir, goldens, name)
    score, n, trees, stats = run_search(xc, p)
    check_against_golden(name, g, p, score, n, trees)
    print("\n%s: %.3f s device, %d nodes (reference %d, %.2f s CPU), %d batches" %
          (name, stats["seconds"], stats["nodes"], g["ref_bandb_nodes"], g["ref_bandb_seconds"], stats["iterations"]))


def test_search_random_alignments_with_ambiguity_codes(xc, tmp_path):
    """Seeded random alignments (IUPAC ambiguity codes, gaps, repeated columns so that weights exceed 1, few
    taxa so that the oracle is instant): host preprocessing + device search against the oracle search."""
    from xmp_b200 import hostlib as xh
    rng = np.random.default_rng(2026)
    alphabet = np.array(list("ACGTACGTACGTACGTRYSWKMBDHVN-?"))
    checked = 0
    for case in range(24):
        n, sites = int(rng.integers(4, 10)), int(rng.integers(8, 60))
        base = rng.integers(0, len(alphabet), (n, sites))
        # evolve columns a little so that they are informative, then repeat some of them
        for t in range(1, n):
            keep = rng.random(sites) < 0.6
            base[t, keep] = base[rng.integers(0, t), keep]
        cols = rng.integers(0, sites, sites * 2)
        text = "%d %d\n" % (n, len(cols)) + "".join("T%-9d%s\n" % (t, "".join(alphabet[base[t, cols]])) for t in range(n))
        f = tmp_path / ("rand%d.phy" % case)
        f.write_text(text)
        a = xh.Alignment(str(f))
        p = xh.Problem(a, options=xh.OPT_USEDISTBASESBOUNDREST | xh.OPT_IMPROVEINITUPPERBOUND | (xh.OPT_USEMSTBOUNDREST if case % 3 == 0 else 0))
        if p.num_sites == 0:
            continue
        want_score, want_n, want_trees, _, _ = ol.search(p)
        score, n_trees, paths, _ = xc.search(p, flags=(0, xc.SEARCH_NO_FUSE, xc.SEARCH_NO_EAGER, xc.SEARCH_NO_FUSE | xc.SEARCH_NO_EAGER)[case % 4])
        assert (score, n_trees) == (want_score, want_n), case
        assert p.newick_list(paths) == want_trees, case       # same trees, and in the reference's DFS order
        checked += 1
    assert checked >= 15


DEEP = ["53humans.24", "53humans.26", "53humans.28"]


@pytest.mark.parametrize("name", DEEP)
def test_search_deep_with_partition_bound_and_tree_cap(xc, alignment_dir, goldens, name):
    """24-28 taxa under the native partition bound: millions of optimal trees pass through the device tree
    buffer; -m keeps the first 1000 in the reference's depth-first order (src/yanbader.c:272-293)."""
    g = goldens[name]
    a, p = prepare(alignment_dir, goldens, name)
    assert [[i, int(p.rest_bound[i])] for i in range(2, p.num_taxa)] == g["rest_bound"]
    cap = int(g["flags"][g["flags"].index("-m") + 1])
    score, n, paths, stats = xc.search(p)
    assert (score, n) == (g["score"], g["num_trees_total"])      # total from a second, uncapped run of the reference
    kept = p.sort_paths_dfs(paths)[:cap]
    check_against_golden(name, g, p, score, cap, [p.path_to_newick(x) + ";" for x in kept])
    print("\n%s: %.3f s device, %d nodes (reference %d, %.2f s CPU), %d batches" %
          (name, stats["seconds"], stats["nodes"], g["ref_bandb_nodes"], g["ref_bandb_seconds"], stats["iterations"]))


def test_search_without_informative_sites(xc, tmp_path):
    """No parsimony-informative site: every tree is optimal at the constant weight, in the reference's depth-first order."""
    from xmp_b200 import hostlib as xh
    f = tmp_path / "zero.phy"
    f.write_text("6 3\nA         AC-\nB         ANG\nC         ACG\nD         TCG\nE         ACG\nF         ACR\n")
    p = xh.Problem(xh.Alignment(str(f)))
    assert (p.num_sites, p.n_blocks) == (0, 1)
    want_score, want_n, want_trees, _, _ = ol.search(p)
    score, n_trees, paths, _ = xc.search(p)
    assert (score, n_trees) == (want_score, want_n) == (p.constant_weight, 105)
    assert p.newick_list(paths) == want_trees


@pytest.mark.parametrize("name", ["53humans.26", "53humans.29"])
def test_search_deep_with_the_cap_applied_inside_the_library(xc, alignment_dir, goldens, name):
    """max_trees handed to the library: a shard that holds more than a million optimal trees cuts its set back to the
    first max_trees in the reference's depth-first order and goes on counting (src/yanbader.c:272-293), so the memory
    of a capped search stays bounded.  53humans.26: 9.8 M optimal trees; 53humans.29: 88.4 M (seven minutes in the
    reference, seventeen to count them all) -- the reference's first 1000 and its total count."""
    g = goldens[name]
    a, p = prepare(alignment_dir, goldens, name)
    cap = int(g["flags"][g["flags"].index("-m") + 1])
    score, n, paths, stats = xc.search(p, max_trees=cap)
    assert (score, n) == (g["score"], g["num_trees_total"])
    assert cap <= len(paths) < n                                 # the paths of most of them were let go
    kept = p.sort_paths_dfs(paths)[:cap]
    check_against_golden(name, g, p, score, cap, [p.path_to_newick(x) + ";" for x in kept])


@pytest.mark.parametrize("name", ["Metazoan", "rbc14x759", "EukaryotesrDNA", "mh1"])
def test_search_classic_layout_on_multi_block_sets(xc, alignment_dir, goldens, name):
    """The layout that keeps MIDPOINT sets (used by default only above 32 blocks per set, i.e. mt-10 here) on
    1-, 5-, 6- and 19-block sets, where the default is eager scoring: same trees either way."""
    g = goldens[name]
    a, p = prepare(alignment_dir, goldens, name)
    score, n, trees, stats = run_search(xc, p, flags=xc.SEARCH_NO_EAGER)
    check_against_golden(name, g, p, score, n, trees)


def test_search_memory_is_reserved_ahead_and_reused(xc, alignment_dir, goldens):
    """xmp_cuda_search_reserve sets up the device memory before the search; the context keeps it, and
    searches of other shapes in the same context (smaller, larger, reserved or not) stay exact."""
    ctx = xc.Context(0)
    try:
        for name, reserve in (("h3", True), ("mh6", False), ("its14", True), ("mt-10", False), ("h3", False)):
            g = goldens[name]
            a, p = prepare(alignment_dir, goldens, name)
            if reserve:
                ctx.search_reserve(p.num_taxa, p.n_blocks)
            ctx.load_problem(p)
            ctx.search_begin(bound=p.bound)
            while not ctx.search_run(0):
                pass
            score, n, paths = ctx.search_result()
            check_against_golden(name, g, p, score, n, p.newick_list(paths))
        with pytest.raises(xc.XmpCudaError):
            ctx.search_reserve(3, 1)
    finally:
        ctx.close()


def test_search_its36_with_loaded_partition_bound(xc, alignment_dir, goldens, tmp_path):
    """The author's own its36 setting (-Bp -b 233 -m 100000): restBound[] comes from the reference's
    restBound.txt through -Bf, as INTEGRATION.md describes."""
    from xmp_b200 import hostlib as xh
    g = goldens["its36"]
    f = tmp_path / "restBound.txt"
    f.write_text("i\tbound\n" + "".join("%d\t%d\n" % (i, b) for i, b in g["rest_bound"]))
    a = xh.Alignment(str(alignment_dir / g["alignment"]))
    p = xh.Problem(a, options=xh.OPT_USELOADRESTBOUND | xh.OPT_IMPROVEINITUPPERBOUND, bound=233, max_trees=100000, rest_bound_file=str(f))
    score, n, trees, stats = run_search(xc, p)
    check_against_golden("its36", g, p, score, n, trees)


def test_search_its36_with_native_partition_bound(xc, alignment_dir, goldens):
    """The same setting with restBound[] computed here (host/partbound.c), no file from the reference involved."""
    from xmp_b200 import hostlib as xh
    g = goldens["its36"]
    a = xh.Alignment(str(alignment_dir / g["alignment"]))
    p = xh.Problem(a, options=xh.OPT_USEPARTBOUND | xh.OPT_IMPROVEINITUPPERBOUND, bound=233, max_trees=100000)
    assert [[i, int(p.rest_bound[i])] for i in range(2, p.num_taxa)] == g["rest_bound"]
    score, n, trees, stats = run_search(xc, p)
    check_against_golden("its36", g, p, score, n, trees)


def test_search_mh3_large_tree_set(xc, alignment_dir, goldens):
    g = goldens["mh3"]
    a, p = prepare(alignment_dir, goldens, "mh3")
    score, n, trees, stats = run_search(xc, p)
    check_against_golden("mh3", g, p, score, n, trees)


def test_search_bound_edge_cases(xc, alignment_dir, goldens):
    a, p = prepare(alignment_dir, goldens, "h3")
    g = goldens["h3"]
    # a bound below the optimum finds nothing and reports the bound (reference: "NO BEST TREE FOUND")
    score, n, trees, _ = run_search(xc, p, bound=g["score"] - 1)
    assert (score, n, trees) == (goldens["h3.b358"]["score"], 0, [])
    # a bound equal to the optimum keeps ties
    score, n, trees, _ = run_search(xc, p, bound=g["score"])
    check_against_golden("h3.b359", goldens["h3.b359"], p, score, n, trees)


def test_search_is_independent_of_scheduling(xc, alignment_dir, goldens):
    """Same tree set without the greedy dive, with a frontier budget that forces small batches, and in
    slices of a few batches at a time."""
    g = goldens["mh6"]
    a, p = prepare(alignment_dir, goldens, "mh6")
    for kw in ({"flags": xc.SEARCH_NO_DIVE}, {"mem_budget": 600 << 20}, {"flags": xc.SEARCH_NO_FUSE},
               {"flags": xc.SEARCH_NO_FUSE | xc.SEARCH_NO_DIVE, "mem_budget": 400 << 20},
               {"flags": xc.SEARCH_NO_EAGER}, {"flags": xc.SEARCH_NO_EAGER | xc.SEARCH_NO_FUSE, "mem_budget": 600 << 20}):
        score, n, trees, stats = run_search(xc, p, **kw)
        check_against_golden("mh6", g, p, score, n, trees)
    # the one-level-per-synchronisation scheduler (kept for A/B measurements) instead of the sweep
    os.environ["XMP_SWEEP"] = "0"
    try:
        score, n, trees, stats = run_search(xc, p)
        check_against_golden("mh6", g, p, score, n, trees)
    finally:
        del os.environ["XMP_SWEEP"]
    ctx = xc.Context(0)
    ctx.load_problem(p)
    ctx.search_begin(bound=p.bound)
    slices = 0
    while not ctx.search_run(3):
        slices += 1
    score, n, paths = ctx.search_result()
    check_against_golden("mh6", g, p, score, n, p.newick_list(paths))
    assert slices > 1
    ctx.close()


def test_sharded_search_union_equals_whole(xc, alignment_dir, goldens):
    """Two shards (run one after the other on this GPU) with bound exchange: the union of their tree
    sets is the reference's set, each tree found exactly once."""
    for name in ("mh6", "h3", "53humans.16"):
        g = goldens[name]
        a, p = prepare(alignment_dir, goldens, name)
        ctxs = [xc.Context(0) for _ in range(3)]
        for i, c in enumerate(ctxs):
            c.load_problem(p)
            c.search_begin(bound=p.bound, shard_index=i, shard_count=3)
        done = [False] * 3
        while not all(done):
            for i, c in enumerate(ctxs):
                if not done[i]:
                    done[i] = c.search_run(8)
            best = min(c.search_get_bound() for c in ctxs)
            for c in ctxs:
                c.search_set_bound(best)
        allpaths = []
        for c in ctxs:
            score, n, paths = c.search_result()
            assert score == g["score"]
            allpaths.append(paths)
            c.close()
        paths = np.concatenate(allpaths)
        check_against_golden(name, g, p, g["score"], len(paths), p.newick_list(paths))


def test_job_export_import_moves_work_between_contexts(xc, alignment_dir, goldens):
    """Work stealing primitive: open subtrees detached from one context as job descriptors and replayed
    in another give the same overall result."""
    g = goldens["mh6"]
    a, p = prepare(alignment_dir, goldens, "mh6")
    c0, c1 = xc.Context(0), xc.Context(0)
    for c in (c0, c1):
        c.load_problem(p)
    c0.search_begin(bound=p.bound)
    c1.search_begin(bound=p.bound, flags=xc.SEARCH_NO_DIVE)
    assert len(c1.search_export_jobs(16)) == 1          # throw c1's own root away: it only gets stolen work
    assert c1.search_run(0)                             # nothing to do: done, and may be woken up again
    c0.search_run(6)
    moved, rounds = 0, 0
    done0 = done1 = False
    while not (done0 and done1):
        jobs = c0.search_export_jobs(64) if rounds % 2 == 0 else np.zeros((0, p.num_taxa), np.uint8)
        moved += len(jobs)
        c1.search_import_jobs(jobs)
        done0 = c0.search_run(5)
        done1 = c1.search_run(5)
        best = min(c0.search_get_bound(), c1.search_get_bound())
        c0.search_set_bound(best)
        c1.search_set_bound(best)
        rounds += 1
    assert moved > 0
    s0, n0, p0 = c0.search_result()
    s1, n1, p1 = c1.search_result()
    assert s0 == s1 == g["score"]
    paths = np.concatenate([p0, p1])
    check_against_golden("mh6", g, p, g["score"], len(paths), p.newick_list(paths))
    c0.close()
    c1.close()


def test_stealing_between_shards_with_a_deep_split_level(xc, alignment_dir, goldens):
    """Sharding and stealing together (round-1 advisor finding): with the frontier dealt out at a deep tree size and
    one-batch slices, nodes below the split level are still open when the first steal happens.  They must never change
    hands (they would be dealt again with the thief's index): the union of the shards is the reference's set, each tree
    exactly once, and a job below the split level is refused on import."""
    g = goldens["mh6"]
    a, p = prepare(alignment_dir, goldens, "mh6")
    ctxs = [xc.Context(0) for _ in range(3)]
    for i, c in enumerate(ctxs):
        c.load_problem(p)
        c.search_begin(bound=p.bound, shard_index=i, shard_count=3, split_level=9)
    done, moved, rounds = [False] * 3, 0, 0
    while not all(done):
        for i, c in enumerate(ctxs):
            if not done[i]:
                done[i] = c.search_run(1)
        donor, thief = rounds % 3, (rounds + 1) % 3
        jobs = ctxs[donor].search_export_jobs(8)
        assert all(j[0] >= 9 for j in jobs)
        if len(jobs):
            ctxs[thief].search_import_jobs(jobs)
            done[thief] = False
            moved += len(jobs)
        best = min(c.search_get_bound() for c in ctxs)
        for c in ctxs:
            c.search_set_bound(best)
        rounds += 1
    assert moved > 0
    bad = np.zeros((1, p.num_taxa), np.uint8)
    bad[0, 0] = 5
    with pytest.raises(xc.XmpCudaError):
        ctxs[0].search_import_jobs(bad)
    paths = []
    for c in ctxs:
        score, n, ps = c.search_result()
        assert score == g["score"]
        paths.append(ps)
        c.close()
    paths = np.concatenate(paths)
    check_against_golden("mh6", g, p, g["score"], len(paths), p.newick_list(paths))


# ------------------------------------------------------------------------------------------ the CLI
@pytest.mark.parametrize("name", ["mt-10", "h3", "mh6", "EukaryotesrDNA", "quick2"])
def test_cli_writes_the_reference_result_file(alignment_dir, goldens, tmp_path, name):
    """fastdnamp_b200 with the reference's measurement flags produces the reference's output file byte
    for byte (for mt-10 that is build-vswin32/mt-10.tre) and its restBound.txt."""
    g = goldens[name]
    out = tmp_path / "out.nex"
    r = subprocess.run([os.path.join(ROOT, "fastdnamp_b200"), "--tbr=N", "--seed=1", "--displaynewbounds=n", "--displaynewtrees=n",
                        "--updatesecs=0", "-o", str(out), str(alignment_dir / g["alignment"])], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert out.read_bytes() == golden_output(name)
    rb = [[int(x) for x in line.split()] for line in (tmp_path / "restBound.txt").read_text().splitlines()[1:]]
    assert rb == g["rest_bound"]
    assert "%d parsimoniously-informative sites." % g["pi_sites"] in r.stderr
    raw = subprocess.run([os.path.join(ROOT, "fastdnamp_b200"), "-q", "-r", "-m", "2", str(alignment_dir / g["alignment"])],
                         cwd=tmp_path, capture_output=True, text=True)
    assert raw.returncode == 0 and raw.stderr == ""
    want = [l.split("] ")[1] for l in golden_output(name).decode().splitlines() if "TREE FASTDNAMP_" in l][:2]
    assert raw.stdout.splitlines() == want


# -------------------------------------------------------------------------------------- two GPUs
def _n_gpus():
    try:
        import xmp_b200.cuda as xc
        return xc.lib().xmp_cuda_device_count()
    except Exception:  # noqa: BLE001
        return 0


@pytest.mark.skipif(_n_gpus() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("name", ["mh6", "h1"])
def test_cli_two_gpus_same_result_file(alignment_dir, goldens, tmp_path, name):
    g = goldens[name]
    out = tmp_path / "out.nex"
    r = subprocess.run([os.path.join(ROOT, "fastdnamp_b200"), "--gpus=2", "--tbr=N", "--seed=1", "--displaynewbounds=n", "--displaynewtrees=n",
                        "--updatesecs=0", "-o", str(out), str(alignment_dir / g["alignment"])], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert out.read_bytes() == golden_output(name)


@pytest.mark.skipif(_n_gpus() < 2, reason="needs two GPUs")
def test_torchrun_sharded_search_two_gpus():
    import sys
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29517", os.path.join(ROOT, "tests", "multi_gpu_check.py"), "mh6", "h3", "mh1", "53humans.16"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count(" OK ") == 8 and "MISMATCH" not in r.stdout


# ----------------------------------------------------------- BASELINE.json sizes: the m-sweep program
def test_sweep_m_check_program(tmp_path):
    """tests/sweep_m_check (the synthetic microbenchmark at m = 4, 8, 16, 63), here on 256 trees in chunks
    of 4 GiB: resident scoring == CPU oracle on the first trees of every chunk == the fused host-buffer path."""
    import json
    exe = os.path.join(ROOT, "tests", "sweep_m_check")
    if not os.path.exists(exe):
        pytest.skip("tests/sweep_m_check not built (make checks)")
    env = dict(os.environ, SWEEP_TREES="256", SWEEP_STATE_GIB="4")
    r = subprocess.run([exe, "6554.2", "4", "8", "16", "63"], capture_output=True, text=True, timeout=600, env=env, cwd=tmp_path)
    if r.returncode in (2, 3):
        pytest.skip("sweep_m_check could not run here (CUDA / library error, not a parity result): " + r.stderr[-300:])
    lines = [json.loads(l) for l in r.stdout.splitlines() if l.startswith("{")]
    assert r.returncode == 0 and [l["m"] for l in lines] == [4, 8, 16, 63], r.stdout + r.stderr
    for l in lines:
        assert l["oracle_ok"] and l["fused_path_equal"] and l["oracle_trees_checked"] >= 2, l
    assert lines[3]["chunks"] == 4


# ----------------------------------------------------------- BASELINE.json sizes: structural properties
def test_full_size_microbench_properties(ctx, torch):
    """64 taxa x 2^20 sites (BASELINE config 2), all weights 1: at this size the oracle is only run on a
    few trees; the rest is checked through size-independent properties of the scores."""
    import bench
    wl = bench.make_workload(torch, n_trees=64, m=8)
    torch.cuda.synchronize()
    ctx.load(wl["d_rows"])
    ctx.build_midpoints(wl["paths"], wl["m"], wl["d_mid"])
    n_pairs = wl["n_pairs"]
    d_cost = torch.zeros(n_pairs, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    ctx.score_batch(wl["d_mid"], n_pairs, wl["E"], wl["m"], d_cost)
    ctx.synchronize()
    cost = d_cost.cpu().numpy().reshape(-1, wl["E"])
    rows = wl["d_rows"].cpu().numpy()
    # (1) oracle on all 64 trees, and one checksum over the whole cost matrix
    want = np.zeros_like(cost)
    for i in range(64):
        mid, _, _, _ = ol.tree_sets(rows[:wl["m"] + 1], wl["m"], wl["paths"][i])
        want[i] = oracle_costs(rows, None, mid, wl["m"])
        assert (cost[i] == want[i]).all(), i
    assert hashlib.sha256(cost.astype("<u4").tobytes()).hexdigest() == hashlib.sha256(want.astype("<u4").tobytes()).hexdigest()
    # (2) scoring a taxon against its own leaf edge costs 0; against another taxon's leaf edge it is
    #     their Hamming distance when the leaf's DOWN set does not help, never more
    d_self = torch.zeros(n_pairs, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    ctx.score_batch(wl["d_mid"], n_pairs, wl["E"], 1, d_self)
    ctx.synchronize()
    assert (d_self.cpu().numpy().reshape(-1, wl["E"])[:, 0] == 0).all()    # edge 0 = leaf of taxon 1
    # (3) additivity over site ranges: scoring the two halves separately sums to the whole
    half = wl["n_blocks"] // 2
    lo = torch.zeros(n_pairs, dtype=torch.int32, device="cuda")
    d_mid = wl["d_mid"].view(n_pairs, wl["n_blocks"] * 16)
    c2 = __import__("xmp_b200.cuda", fromlist=["Context"]).Context(0)
    hi = torch.zeros(n_pairs, dtype=torch.int32, device="cuda")
    parts = [(wl["d_rows"][:, :half * 16].contiguous(), d_mid[:, :half * 16].contiguous(), lo),
             (wl["d_rows"][:, half * 16:].contiguous(), d_mid[:, half * 16:].contiguous(), hi)]
    torch.cuda.synchronize()
    for r, s_, out in parts:
        c2.load(r)
        c2.score_batch(s_, n_pairs, wl["E"], wl["m"], out)
        c2.synchronize()
    assert ((lo + hi).cpu().numpy().reshape(-1, wl["E"]) == cost).all()
    c2.close()
