"""Host C code (host/*.c through xmp_b200.hostlib) against the goldens generated from the reference:
reader, preprocessing, taxon order, bounds, tree text.  No GPU needed."""
import hashlib
import os

import numpy as np
import pytest

from tests.conftest import golden_output, golden_trees
from tests.newick import canonical_newick
from xmp_b200 import hostlib as xh

FAST = ["mt-10", "rbc14x759", "h1", "h2", "h3", "h4", "h5", "mh1", "mh2", "mh3", "mh4", "mh6", "EukaryotesrDNA",
        "Metazoan", "its10", "its12", "its14", "primate", "quick2", "53humans.12", "53humans.16", "e1", "e3", "e4", "e5", "e6"]


def prepare(alignment_dir, goldens, name, **kw):
    g = goldens[name]
    a = xh.Alignment(str(alignment_dir / g["alignment"]))
    opts = xh.OPT_NEXUSFMT | xh.OPT_IMPROVEINITUPPERBOUND
    flags = g["flags"]
    bound = xh.NO_LIMIT
    if "-b" in flags:
        bound = int(flags[flags.index("-b") + 1])
    b = [f for f in flags if f.startswith("-B")]
    letters = b[0][2:] if b else "d"
    if "d" in letters:
        opts |= xh.OPT_USEDISTBASESBOUNDREST
    if "m" in letters:
        opts |= xh.OPT_USEMSTBOUNDREST
    if "p" in letters:
        opts |= xh.OPT_USEPARTBOUND
    return a, xh.Problem(a, options=opts, bound=bound, **kw)


@pytest.mark.parametrize("name", FAST)
def test_preprocessing_matches_reference(alignment_dir, goldens, name):
    g = goldens[name]
    a, p = prepare(alignment_dir, goldens, name)
    assert p.num_sites == g["pi_sites"]
    assert p.n_constant_sites == g["constant_sites"]
    assert p.n_nonpi_sites == g["nonpi_sites"]
    assert p.constant_weight == g["constant_weight"]
    assert [int(t) + 1 for t in p.taxon_map] == g["taxon_order"]
    assert [[i, int(p.rest_bound[i])] for i in range(2, p.num_taxa)] == g["rest_bound"]
    if g["initial_upper_bound"] is not None:
        assert p.mst_upper_bound == g["initial_upper_bound"] and p.bound == g["initial_upper_bound"]
    # packing: 16-byte blocks, padding nibbles 0xF with weight 1, weights descending
    assert p.n_blocks == (p.num_sites + 31) // 32
    w = p.weights
    assert (np.diff(w.astype(np.int64)) <= 0).all() and (w[p.num_sites:] == 1).all()
    nib = np.stack([p.rows & 15, p.rows >> 4], axis=2).reshape(p.num_taxa, -1)
    assert (nib[:, :p.num_sites] == p.sites).all() and (nib[:, p.num_sites:] == 15).all()


@pytest.mark.parametrize("name", ["mt-10.Bm", "h3.Bm", "h3.Bdm"])
def test_mst_lower_bound_matches_reference(alignment_dir, goldens, name):
    _, p = prepare(alignment_dir, goldens, name)
    assert [[i, int(p.rest_bound[i])] for i in range(2, p.num_taxa)] == goldens[name]["rest_bound"]


def test_survey_mst_goldens(alignment_dir, goldens):
    """restBound values quoted in SURVEY.md section 8c (probe build with zero-initialised autos)."""
    _, p = prepare(alignment_dir, goldens, "mt-10.Bm")
    assert list(p.rest_bound[2:10]) == [3297, 1794, 1055, 650, 360, 176, 80, 0]
    _, p = prepare(alignment_dir, goldens, "h3.Bm")
    assert list(p.rest_bound[2:12]) == [99, 59, 36, 18, 9, 6, 5, 2, 1, 0]
    _, p = prepare(alignment_dir, goldens, "mt-10")
    assert list(p.rest_bound[2:10]) == [3106, 2109, 1440, 991, 614, 330, 161, 0]


def test_rest_bound_file_round_trip(alignment_dir, goldens, tmp_path):
    g = goldens["its36"]
    f = tmp_path / "restBound.txt"
    f.write_text("i\tMin weight added by taxa i+1, i+2, ..., n-1.\n" + "".join("%d\t%d\n" % (i, b) for i, b in g["rest_bound"]))
    a = xh.Alignment(str(alignment_dir / g["alignment"]))
    p = xh.Problem(a, options=xh.OPT_USELOADRESTBOUND | xh.OPT_IMPROVEINITUPPERBOUND, bound=233, rest_bound_file=str(f))
    assert [[i, int(p.rest_bound[i])] for i in range(2, p.num_taxa)] == g["rest_bound"]
    assert [int(t) + 1 for t in p.taxon_map] == g["taxon_order"]
    assert p.bound == 233


def test_unsupported_bounds_say_so(alignment_dir, goldens):
    a = xh.Alignment(str(alignment_dir / "h3.phy"))
    with pytest.raises(xh.HostError, match="-Bf"):        # -Bi (incompatibility bound) is not built in
        xh.Problem(a, options=128)


# ---- PHYLIP reader fixtures (reference data/phylip_format_testing; names encode the outcome) ----
def test_phylip_ok_fixtures(alignment_dir):
    d = alignment_dir / "phylip_format_testing"
    a1 = xh.Alignment(str(d / "ok1.phy"))
    for fn in ("ok2_lots_of_tabs_spaces_and_blank_lines.phy", "ok3_missing_final_newline.phy"):
        a = xh.Alignment(str(d / fn))
        assert (a.num_taxa, a.seq_len) == (a1.num_taxa, a1.seq_len)
        assert a.labels == a1.labels
        assert (a.chars == a1.chars).all()
    assert a1.labels == ["T1", "T2", "T3", "T4"]
    assert bytes(a1.chars[0]).decode() == "ACGTA"


def test_phylip_bad_fixtures(alignment_dir):
    d = alignment_dir / "phylip_format_testing"
    with pytest.raises(xh.HostError, match="fewer sites than the first line"):
        xh.Alignment(str(d / "bad1_T3_missing_1_base.phy"))
    with pytest.raises(xh.HostError, match="input only contains"):
        xh.Alignment(str(d / "bad2_T4_missing_final_line.phy"))


def test_phylip_interleaved_and_errors(tmp_path):
    f = tmp_path / "il.phy"
    f.write_text(" 4 12\nAlpha     ACGTAC\nBeta  x   ACGTAC\nGamma     ACGTAC\nDelta     ACGTA-\n\nGTACGT\nGTACGT\ngtacgn\nGTACG?\n")
    a = xh.Alignment(str(f))
    assert a.labels == ["Alpha", "Beta  x", "Gamma", "Delta"]
    assert bytes(a.chars[2]).decode() == "ACGTACGTACGN"
    f.write_text("4 4\nA         AC!T\n")
    with pytest.raises(xh.HostError, match="invalid character '!' on line 2"):
        xh.Alignment(str(f))
    f.write_text("4 4\nA         ACGTA\n")
    with pytest.raises(xh.HostError, match="too many characters on line 2"):
        xh.Alignment(str(f))
    f.write_text("3 4\nA         ACGT\nB         ACGT\nC         ACGT\n")
    with pytest.raises(xh.HostError, match="four or more taxa"):
        xh.Problem(xh.Alignment(str(f)))


# ---- trees ----
def test_dfs_coordinates_round_trip():
    rng = np.random.default_rng(3)
    for n in (4, 7, 12, 36, 100):
        for _ in range(20):
            path = np.zeros(n, np.uint8)
            for t in range(3, n):
                path[t] = rng.integers(0, 2 * t - 3)
            dfs = xh.path_to_dfs(path, n)
            assert all(dfs[t] < 2 * t - 3 for t in range(3, n))
            back = xh.dfs_to_path(dfs, n)
            assert (back[3:] == path[3:]).all()


def test_base_tree_newick(alignment_dir, goldens):
    a, p = prepare(alignment_dir, goldens, "its10")
    p.num_taxa_saved = p.num_taxa
    # a 4-taxon path inserting taxon 3 on each of the three edges of the start tree
    tm = p.taxon_map
    p4 = lambda e: np.array([0, 0, 0, e], np.uint8)
    L = xh.lib()
    import ctypes as C
    buf = C.create_string_buffer(256)
    out = []
    for e in range(3):
        n = L.xmph_path_to_newick(p4(e).ctypes.data, 4, np.ascontiguousarray(tm, np.uint32).ctypes.data, buf)
        out.append(buf.raw[:n].decode())
    t = [int(x) + 1 for x in tm[:4]]
    assert out[0] == "(%d,((%d,%d),%d))" % (t[0], t[3], t[1], t[2])     # above leaf 1: new leaf LEFT
    assert out[1] == "(%d,(%d,(%d,%d)))" % (t[0], t[1], t[3], t[2])     # above leaf 2
    assert out[2] == "(%d,(%d,(%d,%d)))" % (t[0], t[3], t[1], t[2])     # above the internal node


def test_walk_program_derivation(tmp_path):
    """xmp_b200/csrc/walk_prog.h (shared with the CUDA kernels): a child's walk program derived from its
    parent's equals the program built from the child's topology, up to 100 taxa."""
    import subprocess
    exe = str(tmp_path / "walk_check")
    subprocess.run(["g++", "-O2", "-Wall", "-Wextra", "-Werror", "-o", exe, os.path.join(os.path.dirname(os.path.abspath(__file__)), "walk_check.cpp")], check=True)
    r = subprocess.run([exe, "100", "120"], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.startswith("OK "), r.stdout + r.stderr


def test_child_walk_replayed_on_the_host(tmp_path):
    """xmp_b200/csrc/child_walk.cuh -- the per-thread child construction of the search kernels, compiled here for the
    host -- against state sets, programs and insertion costs computed from scratch: 1-3 blocks per set, ambiguity
    codes, weighted and weight-1 costs, eager and classic record layouts, every insertion position up to 36 taxa."""
    import shutil
    import subprocess
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    exe = str(tmp_path / "child_walk_check")
    # the shipped walk, and the A/B variant kept behind XMP_WALK_UL (scripts/build_variant.sh)
    for flags in ([], ["-DXMP_WALK_UL"]):
        subprocess.run([nvcc, "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets"] + flags + ["-o", exe,
                        os.path.join(os.path.dirname(os.path.abspath(__file__)), "child_walk_check.cu")], check=True)
        for taxa, reps in (("12", "24"), ("36", "8")):
            r = subprocess.run([exe, taxa, reps], capture_output=True, text=True)
            assert r.returncode == 0 and r.stdout.startswith("OK "), r.stdout + r.stderr


# ---- partition lower bound (-Bp, reference src/partbound.c) --------------------------------------
def _partbound_goldens():
    import json
    from tests.conftest import GOLDEN
    return json.load(open(os.path.join(GOLDEN, "partbound.json")))


PB = _partbound_goldens()


def _partbound_problem(alignment_dir, case, threads=0):
    a = xh.Alignment(str(alignment_dir / case["alignment"]))
    opts = xh.OPT_NEXUSFMT
    strategy, stale = 0, 2
    for f in case["flags"]:
        if f.startswith("-B"):
            opts |= (xh.OPT_USEDISTBASESBOUNDREST if "d" in f else 0) | (xh.OPT_USEMSTBOUNDREST if "m" in f else 0)
            opts |= xh.OPT_USEPARTBOUND if "p" in f else 0
            strategy = 1 if "C" in f else 2 if "L" in f else 0
        elif f.startswith("--pbmaxiters="):
            stale = int(f.split("=")[1])
        elif f == "-d":
            opts |= xh.OPT_DISCARDMISSAMBIG
    return xh.Problem(a, options=opts, part_strategy=strategy, part_max_stale=stale, threads=threads)


@pytest.mark.parametrize("name", sorted(PB["cases"]))
def test_partition_bound_matches_reference(alignment_dir, name):
    """restBound.txt of `fastdnamp -Bp...` (all strategies, --pbmaxiters, mixed with -Bd / -Bm) is reproduced exactly."""
    case = PB["cases"][name]
    p = _partbound_problem(alignment_dir, case)
    assert [[i, int(p.rest_bound[i])] for i in range(2, p.num_taxa)] == case["rest_bound"]


def test_partition_bound_is_independent_of_thread_count(alignment_dir):
    case = PB["cases"]["Metazoan"]
    for threads in (1, 3):
        p = _partbound_problem(alignment_dir, case, threads=threads)
        assert [[i, int(p.rest_bound[i])] for i in range(2, p.num_taxa)] == case["rest_bound"]


def test_site_pair_costs_equal_the_reference_table():
    """The reference ships a generated 65,535-entry table of two-site lengths (src/sitepaircosts.c); here the
    same numbers come out of the general part bound on first use.  All entries, by hash and by sample."""
    L = xh.lib()
    table = bytes(L.xmph_site_pair_cost(i) for i in range(1, 65536))
    assert hashlib.sha256(table).hexdigest() == PB["site_pair_table_sha256"]
    for i, v in PB["site_pair_table_samples"].items():
        assert table[int(i) - 1] == v
    assert L.xmph_site_pair_cost(0) == -1 and L.xmph_site_pair_cost(65536) == -1


_REF_EXE = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "fastdnamp_ref")


@pytest.mark.skipif(not os.path.exists(_REF_EXE), reason="oracle/_ref not built (needs /root/reference)")
def test_partition_bound_equals_compiled_reference_on_random_alignments(tmp_path):
    """Fuzz: random alignments with ambiguity codes, duplicated columns (weights > 1) and all three
    strategies through the reference's own program (--onlycomputelowerbound=Y) and through host/partbound.c."""
    import subprocess
    rng = np.random.default_rng(int(os.environ.get("XMP_FUZZ_SEED", "20261020")))      # one-off wider sweeps: other seeds / more cases
    alphabet = np.array(list("ACGTACGTACGTACGTACGTRYSWKMBDHVN-?"))
    checked = 0
    for case in range(int(os.environ.get("XMP_FUZZ_CASES", "30"))):
        n, sites = int(rng.integers(5, 13)), int(rng.integers(6, 70))
        base = rng.integers(0, len(alphabet), (n, sites))
        for t in range(1, n):                       # related rows, so that columns are informative
            keep = rng.random(sites) < 0.55
            base[t, keep] = base[rng.integers(0, t), keep]
        cols = rng.integers(0, sites, sites + int(rng.integers(0, sites)))      # repeats give weights > 1
        f = tmp_path / ("pb%d.phy" % case)
        f.write_text("%d %d\n" % (n, len(cols)) + "".join("T%-9d%s\n" % (t, "".join(alphabet[base[t, cols]])) for t in range(n)))
        letters, strategy = [("-Bp", 0), ("-BpC", 1), ("-BpL", 2), ("-Bdp", 0)][case % 4]
        stale = [2, 1, 0][case % 3]
        d = tmp_path / ("run%d" % case)
        d.mkdir()
        try:
            r = subprocess.run([_REF_EXE, "--tbr=N", "--seed=1", "--displaynewbounds=n", "--displaynewtrees=n", "--updatesecs=0",
                                "--onlycomputelowerbound=Y", "--pbmaxiters=%d" % stale, letters, str(f)], cwd=d, capture_output=True, text=True,
                               timeout=120)
        except subprocess.TimeoutExpired:
            continue                                # the reference does not return on some degenerate inputs (no informative site, --pbmaxiters=0)
        if r.returncode != 0:
            continue                                # e.g. fewer than 4 informative taxa left: nothing to compare
        want = [[int(x) for x in l.split()] for l in (d / "restBound.txt").read_text().splitlines()[1:]]
        opts = xh.OPT_NEXUSFMT | xh.OPT_USEPARTBOUND | (xh.OPT_USEDISTBASESBOUNDREST if "d" in letters else 0)
        try:
            p = xh.Problem(xh.Alignment(str(f)), options=opts, part_strategy=strategy, part_max_stale=stale)
        except xh.HostError:
            continue
        assert [[i, int(p.rest_bound[i])] for i in range(2, p.num_taxa)] == want, (case, letters, stale)
        checked += 1
    assert checked >= 20


@pytest.mark.skipif(not os.path.exists(_REF_EXE), reason="oracle/_ref not built (needs /root/reference)")
def test_preprocessing_equals_compiled_reference_on_random_alignments(tmp_path):
    """Fuzz of everything before the search: random alignments (ambiguity codes, gaps, repeated and constant
    columns, few to many taxa) and random settings (-d, --startcol / --endcol, -Bd / -Bm / -Bdm) through the
    reference's own program and through host/prep.c: site counts, constant weight, taxon order, MST upper bound
    and restBound[] must agree."""
    import re
    import subprocess
    zinit = _REF_EXE + "_zinit"          # -Bm reads an uninitialised local in the reference
    rng = np.random.default_rng(int(os.environ.get("XMP_FUZZ_SEED", "424242")))
    alphabet = np.array(list("ACGTACGTACGTACGTACGTACGTRYSWKMBDHVN-?"))
    checked = 0
    for case in range(int(os.environ.get("XMP_FUZZ_CASES", "48"))):
        n, sites = int(rng.integers(4, 18)), int(rng.integers(4, 70))
        base = rng.integers(0, len(alphabet), (n, sites))
        for t in range(1, n):
            keep = rng.random(sites) < rng.uniform(0.3, 0.95)
            base[t, keep] = base[rng.integers(0, t), keep]
        cols = rng.integers(0, sites, sites + int(rng.integers(0, sites)))
        f = tmp_path / ("pre%d.phy" % case)
        f.write_text("%d %d\n" % (n, len(cols)) + "".join("T%-9d%s\n" % (t, "".join(alphabet[base[t, cols]])) for t in range(n)))
        letters = ["d", "m", "dm"][case % 3]
        flags, opts, kw = ["-B" + letters], xh.OPT_NEXUSFMT | xh.OPT_IMPROVEINITUPPERBOUND, {}
        opts |= (xh.OPT_USEDISTBASESBOUNDREST if "d" in letters else 0) | (xh.OPT_USEMSTBOUNDREST if "m" in letters else 0)
        if case % 4 == 1:
            flags.append("-d")
            opts |= xh.OPT_DISCARDMISSAMBIG
        if case % 5 == 2:
            lo, hi = sorted(int(x) for x in rng.integers(1, len(cols) + 1, 2))
            flags += ["--startcol=%d" % lo, "--endcol=%d" % (hi + 1)]
            kw = {"start_col": lo - 1, "end_col": hi}
        d = tmp_path / ("pre_run%d" % case)
        d.mkdir()
        # no reordered.phy with -d: when every site is discarded the reference writes gigabytes of garbage there
        save = [] if "-d" in flags else ["--savereordereddataset=Y"]
        r = subprocess.run([zinit if "m" in letters and os.path.exists(zinit) else _REF_EXE, "--tbr=N", "--seed=1", "--displaynewbounds=Y",
                            "--displaynewtrees=n", "--updatesecs=0", "--onlycomputelowerbound=Y"] + save + flags + [str(f)],
                           cwd=d, capture_output=True, text=True, timeout=60)
        try:
            p = xh.Problem(xh.Alignment(str(f)), options=opts, **kw)
        except xh.HostError:
            assert r.returncode != 0, (case, flags)           # both refuse (e.g. nothing informative left)
            continue
        if r.returncode != 0:
            continue
        if p.num_sites + p.n_constant_sites + p.n_nonpi_sites == 0:
            continue          # -d discarded every site: the reference then reads uninitialised weights (garbage site counts)
        e = r.stderr
        assert p.n_constant_sites == int(re.search(r"(\d+) constant sites", e).group(1)), (case, flags)
        assert p.n_nonpi_sites == int(re.search(r"(\d+) non-parsimoniously-informative sites", e).group(1)), (case, flags)
        assert p.constant_weight == int(re.search(r"having a total weight of (\d+)", e).group(1)), (case, flags)
        assert p.num_sites == int(re.search(r"(\d+) parsimoniously-informative sites", e).group(1)), (case, flags)
        mm = re.search(r"Improving initial upper bound to (\d+)", e)
        if mm:
            assert p.mst_upper_bound == int(mm.group(1)), (case, flags)
        rp = d / "reordered.phy"
        if rp.exists():
            order = [int(l[1:10]) for l in rp.read_text().splitlines()[1:] if l.startswith("T")]
            assert [int(t) + 1 for t in p.taxon_map] == order, (case, flags)
        want = [[int(x) for x in l.split()] for l in (d / "restBound.txt").read_text().splitlines()[1:]]
        assert [[i, int(p.rest_bound[i])] for i in range(2, p.num_taxa)] == want, (case, flags)
        checked += 1
    assert checked >= 30


def test_large_tree_sets_sort_and_write_like_small_ones(alignment_dir, goldens, tmp_path):
    """The multi-threaded paths of host/trees.c (depth-first sort by radix, block-wise formatting), used from
    ~20,000 / 65,536 trees up, against the plain ones: same order as sorting the DFS keys, same bytes as writing
    the set in small pieces, NEXUS numbering intact."""
    import ctypes as C
    from tests import oracle_lib as ol
    a, p = prepare(alignment_dir, goldens, "mh3")
    n, N = p.num_taxa, 100_000
    rng = np.random.default_rng(11)
    paths = ol.random_paths(rng, N, n)
    paths = np.concatenate([paths, paths[:5000]])                       # duplicates keep their multiplicity
    ordered = p.sort_paths_dfs(paths)
    keys = [bytes(xh.path_to_dfs(x, n)) for x in ordered[::37]]
    assert keys == sorted(keys)
    assert sorted(map(bytes, ordered)) == sorted(map(bytes, paths))
    small = p.sort_paths_dfs(paths[:3000])                              # single-threaded path
    assert [bytes(xh.path_to_dfs(x, n)) for x in small] == sorted(bytes(xh.path_to_dfs(x, n)) for x in paths[:3000])

    L = xh.lib()
    libc = C.CDLL(None)
    libc.fopen.restype = C.c_void_p
    libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
    libc.fclose.argtypes = [C.c_void_p]
    L.xmph_write_trees.argtypes = [C.c_void_p, C.c_int, C.c_uint, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
    tm = np.ascontiguousarray(p.taxon_map, dtype=np.uint32)

    def write(arr, nexus):
        arr = np.ascontiguousarray(arr)
        f = libc.fopen(str(tmp_path / "w.out").encode(), b"w")
        assert L.xmph_write_trees(f, nexus, 118, C.byref(a._a), tm.ctypes.data, arr.ctypes.data, len(arr)) == 0
        libc.fclose(f)
        return (tmp_path / "w.out").read_bytes()

    whole = write(ordered, 0)
    assert whole == b"".join(write(ordered[i:i + 9000], 0) for i in range(0, len(ordered), 9000))
    assert whole.split(b"\n")[4321].decode() == p.path_to_newick(ordered[4321]) + ";"
    nx = write(ordered, 1).split(b"\n")
    first = nx.index(b"\tTREE FASTDNAMP_1 = [&U] " + p.path_to_newick(ordered[0]).encode() + b";")
    assert nx[first + 77776].decode() == "\tTREE FASTDNAMP_77777 = [&U] " + p.path_to_newick(ordered[77776]) + ";"
    assert nx[first + len(ordered)] == b"END;"


def test_threaded_preprocessing_equals_the_single_threaded_one(tmp_path):
    """Recoding by site tiles, run-wise sorting with parallel merges, the distinct-bases bound by site chunks and the tiled
    transpose only start threads on large inputs: 24 taxa x 400,000 sites with ambiguity codes and repeated columns,
    all cores against one thread (which is what every small test runs), with and without -d and a column range."""
    rng = np.random.default_rng(77)
    n, sites = 24, 400_000
    alphabet = np.frombuffer(b"ACGTACGTACGTACGTACGTACGTACGTRYKMSWN-?", dtype=np.uint8)
    base = rng.integers(0, len(alphabet), (n, 60_000))
    for t in range(1, n):
        keep = rng.random(base.shape[1]) < 0.7
        base[t, keep] = base[rng.integers(0, t), keep]
    cols = rng.integers(0, base.shape[1], sites)                      # every column ~7 times: weights > 1, a real merge
    f = tmp_path / "big.phy"
    with open(f, "wb") as fh:
        fh.write(b"%d %d\n" % (n, sites))
        for t in range(n):
            fh.write(b"T%-9d" % t + alphabet[base[t, cols]].tobytes() + b"\n")
    a = xh.Alignment(str(f))
    for kw in ({}, {"options": xh.OPT_NEXUSFMT | xh.OPT_USEDISTBASESBOUNDREST | xh.OPT_IMPROVEINITUPPERBOUND | xh.OPT_DISCARDMISSAMBIG},
               {"start_col": 1234, "end_col": 333_333, "options": xh.OPT_NEXUSFMT | xh.OPT_USEDISTBASESBOUNDREST | xh.OPT_USEMSTBOUNDREST}):
        one, many = xh.Problem(a, threads=1, **kw), xh.Problem(a, threads=0, **kw)
        assert one.num_sites == many.num_sites > 1000
        for field in ("sites", "site_weights", "taxon_map", "rest_bound", "rows", "weights"):
            assert (getattr(one, field) == getattr(many, field)).all(), field
        assert (one.constant_weight, one.n_constant_sites, one.n_nonpi_sites, one.bound, one.mst_upper_bound) == \
               (many.constant_weight, many.n_constant_sites, many.n_nonpi_sites, many.bound, many.mst_upper_bound)
