"""Pins the CPU oracle (oracle/*.c): against the reference's own compiled Fitch code when
oracle/_ref is present, against the committed known-answer vectors generated from it, and against the
end-to-end goldens produced by running the reference program (tests/golden/datasets.json).  No GPU."""
import gzip
import hashlib
import io
import json
import os

import numpy as np
import pytest

from tests import oracle_lib as ol
from tests.conftest import GOLDEN, golden_output, golden_trees
from tests.newick import canonical_newick
from tests.test_host import prepare

O = ol.oracle()


def test_kat_vectors():
    cases = json.load(open(os.path.join(GOLDEN, "fitch_kat.json")))
    assert len(cases) >= 20
    for c in cases:
        n = c["nbytes"]
        s1, s2, prev = (np.frombuffer(bytes.fromhex(c[k]), np.uint8).copy() for k in ("s1", "s2", "prev"))
        w = np.array(c["weights"], np.uint32)
        dest = np.zeros(n, np.uint8)
        O.xo_fitch_bases(ol.p(s1), ol.p(s2), ol.p(dest), n)
        assert dest.tobytes().hex() == c["bases"]
        d2 = np.zeros(n, np.uint8)
        assert int(O.xo_fitch_bases_and_compare(ol.p(s1), ol.p(s2), ol.p(d2), ol.p(prev), n) != 0) == c["changed_vs_prev"]
        assert int(O.xo_fitch_bases_and_compare(ol.p(s1), ol.p(s2), ol.p(d2), ol.p(dest), n) != 0) == c["changed_vs_self"]
        assert O.xo_fitch_score_weight1(ol.p(s1), ol.p(s2), n) == c["score_weight1"]
        assert O.xo_fitch_score(ol.p(s1), ol.p(s2), ol.p(w), n) == c["score"]
        assert O.xo_fitch_score_stopearly(ol.p(s1), ol.p(s2), ol.p(w), n, 8, 1, 0x7FFFFFFF) == c["score_fw1"]
        assert O.xo_fitch_score_stopearly(ol.p(s1), ol.p(s2), ol.p(w), n, 8, 1, c["limit"]) == c["score_fw1_stopearly_w8"]
        assert O.xo_fitch_score_weight1_stopearly(ol.p(s1), ol.p(s2), n, 8, c["score_weight1"] // 2) == c["score_weight1_stopearly_w8"]
        if n:
            d3 = np.zeros(n, np.uint8)
            O.xo_fitch_bases_swar64(ol.p(s1), ol.p(s2), ol.p(d3), n // 8)
            assert (d3 == dest).all()
            assert O.xo_fitch_score_weight1_swar64(ol.p(s1), ol.p(s2), n // 8) == c["score_weight1"]
            assert O.xo_fitch_score_swar64(ol.p(s1), ol.p(s2), ol.p(w), n // 8) == c["score"]


@pytest.mark.skipif(not ol.have_ref(), reason="oracle/_ref not built (needs /root/reference)")
def test_oracle_equals_compiled_reference_on_random_inputs():
    R = ol.ref()
    rng = np.random.default_rng(11)
    for trial in range(3000):
        nbytes = int(rng.integers(0, 12)) * 8
        s1, s2, prev = (ol.random_sets(rng, nbytes, rng.choice([0.0, 0.1, 0.5])) for _ in range(3))
        w = ol.random_weights(rng, nbytes * 2, heavy_fraction=rng.choice([0.0, 0.4, 1.0]), wmax=int(rng.choice([2, 9, 3000])))
        n8, n4 = nbytes // 8, nbytes // 4
        want = np.zeros(nbytes, np.uint8)
        got = np.zeros(nbytes, np.uint8)
        R.FitchBases_b2_w8_fastc64(ol.p(s1), ol.p(s2), ol.p(want), n8)
        O.xo_fitch_bases(ol.p(s1), ol.p(s2), ol.p(got), nbytes)
        assert (want == got).all()
        for fam, nb in (("b2_w8_fastc64", n8), ("b2_w4_fastc", n4)):
            assert getattr(R, "FitchScoreWeight1_" + fam)(ol.p(s1), ol.p(s2), nb) == O.xo_fitch_score_weight1(ol.p(s1), ol.p(s2), nbytes)
            assert getattr(R, "FitchScore_" + fam)(ol.p(s1), ol.p(s2), ol.p(w), nb) == O.xo_fitch_score(ol.p(s1), ol.p(s2), ol.p(w), nbytes)
        assert R.FitchScore_b2_w1_basic(ol.p(s1), ol.p(s2), ol.p(w), nbytes) == O.xo_fitch_score(ol.p(s1), ol.p(s2), ol.p(w), nbytes)
        assert O.xo_fitch_score_swar64(ol.p(s1), ol.p(s2), ol.p(w), n8) == O.xo_fitch_score(ol.p(s1), ol.p(s2), ol.p(w), nbytes)
        limit = int(rng.integers(0, 40))
        assert R.FitchScore_fw1_stopearly_b2_w8_fastc64(ol.p(s1), ol.p(s2), ol.p(w), n8, limit) == \
            O.xo_fitch_score_stopearly(ol.p(s1), ol.p(s2), ol.p(w), nbytes, 8, 1, limit)
        assert R.FitchScore_stopearly_b2_w4_fastc(ol.p(s1), ol.p(s2), ol.p(w), n4, limit) == \
            O.xo_fitch_score_stopearly(ol.p(s1), ol.p(s2), ol.p(w), nbytes, 4, 0, limit)
        assert int(R.FitchBasesAndCompare_b2_w8_fastc64(ol.p(s1), ol.p(s2), ol.p(want), ol.p(prev), n8) != 0) == \
            int(O.xo_fitch_bases_and_compare(ol.p(s1), ol.p(s2), ol.p(got), ol.p(prev), nbytes) != 0)


SEARCH_CASES = ["its10", "its12", "its14", "primate", "quick2", "53humans.12", "53humans.16", "mt-10", "EukaryotesrDNA", "e1", "e4", "e6"]


@pytest.mark.parametrize("name", SEARCH_CASES)
def test_oracle_search_reproduces_reference_output(alignment_dir, goldens, name):
    """Score, tree count, canonical tree set and -- through the host writer -- the reference's result
    file byte for byte (for mt-10 that file is the reference's own golden build-vswin32/mt-10.tre)."""
    g = goldens[name]
    a, p = prepare(alignment_dir, goldens, name)
    score, n, trees, nodes, calls = ol.search(p)
    assert (score, n) == (g["score"], g["num_trees"])
    # same depth-first order, same bound updates => the very same nodes; the reference's counter also
    # counts the two FitchScore calls of ConstructBaseTree (src/common.c:623-637)
    assert nodes == g["ref_bandb_nodes"] and calls + 2 == g["ref_score_calls"]
    canon = sorted(canonical_newick(t) for t in trees)
    assert canon == golden_trees(name)
    blob = ("\n".join(canon) + "\n").encode() if canon else b""
    assert hashlib.sha256(blob).hexdigest() == g["canonical_sha256"]
    # the oracle walks the reference's DFS order, so its text is the reference's tree block
    body = "".join("\tTREE FASTDNAMP_%d = [&U] %s\n" % (i + 1, t) for i, t in enumerate(trees))
    assert body.encode() in golden_output(name)


def test_oracle_search_under_the_partition_bound_with_a_tree_cap(alignment_dir, goldens):
    """53humans.24 -Bp -m 1000: restBound[] from host/partbound.c drives the oracle search through exactly the
    reference's nodes (same counters), 1,403,325 optimal trees are counted, and the first 1000 in depth-first
    order are the ones the reference kept (src/yanbader.c:272-293) -- its result file byte for byte."""
    g = goldens["53humans.24"]
    a, p = prepare(alignment_dir, goldens, "53humans.24")
    assert [[i, int(p.rest_bound[i])] for i in range(2, p.num_taxa)] == g["rest_bound"]
    score, n, trees, nodes, calls = ol.search(p, max_trees=1000)
    assert (score, n, len(trees)) == (g["score"], g["num_trees_total"], 1000) and n == 1403325
    assert nodes == g["ref_bandb_nodes"] and calls + 2 == g["ref_score_calls"]
    assert sorted(canonical_newick(t) for t in trees) == golden_trees("53humans.24")
    body = "".join("\tTREE FASTDNAMP_%d = [&U] %s\n" % (i + 1, t) for i, t in enumerate(trees))
    assert body.encode() in golden_output("53humans.24")


def test_mt10_golden_is_the_reference_repo_file(goldens):
    """tests/golden/trees/mt-10.out.gz equals reference build-vswin32/mt-10.tre (CRLF stripped); the sha
    below was taken from that file when the fixture was generated."""
    out = golden_output("mt-10")
    assert hashlib.sha256(out).hexdigest() == goldens["mt-10"]["output_sha256"]
    assert b"Each tree has score 16179" in out and out.count(b"TREE FASTDNAMP_") == 1
    ref = "/root/reference/build-vswin32/mt-10.tre"
    if os.path.exists(ref):
        assert open(ref, "rb").read().replace(b"\r", b"") == out


def test_too_low_bound_finds_nothing(alignment_dir, goldens):
    a, p = prepare(alignment_dir, goldens, "its10")
    score, n, trees, _, _ = ol.search(p, bound=goldens["its10"]["score"] - 1)
    assert n == 0 and trees == [] and score == goldens["its10"]["score"] - 1


_REF_EXE = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "fastdnamp_ref")


@pytest.mark.skipif(not os.path.exists(_REF_EXE), reason="oracle/_ref not built (needs /root/reference)")
def test_oracle_search_equals_compiled_reference_on_random_alignments(tmp_path):
    """Fuzz of the whole pipeline on small random alignments (ambiguity codes, repeated columns): the
    reference's own program (raw Newick output, MEASURE counters) against host preprocessing + oracle search:
    same length, same trees in the same order, same number of BranchAndBound() and FitchScore() calls."""
    import re
    import subprocess
    from xmp_b200 import hostlib as xh
    rng = np.random.default_rng(int(os.environ.get("XMP_FUZZ_SEED", "777")))      # one-off wider sweeps: other seeds / more cases
    alphabet = np.array(list("ACGTACGTACGTACGTACGTRYSWKMBDHVN-?"))
    checked = 0
    for case in range(int(os.environ.get("XMP_FUZZ_CASES", "24"))):
        n, sites = int(rng.integers(4, 10)), int(rng.integers(6, 50))
        base = rng.integers(0, len(alphabet), (n, sites))
        for t in range(1, n):
            keep = rng.random(sites) < 0.6
            base[t, keep] = base[rng.integers(0, t), keep]
        cols = rng.integers(0, sites, sites * 2)
        f = tmp_path / ("fz%d.phy" % case)
        f.write_text("%d %d\n" % (n, len(cols)) + "".join("T%-9d%s\n" % (t, "".join(alphabet[base[t, cols]])) for t in range(n)))
        d = tmp_path / ("fz_run%d" % case)
        d.mkdir()
        r = subprocess.run([_REF_EXE, "--tbr=N", "--seed=1", "--displaynewbounds=n", "--displaynewtrees=n", "--updatesecs=0", "-r",
                            "-o", "out.tre", str(f)], cwd=d, capture_output=True, text=True, timeout=120)
        try:
            p = xh.Problem(xh.Alignment(str(f)))          # default settings, like the reference run: -Bd, MST upper bound
        except xh.HostError:
            continue
        if r.returncode != 0:
            continue                                # (an alignment without any informative site is searched like any other)
        want = (d / "out.tre").read_text().split()
        score, n_trees, trees, nodes, calls = ol.search(p)
        assert trees == want, case
        assert n_trees == len(want), case
        assert nodes == int(re.search(r"BranchAndBound\(\) executed\s+(\d+)", r.stderr).group(1)), case
        assert calls + 2 == int(re.search(r"ScoreFitch\(\) executed\s+(\d+)", r.stderr).group(1)), case
        checked += 1
    assert checked >= 15


@pytest.mark.skipif(not os.path.exists(_REF_EXE), reason="oracle/_ref not built (needs /root/reference)")
def test_alignment_without_informative_sites_is_searched_like_the_reference_does(tmp_path):
    """No parsimony-informative site: the reference's search runs on zero blocks and returns every tree, at the constant
    weight, in depth-first order (and honours -m).  Here the problem keeps one block of padding and takes the same path."""
    import subprocess
    from xmp_b200 import hostlib as xh
    for k, text in enumerate(("5 4\nA         ACGT\nB         ACGT\nC         ACGT\nD         ACGA\nE         ACGT\n",
                              "6 3\nA         AC-\nB         ANG\nC         ACG\nD         TCG\nE         ACG\nF         ACR\n")):
        f = tmp_path / ("zero%d.phy" % k)
        f.write_text(text)
        for extra in ([], ["-m", "4"]):
            r = subprocess.run([_REF_EXE, "--tbr=N", "--seed=1", "--displaynewbounds=n", "--displaynewtrees=n", "--updatesecs=0", "-r",
                                "-o", "out.tre"] + extra + [str(f)], cwd=tmp_path, capture_output=True, text=True, timeout=60)
            assert r.returncode == 0, r.stderr
            want = (tmp_path / "out.tre").read_text().split()
            p = xh.Problem(xh.Alignment(str(f)))
            assert (p.num_sites, p.n_blocks) == (0, 1) and (p.rows == 0xFF).all()
            score, n_trees, trees, _, _ = ol.search(p, max_trees=int(extra[1]) if extra else 0x7FFFFFFF)
            assert score == p.constant_weight
            assert trees == want and n_trees >= len(want)
            assert len(want) == (4 if extra else {5: 15, 6: 105}[p.num_taxa])


@pytest.mark.skipif(not os.path.exists(_REF_EXE), reason="oracle/_ref not built (needs /root/reference)")
def test_nexus_files_equal_the_compiled_references_on_random_alignments(tmp_path):
    """The result FILE, byte for byte, in the default NEXUS format: random alignments with awkward taxon labels (inner
    blanks, punctuation, trailing blanks that the reader trims, src/common.c:2896-2936) and a random -m, through the
    reference's own program and through host reader + preprocessing + oracle search + host writer (host/trees.c)."""
    import ctypes as C
    import subprocess
    from xmp_b200 import hostlib as xh
    L, libc = xh.lib(), C.CDLL(None)
    libc.fopen.restype = C.c_void_p
    libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
    libc.fclose.argtypes = [C.c_void_p]
    L.xmph_write_trees.argtypes = [C.c_void_p, C.c_int, C.c_uint, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
    rng = np.random.default_rng(int(os.environ.get("XMP_FUZZ_SEED", "31337")))
    alphabet = np.array(list("ACGTACGTACGTACGTACGTRYSWKMBDHVN-?"))
    label_chars = list("abcXYZ019_.-#'()[]:, ")
    checked = 0
    for case in range(int(os.environ.get("XMP_FUZZ_CASES", "30"))):
        n, sites = int(rng.integers(4, 9)), int(rng.integers(6, 40))
        base = rng.integers(0, len(alphabet), (n, sites))
        for t in range(1, n):
            keep = rng.random(sites) < 0.6
            base[t, keep] = base[rng.integers(0, t), keep]
        labels = []
        for t in range(n):
            k = int(rng.integers(1, 11))
            lab = "".join(rng.choice(label_chars, k)).strip() or "t"
            labels.append(("%d%s" % (t, lab))[:10].ljust(10))           # unique, exactly the 10-character field
        f = tmp_path / ("nx%d.phy" % case)
        f.write_text("%d %d\n" % (n, sites) + "".join(labels[t] + "".join(alphabet[base[t]]) + "\n" for t in range(n)))
        cap = [0x7FFFFFFF, 1, 3][case % 3]
        extra = [] if cap == 0x7FFFFFFF else ["-m", str(cap)]
        d = tmp_path / ("nx_run%d" % case)
        d.mkdir()
        r = subprocess.run([_REF_EXE, "--tbr=N", "--seed=1", "--displaynewbounds=n", "--displaynewtrees=n", "--updatesecs=0",
                            "-o", "out.nex"] + extra + [str(f)], cwd=d, capture_output=True, text=True, timeout=120)
        if r.returncode != 0:
            continue
        want = (d / "out.nex").read_bytes()
        a = xh.Alignment(str(f))
        p = xh.Problem(a)
        score, n_trees, trees, _, _ = ol.search(p, max_trees=cap)
        # the same trees as insertion paths: everything the oracle finds at the optimum below the 3-taxon tree, in DFS order
        _, all_paths, _ = ol.search_subtree(p, np.zeros(p.num_taxa, np.uint8), 3, score)
        paths = np.ascontiguousarray(p.sort_paths_dfs(all_paths)[:len(trees)])
        assert p.newick_list(paths) == trees and n_trees == len(all_paths)
        tm = np.ascontiguousarray(p.taxon_map, dtype=np.uint32)
        fh = libc.fopen(str(d / "ours.nex").encode(), b"w")
        assert L.xmph_write_trees(fh, 1, score, C.byref(a._a), tm.ctypes.data, paths.ctypes.data, len(paths)) == 0
        libc.fclose(fh)
        assert (d / "ours.nex").read_bytes() == want, case
        checked += 1
    assert checked >= 20

