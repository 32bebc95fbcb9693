"""ctypes access to the CPU oracle (oracle/liboracle.so) and, when present, to the reference's own
Fitch kernels compiled from its sources (oracle/_ref/libxmpref_fitch.so).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "liboracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libxmpref_fitch.so")
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "fastdnamp_ref")
_o = _r = None


class XoProblem(C.Structure):
    _fields_ = [("num_taxa", C.c_uint), ("nbytes", C.c_size_t), ("char_mat", C.c_void_p), ("weights", C.c_void_p),
                ("rest_bound", C.c_void_p), ("taxon_map", C.c_void_p), ("constant_weight", C.c_uint),
                ("bound", C.c_uint), ("max_trees", C.c_uint)]


class XoResult(C.Structure):
    _fields_ = [("score", C.c_uint), ("num_trees", C.c_ulonglong), ("nodes", C.c_ulonglong),
                ("score_calls", C.c_ulonglong), ("newick", C.c_void_p), ("newick_len", C.c_size_t),
                ("paths", C.c_void_p), ("paths_len", C.c_size_t)]


def oracle():
    global _o
    if _o is None:
        L = C.CDLL(ORACLE_SO)
        vp, sz, u = C.c_void_p, C.c_size_t, C.c_uint
        L.xo_fitch_bases.argtypes = [vp, vp, vp, sz]
        L.xo_fitch_bases.restype = None
        L.xo_fitch_bases_and_compare.argtypes = [vp, vp, vp, vp, sz]
        L.xo_fitch_score_weight1.argtypes = [vp, vp, sz]
        L.xo_fitch_score.argtypes = [vp, vp, vp, sz]
        L.xo_fitch_score_stopearly.argtypes = [vp, vp, vp, sz, u, C.c_int, u]
        L.xo_fitch_score_weight1_stopearly.argtypes = [vp, vp, sz, u, u]
        L.xo_fitch_bases_swar64.argtypes = [vp, vp, vp, sz]
        L.xo_fitch_bases_swar64.restype = None
        L.xo_fitch_score_weight1_swar64.argtypes = [vp, vp, sz]
        L.xo_fitch_score_swar64.argtypes = [vp, vp, vp, sz]
        L.xo_tree_sets.argtypes = [vp, sz, vp, u, vp, vp, vp, vp]
        L.xo_search.argtypes = [C.POINTER(XoProblem), C.POINTER(XoResult)]
        L.xo_search_subtree.argtypes = [C.POINTER(XoProblem), vp, u, C.POINTER(XoResult)]
        L.xo_result_free.argtypes = [C.POINTER(XoResult)]
        for f in ("xo_fitch_bases_and_compare", "xo_fitch_score_weight1", "xo_fitch_score", "xo_fitch_score_stopearly",
                  "xo_fitch_score_weight1_stopearly", "xo_fitch_score_weight1_swar64", "xo_fitch_score_swar64", "xo_tree_sets"):
            getattr(L, f).restype = u
        _o = L
    return _o


def have_ref():
    return os.path.exists(REF_SO)


def ref():
    """The reference's own fastc64 / fastc / basic Fitch functions (BYTESPERBLOCK 8 / 4 / 1)."""
    global _r
    if _r is None:
        L = C.CDLL(REF_SO)
        vp, u = C.c_void_p, C.c_uint
        for suffix in ("b2_w8_fastc64", "b2_w4_fastc"):
            getattr(L, "FitchBases_" + suffix).argtypes = [vp, vp, vp, u]
            getattr(L, "FitchBases_" + suffix).restype = None
            getattr(L, "FitchBasesAndCompare_" + suffix).argtypes = [vp, vp, vp, vp, u]
            getattr(L, "FitchScoreWeight1_" + suffix).argtypes = [vp, vp, u]
            getattr(L, "FitchScoreWeight1_stopearly_" + suffix).argtypes = [vp, vp, u, u]
            getattr(L, "FitchScore_" + suffix).argtypes = [vp, vp, vp, u]
            getattr(L, "FitchScore_fw1_" + suffix).argtypes = [vp, vp, vp, u]
            getattr(L, "FitchScore_stopearly_" + suffix).argtypes = [vp, vp, vp, u, u]
            getattr(L, "FitchScore_fw1_stopearly_" + suffix).argtypes = [vp, vp, vp, u, u]
        L.FitchBases_b2_w1_basic.argtypes = [vp, vp, vp, u]
        L.FitchBases_b2_w1_basic.restype = None
        L.FitchScoreWeight1_b2_w1_basic.argtypes = [vp, vp, u]
        L.FitchScore_b2_w1_basic.argtypes = [vp, vp, vp, u]
        _r = L
    return _r


def p(a):
    return a.ctypes.data


def random_sets(rng, nbytes, ambiguity=0.15):
    """Random packed state sets: mostly single states, some ambiguous, never empty."""
    n = nbytes * 2
    single = (1 << rng.integers(0, 4, n)).astype(np.uint8)
    amb = rng.integers(1, 16, n).astype(np.uint8)
    nib = np.where(rng.random(n) < ambiguity, amb, single).astype(np.uint8)
    return (nib[0::2] | (nib[1::2] << 4)).astype(np.uint8)


def random_weights(rng, nsites, heavy_fraction=0.3, wmax=40):
    """Descending weights with a weight-1 tail, as the reference's preprocessing produces."""
    k = int(nsites * heavy_fraction)
    w = np.ones(nsites, dtype=np.uint32)
    if k:
        w[:k] = np.sort(rng.integers(2, wmax + 1, k))[::-1]
    return w


def tree_sets(rows, m, path, weights=None):
    """Oracle MIDPOINT / UP / DOWN sets ([2m-3][nbytes] each) and Fitch length of one partial tree."""
    rows = np.ascontiguousarray(rows, dtype=np.uint8)
    nbytes = rows.shape[1]
    E = 2 * m - 3
    mid = np.zeros((E, nbytes), np.uint8)
    up = np.zeros((E, nbytes), np.uint8)
    down = np.zeros((E, nbytes), np.uint8)
    path = np.ascontiguousarray(path, dtype=np.uint8)
    w = None if weights is None else np.ascontiguousarray(weights, dtype=np.uint32)
    length = oracle().xo_tree_sets(p(rows), nbytes, None if w is None else p(w), m, p(path), p(mid), p(up), p(down))
    return mid, up, down, int(length)


def _problem(problem, bound, max_trees):
    keep = (np.ascontiguousarray(problem.rows, dtype=np.uint8), np.ascontiguousarray(problem.weights, dtype=np.uint32),
            np.ascontiguousarray(problem.rest_bound, dtype=np.uint32), np.ascontiguousarray(problem.taxon_map, dtype=np.uint32))
    rows, w, rb, tm = keep
    pr = XoProblem(problem.num_taxa, rows.shape[1], p(rows), p(w), p(rb), p(tm), problem.constant_weight,
                   problem.bound if bound is None else bound, max_trees)
    return pr, keep


def search(problem, bound=None, max_trees=0x7FFFFFFF):
    """Oracle exact search of a prepared host problem (xmp_b200.hostlib.Problem).
    Returns (score, num_trees, [newick strings in DFS order], nodes, score_calls)."""
    pr, keep = _problem(problem, bound, max_trees)
    res = XoResult()
    rc = oracle().xo_search(C.byref(pr), C.byref(res))
    assert rc == 0
    text = C.string_at(res.newick, res.newick_len).decode() if res.newick_len else ""
    out = (res.score, res.num_trees, text.split(), res.nodes, res.score_calls)
    oracle().xo_result_free(C.byref(res))
    return out


def search_subtree(problem, path, m, bound):
    """Oracle search below one partial tree.  Returns (bound afterwards, paths of the trees found at
    that bound [k][num_taxa], nodes)."""
    pr, keep = _problem(problem, bound, 0x7FFFFFFF)
    res = XoResult()
    path = np.ascontiguousarray(path, dtype=np.uint8)
    rc = oracle().xo_search_subtree(C.byref(pr), p(path), m, C.byref(res))
    assert rc == 0
    n = problem.num_taxa
    paths = np.frombuffer(C.string_at(res.paths, res.paths_len), np.uint8).reshape(-1, n).copy() if res.paths_len else np.zeros((0, n), np.uint8)
    out = (res.score, paths, res.nodes)
    oracle().xo_result_free(C.byref(res))
    return out


def random_paths(rng, n_trees, m, stride=None):
    """Random stepwise-addition topologies on taxa 0..m-1 as insertion paths (byte t = edge of taxon t)."""
    stride = stride or m
    paths = np.zeros((n_trees, stride), dtype=np.uint8)
    for t in range(3, m):
        paths[:, t] = rng.integers(0, 2 * t - 3, n_trees)
    return paths
