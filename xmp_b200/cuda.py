"""ctypes mirror of include/xmp_cuda.h (libxmp_b200.so, hand-written sm_100a CUDA behind a C ABI).

There is no fallback of any kind here: if the library is not built, or there is no B200-class device,
every entry point raises.  Device buffers may be torch CUDA tensors (their ``data_ptr()`` is passed
through) -- torch is plumbing for memory and streams, the kernels are ours.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libxmp_b200.so")
_LIB = None
NO_LIMIT = 0x7FFFFFFF
SEARCH_NO_DIVE = 1
SEARCH_NO_FUSE = 2
SEARCH_NO_EAGER = 4
MAX_TAXA = 100

EXPORTS = [
    "xmp_cuda_last_error", "xmp_cuda_device_count", "xmp_cuda_init", "xmp_cuda_destroy", "xmp_cuda_set_stream",
    "xmp_cuda_synchronize",
    "FitchBases_b2_w16_cuda", "FitchBasesAndCompare_b2_w16_cuda", "FitchScoreWeight1_b2_w16_cuda",
    "FitchScoreWeight1_stopearly_b2_w16_cuda", "FitchScore_b2_w16_cuda", "FitchScore_stopearly_b2_w16_cuda",
    "FitchScore_fw1_b2_w16_cuda", "FitchScore_fw1_stopearly_b2_w16_cuda",
    "xmp_cuda_load", "xmp_cuda_load_device_rows", "xmp_cuda_build_midpoints", "xmp_cuda_score_batch",
    "xmp_cuda_score_insertions_host",
    "xmp_cuda_search_reserve", "xmp_cuda_search_begin", "xmp_cuda_search_run", "xmp_cuda_search_get_bound", "xmp_cuda_search_set_bound",
    "xmp_cuda_search_pending", "xmp_cuda_search_export_jobs", "xmp_cuda_search_import_jobs",
    "xmp_cuda_search_result", "xmp_cuda_search_stats", "xmp_cuda_int_pipe_peak",
    "xmp_cuda_search_stored", "xmp_cuda_paths_keep_first_dfs",
    "xmp_cuda_exchange_handle", "xmp_cuda_exchange_connect", "xmp_cuda_exchange_connect_local", "xmp_cuda_exchange_disconnect",
    "xmp_cuda_exchange_reset", "xmp_cuda_search_steal",
]
EXCHANGE_HANDLE_BYTES = 64
STEAL_GOT_WORK = 1
STEAL_FINISHED = 2


class XmpCudaError(RuntimeError):
    pass


class SearchParams(C.Structure):
    _fields_ = [("bound", C.c_uint), ("max_trees", C.c_uint), ("mem_budget", C.c_size_t),
                ("shard_index", C.c_uint), ("shard_count", C.c_uint), ("split_level", C.c_uint), ("flags", C.c_uint)]


class SearchStats(C.Structure):
    _fields_ = [("nodes", C.c_ulonglong), ("score_calls", C.c_ulonglong), ("children", C.c_ulonglong),
                ("full_trees", C.c_ulonglong), ("site_merges", C.c_ulonglong), ("launches", C.c_ulonglong),
                ("iterations", C.c_ulonglong), ("seconds", C.c_double),
                ("nodes_by_depth", C.c_ulonglong * (MAX_TAXA + 1)), ("edge_evals_by_depth", C.c_ulonglong * (MAX_TAXA + 1)),
                ("batches_by_depth", C.c_ulonglong * (MAX_TAXA + 1)),
                ("steals_served", C.c_ulonglong), ("steals_received", C.c_ulonglong), ("begin_seconds", C.c_double)]

    def as_dict(self):
        return {k: (list(getattr(self, k)) if k.endswith("_by_depth") else getattr(self, k)) for k, _ in self._fields_}


def lib():
    """Loads libxmp_b200.so.  Raises if it has not been built -- never substitutes anything else."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise XmpCudaError("xmp_b200/libxmp_b200.so is missing: run `make cuda` (or __graft_entry__.build()). "
                               "There is no CPU fallback.")
        _LIB = bind(C.CDLL(LIB_PATH))
    return _LIB


def bind(L, partial=False):
    """Declares the argument types of every entry point of include/xmp_cuda.h on a loaded library.  partial: skip what
    the library does not export (the CPU tests bind their test double of the search entry points this way)."""
    vp, u, sz, ull = C.c_void_p, C.c_uint, C.c_size_t, C.c_ulonglong
    pu = C.POINTER(u)
    sig = {
        "xmp_cuda_last_error": ([], C.c_char_p), "xmp_cuda_device_count": ([], C.c_int),
        "xmp_cuda_init": ([C.c_int, C.POINTER(vp)], C.c_int), "xmp_cuda_destroy": ([vp], None),
        "xmp_cuda_set_stream": ([vp, vp], C.c_int), "xmp_cuda_synchronize": ([vp], C.c_int),
        "xmp_cuda_load": ([vp, vp, vp, u, u, vp, u], C.c_int), "xmp_cuda_load_device_rows": ([vp, vp, vp, u, u, vp, u], C.c_int),
        "xmp_cuda_build_midpoints": ([vp, vp, sz, u, u, vp, vp], C.c_int),
        "xmp_cuda_score_batch": ([vp, vp, sz, u, u, vp, u, vp, vp, vp], C.c_int),
        "xmp_cuda_score_insertions_host": ([vp, vp, sz, u, u, vp], C.c_int),
        "xmp_cuda_search_reserve": ([vp, u, u, sz], C.c_int), "xmp_cuda_search_begin": ([vp, C.POINTER(SearchParams)], C.c_int),
        "xmp_cuda_search_run": ([vp, u, C.POINTER(C.c_int)], C.c_int),
        "xmp_cuda_search_get_bound": ([vp, pu], C.c_int), "xmp_cuda_search_set_bound": ([vp, u], C.c_int),
        "xmp_cuda_search_pending": ([vp, vp], C.c_int), "xmp_cuda_search_export_jobs": ([vp, u, vp, pu], C.c_int),
        "xmp_cuda_search_import_jobs": ([vp, vp, u], C.c_int),
        "xmp_cuda_search_result": ([vp, pu, C.POINTER(ull), vp, sz], C.c_int),
        "xmp_cuda_search_stats": ([vp, C.POINTER(SearchStats)], C.c_int), "xmp_cuda_search_stored": ([vp, C.POINTER(ull)], C.c_int),
        "xmp_cuda_paths_keep_first_dfs": ([vp, sz, u, sz], C.c_int), "xmp_cuda_int_pipe_peak": ([vp, C.POINTER(C.c_double)], C.c_int),
        "xmp_cuda_exchange_handle": ([vp, vp], C.c_int), "xmp_cuda_exchange_connect": ([vp, u, u, vp], C.c_int),
        "xmp_cuda_exchange_connect_local": ([vp, u, u, vp], C.c_int), "xmp_cuda_exchange_disconnect": ([vp], C.c_int),
        "xmp_cuda_exchange_reset": ([vp], C.c_int), "xmp_cuda_search_steal": ([vp, C.POINTER(C.c_int)], C.c_int),
        "FitchBases_b2_w16_cuda": ([vp, vp, vp, u], None), "FitchBasesAndCompare_b2_w16_cuda": ([vp, vp, vp, vp, u], u),
        "FitchScoreWeight1_b2_w16_cuda": ([vp, vp, u], u), "FitchScoreWeight1_stopearly_b2_w16_cuda": ([vp, vp, u, u], u),
        "FitchScore_b2_w16_cuda": ([vp, vp, vp, u], u), "FitchScore_stopearly_b2_w16_cuda": ([vp, vp, vp, u, u], u),
        "FitchScore_fw1_b2_w16_cuda": ([vp, vp, vp, u], u), "FitchScore_fw1_stopearly_b2_w16_cuda": ([vp, vp, vp, u, u], u),
    }
    assert sorted(sig) == sorted(EXPORTS)
    for name, (args, res) in sig.items():
        if partial and not hasattr(L, name):
            continue
        f = getattr(L, name)
        f.argtypes, f.restype = args, res
    return L


def _ptr(x):
    """Device or host address of a torch tensor / numpy array / int / None."""
    if x is None:
        return None
    if isinstance(x, int):
        return x
    if isinstance(x, np.ndarray):
        return x.ctypes.data
    return x.data_ptr()


def _check(rc):
    if rc != 0:
        raise XmpCudaError("libxmp_b200 error %d: %s" % (rc, lib().xmp_cuda_last_error().decode()))


class Context:
    """One device context (reference analogue: the global TreeData of a process, src/common.c:42)."""

    def __init__(self, device=0, stream=None):
        L = lib()
        h = C.c_void_p()
        _check(L.xmp_cuda_init(device, C.byref(h)))
        self._h = h
        self._destroy = L.xmp_cuda_destroy     # kept for interpreter shutdown, when module globals are gone
        self.device = device
        self.num_taxa = self.n_blocks = 0
        if stream is not None:
            self.set_stream(stream)

    def close(self):
        if getattr(self, "_h", None):
            self._destroy(self._h)
            self._h = None

    __del__ = close

    def set_stream(self, stream):
        """stream: a raw cudaStream_t value (e.g. torch.cuda.current_stream().cuda_stream) or None."""
        _check(lib().xmp_cuda_set_stream(self._h, stream))

    def synchronize(self):
        _check(lib().xmp_cuda_synchronize(self._h))

    # ---- problem ----
    def load(self, rows, weights=None, rest_bound=None, constant_weight=0):
        """rows: host uint8 [num_taxa][n_blocks*16] (numpy), or a torch CUDA uint8 tensor of that shape."""
        num_taxa, nbytes = rows.shape
        assert nbytes % 16 == 0
        w = None if weights is None else np.ascontiguousarray(weights, dtype=np.uint32)
        rb = None if rest_bound is None else np.ascontiguousarray(rest_bound, dtype=np.uint32)
        if isinstance(rows, np.ndarray):
            rows = np.ascontiguousarray(rows, dtype=np.uint8)
            _check(lib().xmp_cuda_load(self._h, _ptr(rows), _ptr(w), num_taxa, nbytes // 16, _ptr(rb), constant_weight))
        else:
            assert rows.is_cuda and rows.is_contiguous()
            self._rows_keepalive = rows
            _check(lib().xmp_cuda_load_device_rows(self._h, _ptr(rows), _ptr(w), num_taxa, nbytes // 16, _ptr(rb), constant_weight))
        self.num_taxa, self.n_blocks = num_taxa, nbytes // 16

    def load_problem(self, problem):
        self.load(problem.rows, problem.weights, problem.rest_bound, problem.constant_weight)

    # ---- batched scoring ----
    def build_midpoints(self, paths, m, d_mid, d_tree_len=None):
        paths = np.ascontiguousarray(paths, dtype=np.uint8)
        _check(lib().xmp_cuda_build_midpoints(self._h, _ptr(paths), paths.shape[1], paths.shape[0], m, _ptr(d_mid), _ptr(d_tree_len)))

    def score_batch(self, d_sets, n_pairs, pairs_per_tree, taxon, d_cost, d_base_score=None, bound=NO_LIMIT,
                    d_survivors=None, d_n_survivors=None):
        _check(lib().xmp_cuda_score_batch(self._h, _ptr(d_sets), n_pairs, pairs_per_tree, taxon, _ptr(d_base_score), bound,
                                          _ptr(d_cost), _ptr(d_survivors), _ptr(d_n_survivors)))

    def score_insertions_host(self, paths, m, out=None):
        paths = np.ascontiguousarray(paths, dtype=np.uint8)
        n = paths.shape[0]
        if out is None:
            out = np.empty((n, 2 * m - 3), dtype=np.uint32)
        _check(lib().xmp_cuda_score_insertions_host(self._h, _ptr(paths), paths.shape[1], n, m, _ptr(out)))
        return out

    def int_pipe_peak(self):
        """G site-operations/s the integer pipe sustains: {'cost_w1', 'merge', 'merge_cost'}."""
        out = (C.c_double * 3)()
        _check(lib().xmp_cuda_int_pipe_peak(self._h, out))
        return {"cost_w1": out[0], "merge": out[1], "merge_cost": out[2]}

    # ---- search ----
    def search_reserve(self, num_taxa, n_blocks, mem_budget=0):
        """Allocates the search's device memory now, so that search_begin does not have to."""
        _check(lib().xmp_cuda_search_reserve(self._h, num_taxa, n_blocks, mem_budget))

    def search_begin(self, bound=NO_LIMIT, max_trees=NO_LIMIT, mem_budget=0, shard_index=0, shard_count=1,
                     split_level=0, flags=0):
        p = SearchParams(bound, max_trees, mem_budget, shard_index, shard_count, split_level, flags)
        _check(lib().xmp_cuda_search_begin(self._h, C.byref(p)))

    def search_run(self, max_batches=0):
        done = C.c_int(0)
        _check(lib().xmp_cuda_search_run(self._h, max_batches, C.byref(done)))
        return bool(done.value)

    def search_get_bound(self):
        b = C.c_uint(0)
        _check(lib().xmp_cuda_search_get_bound(self._h, C.byref(b)))
        return b.value

    def search_set_bound(self, bound):
        _check(lib().xmp_cuda_search_set_bound(self._h, int(bound)))

    def search_pending(self):
        a = np.zeros(101, dtype=np.uint64)
        _check(lib().xmp_cuda_search_pending(self._h, _ptr(a)))
        return a

    def search_export_jobs(self, max_nodes):
        jobs = np.zeros((max_nodes, self.num_taxa), dtype=np.uint8)
        n = C.c_uint(0)
        _check(lib().xmp_cuda_search_export_jobs(self._h, max_nodes, _ptr(jobs), C.byref(n)))
        return jobs[:n.value].copy()

    def search_import_jobs(self, jobs):
        jobs = np.ascontiguousarray(jobs, dtype=np.uint8)
        if len(jobs):
            _check(lib().xmp_cuda_search_import_jobs(self._h, _ptr(jobs), len(jobs)))

    def search_result(self):
        """(optimal length, number of optimal trees, their insertion paths [n_stored][num_taxa]); n_stored is the
        number of optimal trees unless a finite max_trees made the library cut the set back (xmp_cuda_search_stored)."""
        score, n, stored = C.c_uint(0), C.c_ulonglong(0), C.c_ulonglong(0)
        _check(lib().xmp_cuda_search_result(self._h, C.byref(score), C.byref(n), None, 0))
        _check(lib().xmp_cuda_search_stored(self._h, C.byref(stored)))
        paths = np.zeros((stored.value, self.num_taxa), dtype=np.uint8)
        if stored.value:
            _check(lib().xmp_cuda_search_result(self._h, C.byref(score), C.byref(n), _ptr(paths), stored.value))
        return score.value, n.value, paths

    # ---- several GPUs on one search (include/xmp_cuda.h 2d) ----
    def exchange_handle(self):
        h = np.zeros(EXCHANGE_HANDLE_BYTES, dtype=np.uint8)
        _check(lib().xmp_cuda_exchange_handle(self._h, _ptr(h)))
        return h

    def exchange_connect(self, rank, world, handles):
        handles = np.ascontiguousarray(handles, dtype=np.uint8)
        assert handles.size == world * EXCHANGE_HANDLE_BYTES
        _check(lib().xmp_cuda_exchange_connect(self._h, rank, world, _ptr(handles)))

    def exchange_disconnect(self):
        _check(lib().xmp_cuda_exchange_disconnect(self._h))

    def exchange_reset(self):
        _check(lib().xmp_cuda_exchange_reset(self._h))

    def search_steal(self):
        st = C.c_int(0)
        _check(lib().xmp_cuda_search_steal(self._h, C.byref(st)))
        return st.value

    def search_stats(self):
        st = SearchStats()
        _check(lib().xmp_cuda_search_stats(self._h, C.byref(st)))
        return st.as_dict()


def paths_keep_first_dfs(paths, keep):
    """The first `keep` insertion paths in the reference's depth-first order, in that order (host only)."""
    paths = np.ascontiguousarray(paths, dtype=np.uint8).copy()
    if len(paths):
        _check(lib().xmp_cuda_paths_keep_first_dfs(_ptr(paths), len(paths), paths.shape[1], keep))
    return paths[:keep]


def search(problem, device=0, bound=None, **kw):
    """Whole exact search of a prepared host problem on one GPU.  Returns (score, n_trees, paths, stats)."""
    ctx = Context(device)
    try:
        ctx.load_problem(problem)
        ctx.search_begin(bound=problem.bound if bound is None else bound, **kw)
        while not ctx.search_run(0):
            pass
        score, n, paths = ctx.search_result()
        return score, n, paths, ctx.search_stats()
    finally:
        ctx.close()


# ---- the per-pair family, numpy in / numpy out (parity tests) ----
def _pair_args(*arrays):
    out = []
    for a in arrays:
        a = np.ascontiguousarray(a, dtype=np.uint8)
        assert a.size % 16 == 0
        out.append(a)
    return out


def fitch_bases(s1, s2):
    s1, s2 = _pair_args(s1, s2)
    dest = np.empty_like(s1)
    lib().FitchBases_b2_w16_cuda(_ptr(s1), _ptr(s2), _ptr(dest), s1.size // 16)
    return dest


def fitch_bases_and_compare(s1, s2, prev):
    s1, s2, prev = _pair_args(s1, s2, prev)
    dest = np.empty_like(s1)
    ch = lib().FitchBasesAndCompare_b2_w16_cuda(_ptr(s1), _ptr(s2), _ptr(dest), _ptr(prev), s1.size // 16)
    return dest, int(ch)


def fitch_score_weight1(s1, s2, max_score=None):
    s1, s2 = _pair_args(s1, s2)
    if max_score is None:
        return int(lib().FitchScoreWeight1_b2_w16_cuda(_ptr(s1), _ptr(s2), s1.size // 16))
    return int(lib().FitchScoreWeight1_stopearly_b2_w16_cuda(_ptr(s1), _ptr(s2), s1.size // 16, max_score))


def fitch_score(s1, s2, weights, fw1=False, max_score=None):
    s1, s2 = _pair_args(s1, s2)
    w = np.ascontiguousarray(weights, dtype=np.uint32)
    assert w.size == s1.size * 2
    L, n = lib(), s1.size // 16
    if fw1:
        if max_score is None:
            return int(L.FitchScore_fw1_b2_w16_cuda(_ptr(s1), _ptr(s2), _ptr(w), n))
        return int(L.FitchScore_fw1_stopearly_b2_w16_cuda(_ptr(s1), _ptr(s2), _ptr(w), n, max_score))
    if max_score is None:
        return int(L.FitchScore_b2_w16_cuda(_ptr(s1), _ptr(s2), _ptr(w), n))
    return int(L.FitchScore_stopearly_b2_w16_cuda(_ptr(s1), _ptr(s2), _ptr(w), n, max_score))
