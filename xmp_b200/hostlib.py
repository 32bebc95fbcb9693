"""ctypes mirror of host/xmp_host.h (PHYLIP reader, preprocessing, tree text).

Names and meaning follow the C header, which in turn follows the reference's InitTreeData /
OutputResults (reference src/common.c:1113-1224, :371-454).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
_LIB = None

OPT_NEXUSFMT = 1
OPT_DISCARDMISSAMBIG = 2
OPT_DISPLAYNEWBOUNDS = 4
OPT_DISPLAYNEWTREES = 8
OPT_QUIET = 16
OPT_USELOADRESTBOUND = 32
OPT_USEDISTBASESBOUNDREST = 64
OPT_USEMSTBOUNDREST = 256
OPT_IMPROVEINITUPPERBOUND = 4096
OPT_USEPARTBOUND = 8192
OPT_TBR = 131072
NO_LIMIT = 0x7FFFFFFF


class _Alignment(C.Structure):
    _fields_ = [("num_taxa", C.c_uint), ("seq_len", C.c_uint),
                ("chars", C.POINTER(C.c_ubyte)), ("labels", C.POINTER(C.c_char_p))]


class _Settings(C.Structure):
    _fields_ = [("options", C.c_uint), ("bound", C.c_uint), ("max_trees", C.c_uint),
                ("start_col", C.c_uint), ("end_col", C.c_uint), ("rest_bound_file", C.c_char_p),
                ("part_strategy", C.c_uint), ("part_max_stale", C.c_uint), ("threads", C.c_uint)]


class _Problem(C.Structure):
    _fields_ = [("num_taxa", C.c_uint), ("num_sites", C.c_uint), ("n_blocks", C.c_uint),
                ("rows", C.POINTER(C.c_ubyte)), ("weights", C.POINTER(C.c_uint)),
                ("sites", C.POINTER(C.c_ubyte)), ("site_weights", C.POINTER(C.c_uint)),
                ("taxon_map", C.POINTER(C.c_uint)), ("rest_bound", C.POINTER(C.c_uint)),
                ("constant_weight", C.c_uint), ("n_constant_sites", C.c_uint), ("n_nonpi_sites", C.c_uint),
                ("bound", C.c_uint), ("mst_upper_bound", C.c_uint)]


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_ROOT, "host", "libxmp_host.so")
        if not os.path.exists(path):
            raise RuntimeError("host/libxmp_host.so is missing: run `make host` (or __graft_entry__.build())")
        L = C.CDLL(path)
        L.xmph_last_error.restype = C.c_char_p
        L.xmph_site_pair_cost.argtypes = [C.c_uint]
        L.xmph_path_to_newick.restype = C.c_size_t
        L.xmph_path_to_newick.argtypes = [C.c_void_p, C.c_uint, C.c_void_p, C.c_char_p]
        L.xmph_path_to_dfs.argtypes = [C.c_void_p, C.c_uint, C.c_void_p]
        L.xmph_dfs_to_path.argtypes = [C.c_void_p, C.c_uint, C.c_void_p]
        L.xmph_sort_paths_dfs.argtypes = [C.c_void_p, C.c_size_t, C.c_uint]
        _LIB = L
    return _LIB


class HostError(RuntimeError):
    pass


_libc = C.CDLL(None)
_libc.fopen.restype = C.c_void_p
_libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
_libc.fclose.argtypes = [C.c_void_p]


class Alignment:
    def __init__(self, path):
        L = lib()
        self._a = _Alignment()
        f = _libc.fopen(os.fsencode(path), b"r")
        if not f:
            raise HostError("File '%s' could not be opened for input, aborting." % path)
        try:
            rc = L.xmph_read_phylip(C.c_void_p(f), C.byref(self._a))
        finally:
            _libc.fclose(f)
        if rc:
            raise HostError(L.xmph_last_error().decode())
        self.num_taxa = self._a.num_taxa
        self.seq_len = self._a.seq_len
        self.labels = [self._a.labels[i].decode("latin-1") for i in range(self.num_taxa)]
        self.chars = np.ctypeslib.as_array(self._a.chars, shape=(self.num_taxa, self.seq_len)).copy()

    def __del__(self):
        if getattr(self, "_a", None) is not None and _LIB is not None:
            _LIB.xmph_free_alignment(C.byref(self._a))
            self._a = None


class Problem:
    """The packed search problem (reference: struct TreeData hot fields, src/common.h:152-237)."""

    def __init__(self, alignment, options=None, bound=NO_LIMIT, max_trees=NO_LIMIT, start_col=0,
                 end_col=NO_LIMIT, rest_bound_file=None, part_strategy=0, part_max_stale=2, threads=0):
        L = lib()
        s = _Settings()
        L.xmph_default_settings(C.byref(s))
        if options is not None:
            s.options = options
        s.bound, s.max_trees, s.start_col, s.end_col = bound, max_trees, start_col, end_col
        s.part_strategy, s.part_max_stale, s.threads = part_strategy, part_max_stale, threads
        self._rbf = os.fsencode(rest_bound_file) if rest_bound_file else None
        s.rest_bound_file = self._rbf
        self.alignment = alignment
        self._p = _Problem()
        if L.xmph_prepare(C.byref(alignment._a), C.byref(s), C.byref(self._p)):
            self._p = None
            raise HostError(L.xmph_last_error().decode())
        p = self._p
        self.options = s.options
        self.max_trees = max_trees
        self.num_taxa, self.num_sites, self.n_blocks = p.num_taxa, p.num_sites, p.n_blocks
        as_array = np.ctypeslib.as_array
        nb = max(p.n_blocks, 1)
        self.rows = as_array(p.rows, shape=(p.num_taxa, nb * 16))[:, :p.n_blocks * 16].copy()
        self.weights = as_array(p.weights, shape=(nb * 32,))[:p.n_blocks * 32].copy()
        ns = max(p.num_sites, 1)
        self.sites = as_array(p.sites, shape=(p.num_taxa, ns))[:, :p.num_sites].copy()
        self.site_weights = as_array(p.site_weights, shape=(ns,))[:p.num_sites].copy()
        self.taxon_map = as_array(p.taxon_map, shape=(p.num_taxa,)).copy()
        self.rest_bound = as_array(p.rest_bound, shape=(p.num_taxa,)).copy()
        self.constant_weight = p.constant_weight
        self.n_constant_sites, self.n_nonpi_sites = p.n_constant_sites, p.n_nonpi_sites
        self.bound, self.mst_upper_bound = p.bound, p.mst_upper_bound

    def __del__(self):
        if getattr(self, "_p", None) is not None and _LIB is not None:
            _LIB.xmph_free_problem(C.byref(self._p))
            self._p = None

    # ---- trees ----
    def path_to_newick(self, path):
        path = np.ascontiguousarray(path, dtype=np.uint8)
        buf = C.create_string_buffer(16 * self.num_taxa + 16)
        tm = np.ascontiguousarray(self.taxon_map, dtype=np.uint32)
        n = lib().xmph_path_to_newick(path.ctypes.data, self.num_taxa, tm.ctypes.data, buf)
        return buf.raw[:n].decode()

    def sort_paths_dfs(self, paths):
        paths = np.ascontiguousarray(paths, dtype=np.uint8).copy()
        if len(paths):
            lib().xmph_sort_paths_dfs(paths.ctypes.data, len(paths), self.num_taxa)
        return paths

    def newick_list(self, paths):
        """Trees as "<newick>;" strings in the reference's depth-first discovery order."""
        return [self.path_to_newick(p) + ";" for p in self.sort_paths_dfs(paths)]


def path_to_dfs(path, num_taxa):
    path = np.ascontiguousarray(path, dtype=np.uint8)
    out = np.zeros(num_taxa, dtype=np.uint8)
    lib().xmph_path_to_dfs(path.ctypes.data, num_taxa, out.ctypes.data)
    return out


def dfs_to_path(dfs, num_taxa):
    dfs = np.ascontiguousarray(dfs, dtype=np.uint8)
    out = np.zeros(num_taxa, dtype=np.uint8)
    lib().xmph_dfs_to_path(dfs.ctypes.data, num_taxa, out.ctypes.data)
    return out
