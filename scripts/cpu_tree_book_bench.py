"""CPU measurement of the search's host-side tree bookkeeping under a -m cap (xmp_b200/csrc/tree_book.h), with and
without the threshold pre-filter: N optimal trees on `taxa` taxa arrive in drains of 4 M records (the size of the
device's tree buffer) with max_trees = 1000 -- the shape of 53humans.29 under -Bp -m 1000 (29 taxa, 88.4 M optimal
trees).  Test infrastructure: builds tests/stub/selftest_tree_cap.cpp twice into a scratch directory.

    python scripts/cpu_tree_book_bench.py [million_trees] [taxa]"""
import ctypes as C
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import oracle_lib as ol  # noqa: E402

millions = int(sys.argv[1]) if len(sys.argv) > 1 else 24
n = int(sys.argv[2]) if len(sys.argv) > 2 else 29
trec = 4 + ((n + 3) & ~3)
rng = np.random.default_rng(29)
paths = ol.random_paths(rng, millions * 1_000_000, n, n)
rec = np.zeros((len(paths), trec), np.uint8)
rec[:, :4] = np.full(len(paths), 91, dtype="<u4").view(np.uint8).reshape(-1, 4)
rec[:, 4:4 + n] = paths
drain = 4_000_000
sizes = np.array([min(drain, len(paths) - i) for i in range(0, len(paths), drain)], np.uint32)
bounds = np.full(len(sizes), 91, np.uint32)
results = {}
with tempfile.TemporaryDirectory() as tmp:
    for label, flags in (("with pre-filter (shipped)", []), ("without (before commit 67e64d6)", ["-DXMP_TREEBOOK_NO_PREFILTER"])):
        so = os.path.join(tmp, "lib_%d.so" % len(results))
        subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-pthread"] + flags + ["-o", so,
                        os.path.join(ROOT, "tests", "stub", "selftest_tree_cap.cpp")], check=True)
        L = C.CDLL(so)
        vp = C.c_void_p
        L.xmp_selftest_tree_cap.argtypes = [vp, vp, vp, C.c_uint, C.c_uint, C.c_uint, C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong), vp, C.c_size_t]
        total, stored = C.c_ulonglong(0), C.c_ulonglong(0)
        cap = 5_000_000
        out = np.zeros((cap, n), np.uint8)
        t0 = time.perf_counter()
        rc = L.xmp_selftest_tree_cap(rec.ctypes.data, sizes.ctypes.data, bounds.ctypes.data, len(sizes), n, 1000, C.byref(total), C.byref(stored), out.ctypes.data, cap)
        dt = time.perf_counter() - t0
        assert rc == 0 and total.value == len(paths)
        results[label] = (dt, stored.value, out[:stored.value].copy())
        print("%-34s %6.2f s for %d M trees in %d drains (%.1f M trees/s), %d paths held at the end, %d threads" %
              (label, dt, millions, len(sizes), millions / dt, stored.value, os.cpu_count()), flush=True)
a, b = list(results.values())
from xmp_b200 import hostlib as xh  # noqa: E402


def first1000(x):
    x = np.ascontiguousarray(x).copy()
    xh.lib().xmph_sort_paths_dfs(x.ctypes.data, len(x), n)
    return x[:1000]


assert (first1000(a[2]) == first1000(b[2])).all(), "the two builds keep different trees"
print("both builds keep the same first 1000 trees of the depth-first order")
