"""The C-ABI library loads on a machine without a GPU and exports every symbol include/xmp_cuda.h
declares; entry points that need a device fail loudly instead of falling back.  No GPU needed."""
import ctypes
import os
import re
import subprocess

import pytest

from tests.conftest import ROOT
import xmp_b200.cuda as xc


def declared_functions():
    text = open(os.path.join(ROOT, "include", "xmp_cuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return re.findall(r"^(?:const\s+)?[A-Za-z_][\w\s\*]*?\b(\w+)\s*\([^;{]*\)\s*;", text, flags=re.M)


def test_header_and_library_agree():
    names = declared_functions()
    assert len(names) >= 28 and "FitchScore_fw1_stopearly_b2_w16_cuda" in names and "xmp_cuda_search_run" in names
    L = ctypes.CDLL(xc.LIB_PATH)
    for n in names:
        assert hasattr(L, n), "libxmp_b200.so does not export %s" % n
    assert sorted(names) == sorted(xc.EXPORTS)


def test_library_is_sm100a_only_and_self_contained():
    out = subprocess.run(["cuobjdump", "-lelf", xc.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs
    needed = subprocess.run(["readelf", "-d", xc.LIB_PATH], capture_output=True, text=True).stdout
    assert "libcudart" not in needed and "oracle" not in needed and "torch" not in needed


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("this check is for the GPU-less build container")
    assert xc.lib().xmp_cuda_device_count() == 0
    with pytest.raises(xc.XmpCudaError, match="no CPU fallback"):
        xc.Context(0)


def test_product_never_touches_the_oracle():
    """Nothing under xmp_b200/ or host/ may import, link or execute oracle/."""
    for base in ("xmp_b200", "host", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".c", ".h", ".cu", ".cuh")):
                    text = open(os.path.join(dirpath, f), errors="replace").read()
                    assert "oracle" not in text.lower(), os.path.join(dirpath, f)
    mk = open(os.path.join(ROOT, "Makefile")).read()
    cuda_rule = mk[mk.index("xmp_b200/libxmp_b200.so:"):mk.index("cli:")]
    assert "oracle" not in cuda_rule


def test_keep_first_dfs_is_the_head_of_the_depth_first_order():
    """xmp_cuda_paths_keep_first_dfs (the host-side half of the -m cap: the reference keeps the FIRST maxTrees trees
    its depth-first search meets, src/yanbader.c:272-293) against the host library's full depth-first sort
    (host/trees.c), on random path sets with duplicates-free keys, small and large enough for the threaded path."""
    import numpy as np
    from tests import oracle_lib as ol
    from xmp_b200 import hostlib as xh
    rng = np.random.default_rng(99)
    for n, count, keep in ((5, 40, 7), (12, 3000, 100), (28, 200000, 1000), (36, 70000, 70000), (100, 500, 1)):
        paths = np.unique(ol.random_paths(rng, count, n, n), axis=0)
        rng.shuffle(paths)
        full = paths.copy()
        xh.lib().xmph_sort_paths_dfs(full.ctypes.data, len(full), n)
        got = xc.paths_keep_first_dfs(paths, keep)
        assert got.shape == (min(keep, len(paths)), n)
        assert (got == full[:keep]).all(), (n, count, keep)
        # the order itself: keys of consecutive trees are non-decreasing in the reference's job coordinates
        keys = [tuple(xh.path_to_dfs(p, n)[3:]) for p in got[:50]]
        assert keys == sorted(keys)


def test_tree_bookkeeping_with_and_without_the_cap(tmp_path):
    """tests/stub/selftest_tree_cap.cpp replays drains through the very code the search uses (xmp_b200/csrc/tree_book.h,
    compiled here into a test-only library: the product library has no such entry point): records above the bound are dropped, a falling bound discards what was held (and what
    had been cut and counted), and with a finite max_trees the first max_trees in depth-first order survive while the
    count stays complete -- the reference's -m semantics (src/yanbader.c:248-294)."""
    import ctypes as C
    import numpy as np
    from tests import oracle_lib as ol
    from xmp_b200 import hostlib as xh
    so = str(tmp_path / "libxmp_selftest.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-pthread", "-o", so,
                    os.path.join(ROOT, "tests", "stub", "selftest_tree_cap.cpp")], check=True)
    L = C.CDLL(so)
    vp = C.c_void_p
    L.xmp_selftest_tree_cap.argtypes = [vp, vp, vp, C.c_uint, C.c_uint, C.c_uint, C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong), vp, C.c_size_t]
    rng = np.random.default_rng(4242)
    n = 26
    trec = 4 + ((n + 3) & ~3)

    def records(paths, scores):
        rec = np.zeros((len(paths), trec), np.uint8)
        rec[:, :4] = np.asarray(scores, dtype="<u4").view(np.uint8).reshape(-1, 4)
        rec[:, 4:4 + n] = paths
        return rec

    def replay(batches, max_trees, cap):
        recs = np.concatenate([records(p, s) for p, s, _ in batches])
        sizes = np.array([len(p) for p, _, _ in batches], np.uint32)
        bounds = np.array([b for _, _, b in batches], np.uint32)
        total, stored = C.c_ulonglong(0), C.c_ulonglong(0)
        out = np.zeros((cap, n), np.uint8)
        rc = L.xmp_selftest_tree_cap(recs.ctypes.data, sizes.ctypes.data, bounds.ctypes.data, len(batches), n, max_trees,
                                          C.byref(total), C.byref(stored), out.ctypes.data, cap)
        assert rc == 0
        return total.value, stored.value, out[:min(cap, stored.value)]

    def dfs_sorted(paths):
        full = np.ascontiguousarray(paths).copy()
        xh.lib().xmph_sort_paths_dfs(full.ctypes.data, len(full), n)
        return full

    all_paths = np.unique(ol.random_paths(rng, 3_400_000, n, n), axis=0)
    rng.shuffle(all_paths)
    a, b, c, d = np.split(all_paths[:3_200_000], [900_000, 1_500_000, 2_300_000])
    mixed = np.where(rng.random(len(b)) < 0.5, 101, 100)                # batch 2: the bound falls to 100 inside it
    batches = [(a, np.full(len(a), 101), 101),                           # 900 K trees of length 101
               (b, mixed, 100),                                          # ... superseded: half of these have length 100
               (c, np.full(len(c), 100), 100), (d, np.where(rng.random(len(d)) < 0.1, 103, 100), 100)]
    at100 = np.concatenate([b[mixed == 100], c, d[batches[3][1] == 100]])
    # no cap: everything at the final bound is held
    total, stored, got = replay(batches, 0x7FFFFFFF, len(at100))
    assert total == stored == len(at100)
    assert (dfs_sorted(got) == dfs_sorted(at100)).all()
    # -m 1000: the count is complete, the held paths contain exactly the first 1000 of the depth-first order
    total, stored, got = replay(batches, 1000, 1_000_000)
    assert total == len(at100) and 1000 <= stored < total and len(got) == stored
    assert (dfs_sorted(got)[:1000] == dfs_sorted(at100)[:1000]).all()
    # the cut had already happened at length 101 (900 K < 2^20 held, then 1.2 M): a later better bound forgets all of it
    total, stored, got = replay(batches[:1] + [(b, np.full(len(b), 101), 101)] + [(c[:10], np.full(10, 99), 99)], 1000, 100)
    assert (total, stored) == (10, 10) and (dfs_sorted(got) == dfs_sorted(c[:10])).all()
    total, stored, _ = replay(batches[:1] + [(b, np.full(len(b), 101), 101)], 1000, 10)
    assert total == len(a) + len(b) and stored == 1000
    # after a cut, trees that do not come before the last kept one are counted without being stored (the threshold
    # pre-filter): the same 3.2 M trees, all of one length, arriving in 80 drains -- several cuts, each with a lower
    # threshold -- must still leave exactly the first 500 of the depth-first order and the complete count
    everything = all_paths[:3_200_000]
    pieces = np.array_split(everything, 80)
    total, stored, got = replay([(p, np.full(len(p), 77), 77) for p in pieces], 500, 3_000_000)
    assert total == len(everything) and 500 <= stored < 1_100_000
    assert (dfs_sorted(got)[:500] == dfs_sorted(everything)[:500]).all()


def test_bench_reference_arm_contract_under_torchrun(tmp_path):
    """`bench.py --impl reference` is the driver's CPU arm: the reference's own Fitch code (oracle/_ref) on the host cores, no
    GPU involved, so it runs here.  Launched the way the driver launches N > 1 (torchrun, one process per rank): rank 0
    alone works and prints ONE JSON line with the contract's keys, the other rank exits 0 without output."""
    import json
    import sys
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libxmpref_fitch.so")):
        pytest.skip("oracle/_ref not built")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29641", os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                       cwd=tmp_path, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout
    out = json.loads(lines[0])
    assert out["impl"] == "reference" and out["n_gpus"] == 2 and out["steps"] == 1 and out["higher_is_better"] is True
    assert out["metric"] == "fitch_insertion_scoring_site_merges_per_s" and out["unit"] == "Gsite-merges/s" and out["value"] > 0
    assert out["cpu_baseline"]["kind"] == "reference" and out["cpu_baseline"]["cores"] >= 1 and out["cpu_baseline"]["value"] == out["value"]
    assert out["e2e"] == {"value": out["value"], "unit": out["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert out["config"]["taxa"] == 64 and out["config"]["sites"] == 1 << 20 and out["config"]["m"] == 8
