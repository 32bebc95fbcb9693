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
