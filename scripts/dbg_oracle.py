import faulthandler, sys, os
faulthandler.enable()
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests import oracle_lib as ol
import bench, torch
paths = bench.make_paths(64, 8)
rows = bench.make_rows(torch, "cpu", 1).numpy()
print("A before cuda", ol.tree_sets(rows[:9], 8, paths[0])[3], flush=True)
d = bench.make_rows(torch, "cuda", 1)
rows2 = d.cpu().numpy()
print("B rows from cuda", rows2.shape, rows2.ctypes.data % 64, flush=True)
for i in (0, 31, 63):
    print("C", i, ol.tree_sets(rows2[:9], 8, paths[i])[3], flush=True)
import xmp_b200.cuda as xc
ctx = xc.Context(0)
ctx.load(d)
wl = bench.make_workload(torch, n_trees=64, m=8)
ctx.build_midpoints(wl["paths"], 8, wl["d_mid"])
ctx.synchronize()
print("D after ctx", flush=True)
rows3 = wl["d_rows"].cpu().numpy()
for i in (0, 31, 63):
    print("E", i, ol.tree_sets(rows3[:9], 8, wl["paths"][i])[3], flush=True)
