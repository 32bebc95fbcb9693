"""How much could a tighter initial upper bound (the reference's TBR) save?  For every dataset:
the bound the greedy beam dive finds, the optimum, and nodes / device time with the dive's bound versus
with the optimum given up front (-b <optimum>)."""
import gzip, json, os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from xmp_b200 import cuda as xc, hostlib as xh
bundle = json.loads(gzip.open(os.path.join(ROOT, "tests/golden/alignments.json.gz")).read())
gold = json.load(open(os.path.join(ROOT, "tests/golden/datasets.json")))
names = sys.argv[1:] or ["mt-10", "h1", "h2", "h3", "h4", "h5", "mh1", "mh2", "mh3", "mh4", "mh6", "rbc14x759", "Metazoan", "EukaryotesrDNA"]
print("| dataset | optimum | dive bound | nodes (dive) | device s (dive) | nodes (-b optimum) | device s (-b optimum) |\n|---|---:|---:|---:|---:|---:|---:|")
for name in names:
    g = gold[name]
    with tempfile.NamedTemporaryFile("wb", suffix=".phy", delete=False) as f:
        f.write(bundle[g["alignment"]].encode("latin-1"))
    a = xh.Alignment(f.name); os.unlink(f.name)
    p = xh.Problem(a)
    row = []
    for given in (None, g["score"]):
        for rep in range(2):      # second repetition is the warm one
            ctx = xc.Context(0)
            ctx.load_problem(p)
            ctx.search_begin(bound=p.bound if given is None else given)
            dive = ctx.search_get_bound()
            while not ctx.search_run(0): pass
            score, n, paths = ctx.search_result(); st = ctx.search_stats()
            ctx.close()
            assert (score, n) == (g["score"], g["num_trees"]), (name, score, n)
        row.append((dive, st["nodes"], st["seconds"]))
    print("| %s | %d | %d | %d | %.3f | %d | %.3f |" % (name, g["score"], row[0][0], row[0][1], row[0][2], row[1][1], row[1][2]), flush=True)
