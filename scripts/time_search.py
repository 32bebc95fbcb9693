"""Wall-clock breakdown of one exact search: context, upload, search_begin (allocation + dive), run."""
import gzip, json, os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from xmp_b200 import cuda as xc, hostlib as xh
bundle = json.loads(gzip.open(os.path.join(ROOT, "tests/golden/alignments.json.gz")).read())
gold = json.load(open(os.path.join(ROOT, "tests/golden/datasets.json")))
names = sys.argv[1:] or ["mt-10", "h1", "h4", "mh1", "Metazoan"]
for rep in range(2):
    for name in names:
        g = gold[name]
        with tempfile.NamedTemporaryFile("wb", suffix=".phy", delete=False) as f:
            f.write(bundle[g["alignment"]].encode("latin-1"))
        a = xh.Alignment(f.name); os.unlink(f.name)
        opts = None
        if "-Bp" in g["flags"]:
            opts = xh.OPT_USEPARTBOUND | xh.OPT_IMPROVEINITUPPERBOUND
        bound = int(g["flags"][g["flags"].index("-b") + 1]) if "-b" in g["flags"] else xh.NO_LIMIT
        t0 = time.perf_counter(); p = xh.Problem(a, options=opts, bound=bound); t1 = time.perf_counter()
        ctx = xc.Context(0); t2 = time.perf_counter()
        budget = int(float(os.environ.get('XMP_BUDGET_GB', '0')) * 2**30)
        tr = time.perf_counter()
        if not os.environ.get("XMP_NO_RESERVE"):
            ctx.search_reserve(p.num_taxa, p.n_blocks, budget)
        t2r = time.perf_counter(); reserve_ms = 1e3 * (t2r - tr); t2 = t2r
        ctx.load_problem(p); t3 = time.perf_counter()
        ctx.search_begin(bound=p.bound, mem_budget=int(float(os.environ.get('XMP_BUDGET_GB', '0')) * 2**30), flags=int(os.environ.get('XMP_FLAGS', '0'))); t4 = time.perf_counter()
        while not ctx.search_run(0): pass
        t5 = time.perf_counter()
        score, n, paths = ctx.search_result(); st = ctx.search_stats(); t6 = time.perf_counter()
        ctx.close(); t7 = time.perf_counter()
        print("%-10s rep%d reserve %.1f ms | prep %.1f ms | ctx %.1f | load %.1f | begin %.1f | run %.1f | result %.1f | close %.1f | device %.1f ms, %d batches, %d launches, nodes %d %s" % (
            name, rep, reserve_ms, 1e3*(t1-t0), 1e3*(tr-t1), 1e3*(t3-t2), 1e3*(t4-t3), 1e3*(t5-t4), 1e3*(t6-t5), 1e3*(t7-t6), 1e3*st["seconds"], st["iterations"], st["launches"], st["nodes"],
            "OK" if score == g["score"] and (n == g["num_trees"] or "-m" in g["flags"]) else "MISMATCH"), flush=True)
