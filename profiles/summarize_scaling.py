#!/usr/bin/env python3
"""Tabulates the exact-search legs of bench.py runs kept under profiles/ (r02_bandb_*.json, r02_bench_n8_first.json):
time to optimum, efficiency against the one-GPU pass of the same run, node inflation, busy fraction, steals."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def load(name):
    p = os.path.join(HERE, name)
    if not os.path.exists(p):
        return None
    d = json.loads([l for l in open(p) if l.startswith("{")][-1])
    recs = d["datasets"] if "datasets" in d else d["bandb"]["datasets"]
    return d.get("n_gpus", d.get("bandb", {}).get("n_gpus")), {r["dataset"]: r for r in recs}


def main():
    runs = [("first peer exchange (one request word, half of the shallowest level, deal at 64 subtrees per shard)", "r02_bench_n8_first.json"),
            ("+ one slot per thief, share of every level (64 per thief), beam 4096", "r02_bandb_n8_v2.json"),
            ("+ deal at 4096 subtrees per shard, steals from the shallowest levels only", "r02_bandb_n8_v3.json"),
            ("+ work-weighted hint", "r02_bandb_n8_v4.json"),
            ("+ work-proportional steals, warm-up outside the clock", "r02_bandb_n8_v5.json"),
            ("+ flatter level weights, 16 K-job mailbox, sparse frontiers sweep every level", "r02_bandb_n8_v6.json"),
            ("+ deal at 65536 subtrees per shard (shipped)", "r02_bandb_n8_split64k.json"),
            ("deal at 2^20 subtrees per shard (for comparison)", "r02_bandb_n8_split1m.json")]
    names = ["h4", "mh3", "Metazoan", "its36", "its36.Bd"]
    print("| 8 x B200, seconds to the proven optimum (efficiency vs one GPU in the same run) | " + " | ".join(names) + " |")
    print("|---|" + "---:|" * len(names))
    for label, f in runs:
        r = load(f)
        if not r:
            continue
        cells = []
        for n in names:
            x = r[1].get(n)
            if not x:
                cells.append("")
                continue
            eff = x.get("efficiency")
            cells.append("%.4f%s" % (x["time_to_optimum_s"], " (%.2f)" % eff if eff else ""))
        print("| %s | %s |" % (label, " | ".join(cells)))
    for f in sys.argv[1:]:
        r = load(f)
        if not r:
            continue
        print("\n%s (N = %s)\n" % (f, r[0]))
        print("| dataset | s | one GPU, same run | speed-up | efficiency | node inflation | busy fraction | steals | slowest / fastest rank run time |")
        print("|---|---:|---:|---:|---:|---:|---:|---:|---:|")
        for n, x in r[1].items():
            pr = x.get("per_rank") or {}
            run = pr.get("run_s") or []
            print("| %s | %.4f | %s | %s | %s | %s | %.2f | %d | %s |" % (
                n, x["time_to_optimum_s"], "%.4f" % x["one_gpu_same_run_s"] if "one_gpu_same_run_s" in x else "",
                "%.2f" % x["speedup_vs_one_gpu"] if "speedup_vs_one_gpu" in x else "", "%.2f" % x["efficiency"] if "efficiency" in x else "",
                "%.3f" % x["node_inflation"] if "node_inflation" in x else "", x["busy_fraction"], x.get("steals", 0),
                "%.4f / %.4f" % (max(run), min(run)) if run else ""))


if __name__ == "__main__":
    main()
