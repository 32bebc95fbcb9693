#!/usr/bin/env python3
"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference here.

TEST INFRASTRUCTURE ONLY.  Run in the build container, where /root/reference exists:

    make -C oracle ref && python tests/golden/make_goldens.py [--only NAME ...] [--jobs N]

What it does
  * runs oracle/_ref/fastdnamp_ref (the reference's own serial fastc64 program, compiled in place
    by oracle/Makefile) on every dataset BASELINE.json names, with the author's measurement flags
    (reference tools/orac_perf_test.pl:40): --seed=1 --displaynewtrees=n --updatesecs=0, plus
    --tbr=N (BASELINE config 1) and --savereordereddataset=Y so the taxon order is observable;
  * records, per dataset: optimal score, number of MP trees, sha256 of the output file, sha256 of the
    sorted canonical Newick set (and the set itself when small), restBound[], the preprocessing
    report (constant / non-PI / PI counts), the MST initial upper bound, the reordered taxon order
    and column weights, the reference's node counters and its own B&B seconds;
  * bundles the input alignments (measure/*.phy, data/phylip_format_testing/*.phy) into one
    compressed fixture, because /root/reference does not exist on the GPU box.

Outputs (committed): datasets.json, alignments.json.gz, trees/<name>.txt.gz (small sets only).
"""
import argparse
import concurrent.futures as cf
import gzip
import hashlib
import json
import os
import re
import subprocess
import sys
import tempfile
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from tests.newick import canonical_newick  # noqa: E402

REF = os.environ.get("XMP_REFERENCE", "/root/reference")
BIN = os.path.join(ROOT, "oracle", "_ref", "fastdnamp_ref")
BIN_ZINIT = os.path.join(ROOT, "oracle", "_ref", "fastdnamp_ref_zinit")
COMMON = ["--tbr=N", "--seed=1", "--displaynewbounds=Y", "--displaynewtrees=n", "--updatesecs=0",
          "--savereordereddataset=Y"]

# name -> (alignment path relative to REF, extra flags)
DATASETS = {
    "mt-10": ("measure/mt-10.phy", []),
    "rbc14x759": ("measure/rbc14x759.phy", []),
    "h1": ("measure/h1.phy", []), "h2": ("measure/h2.phy", []), "h3": ("measure/h3.phy", []),
    "h4": ("measure/h4.phy", []), "h5": ("measure/h5.phy", []),
    "mh1": ("measure/mh1.phy", []), "mh2": ("measure/mh2.phy", []), "mh3": ("measure/mh3.phy", []),
    "mh4": ("measure/mh4.phy", []), "mh6": ("measure/mh6.phy", []),
    "EukaryotesrDNA": ("measure/EukaryotesrDNA.phy", []),
    "Metazoan": ("measure/Metazoan.phy", []),
    # the author's own setting for its36 (tools/FastdnampPerfTesting.pm:31-34)
    "its36": ("measure/its36.phy", ["-Bp", "-b", "233", "-m", "100000"]),
    # the five simulated 24-taxon alignments of measure/ (0.1-4 M nodes each): CPU-side parity cases
    "e1": ("measure/e1.phy", []), "e3": ("measure/e3.phy", []), "e4": ("measure/e4.phy", []),
    "e5": ("measure/e5.phy", []), "e6": ("measure/e6.phy", []),
    # small extra cases: quick parity runs and ambiguity codes
    "mt-10.Bm": ("measure/mt-10.phy", ["-Bm"]),
    "h3.Bm": ("measure/h3.phy", ["-Bm"]),
    "h3.Bdm": ("measure/h3.phy", ["-Bdm"]),
    "h3.b359": ("measure/h3.phy", ["-b", "359"]),
    "h3.b358": ("measure/h3.phy", ["-b", "358"]),
    "its10": ("data/its/its10.phy", []),
    "its12": ("data/its/its12.phy", []),
    "its14": ("data/its/its14.phy", []),
    "primate": ("data/from_paup/primate-mtDNA.phy", []),
    "quick2": ("data/quick_test/two_incompat_sites_of_weight_2.phy", []),
    "53humans.12": ("data/from_minmaxsqueeze/53humans.12.phy", []),
    "53humans.16": ("data/from_minmaxsqueeze/53humans.16.phy", []),
    # deep searches (24-28 tree sizes in flight) under the partition bound; 1.4 M / 9.8 M / 9.8 M optimal trees,
    # of which -m keeps the first 1000 the depth-first search finds (src/yanbader.c:272-293)
    "53humans.24": ("data/from_minmaxsqueeze/53humans.24.phy", ["-Bp", "-m", "1000"]),
    "53humans.26": ("data/from_minmaxsqueeze/53humans.26.phy", ["-Bp", "-m", "1000"]),
    "53humans.28": ("data/from_minmaxsqueeze/53humans.28.phy", ["-Bp", "-m", "1000"]),
    # 88 M optimal trees: the search the bounded-memory -m cut exists for (10 minutes in the reference; the second,
    # uncapped run writes all of them -- ~13 GB -- just to count the lines)
    "53humans.29": ("data/from_minmaxsqueeze/53humans.29.phy", ["-Bp", "-m", "1000"]),
}
PARSER_FIXTURES = "data/phylip_format_testing"
# inputs of other golden files (partbound.json): bundled here, not run by this script
EXTRA_ALIGNMENTS = ["measure/53humans.32.phy"] + ["measure/e%d.phy" % i for i in (1, 3, 4, 5, 6)]
SMALL_SET = 2000  # full canonical tree list is committed when the set has at most this many trees


def parse_counter(err, label):
    m = re.search(re.escape(label) + r"\s+(-?\d+)", err)
    return int(m.group(1)) if m else None


def run_one(name):
    rel, extra = DATASETS[name]
    src = os.path.join(REF, rel)
    if not os.path.exists(src):
        return name, None
    needs_zinit = any(f.startswith("-B") and "m" in f[2:] for f in extra)
    exe = BIN_ZINIT if needs_zinit else BIN
    with tempfile.TemporaryDirectory(prefix="xmpgold_") as cwd:
        out = os.path.join(cwd, "out.nex")
        t0 = time.time()
        p = subprocess.run([exe] + COMMON + extra + ["-o", out, src], cwd=cwd, capture_output=True, text=True)
        wall = time.time() - t0
        err = p.stderr
        if p.returncode != 0:
            return name, {"error": "exit %d" % p.returncode, "stderr_tail": err[-2000:]}
        raw = open(out, "rb").read()
        text = raw.decode()
        trees = re.findall(r"= \[&U\] (\S+;)", text)
        m = re.search(r"Each tree has score (\d+)", text)
        score = int(m.group(1))
        canon = sorted(canonical_newick(t) for t in trees)
        canon_blob = ("\n".join(canon) + "\n").encode() if canon else b""
        rb = []
        for line in open(os.path.join(cwd, "restBound.txt")).read().splitlines()[1:]:
            i, b = line.split()
            rb.append([int(i), int(b)])
        order, weights = [], None
        rp = os.path.join(cwd, "reordered.phy")
        if os.path.exists(rp):
            lines = open(rp).read().splitlines()
            order = [int(l[1:10]) for l in lines[1:] if l.startswith("T")]
        mm = re.search(r"Improving initial upper bound to (\d+)", err)
        rec = {
            "alignment": os.path.basename(rel),
            "flags": extra,
            "score": score,
            "num_trees": len(trees),
            "output_sha256": hashlib.sha256(raw).hexdigest(),
            "canonical_sha256": hashlib.sha256(canon_blob).hexdigest(),
            "rest_bound": rb,
            "taxon_order": order,   # original 1-based taxon numbers in search order (taxonMap[i] + 1)
            "constant_sites": int(re.search(r"(\d+) constant sites", err).group(1)),
            "nonpi_sites": int(re.search(r"(\d+) non-parsimoniously-informative sites", err).group(1)),
            "constant_weight": int(re.search(r"having a total weight of (\d+)", err).group(1)),
            "pi_sites": int(re.search(r"(\d+) parsimoniously-informative sites", err).group(1)),
            "initial_upper_bound": int(mm.group(1)) if mm else None,
            "new_bounds": [int(x) for x in re.findall(r"New bound = (\d+)", err)],
            "ref_bandb_nodes": parse_counter(err, "BranchAndBound() executed"),
            "ref_score_calls": parse_counter(err, "ScoreFitch() executed"),
            "ref_bases_calls": parse_counter(err, "FitchBases() executed"),
            "ref_full_trees": parse_counter(err, "Number of full trees considered:"),
            "ref_bandb_seconds": float(re.search(r"Time to perform branch and bound search:\s+([\d.]+)s", err).group(1)),
            "ref_wall_seconds": round(wall, 2),
        }
        if "-m" in extra:
            # the cap hides how many optimal trees there are: count them in a second run without it
            uncapped = [f for i, f in enumerate(extra) if f != "-m" and (i == 0 or extra[i - 1] != "-m")]
            out2 = os.path.join(cwd, "all.tre")
            t0 = time.time()
            p2 = subprocess.run([exe] + COMMON + uncapped + ["-r", "-o", out2, src], cwd=cwd, capture_output=True, text=True)
            if p2.returncode == 0:
                with open(out2, "rb") as fh:
                    rec["num_trees_total"] = sum(1 for _ in fh)
                rec["ref_bandb_seconds_uncapped"] = float(re.search(r"Time to perform branch and bound search:\s+([\d.]+)s", p2.stderr).group(1))
        if len(trees) <= SMALL_SET:
            os.makedirs(os.path.join(HERE, "trees"), exist_ok=True)
            with gzip.GzipFile(os.path.join(HERE, "trees", name + ".txt.gz"), "wb", mtime=0) as f:
                f.write(canon_blob)
            with gzip.GzipFile(os.path.join(HERE, "trees", name + ".out.gz"), "wb", mtime=0) as f:
                f.write(raw)
        return name, rec


def bundle_alignments():
    bundle = {}
    for rel in [rel for rel, _ in DATASETS.values()] + EXTRA_ALIGNMENTS:
        p = os.path.join(REF, rel)
        if os.path.exists(p):
            bundle[os.path.basename(rel)] = open(p, "rb").read().decode("latin-1")
    d = os.path.join(REF, PARSER_FIXTURES)
    for fn in sorted(os.listdir(d)):
        if fn.endswith(".phy"):
            bundle["phylip_format_testing/" + fn] = open(os.path.join(d, fn), "rb").read().decode("latin-1")
    with gzip.GzipFile(os.path.join(HERE, "alignments.json.gz"), "wb", mtime=0) as f:
        f.write(json.dumps(bundle, sort_keys=True).encode())
    return len(bundle)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", nargs="*")
    ap.add_argument("--jobs", type=int, default=6)
    a = ap.parse_args()
    names = a.only or list(DATASETS)
    path = os.path.join(HERE, "datasets.json")
    db = json.load(open(path)) if os.path.exists(path) else {}
    with cf.ThreadPoolExecutor(a.jobs) as ex:
        for name, rec in ex.map(run_one, names):
            if rec is None:
                print("skip", name, "(alignment missing)")
                continue
            db[name] = rec
            print(name, {k: rec.get(k) for k in ("score", "num_trees", "ref_bandb_seconds", "error")}, flush=True)
            json.dump(db, open(path, "w"), indent=1, sort_keys=True)
    print("bundled", bundle_alignments(), "alignments")


if __name__ == "__main__":
    main()
