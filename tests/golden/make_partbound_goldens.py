#!/usr/bin/env python3
"""Golden restBound[] arrays for the partition lower bound (-Bp), from the UNMODIFIED reference.

TEST INFRASTRUCTURE ONLY.  Run in the build container, where /root/reference exists:

    make -C oracle ref && python tests/golden/make_partbound_goldens.py

Runs oracle/_ref/fastdnamp_ref (the reference's own program, compiled in place by oracle/Makefile)
with --onlycomputelowerbound=Y and records the restBound.txt it writes (reference src/bound.c:47-56)
for -Bp, its -BpC / -BpL strategies, --pbmaxiters and combinations with -Bd / -Bm.  Also records the
sha256 of the reference's generated two-site table (src/sitepaircosts.c), extracted from the text of
that file.  Output (committed): partbound.json; alignments.json.gz is re-bundled (make_goldens.EXTRA_ALIGNMENTS
lists the inputs only this file needs).  ref_wall_seconds is the whole reference process, timed while
the other cases run beside it: an upper bound.
"""
import concurrent.futures as cf
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
REF = os.environ.get("XMP_REFERENCE", "/root/reference")
COMMON = ["--tbr=N", "--seed=1", "--displaynewbounds=n", "--displaynewtrees=n", "--updatesecs=0",
          "--onlycomputelowerbound=Y"]

M = "measure/%s.phy"
CASES = {name: (M % name, ["-Bp"]) for name in
         ["mt-10", "rbc14x759", "h1", "h3", "h4", "mh1", "mh2", "mh3", "mh4", "mh6", "EukaryotesrDNA", "Metazoan",
          "its36", "53humans.32", "e1", "e3", "e4", "e5", "e6"]}
CASES.update({
    "its14": ("data/its/its14.phy", ["-Bp"]),
    "primate": ("data/from_paup/primate-mtDNA.phy", ["-Bp"]),
    "53humans.16": ("data/from_minmaxsqueeze/53humans.16.phy", ["-Bp"]),
    "h1.C": (M % "h1", ["-BpC"]), "h1.L": (M % "h1", ["-BpL"]),
    "Metazoan.C": (M % "Metazoan", ["-BpC"]), "Metazoan.L": (M % "Metazoan", ["-BpL"]),
    "its36.C": (M % "its36", ["-BpC"]), "its36.L": (M % "its36", ["-BpL"]),
    "primate.C": ("data/from_paup/primate-mtDNA.phy", ["-BpC"]),
    "primate.L": ("data/from_paup/primate-mtDNA.phy", ["-BpL"]),
    "Metazoan.it0": (M % "Metazoan", ["-Bp", "--pbmaxiters=0"]),
    "Metazoan.it1": (M % "Metazoan", ["-Bp", "--pbmaxiters=1"]),
    "Metazoan.dp": (M % "Metazoan", ["-Bdp"]),
    "its36.dp": (M % "its36", ["-Bdp"]),
    "EukaryotesrDNA.d": (M % "EukaryotesrDNA", ["-Bp", "-d"]),      # ambiguity codes dropped vs kept
})
# -Bm reads an uninitialised local in the reference: the zero-initialised build is the defined one
ZINIT = {"its36.dmp": (M % "its36", ["-Bdmp"])}


def run(item):
    name, (rel, flags), binary = item
    with tempfile.TemporaryDirectory() as d:
        t0 = time.time()
        r = subprocess.run([binary] + COMMON + flags + [os.path.join(REF, rel)], cwd=d, capture_output=True, text=True)
        assert r.returncode == 0, (name, r.stderr[-400:])
        rows = open(os.path.join(d, "restBound.txt")).read().splitlines()[1:]
        secs = time.time() - t0
    return name, {"alignment": os.path.basename(rel), "flags": flags,
                  "rest_bound": [[int(x) for x in row.split()] for row in rows],
                  "ref_wall_seconds": round(secs, 2)}


def main():
    items = [(k, v, os.path.join(ROOT, "oracle", "_ref", "fastdnamp_ref")) for k, v in CASES.items()]
    items += [(k, v, os.path.join(ROOT, "oracle", "_ref", "fastdnamp_ref_zinit")) for k, v in ZINIT.items()]
    with cf.ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        out = dict(ex.map(run, items))
    table = bytes(int(x) for x in re.findall(r"^\t(\d+),?\t", open(os.path.join(REF, "src", "sitepaircosts.c")).read(), re.M))
    assert len(table) == 65535
    doc = {"_generated_by": "tests/golden/make_partbound_goldens.py", "cases": dict(sorted(out.items())),
           "site_pair_table_sha256": hashlib.sha256(table).hexdigest(),
           "site_pair_table_samples": {str(i): table[i - 1] for i in (1, 3, 0x8421, 0x1248, 0xffff, 0x0f0f, 0x8001, 12345, 54321)}}
    with open(os.path.join(HERE, "partbound.json"), "w") as f:
        json.dump(doc, f, indent=1, sort_keys=True)
        f.write("\n")
    sys.path.insert(0, HERE)
    import make_goldens
    print("bundled", make_goldens.bundle_alignments(), "alignments")
    for k, v in sorted(out.items()):
        print("%-18s %-22s ref %.2fs  %s" % (k, " ".join(v["flags"]), v["ref_wall_seconds"], [r[1] for r in v["rest_bound"]][:12]))


if __name__ == "__main__":
    main()
