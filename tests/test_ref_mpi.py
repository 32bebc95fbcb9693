"""The reference's PARALLEL program as a CPU baseline (test infrastructure, not the product).

There is no MPI in this image, so the reference's TARGETMULTI build (src/mpiboss.c, src/mpiworker.c, the polling hook in
src/yanbader.c:78-94) "does not compile here" as it stands.  oracle/shim/mpi provides the seventeen MPI-1 calls it makes
 for the ranks of one node; `make -C oracle ref_mpi` compiles the reference's unmodified sources against it.
These tests pin the shim (tests/xmpi_check.c) and then the program: the same score, the same number of trees and the same
canonical tree set as the serial reference, whatever the number of workers.  bench.py times it beside the GPU searches."""
import hashlib
import os
import re
import subprocess

import pytest

from tests.conftest import ROOT
from tests.newick import canonical_newick

REF = "/root/reference"
SHIM = os.path.join(ROOT, "oracle", "shim", "mpi")
EXE = os.path.join(ROOT, "oracle", "_ref", "fastdnamp_ref_mpi")
RUN = os.path.join(ROOT, "oracle", "_ref", "xmpirun")
FLAGS = ["--tbr=N", "--seed=1", "--displaynewbounds=n", "--displaynewtrees=n", "--updatesecs=0"]


@pytest.mark.parametrize("ranks", [3, 6])
def test_shim_semantics(tmp_path, ranks):
    """Ordering per pair, matching by tag, wildcards, zero-length messages, 2 MiB messages through full rings, the
    Wait / Test family with null handles."""
    for out, srcs in (("xmpirun", ["xmpirun.c"]), ("xmpi_check", [os.path.join(ROOT, "tests", "xmpi_check.c"), "xmpi.c"])):
        r = subprocess.run(["gcc", "-O1", "-g", "-std=gnu11", "-Wall", "-Wextra", "-I" + SHIM, "-o", str(tmp_path / out)] +
                           [s if os.path.isabs(s) else os.path.join(SHIM, s) for s in srcs], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-3000:]
    r = subprocess.run([str(tmp_path / "xmpirun"), "-n", str(ranks), str(tmp_path / "xmpi_check")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "xmpi_check ok: %d ranks" % ranks in r.stdout, r.stdout[-1000:] + r.stderr[-3000:]


def test_launcher_reports_a_failing_rank(tmp_path):
    r = subprocess.run(["gcc", "-O1", "-std=gnu11", "-I" + SHIM, "-o", str(tmp_path / "xmpirun"), os.path.join(SHIM, "xmpirun.c")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    # rank 1 exits with an error while the others would wait for ever: the launcher kills them and returns the failure
    script = tmp_path / "rank.sh"
    script.write_text("#!/bin/sh\nif [ \"$XMPI_RANK\" = 1 ]; then exit 7; fi\nexec sleep 600\n")
    script.chmod(0o755)
    r = subprocess.run([str(tmp_path / "xmpirun"), "-n", "3", str(script)], capture_output=True, text=True, timeout=60)
    assert r.returncode == 7


def _build():
    if os.path.exists(os.path.join(REF, "src", "mpiboss.c")):
        r = subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "ref_mpi"], capture_output=True, text=True)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    if not (os.path.exists(EXE) and os.path.exists(RUN)):
        pytest.skip("oracle/_ref/fastdnamp_ref_mpi not built and the reference sources are not on this machine")


def test_program_needs_the_launcher(tmp_path):
    _build()
    r = subprocess.run([EXE, "-h"], cwd=tmp_path, capture_output=True, text=True, timeout=60)
    assert r.returncode != 0 and "xmpirun" in r.stderr


@pytest.mark.parametrize("name,ranks", [("quick2", 2), ("its12", 4), ("mt-10", 5), ("h3", 9), ("mh6", 7), ("h3.b358", 4), ("e3", 5), ("e5", 9)])
def test_parallel_reference_finds_the_serial_reference_set(tmp_path, alignment_dir, goldens, name, ranks):
    """One boss and ranks - 1 workers: score, number of trees and canonical tree set of the serial program's golden."""
    _build()
    g = goldens[name]
    r = subprocess.run([RUN, "-n", str(ranks), EXE] + FLAGS + list(g["flags"]) + ["-r", "-o", "out.tre", str(alignment_dir / g["alignment"])],
                       cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert re.search(r"Time to perform branch and bound search:\s+[\d.]+s", r.stderr)
    trees = [t for t in (tmp_path / "out.tre").read_text().split() if t.startswith("(")]
    assert len(trees) == g["num_trees"]
    if trees:
        canon = sorted(canonical_newick(t) for t in trees)
        assert hashlib.sha256(("\n".join(canon) + "\n").encode()).hexdigest() == g["canonical_sha256"]
