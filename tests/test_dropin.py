"""The drop-in proof: the REFERENCE's own program, its sources unmodified, built with the B200 Fitch family selected
through the reference's family seam (include/seq_b2_w16_cuda.h; src/seq.h:5-7,17-20, src/seq_b2_sse2asm.h:23-32,
src/CMakeLists.txt:264-283) and linked against libxmp_b200.so.  oracle/Makefile's `ref_cuda` target generates the two
lines a maintainer would add to seq.h / switches.h at build time; nothing of the reference is copied.

not gpu: the build succeeds, the binary imports the family from the library, and without a device it fails loudly
(no CPU fallback).  gpu: its result files equal those of the reference's stock fastc64 build, byte for byte."""
import os
import subprocess

import pytest

from tests.conftest import ROOT, golden_output

REF = "/root/reference"
EXE = os.path.join(ROOT, "oracle", "_ref", "fastdnamp_ref_cuda")
FLAGS = ["--tbr=N", "--seed=1", "--displaynewbounds=n", "--displaynewtrees=n", "--updatesecs=0"]


def test_reference_builds_and_links_against_the_library():
    if not os.path.exists(os.path.join(REF, "src", "seq.h")):
        pytest.skip("the reference sources are not on this machine (oracle/_ref is used as prebuilt)")
    r = subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "ref_cuda"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    gen = os.path.join(ROOT, "oracle", "_ref", "src_cuda")
    sw = open(os.path.join(gen, "switches.h")).read()
    assert "#define CUDAFITCH" in sw and "#define BYTESPERBLOCK 16" in sw and "#define FASTC64FITCH" not in sw
    # every reference source is a link to the file where it lies; only the two generated headers are real files
    real = sorted(f for f in os.listdir(gen) if not os.path.islink(os.path.join(gen, f)))
    assert real == ["seq.h", "switches.h"], real
    ours = subprocess.run(["diff", os.path.join(REF, "src", "seq.h"), os.path.join(gen, "seq.h")], capture_output=True, text=True).stdout
    changed = [l for l in ours.splitlines() if l.startswith((">", "<"))]
    assert len(changed) == 5 and any("seq_b2_w16_cuda.h" in l for l in changed) and sum("defined(CUDAFITCH)" in l for l in changed) == 1, ours
    syms = subprocess.run(["nm", "-D", "--undefined-only", EXE], capture_output=True, text=True).stdout
    for name in ("FitchBases_b2_w16_cuda", "FitchBasesAndCompare_b2_w16_cuda", "FitchScore_fw1_stopearly_b2_w16_cuda"):
        assert name in syms, syms
    assert "fastc64" not in syms
    needed = subprocess.run(["readelf", "-d", EXE], capture_output=True, text=True).stdout
    assert "libxmp_b200.so" in needed


def test_reference_binary_fails_loudly_without_a_device(tmp_path, alignment_dir, goldens):
    if not os.path.exists(EXE):
        pytest.skip("oracle/_ref/fastdnamp_ref_cuda not built")
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present: covered by the gpu test")
    except ImportError:
        pass
    r = subprocess.run([EXE] + FLAGS + ["-o", "out.nex", str(alignment_dir / goldens["quick2"]["alignment"])], cwd=tmp_path, capture_output=True, text=True, timeout=120)
    assert r.returncode != 0 and "CUDA" in r.stderr, r.stderr[-500:]


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["quick2", "its10", "its12"])
def test_reference_over_the_library_writes_the_reference_result_file(tmp_path, alignment_dir, goldens, name):
    """One kernel launch per Fitch call -- never the fast path, but the reference's own search loop driving the library
    through its own call surface: the same result file as its stock build (the golden output)."""
    if not os.path.exists(EXE):
        pytest.skip("oracle/_ref/fastdnamp_ref_cuda not built (make -C oracle ref_cuda, where the reference sources are)")
    g = goldens[name]
    r = subprocess.run([EXE] + FLAGS + ["-o", "out.nex", str(alignment_dir / g["alignment"])], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert (tmp_path / "out.nex").read_bytes() == golden_output(name)
    import re
    # the reference counts FitchBases / ScoreFitch calls INSIDE its own implementations (INCMEASURE in
    # src/seq_b2_w8_fastc64.c, src/fitchtemplate_b2_w8_fastc64.inc): with the B200 family selected none of those bodies
    # ran, although the search visited its nodes
    assert int(re.search(r"FitchBases\(\) executed\s+(\d+) times", r.stderr).group(1)) == 0
    assert int(re.search(r"ScoreFitch\(\) executed\s+(\d+) times", r.stderr).group(1)) == 0
    assert int(re.search(r"BranchAndBound\(\) executed\s+(\d+) times", r.stderr).group(1)) == g["ref_bandb_nodes"]
