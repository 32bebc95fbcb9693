"""The fastdnamp_b200 command line where no GPU is needed: option handling, settings report, lower-bound
only runs (restBound.txt), reader errors, and the loud failure when a search is requested without a
device.  The search itself is covered by tests/test_gpu.py."""
import os
import subprocess

import pytest

from tests.conftest import ROOT

EXE = os.path.join(ROOT, "fastdnamp_b200")


def run(args, cwd):
    return subprocess.run([EXE] + args, cwd=cwd, capture_output=True, text=True)


def test_usage_and_bad_options(tmp_path):
    r = run([], tmp_path)
    assert r.returncode == 0 and "Syntax: fastDNAmp [<options>] <phylipfile>" in r.stderr and "--gpus=<nn>" in r.stderr
    r = run(["-h"], tmp_path)
    assert r.returncode == 0 and "-m <nn>      Maximum number of trees to save" in r.stderr
    r = run(["--nosuchoption=3", "x.phy"], tmp_path)
    assert r.returncode == 1 and "Unrecognised option '--nosuchoption=3', aborting." in r.stderr
    r = run(["-Z"], tmp_path)
    assert r.returncode == 1 and "Unrecognised option '-Z', aborting." in r.stderr
    r = run(["missing.phy"], tmp_path)
    assert r.returncode == 1 and "File 'missing.phy' could not be opened for input, aborting." in r.stderr
    r = run(["-Bx", "missing.phy"], tmp_path)
    assert r.returncode == 1 and "Unknown bound type 'x'." in r.stderr


@pytest.mark.parametrize("name", ["mt-10", "h3", "EukaryotesrDNA"])
def test_lower_bound_only_run_matches_reference(alignment_dir, goldens, tmp_path, name):
    """--onlycomputelowerbound=Y: everything up to the search, which is all host C: the settings report,
    the preprocessing report, the MST upper bound line and restBound.txt must be the reference's."""
    g = goldens[name]
    r = run(["--tbr=N", "--seed=1", "--displaynewtrees=n", "--updatesecs=0", "--onlycomputelowerbound=Y", "--savereordereddataset=Y",
             str(alignment_dir / g["alignment"])], tmp_path)
    assert r.returncode == 0, r.stderr
    e = r.stderr
    assert e.startswith("fastDNAmp: Exact Maximum Parsimony Phylogenetic Tree Search v1.1\n\n")
    for line in ("No initial bound specified.", "All optimal trees will be saved.", "All sites will be processed.",
                 "Sites containing gaps or ambiguous characters will be retained.", "Trees will be output in NEXUS format.",
                 "New trees will not be reported as they occur.", "The distinct-bases-per-site bound will be used.",
                 "The program will exit after computing lower bounds.",
                 "Using random number seed 1.  First 3 random numbers: 1804289383, 846930886, 1681692777.",
                 "%d constant sites." % g["constant_sites"],
                 "%d non-parsimoniously-informative sites, having a total weight of %d." % (g["nonpi_sites"], g["constant_weight"]),
                 "%d parsimoniously-informative sites." % g["pi_sites"],
                 "Improving initial upper bound to %d, based on minimum spanning tree weight." % g["initial_upper_bound"]):
        assert line in e, line
    rb = [[int(x) for x in l.split()] for l in (tmp_path / "restBound.txt").read_text().splitlines()[1:]]
    assert rb == g["rest_bound"]
    order = [int(l[1:10]) for l in (tmp_path / "reordered.phy").read_text().splitlines()[1:]]
    assert order == g["taxon_order"]


@pytest.mark.parametrize("name", ["its36", "h1.C", "Metazoan.it1", "its36.L"])
def test_partition_bound_run_matches_reference(alignment_dir, tmp_path, name):
    """-Bp and its variants from the command line: the settings lines and restBound.txt of the reference."""
    import json
    case = json.load(open(os.path.join(ROOT, "tests", "golden", "partbound.json")))["cases"][name]
    r = run(["--tbr=N", "--seed=1", "--updatesecs=0", "--onlycomputelowerbound=Y"] + case["flags"] + [str(alignment_dir / case["alignment"])], tmp_path)
    assert r.returncode == 0, r.stderr
    kind = "Clipped scores" if name.endswith(".C") else "Lower bounds" if name.endswith(".L") else "Final scores"
    assert "The PARTBOUND bound will be used.  %s will be maximised." % kind in r.stderr
    stale = 1 if name.endswith(".it1") else 2
    assert "PARTBOUND iterations will stop after %d iterations with no improvement." % stale in r.stderr
    assert "The distinct-bases-per-site bound will be used." not in r.stderr
    rb = [[int(x) for x in l.split()] for l in (tmp_path / "restBound.txt").read_text().splitlines()[1:]]
    assert rb == case["rest_bound"]


def test_reader_errors_reach_the_user(alignment_dir, tmp_path):
    r = run([str(alignment_dir / "phylip_format_testing" / "bad1_T3_missing_1_base.phy")], tmp_path)
    assert r.returncode == 1 and "contains 1 fewer sites than the first line in this group, aborting." in r.stderr
    f = tmp_path / "three.phy"
    f.write_text("3 4\nA         ACGT\nB         ACGT\nC         ACGT\n")
    r = run([str(f)], tmp_path)
    assert r.returncode == 1 and "You must have four or more taxa to perform this analysis." in r.stderr
    r = run(["-Bi", str(alignment_dir / "h3.phy")], tmp_path)
    assert r.returncode == 1 and "-Bf" in r.stderr
    r = run(["-BC", str(alignment_dir / "h3.phy")], tmp_path)
    assert r.returncode == 1 and "you must have previously specified -Bp" in r.stderr


def test_search_without_a_device_fails_loudly(alignment_dir, tmp_path):
    import xmp_b200.cuda as xc
    if xc.lib().xmp_cuda_device_count() > 0:
        pytest.skip("this check is for machines without a GPU")
    r = run(["-q", str(alignment_dir / "h3.phy")], tmp_path)
    assert r.returncode == 1 and "No CUDA device found.  This program has no CPU search path, aborting." in r.stderr
    assert r.stdout == ""


_REF_EXE = os.path.join(ROOT, "oracle", "_ref", "fastdnamp_ref")


@pytest.mark.skipif(not os.path.exists(_REF_EXE), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("flags", [["-d"], ["--startcol=3", "--endcol=40"], ["-d", "--startcol=10", "--endcol=300", "-b", "5000"],
                                   ["-m", "7", "-q"]])
def test_option_plumbing_matches_the_compiled_reference(alignment_dir, tmp_path, flags):
    """-d, --startcol / --endcol, -b, -m, -q and reading the alignment from standard input ('-'): the same command line
    through the reference's own program and through fastdnamp_b200, up to the search: identical restBound.txt,
    reordered.phy and site-count lines (with -q: identical, i.e. empty, stderr)."""
    src = (alignment_dir / "Metazoan.phy").read_bytes()
    common = ["--tbr=N", "--seed=1", "--updatesecs=0", "--displaynewtrees=n", "--onlycomputelowerbound=Y", "--savereordereddataset=Y"] + flags + ["-"]
    outs = []
    for exe, name in ((_REF_EXE, "ref"), (EXE, "ours")):
        d = tmp_path / name
        d.mkdir()
        r = subprocess.run([exe] + common, cwd=d, input=src, capture_output=True, timeout=120)
        assert r.returncode == 0, r.stderr
        err = r.stderr.decode()
        counts = [l for l in err.splitlines() if "sites" in l and ("constant" in l or "informative" in l)]
        outs.append(((d / "restBound.txt").read_bytes(), (d / "reordered.phy").read_bytes(), counts,
                     [l for l in err.splitlines() if l.startswith(("Initial bound", "Up to", "Sites ", "Sites containing", "Reading data"))]))
        if "-q" in flags:
            assert err == ""
    assert outs[0] == outs[1]


@pytest.mark.skipif(not os.path.exists(_REF_EXE), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("flags", [[], ["-Bdm", "-b", "400", "-m", "10", "-r", "-o", "o.tre", "--keeptempfiles=Y", "--displaynewbounds=n"],
                                   ["-Bp", "--pbmaxiters=1", "-d", "--startcol=2", "--endcol=60", "--updatesecs=7"],
                                   ["-BpC", "--displaynewtrees=n", "--savereordereddataset=Y"]])
def test_settings_and_preprocessing_report_is_the_references(alignment_dir, tmp_path, flags):
    """Everything the two programs print before the search (ReportSettings, src/common.c:252-369, and the preprocessing
    lines) is the same text, except the one line that names the build ("Single-processor version.")."""
    outs = []
    for exe, name in ((_REF_EXE + "_zinit" if any("m" in f[2:] for f in flags if f.startswith("-B")) else _REF_EXE, "ref"), (EXE, "ours")):
        d = tmp_path / name
        d.mkdir()
        r = subprocess.run([exe, "--tbr=N", "--seed=1", "--onlycomputelowerbound=Y"] + flags + [str(alignment_dir / "h3.phy")], cwd=d, capture_output=True,
                           text=True, timeout=300)
        assert r.returncode == 0, r.stderr
        lines = r.stderr.splitlines()
        build = [l for l in lines if l in ("Single-processor version.", "B200 (CUDA) version searching on 1 GPU.")]
        assert len(build) == 1
        outs.append([l for l in lines if l not in build])
    assert outs[0] == outs[1]


@pytest.mark.skipif(not os.path.exists(_REF_EXE), reason="oracle/_ref not built (needs /root/reference)")
def test_phylip_reader_fuzz_against_the_compiled_reference(tmp_path):
    """The same alignment laid out at random -- interleaved blocks of random width, blank lines, lower case, U for T,
    blanks / tabs / digits inside the sequences, CR LF line ends -- and sometimes damaged (bad character, a line too long
    or too short, truncated file): both programs must say the same thing up to the search (src/common.c:2850-2948)."""
    import numpy as np
    rng = np.random.default_rng(int(os.environ.get("XMP_FUZZ_SEED", "8128")))
    states = list("ACGTACGTACGTURYSWKMBDHV-N?")
    ok = bad = 0
    for case in range(int(os.environ.get("XMP_FUZZ_CASES", "60"))):
        n, sites = int(rng.integers(4, 9)), int(rng.integers(5, 60))
        seqs = [[str(rng.choice(states)) for _ in range(sites)] for _ in range(n)]
        for t in range(1, n):
            for s in range(sites):
                if rng.random() < 0.6:
                    seqs[t][s] = seqs[int(rng.integers(0, t))][s]
        eol = "\r\n" if case % 5 == 0 else "\n"
        lines, pos, first = ["%d %d" % (n, sites)], 0, True
        damage = ["", "badchar", "long", "short", "truncate"][case % 5 if case % 2 else 0]
        while pos < sites:
            width = int(rng.integers(1, sites - pos + 1))
            for t in range(n):
                chunk = seqs[t][pos:pos + width]
                if rng.random() < 0.3:
                    chunk = [c.lower() for c in chunk]
                text = ""
                for c in chunk:
                    text += c
                    r = rng.random()
                    if r < 0.08:
                        text += " "
                    elif r < 0.11:
                        text += "\t"
                    elif r < 0.14:
                        text += str(int(rng.integers(0, 10)))
                label = ("T%d x" % t).ljust(10) if first else ""
                lines.append(label + text)
            if rng.random() < 0.5:
                lines.append("" if rng.random() < 0.5 else "   ")
            pos += width
            first = False
        if damage == "badchar":
            k = int(rng.integers(1, len(lines)))
            lines[k] += "!"
        elif damage == "long":
            lines[-1 if lines[-1].strip() else -2] += "A"
        elif damage == "short":
            k = n if len(lines) > n else 1
            if len(lines[k].rstrip()) > 11:
                lines[k] = lines[k].rstrip()[:-1]
        elif damage == "truncate":
            lines = lines[:max(2, len(lines) - int(rng.integers(1, n + 1)))]
        f = tmp_path / ("rd%d.phy" % case)
        f.write_bytes((eol.join(lines) + (eol if rng.random() < 0.8 else "")).encode())
        res = []
        for exe, name in ((_REF_EXE, "ref"), (EXE, "ours")):
            d = tmp_path / ("%s%d" % (name, case))
            d.mkdir()
            r = subprocess.run([exe, "--tbr=N", "--seed=1", "--updatesecs=0", "--onlycomputelowerbound=Y", "--savereordereddataset=Y", str(f)],
                               cwd=d, capture_output=True, timeout=60)
            err = [l for l in r.stderr.decode(errors="replace").splitlines()
                   if l not in ("Single-processor version.", "B200 (CUDA) version searching on 1 GPU.")]
            files = tuple((d / x).read_bytes() if (d / x).exists() else None for x in ("restBound.txt", "reordered.phy"))
            res.append((r.returncode, err, files))
        if res[0][0] not in (0, 1):
            continue                                   # the reference itself crashed on this input (it reads past short files)
        assert res[0] == res[1], (case, damage, res[0][1][-2:], res[1][1][-2:])
        ok += res[0][0] == 0
        bad += res[0][0] == 1
    assert ok >= 20 and bad >= 5


@pytest.mark.skipif(not os.path.exists(_REF_EXE), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("body, quiet", [("2\t99\n3\t60\n5 7\n", False), ("2\t99\n3\t60\n", True), ("2\t9\nx y\n", False), ("", False),
                                         (None, False)])
def test_loading_rest_bound_files_like_the_reference(alignment_dir, tmp_path, body, quiet):
    """-Bf <file> (src/common.c:1530-1554): same restBound.txt afterwards, same 'Read N lines' / parse-error text on
    stderr (the reference prints the former even under -q), same exit code."""
    rb = tmp_path / "in.txt"
    if body is not None:                                       # None: the file does not exist
        rb.write_text("i\twhatever the header says\n" + body)
    res = []
    for exe, name in ((_REF_EXE, "ref"), (EXE, "ours")):
        d = tmp_path / name
        d.mkdir()
        r = subprocess.run([exe, "--tbr=N", "--seed=1", "--onlycomputelowerbound=Y", "-Bf", str(rb)] + (["-q"] if quiet else []) +
                           [str(alignment_dir / "h3.phy")], cwd=d, capture_output=True, text=True, timeout=60)
        err = [l for l in r.stderr.splitlines() if l not in ("Single-processor version.", "B200 (CUDA) version searching on 1 GPU.")]
        res.append((r.returncode, err, (d / "restBound.txt").read_text() if (d / "restBound.txt").exists() else None))
    assert res[0] == res[1]
    assert res[0][0] == (1 if body is None or "x y" in body else 0)


@pytest.mark.skipif(not os.path.exists(_REF_EXE), reason="oracle/_ref not built (needs /root/reference)")
def test_command_line_fuzz_against_the_compiled_reference(alignment_dir, tmp_path):
    """Random option sets -- valid, malformed and unknown -- through both programs with --onlycomputelowerbound=Y:
    same exit code, same stderr (but for the line naming the build), same stdout, same files with the same bytes.
    Left out on purpose: column ranges that are empty or outside the alignment (the reference reads out of bounds there)
    and --tbr=Y (its greedy-tree line; the dive replaces it)."""
    import random
    import shutil
    rnd = random.Random(int(os.environ.get("XMP_FUZZ_SEED", "5")))
    pool = [["-b", "400"], ["-b", "0"], ["-b", "abc"], ["-m", "5"], ["-m", "0"], ["-r"], ["-d"], ["-q"], ["-Bd"], ["-Bm"], ["-Bdm"], ["-Bp"],
            ["-BpC"], ["-BpL"], ["-Bdp"], ["-BC"], ["-BL"], ["-Bx"], ["-B"], ["--startcol=3"], ["--endcol=20"], ["--endcol=99999"],
            ["--startcol=5", "--endcol=64"], ["--displaynewbounds=Y"], ["--displaynewbounds=N"], ["--displaynewbounds=x"],
            ["--displaynewbounds"], ["--displaynewtrees=n"], ["--updatesecs=5"], ["--updatesecs=x"], ["--seed=7"], ["--pbmaxiters=1"],
            ["--improveinitupperbound=N"], ["--keeptempfiles=Y"], ["--savereordereddataset=Y"], ["--bogus=1"], ["-Z"], ["-o", "res.out"]]
    checked = 0
    for case in range(int(os.environ.get("XMP_FUZZ_CASES", "60"))):
        flags = [x for f in rnd.sample(pool, rnd.randint(0, 4)) for x in f]
        part = any(f.startswith("-Bp") or f == "-Bdp" for f in flags)
        aln = "h3.phy" if part else rnd.choice(["h3.phy", "mt-10.phy"])
        use_m = any(f.startswith("-B") and "m" in f[2:] for f in flags)
        res = []
        for exe, name in ((_REF_EXE + "_zinit" if use_m else _REF_EXE, "ref"), (EXE, "ours")):
            d = tmp_path / name
            shutil.rmtree(d, ignore_errors=True)
            d.mkdir()
            r = subprocess.run([exe, "--tbr=N", "--seed=1", "--onlycomputelowerbound=Y"] + flags + [str(alignment_dir / aln)], cwd=d,
                               capture_output=True, timeout=120)
            err = [l for l in r.stderr.decode(errors="replace").splitlines()
                   if l not in ("Single-processor version.", "B200 (CUDA) version searching on 1 GPU.")]
            res.append((r.returncode, err, {x: (d / x).read_bytes() for x in sorted(os.listdir(d))}, r.stdout))
        assert res[0] == res[1], (case, flags, aln)
        checked += 1
    assert checked >= 50


# ---- the command line end to end on the CPU, over a test double of the device library ------------------------------
def build_stub(d, extra=()):
    """tests/stub: the entry points host/fastdnamp.c calls, on top of the oracle's subtree search, plus the exchange
    entry points running the product's protocol source (xmp_b200/csrc/exchange_proto.h) over host memory.  Test
    infrastructure; the product never loads it."""
    inc = ["-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "oracle"), "-I" + os.path.join(ROOT, "tests", "stub")]
    subprocess.run(["gcc", "-O2", "-std=c99", "-fPIC", "-c", *extra, *inc, "-o", str(d / "xmp_stub.o"), os.path.join(ROOT, "tests", "stub", "xmp_stub.c")], check=True)
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-c", *extra, *inc, "-o", str(d / "xmp_stub_exchange.o"),
                    os.path.join(ROOT, "tests", "stub", "xmp_stub_exchange.cpp")], check=True)
    subprocess.run(["g++", "-shared", *extra, "-o", str(d / "libxmp_b200.so"), str(d / "xmp_stub.o"), str(d / "xmp_stub_exchange.o"),
                    "-L" + os.path.join(ROOT, "oracle"), "-loracle", "-Wl,-rpath," + os.path.join(ROOT, "oracle"), "-lpthread"], check=True)
    return d


@pytest.fixture(scope="module")
def stub_dir(tmp_path_factory):
    """Lets the host program's own logic run here: scheduler threads over the exchange protocol (shared bound, steal
    requests, termination), gather, depth-first sort, -m cap, writers, final report."""
    return build_stub(tmp_path_factory.mktemp("stub"))


def _run_over_stub(stub_dir, exe, args, cwd, devices=1, skew=False):
    env = dict(os.environ, LD_LIBRARY_PATH=str(stub_dir), XMP_STUB_DEVICES=str(devices), XMP_STUB_LOG="1")
    if skew:
        env["XMP_STUB_SKEW"] = "1"
    return subprocess.run([exe, "--gpus=%d" % devices] + args, cwd=cwd, capture_output=True, text=True, env=env, timeout=300)


@pytest.mark.parametrize("name", ["its10", "quick2", "primate", "53humans.12", "its14", "e4"])
@pytest.mark.parametrize("devices, skew", [(1, False), (3, False), (4, True), (8, True), (8, False)])
def test_whole_program_over_the_test_double(stub_dir, alignment_dir, goldens, tmp_path, name, devices, skew):
    """The reference's result file, byte for byte, from fastdnamp_b200's host side with 1, 3, 4 and 8 scheduler threads over
    the exchange protocol -- with `skew` every subtree starts on shard 0 and the others live on stolen work."""
    from tests.conftest import golden_output
    g = goldens[name]
    out = tmp_path / "out.nex"
    r = _run_over_stub(stub_dir, EXE, ["--tbr=N", "--seed=1", "--displaynewbounds=n", "--displaynewtrees=n", "--updatesecs=0", "-o", str(out),
                                       str(alignment_dir / g["alignment"])], tmp_path, devices, skew)
    assert r.returncode == 0, r.stderr
    assert out.read_bytes() == golden_output(name)
    import re
    assert re.search(r"Optimal trees found:\s+%d\n" % g["num_trees"], r.stderr) and "Time to perform branch and bound search:" in r.stderr
    if skew:
        assert r.stderr.count("imports") >= 1                  # idle shards were fed (how many depends on how long the search lasts)
        assert re.search(r"Steals served / received\s+[1-9]\d* / [1-9]", r.stderr)
    # raw format, quiet, capped: the reference's first trees in its order, nothing on stderr but the test double's log
    r = _run_over_stub(stub_dir, EXE, ["-q", "-r", "-m", "3", str(alignment_dir / g["alignment"])], tmp_path, devices, skew)
    want = [l.split("] ")[1] for l in golden_output(name).decode().splitlines() if "TREE FASTDNAMP_" in l][:3]
    assert r.returncode == 0 and r.stdout.splitlines() == want
    assert all(l.startswith("stub: ") for l in r.stderr.splitlines())


def test_scheduler_threads_under_thread_sanitizer(stub_dir, alignment_dir, goldens, tmp_path):
    """host/fastdnamp.c (with the host library's sources) compiled with -fsanitize=thread, four scheduler threads, skewed
    start: no data race reported in the shared bound / mailbox / termination logic, and still the reference's file."""
    from tests.conftest import golden_output
    exe = str(tmp_path / "cli_tsan")
    src = [os.path.join(ROOT, "host", f) for f in ("fastdnamp.c", "phylip.c", "prep.c", "trees.c", "partbound.c")]
    try:
        # the test double is instrumented too: the exchange protocol (steal requests, replies, the busy counter) runs in it
        stub_dir = build_stub(tmp_path, extra=("-fsanitize=thread", "-g"))
    except subprocess.CalledProcessError:
        pytest.skip("no thread sanitizer runtime here")
    c = subprocess.run(["gcc", "-O1", "-g", "-mpopcnt", "-std=c99", "-D_POSIX_C_SOURCE=200809L", "-fsanitize=thread", "-o", exe] + src +
                       ["-I" + os.path.join(ROOT, "host"), "-I" + os.path.join(ROOT, "include"), "-L" + str(stub_dir), "-lxmp_b200", "-lpthread"],
                       capture_output=True, text=True)
    if c.returncode != 0:
        pytest.skip("no thread sanitizer runtime here: " + c.stderr[-200:])
    for name in ("53humans.12", "its14"):
        out = tmp_path / (name + ".nex")
        r = _run_over_stub(stub_dir, exe, ["--tbr=N", "--seed=1", "--displaynewbounds=y", "--displaynewtrees=n", "--updatesecs=0", "-o", str(out),
                                           str(alignment_dir / goldens[name]["alignment"])], tmp_path, 4, True)
        if "unexpected memory mapping" in r.stderr:
            pytest.skip("the thread sanitizer cannot map its shadow memory in this container")
        assert r.returncode == 0 and "ThreadSanitizer" not in r.stderr, r.stderr[-2000:]
        assert out.read_bytes() == golden_output(name)


@pytest.mark.skipif(not os.path.exists(_REF_EXE), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("flags", [["-b", "1"], ["-b", "694"], ["-m", "2", "-r"], ["-d"], ["-Bdm"], ["--startcol=100", "--endcol=400", "-r"]])
def test_whole_program_with_options_against_the_compiled_reference(stub_dir, alignment_dir, tmp_path, flags):
    """Whole runs with options that change the search or the output (a bound that is too low: an empty tree block and
    'NO BEST TREE FOUND'; a bound equal to the optimum; a cap; -d; another lower bound; a column range): the reference's
    program and fastdnamp_b200 over the test double write the same file and the same messages."""
    res = []
    zinit = any(f.startswith("-B") and "m" in f[2:] for f in flags)
    for exe, name in ((_REF_EXE + "_zinit" if zinit else _REF_EXE, "ref"), (EXE, "ours")):
        d = tmp_path / name
        d.mkdir()
        args = ["--tbr=N", "--seed=1", "--displaynewbounds=n", "--displaynewtrees=n", "--updatesecs=0", "-o", "out.txt"] + flags + [str(alignment_dir / "its10.phy")]
        r = _run_over_stub(stub_dir, exe, args, d) if name == "ours" else subprocess.run([exe] + args, cwd=d, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr
        said = [l for l in r.stderr.splitlines() if l.startswith(("NO BEST TREE", "Improving", "Beginning", "Read ")) or "informative" in l]
        res.append(((d / "out.txt").read_bytes(), (d / "restBound.txt").read_bytes(), said))
    assert res[0] == res[1]


@pytest.mark.skipif(not os.path.exists(_REF_EXE), reason="oracle/_ref not built (needs /root/reference)")
def test_whole_program_fuzz_against_the_compiled_reference(stub_dir, tmp_path):
    """Random small alignments (ambiguity codes, repeated columns) and random option sets through the reference's program
    and through fastdnamp_b200 over the test double, with one to three scheduler threads: same result file, same
    restBound.txt, same number of optimal trees reported."""
    import re
    import shutil
    import numpy as np
    rng = np.random.default_rng(int(os.environ.get("XMP_FUZZ_SEED", "271828")))
    alphabet = np.array(list("ACGTACGTACGTACGTACGTRYSWKMBDHVN-?"))
    pool = [["-r"], ["-d"], ["-m", "1"], ["-m", "4"], ["-Bdm"], ["-Bm"], ["-Bp"], ["-BpC"], ["--startcol=2", "--endcol=30"], ["-q"], ["-b", "?"]]
    checked = 0
    for case in range(int(os.environ.get("XMP_FUZZ_CASES", "40"))):
        lo, hi = (int(x) for x in os.environ.get("XMP_FUZZ_TAXA", "4,9").split(","))
        n, sites = int(rng.integers(lo, hi)), int(rng.integers(32, 60))
        picks = [pool[i] for i in rng.choice(len(pool), int(rng.integers(0, 4)), replace=False)]
        picks = [p for i, p in enumerate(picks) if not (p[0].startswith("-B") and any(q[0].startswith("-B") for q in picks[:i]))]
        flags = [x if x != "?" else str(int(rng.integers(5, 90))) for p in picks for x in p]     # -b below, at or above the optimum
        # The partition bound is not a valid lower bound when ambiguity codes abound: the reference's own
        # answer then depends on its visiting order and need not be the optimum.  -Bp cases get unambiguous data.
        part = any(x.startswith("-Bp") for x in flags)
        base = rng.integers(0, 4 if part else len(alphabet), (n, sites))
        for t in range(1, n):
            keep = rng.random(sites) < 0.6
            base[t, keep] = base[rng.integers(0, t), keep]
        cols = rng.integers(0, sites, sites + int(rng.integers(0, sites)))
        f = tmp_path / ("wp%d.phy" % case)
        f.write_text("%d %d\n" % (n, len(cols)) + "".join("T%-9d%s\n" % (t, "".join(alphabet[base[t, cols]])) for t in range(n)))
        zinit = any(x.startswith("-B") and "m" in x[2:] for x in flags)
        res = []
        for exe, name in ((_REF_EXE + "_zinit" if zinit else _REF_EXE, "ref"), (EXE, "ours")):
            d = tmp_path / name
            shutil.rmtree(d, ignore_errors=True)
            d.mkdir()
            args = ["--tbr=N", "--seed=1", "--displaynewbounds=n", "--displaynewtrees=n", "--updatesecs=0", "-o", "out.txt"] + flags + [str(f)]
            if name == "ours":
                r = _run_over_stub(stub_dir, exe, args, d, devices=1 + case % 3, skew=case % 2 == 1)
            else:
                r = subprocess.run([exe] + args, cwd=d, capture_output=True, text=True, timeout=300)
            rb = (d / "restBound.txt").read_bytes() if (d / "restBound.txt").exists() else None
            out = (d / "out.txt").read_bytes() if (d / "out.txt").exists() else None
            res.append((r.returncode, out, rb, "NO BEST TREE" in r.stderr))
        if res[0][0] != 0 and res[1][0] != 0:
            continue                                  # e.g. fewer than four taxa survive: both refuse
        m = re.search(rb"Each tree has score (\d+)", res[0][1] or b"")
        if m and int(m.group(1)) > 1_000_000:
            continue                                  # -d left no site at all and -Bm made the reference print an uninitialised length
        assert res[0] == res[1], (case, flags)
        checked += 1
    assert checked >= 30


def test_progress_reports(stub_dir, alignment_dir, goldens, tmp_path):
    """--updatesecs=<n> (reference src/yanbader.c:28-44): every n seconds the share of the search space already settled --
    here 1 minus what the open subtrees stand for -- the best length so far and a time estimate; none with 0 or -q."""
    import re
    from tests.conftest import golden_output
    aln = str(alignment_dir / goldens["mt-10"]["alignment"])
    out = tmp_path / "out.nex"
    r = _run_over_stub(stub_dir, EXE, ["--tbr=N", "--seed=1", "--displaynewbounds=n", "--displaynewtrees=n", "--updatesecs=1", "-o", str(out), aln],
                       tmp_path, devices=2)
    assert r.returncode == 0 and out.read_bytes() == golden_output("mt-10")
    pct = [float(x) for x in re.findall(r"^([\d.]+)% complete\.  Best length so far: \d+\.  Estimated time remaining: ", r.stderr, re.M)]
    assert pct and pct == sorted(pct) and 0 <= pct[0] and pct[-1] <= 100
    for quiet in (["--updatesecs=0"], ["-q"]):
        r = _run_over_stub(stub_dir, EXE, ["--tbr=N", "--seed=1", "--displaynewbounds=n", "--displaynewtrees=n", "-o", str(out)] + quiet + [aln], tmp_path)
        assert r.returncode == 0 and "% complete" not in r.stderr
