import gzip
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session", autouse=True)
def built():
    """Build everything once per session (host lib, oracle, CUDA lib); cheap when up to date."""
    import __graft_entry__
    __graft_entry__.build()


@pytest.fixture(scope="session")
def goldens():
    return json.load(open(os.path.join(GOLDEN, "datasets.json")))


@pytest.fixture(scope="session")
def alignment_dir(tmp_path_factory):
    """Unpacks the bundled input alignments (tests/golden/alignments.json.gz) into a temp dir."""
    d = tmp_path_factory.mktemp("alignments")
    bundle = json.loads(gzip.open(os.path.join(GOLDEN, "alignments.json.gz")).read())
    for name, text in bundle.items():
        p = d / name
        p.parent.mkdir(parents=True, exist_ok=True)
        p.write_bytes(text.encode("latin-1"))
    return d


def golden_trees(name):
    return gzip.open(os.path.join(GOLDEN, "trees", name + ".txt.gz")).read().decode().split()


def golden_output(name):
    return gzip.open(os.path.join(GOLDEN, "trees", name + ".out.gz")).read()


@pytest.fixture(autouse=True)
def _rss_report(request):
    """With XMP_TEST_RSS=1 prints the resident set size after every test (debugging aid)."""
    yield
    if os.environ.get("XMP_TEST_RSS"):
        import resource
        print("\n[rss] %s: %d MiB" % (request.node.name, resource.getrusage(resource.RUSAGE_SELF).ru_maxrss // 1024), flush=True)
