"""A CPU stand-in for xmp_b200.cuda.Context's search_* interface, backed by the oracle.  TEST
INFRASTRUCTURE ONLY: it lets the multi-rank protocol of xmp_b200.distributed run over gloo on a
machine without GPUs.  A "batch" here is one job (one open subtree) searched to completion."""
import itertools

import numpy as np

from tests import oracle_lib as ol


class OracleEngine:
    def __init__(self, problem, lazy_done=True):
        self.p = problem
        self.lazy_done = lazy_done      # like the CUDA engine before it learnt to report an empty frontier at once
        self.wrapped_up = False
        self.n = problem.num_taxa
        self.jobs = []          # (m, path) open subtrees
        self.found = []         # (score, path)
        self.bound = problem.bound
        self.nodes = 0

    def search_begin(self, bound, shard_index=0, shard_count=1, split_level=5):
        self.bound = bound
        s = min(split_level, self.n - 1)
        k = 0
        for choice in itertools.product(*[range(2 * t - 3) for t in range(3, s)]):
            if k % shard_count == shard_index:
                path = np.zeros(self.n, np.uint8)
                path[3:s] = choice
                self.jobs.append((s, path))
            k += 1

    def search_run(self, max_batches):
        done = 0
        while self.jobs and (max_batches == 0 or done < max_batches):
            m, path = self.jobs.pop()
            score, paths, nodes = ol.search_subtree(self.p, path, m, self.bound)
            self.nodes += nodes
            if score < self.bound:
                self.bound = score
            self.found.extend((score, q) for q in paths)
            done += 1
        if self.lazy_done and done:
            return False          # an engine may notice that it ran dry only on the next call
        self.wrapped_up = not self.jobs
        return not self.jobs

    def search_get_bound(self):
        return self.bound

    def search_set_bound(self, b):
        self.bound = min(self.bound, b)

    def search_pending(self):
        a = np.zeros(101, np.uint64)
        for m, _ in self.jobs:
            a[m] += 1
        return a

    def search_export_jobs(self, max_nodes):
        k = min(max_nodes, len(self.jobs))
        out = np.zeros((k, self.n), np.uint8)
        for i in range(k):
            m, path = self.jobs.pop(0)       # shallow end first, like the reference's thief
            out[i] = path
            out[i, 0] = m
        return out

    def search_import_jobs(self, jobs):
        for j in jobs:
            path = np.array(j, np.uint8)
            m = int(path[0])
            path[0] = 0
            self.jobs.append((m, path))
        self.wrapped_up = False

    def search_result(self):
        assert self.wrapped_up and not self.jobs, "search_result before search_run reported done"
        paths = [q for s, q in self.found if s == self.bound]
        arr = np.stack(paths) if paths else np.zeros((0, self.n), np.uint8)
        return self.bound, len(paths), arr
