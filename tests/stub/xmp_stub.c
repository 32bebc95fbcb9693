/* TEST INFRASTRUCTURE ONLY -- a CPU test double for libxmp_b200.so.
 *
 * Implements the part of include/xmp_cuda.h that host/fastdnamp.c calls, on top of the oracle's subtree search
 * (oracle/search_oracle.c), so that the command line's own logic -- the one-thread-per-GPU scheduler with its shared
 * bound and job mailbox, the gather, the depth-first sort, the -m cap and the writers -- can run, and be put under
 * sanitizers, on a machine without a GPU.  It is never installed, linked or loaded by the product: tests put its
 * directory in LD_LIBRARY_PATH for one child process (tests/test_cli.py).  XMP_STUB_DEVICES = how many "GPUs" it reports;
 * XMP_STUB_SKEW = deal every subtree to shard 0, so that the others live on stolen work; XMP_STUB_LOG = report steals.
 *
 * A "batch" is one open subtree searched to completion; search_begin deals the partial trees of a shallow tree size to
 * the shards by index, like the library's shard_index / shard_count. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "xmp_cuda.h"
#include "oracle.h"

#include "xmp_stub.h"

static char g_err[256] = "";
const char *xmp_cuda_last_error(void) { return g_err; }
int stub_fail(int code, const char *msg) {
	snprintf(g_err, sizeof g_err, "%s", msg);
	return code;
}
#define fail stub_fail
int xmp_cuda_device_count(void) {
	const char *e = getenv("XMP_STUB_DEVICES");
	return e ? atoi(e) : 1;
}
int xmp_cuda_init(int device, xmp_ctx **out) {
	if (device < 0 || device >= xmp_cuda_device_count()) return fail(XMP_ERR_NODEVICE, "no such device");
	*out = (xmp_ctx *) calloc(1, sizeof(xmp_ctx));
	(*out)->device = device;
	return XMP_OK;
}
void xmp_cuda_destroy(xmp_ctx *c) {
	if (!c) return;
	free(c->rows); free(c->weights); free(c->rest_bound); free(c->taxon_map); free(c->jobs); free(c->trees); free(c->staged);
	free(c);
}
int xmp_cuda_search_reserve(xmp_ctx *c, unsigned num_taxa, unsigned n_blocks, size_t mem_budget) {
	(void) mem_budget;
	return c && num_taxa >= 4 && n_blocks ? XMP_OK : fail(XMP_ERR_ARG, "cannot reserve");
}
int xmp_cuda_load(xmp_ctx *c, const void *rows, const unsigned *weights, unsigned num_taxa, unsigned n_blocks,
                  const unsigned *rest_bound, unsigned constant_weight) {
	const size_t nbytes = (size_t) n_blocks * 16;
	free(c->rows); free(c->weights); free(c->rest_bound); free(c->taxon_map);
	c->n = num_taxa;
	c->n_blocks = n_blocks;
	c->constant_weight = constant_weight;
	c->rows = (uint8_t *) malloc(nbytes * num_taxa + 1);
	memcpy(c->rows, rows, nbytes * num_taxa);
	c->weights = (unsigned *) malloc((2 * nbytes + 1) * sizeof(unsigned));
	for (size_t i = 0; i < 2 * nbytes; ++i) c->weights[i] = weights ? weights[i] : 1;
	c->rest_bound = (unsigned *) calloc(num_taxa, sizeof(unsigned));
	if (rest_bound) memcpy(c->rest_bound, rest_bound, num_taxa * sizeof(unsigned));
	c->taxon_map = (unsigned *) malloc(num_taxa * sizeof(unsigned));
	for (unsigned i = 0; i < num_taxa; ++i) c->taxon_map[i] = i;
	return XMP_OK;
}
#define push_job stub_push_job
void stub_push_job(xmp_ctx *c, const uint8_t *path, unsigned m) {
	if (c->n_jobs == c->cap_jobs) {
		c->cap_jobs = c->cap_jobs ? 2 * c->cap_jobs : 256;
		c->jobs = (uint8_t *) realloc(c->jobs, c->cap_jobs * c->n);
	}
	uint8_t *j = c->jobs + c->n_jobs++ * c->n;
	memcpy(j, path, c->n);
	j[0] = (uint8_t) m;
}
int xmp_cuda_search_begin(xmp_ctx *c, const xmp_search_params *prm) {
	if (!c || !c->rows) return fail(XMP_ERR_ARG, "no problem loaded");
	const unsigned shards = prm->shard_count ? prm->shard_count : 1;
	c->bound = prm->bound;
	c->n_jobs = c->head = c->n_trees = 0;
	memset(&c->st, 0, sizeof c->st);
	/* split level: the first tree size with at least 8 partial trees per shard (at most n - 1 taxa) */
	unsigned L = 3;
	unsigned long long count = 1;
	while (L + 1 < c->n && count < 8ull * shards) { count *= 2 * L - 3; ++L; }
	uint8_t path[XMP_MAX_TAXA + 1];
	memset(path, 0, sizeof path);
	unsigned long long k = 0;
	for (;;) {
		/* XMP_STUB_SKEW: shard 0 gets everything, the others start idle and must be fed through the mailbox */
		if (getenv("XMP_STUB_SKEW") ? prm->shard_index == 0 : k % shards == prm->shard_index) push_job(c, path, L);
		++k;
		unsigned t = L;                                   /* next combination, last taxon fastest */
		while (t > 3) {
			--t;
			if (++path[t] < 2 * t - 3) break;
			path[t] = 0;
			if (t == 3) { t = 0; break; }
		}
		if (L == 3 || t == 0) break;
	}
	c->begun = 1;
	c->steals_served = c->steals_received = 0;
	if (c->xch_connected) stub_xch_bound_local(c, c->bound);
	return XMP_OK;
}
int xmp_cuda_search_run(xmp_ctx *c, unsigned max_batches, int *done) {
	if (!c || !c->begun) return fail(XMP_ERR_ARG, "no search in progress");
	unsigned batches = 0;
	while (c->n_jobs > c->head && (max_batches == 0 || batches < max_batches)) {
		uint8_t path[XMP_MAX_TAXA + 1];
		if (c->xch_connected) {                       /* what the kernels do: read the (peer-lowered) bound word per launch */
			const unsigned b = stub_xch_bound_in(c);
			if (b < c->bound) { c->bound = b; c->n_trees = 0; }
		}
		memcpy(path, c->jobs + --c->n_jobs * c->n, c->n);
		const unsigned m = path[0];
		path[0] = 0;
		xo_problem p = { c->n, (size_t) c->n_blocks * 16, c->rows, c->weights, c->rest_bound, c->taxon_map, c->constant_weight, c->bound, 0x7fffffffu };
		xo_result r;
		memset(&r, 0, sizeof r);
		if (xo_search_subtree(&p, path, m, &r)) return fail(XMP_ERR_ARG, "oracle search failed");
		if (r.score < c->bound) {
			c->bound = r.score;
			c->n_trees = 0;
			if (c->xch_connected) stub_xch_bound_out(c, r.score);
		}
		const size_t k = r.paths_len / c->n;
		if (k && r.score == c->bound) {
			if (c->n_trees + k > c->cap_trees) {
				c->cap_trees = 2 * (c->n_trees + k);
				c->trees = (uint8_t *) realloc(c->trees, c->cap_trees * c->n);
			}
			memcpy(c->trees + c->n_trees * c->n, r.paths, k * c->n);
			c->n_trees += k;
		}
		c->st.nodes += r.nodes;
		c->st.score_calls += r.score_calls;
		c->st.full_trees += k;
		c->st.iterations += 1;
		xo_result_free(&r);
		++batches;
		if (c->xch_connected && stub_xch_after_job(c)) return fail(XMP_ERR_ARG, "exchange failed");
	}
	if (done) *done = c->n_jobs == c->head;
	return XMP_OK;
}
int xmp_cuda_search_get_bound(xmp_ctx *c, unsigned *bound) { *bound = c->bound; return XMP_OK; }
int xmp_cuda_search_set_bound(xmp_ctx *c, unsigned bound) {
	if (bound < c->bound) { c->bound = bound; c->n_trees = 0; }
	return XMP_OK;
}
int xmp_cuda_search_pending(xmp_ctx *c, unsigned long long *per_level) {
	memset(per_level, 0, (XMP_MAX_TAXA + 1) * sizeof *per_level);
	for (size_t i = c->head; i < c->n_jobs; ++i) per_level[c->jobs[i * c->n]]++;
	return XMP_OK;
}
int xmp_cuda_search_export_jobs(xmp_ctx *c, unsigned max_nodes, uint8_t *jobs, unsigned *n_jobs) {
	unsigned k = 0;
	while (k < max_nodes && c->n_jobs - c->head > 1) {        /* oldest first; keep one to work on */
		memcpy(jobs + (size_t) k++ * c->n, c->jobs + c->head++ * c->n, c->n);
	}
	*n_jobs = k;
	return XMP_OK;
}
int xmp_cuda_search_import_jobs(xmp_ctx *c, const uint8_t *jobs, unsigned n_jobs) {
	if (getenv("XMP_STUB_LOG")) fprintf(stderr, "stub: device %d imports %u jobs\n", c->device, n_jobs);
	for (unsigned i = 0; i < n_jobs; ++i) push_job(c, jobs + (size_t) i * c->n, jobs[(size_t) i * c->n]);
	return XMP_OK;
}
int xmp_cuda_search_result(xmp_ctx *c, unsigned *score, unsigned long long *n_trees, uint8_t *paths, size_t cap) {
	if (score) *score = c->bound;
	if (n_trees) *n_trees = c->n_trees;
	if (paths) memcpy(paths, c->trees, (cap < c->n_trees ? cap : c->n_trees) * c->n);
	return XMP_OK;
}
int xmp_cuda_search_stored(xmp_ctx *c, unsigned long long *n_stored) { *n_stored = c->n_trees; return XMP_OK; }
int xmp_cuda_search_stats(xmp_ctx *c, xmp_search_stats *out) {
	*out = c->st;
	out->steals_served = c->steals_served;
	out->steals_received = c->steals_received;
	return XMP_OK;
}
