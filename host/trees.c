/* Trees as insertion paths: reconstruction, the reference's Newick text, its depth-first order and
 * its result-file format (reference src/common.c:371-454 OutputResults, :933-974 PrintTree,
 * src/yanbader.c:103,238-243 edge order).
 *
 * Edge ids (include/xmp_cuda.h): a tree on taxa 0..m-1 has edges 0..2m-4, each named after the node
 * below it.  Edge 0 / 1: leaves 1 / 2; edge 2: the first internal node (LEFT = leaf 1, RIGHT = leaf 2),
 * child of the root, which is taxon 0.  Inserting taxon t on edge v adds leaf edge 2t-3 and internal
 * edge 2t-2, whose LEFT child is the new leaf and whose RIGHT child is v (src/common.c:735-736).
 */
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include "xmp_host.h"

int xmph_fail(const char *fmt, ...);

#define NONE 0xFFu
#define MAXE (2 * XMPH_MAXTAXA)

typedef struct {
	unsigned char par[MAXE], kid[MAXE][2];
	unsigned top;     /* the root's only child */
} topo;

static int edge_is_leaf(unsigned e) { return e == 0 || (e & 1); }
static unsigned edge_taxon(unsigned e) { return (e + 3) / 2; }

static void topo_base(topo *t) {
	t->par[0] = t->par[1] = 2;
	t->par[2] = NONE;
	t->kid[0][0] = t->kid[0][1] = t->kid[1][0] = t->kid[1][1] = NONE;
	t->kid[2][0] = 0;
	t->kid[2][1] = 1;
	t->top = 2;
}

static void topo_insert(topo *t, unsigned taxon, unsigned v) {
	const unsigned a = 2 * taxon - 3, b = 2 * taxon - 2, pv = t->par[v];
	t->par[b] = (unsigned char) pv;
	if (pv == NONE) t->top = b; else t->kid[pv][t->kid[pv][1] == v] = (unsigned char) b;
	t->kid[b][0] = (unsigned char) a;
	t->kid[b][1] = (unsigned char) v;
	t->par[a] = t->par[v] = (unsigned char) b;
	t->kid[a][0] = t->kid[a][1] = NONE;
}

/* Pre-order list of the edges: own edge, LEFT subtree, RIGHT subtree. */
static unsigned topo_preorder(const topo *t, unsigned char *out) {
	unsigned char stack[MAXE];
	unsigned sp = 0, n = 0;
	stack[sp++] = (unsigned char) t->top;
	while (sp) {
		unsigned e = stack[--sp];
		out[n++] = (unsigned char) e;
		if (!edge_is_leaf(e)) {
			stack[sp++] = t->kid[e][1];
			stack[sp++] = t->kid[e][0];
		}
	}
	return n;
}

static char *put_unsigned(char *o, unsigned v) {
	char tmp[12];
	int n = 0;
	do { tmp[n++] = (char) ('0' + v % 10); v /= 10; } while (v);
	while (n) *o++ = tmp[--n];
	return o;
}

static char *put_subtree(const topo *t, unsigned e, const unsigned *taxon_map, char *o) {
	if (edge_is_leaf(e)) return put_unsigned(o, taxon_map[edge_taxon(e)] + 1);
	*o++ = '(';
	o = put_subtree(t, t->kid[e][0], taxon_map, o);
	*o++ = ',';
	o = put_subtree(t, t->kid[e][1], taxon_map, o);
	*o++ = ')';
	return o;
}

size_t xmph_path_to_newick(const unsigned char *path, unsigned num_taxa, const unsigned *taxon_map, char *out) {
	topo t;
	char *o = out;
	topo_base(&t);
	for (unsigned k = 3; k < num_taxa; ++k) topo_insert(&t, k, path[k]);
	*o++ = '(';
	o = put_unsigned(o, taxon_map[0] + 1);
	*o++ = ',';
	o = put_subtree(&t, t.top, taxon_map, o);
	*o++ = ')';
	*o = 0;
	return (size_t) (o - out);
}

void xmph_path_to_dfs(const unsigned char *path, unsigned num_taxa, unsigned char *dfs) {
	topo t;
	unsigned char order[MAXE];
	topo_base(&t);
	dfs[0] = dfs[1] = dfs[2] = 0;
	for (unsigned k = 3; k < num_taxa; ++k) {
		unsigned n = topo_preorder(&t, order), r = 0;
		while (r < n && order[r] != path[k]) ++r;
		dfs[k] = (unsigned char) r;
		topo_insert(&t, k, path[k]);
	}
}

void xmph_dfs_to_path(const unsigned char *dfs, unsigned num_taxa, unsigned char *path) {
	topo t;
	unsigned char order[MAXE];
	topo_base(&t);
	path[0] = path[1] = path[2] = 0;
	for (unsigned k = 3; k < num_taxa; ++k) {
		topo_preorder(&t, order);
		path[k] = order[dfs[k]];
		topo_insert(&t, k, path[k]);
	}
}

/* ---- depth-first order of a tree set ------------------------------------------------------------------
 * mh3 has 408,945 optimal trees, 53humans.26 9.8 million: computing the keys (a pre-order walk per taxon per
 * tree) runs on all cores, and the sort is an MSD radix sort on the key bytes -- byte k is below 2k-3, the
 * first three are zero -- over a permutation, so that records move once. */
typedef struct { const unsigned char *paths; unsigned char *keys; size_t n; unsigned num_taxa, next, chunk; } key_job;

static void *key_thread(void *x) {
	key_job *j = (key_job *) x;
	for (;;) {
		size_t lo = (size_t) __atomic_fetch_add(&j->next, 1u, __ATOMIC_RELAXED) * j->chunk;
		if (lo >= j->n) return NULL;
		size_t hi = lo + j->chunk < j->n ? lo + j->chunk : j->n;
		for (size_t i = lo; i < hi; ++i) xmph_path_to_dfs(j->paths + i * j->num_taxa, j->num_taxa, j->keys + i * j->num_taxa);
	}
}

static void radix_sort_keys(const unsigned char *keys, unsigned key_len, unsigned *idx, unsigned *tmp, size_t n, unsigned depth) {
	while (depth < key_len) {
		if (n < 2) return;
		if (n <= 24) {                                   /* insertion sort on the remaining bytes */
			for (size_t i = 1; i < n; ++i) {
				unsigned v = idx[i];
				size_t k = i;
				while (k > 0 && memcmp(keys + (size_t) idx[k - 1] * key_len + depth, keys + (size_t) v * key_len + depth, key_len - depth) > 0) {
					idx[k] = idx[k - 1];
					--k;
				}
				idx[k] = v;
			}
			return;
		}
		size_t count[257];
		memset(count, 0, sizeof count);
		for (size_t i = 0; i < n; ++i) ++count[keys[(size_t) idx[i] * key_len + depth] + 1];
		if (count[keys[(size_t) idx[0] * key_len + depth] + 1] == n) { ++depth; continue; }      /* all equal at this byte */
		for (unsigned b = 1; b <= 256; ++b) count[b] += count[b - 1];
		size_t start[257];
		memcpy(start, count, sizeof start);
		for (size_t i = 0; i < n; ++i) tmp[count[keys[(size_t) idx[i] * key_len + depth]]++] = idx[i];
		memcpy(idx, tmp, n * sizeof(unsigned));
		for (unsigned b = 0; b < 256; ++b) {
			const size_t lo = start[b], hi = start[b + 1];
			if (hi - lo > 1) radix_sort_keys(keys, key_len, idx + lo, tmp + lo, hi - lo, depth + 1);
		}
		return;
	}
}

void xmph_sort_paths_dfs(unsigned char *paths, size_t n, unsigned num_taxa) {
	if (n < 2) return;
	unsigned char *keys = (unsigned char *) malloc(n * num_taxa);
	unsigned char *sorted = (unsigned char *) malloc(n * num_taxa);
	unsigned *idx = (unsigned *) malloc(n * sizeof(unsigned)), *tmp = (unsigned *) malloc(n * sizeof(unsigned));
	if (!keys || !sorted || !idx || !tmp || n > 0xffffffffu) { free(keys); free(sorted); free(idx); free(tmp); return; }
	key_job job = { paths, keys, n, num_taxa, 0, 4096 };
	long cores = sysconf(_SC_NPROCESSORS_ONLN);
	unsigned threads = cores > 0 ? (unsigned) cores : 1, started = 0;
	pthread_t th[64];
	if (threads > 64) threads = 64;
	if (n < 20000) threads = 1;
	for (; started + 1 < threads; ++started) if (pthread_create(&th[started], NULL, key_thread, &job)) break;
	key_thread(&job);
	for (unsigned i = 0; i < started; ++i) pthread_join(th[i], NULL);
	for (size_t i = 0; i < n; ++i) idx[i] = (unsigned) i;
	radix_sort_keys(keys, num_taxa, idx, tmp, n, 3);         /* bytes 0..2 are zero for every tree */
	for (size_t i = 0; i < n; ++i) memcpy(sorted + i * num_taxa, paths + (size_t) idx[i] * num_taxa, num_taxa);
	memcpy(paths, sorted, n * num_taxa);
	free(keys); free(sorted); free(idx); free(tmp);
}

typedef struct {
	const unsigned char *paths; const unsigned *taxon_map; unsigned num_taxa; int nexus;
	size_t first, count, len;
	char *buf;
} fmt_job;

static void *fmt_thread(void *x) {
	fmt_job *j = (fmt_job *) x;
	char *o = j->buf;
	for (size_t k = j->first; k < j->first + j->count; ++k) {
		if (j->nexus) {
			memcpy(o, "\tTREE FASTDNAMP_", 16); o += 16;
			size_t v = k + 1;
			char digits[24];
			int nd = 0;
			do { digits[nd++] = (char) ('0' + v % 10); v /= 10; } while (v);
			while (nd) *o++ = digits[--nd];
			memcpy(o, " = [&U] ", 8); o += 8;
		}
		o += xmph_path_to_newick(j->paths + k * j->num_taxa, j->num_taxa, j->taxon_map, o);
		*o++ = ';';
		*o++ = '\n';
	}
	j->len = (size_t) (o - j->buf);
	return NULL;
}

int xmph_write_trees(FILE *out, int nexus, unsigned score, const xmph_alignment *a, const unsigned *taxon_map,
                     const unsigned char *paths, size_t n_trees) {
	char *buf = (char *) malloc(16 * (size_t) a->num_taxa + 16);
	if (nexus) {
		fprintf(out,
		        "#NEXUS\n"
		        "[ Trees produced by fastDNAmp: Exact Maximum Parsimony Phylogenetic Tree Search v1.1\n"
		        "> Each tree has score %u. ]\n"
		        "BEGIN TREES;\n"
		        "\tTRANSLATE\n", score);
		for (unsigned i = 0; i < a->num_taxa; ++i) {
			fprintf(out, "\t\t%u\t%s%c\n", i + 1, a->labels[i], i + 1 < a->num_taxa ? ',' : ';');
		}
	}
	(void) buf;
	/* tree lines: formatted a block of trees per thread into memory, written out in order */
	{
		enum { BLOCK = 16384 };
		long cores = sysconf(_SC_NPROCESSORS_ONLN);
		unsigned threads = cores > 0 ? (unsigned) cores : 1;
		if (threads > 32) threads = 32;
		if (n_trees < 4 * BLOCK) threads = 1;
		const size_t line_max = 16 * (size_t) a->num_taxa + 64;
		fmt_job jobs[32];
		pthread_t th[32];
		int failed = 0;
		for (unsigned t = 0; t < threads; ++t) {
			jobs[t].buf = (char *) malloc(BLOCK * line_max);
			if (!jobs[t].buf) failed = 1;
		}
		for (size_t base = 0; base < n_trees && !failed; base += (size_t) threads * BLOCK) {
			unsigned used = 0;
			for (unsigned t = 0; t < threads; ++t) {
				const size_t first = base + (size_t) t * BLOCK;
				if (first >= n_trees) break;
				jobs[t].paths = paths; jobs[t].taxon_map = taxon_map; jobs[t].num_taxa = a->num_taxa; jobs[t].nexus = nexus;
				jobs[t].first = first;
				jobs[t].count = n_trees - first < BLOCK ? n_trees - first : BLOCK;
				++used;
			}
			unsigned started = 0;
			for (unsigned t = 1; t < used; ++t) {
				if (pthread_create(&th[t], NULL, fmt_thread, &jobs[t])) break;
				started = t;
			}
			fmt_thread(&jobs[0]);
			for (unsigned t = started + 1; t < used; ++t) fmt_thread(&jobs[t]);      /* threads that could not be started */
			for (unsigned t = 1; t <= started; ++t) pthread_join(th[t], NULL);
			for (unsigned t = 0; t < used; ++t) fwrite(jobs[t].buf, 1, jobs[t].len, out);
		}
		for (unsigned t = 0; t < threads; ++t) free(jobs[t].buf);
		if (failed) { free(buf); return xmph_fail("Out of memory while writing result trees."); }
	}
	if (nexus) fprintf(out, "END;\n");
	free(buf);
	return ferror(out) ? xmph_fail("Error writing result trees.") : 0;
}
