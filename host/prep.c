/* Preprocessing: alignment -> packed, ordered, bounded search problem.
 *
 * Mirrors the reference's InitTreeData pipeline (reference src/common.c:1113-1224) stage by stage so
 * that the search starts from exactly the same taxon order, column order, weights, upper bound and
 * restBound[] as fastdnamp does:
 *
 *   recode columns        src/common.c:1293-1388   state bits in order of first appearance per column
 *   merge equal columns   src/common.c:1159-1160, 1418-1438, 2813-2815
 *   sort by weight        src/common.c:1163, 2821-2832   descending weight, then ascending column bytes
 *   drop uninformative    src/common.c:1445-1496, 2660-2701
 *   order taxa            src/ordertaxa.c:337-477  (MAXMINIORDER)
 *   MST upper bound       src/common.c:2597-2604, src/seq_b1.c:58-152,260-275
 *   lower bounds          src/bound.c:12-41, src/distbasesbound.c:21-160, src/mstbound.c:67-93
 *   pack                  src/common.c:1034-1109, here with BYTESPERBLOCK 16
 */
#include <ctype.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include "xmp_host.h"

int xmph_fail(const char *fmt, ...);

void xmph_default_settings(xmph_settings *s) {
	s->options = XMPH_OPT_NEXUSFMT | XMPH_OPT_DISPLAYNEWBOUNDS | XMPH_OPT_DISPLAYNEWTREES |
	             XMPH_OPT_USEDISTBASESBOUNDREST | XMPH_OPT_IMPROVEINITUPPERBOUND | XMPH_OPT_TBR;
	s->bound = XMPH_NO_LIMIT;
	s->max_trees = XMPH_NO_LIMIT;
	s->start_col = 0;
	s->end_col = XMPH_NO_LIMIT;
	s->rest_bound_file = NULL;
	s->part_strategy = 0;     /* PBS_CLIPPEDANDADDITIVE, src/common.c:64 */
	s->part_max_stale = 2;    /* src/common.c:65 */
	s->threads = 0;
}

/* IUPAC code -> set over {A=1, C=2, G=4, T=8}; gap and '?' are fully ambiguous (src/common.c:891-916). */
static int base_set(int ch) {
	switch (toupper(ch)) {
	case 'A': return 1;
	case 'C': return 2;
	case 'G': return 4;
	case 'U': case 'T': return 8;
	case 'R': return 1 | 4;
	case 'Y': return 2 | 8;
	case 'S': return 2 | 4;
	case 'W': return 1 | 8;
	case 'K': return 4 | 8;
	case 'M': return 1 | 2;
	case 'B': return 2 | 4 | 8;
	case 'D': return 1 | 4 | 8;
	case 'H': return 1 | 2 | 8;
	case 'V': return 1 | 2 | 4;
	case '-': case '?': case 'N': return 15;
	default: return -1;
	}
}

/* ---- columns ------------------------------------------------------------------------------ */
/* A column table is site-major: col[i] holds num_taxa recoded state sets, wt[i] its weight. */
typedef struct {
	unsigned n, num_taxa;
	unsigned char *col;
	unsigned *wt;
} columns;

/* Re-labels the states of every column by order of first appearance (taxon order, A<C<G<T inside an
 * ambiguous set), so that columns with the same partition structure become byte-identical. */
/* A tiny parallel-for over independent items (site tiles, sort runs, pair distances); threads == 0 means all cores. */
typedef struct { void (*fn)(void *, unsigned); void *arg; unsigned n, next; } par_for;

static void *par_for_thread(void *x) {
	par_for *pf = (par_for *) x;
	for (;;) {
		unsigned i = __atomic_fetch_add(&pf->next, 1u, __ATOMIC_RELAXED);
		if (i >= pf->n) return NULL;
		pf->fn(pf->arg, i);
	}
}

static void parallel_for(unsigned n, void (*fn)(void *, unsigned), void *arg, unsigned threads, size_t work_per_item) {
	par_for pf = { fn, arg, n, 0 };
	pthread_t th[64];
	unsigned started = 0;
	if (!threads) {
		long cores = sysconf(_SC_NPROCESSORS_ONLN);
		threads = cores > 0 ? (unsigned) cores : 1;
	}
	if (threads > 64) threads = 64;
	if (threads > n) threads = n;
	if ((double) work_per_item * n < 4e6) threads = 1;           /* not worth a thread start */
	for (; started + 1 < threads; ++started) if (pthread_create(&th[started], NULL, par_for_thread, &pf)) break;
	par_for_thread(&pf);
	for (unsigned i = 0; i < started; ++i) pthread_join(th[i], NULL);
}

/* Recoding is independent per site: tiles of sites go to threads, every kept or dropped site is written at its
 * place in the range first and dropped ones (-d) are squeezed out afterwards. */
enum { RECODE_TILE = 4096 };
typedef struct {
	const xmph_alignment *a;
	const xmph_settings *s;
	unsigned char *col, *keep;
	unsigned start, end;
	unsigned bad_site;          /* lowest site with a character that is no state set (UINT_MAX: none) */
} recode_job;

static void recode_tile(void *arg, unsigned tile) {
	recode_job *j = (recode_job *) arg;
	const xmph_alignment *a = j->a;
	const unsigned lo = j->start + tile * RECODE_TILE, hi = lo + RECODE_TILE < j->end ? lo + RECODE_TILE : j->end;
	for (unsigned site = lo; site < hi; ++site) {
		int relabel[4] = { -1, -1, -1, -1 };
		int next = 0, keep = 1;
		unsigned char *out = j->col + (size_t) (site - j->start) * a->num_taxa;
		for (unsigned t = 0; t < a->num_taxa; ++t) {
			int set = base_set(a->chars[(size_t) t * a->seq_len + site]);
			if (set < 0) {
				unsigned seen = __atomic_load_n(&j->bad_site, __ATOMIC_RELAXED);
				while (site < seen && !__atomic_compare_exchange_n(&j->bad_site, &seen, site, 0, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
				return;
			}
			if ((set & (set - 1)) && (j->s->options & XMPH_OPT_DISCARDMISSAMBIG)) {
				keep = 0;
				break;
			}
			int coded = 0;
			for (int b = 0; b < 4; ++b) {
				if (!(set & (1 << b))) continue;
				if (relabel[b] < 0) relabel[b] = next++;
				coded |= 1 << relabel[b];
			}
			out[t] = (unsigned char) coded;
		}
		j->keep[site - j->start] = (unsigned char) keep;
	}
}

static int recode(const xmph_alignment *a, const xmph_settings *s, columns *c) {
	unsigned end = s->end_col > a->seq_len ? a->seq_len : s->end_col;
	const unsigned start = s->start_col < end ? s->start_col : end, range = end - start;
	c->num_taxa = a->num_taxa;
	c->col = (unsigned char *) malloc((size_t) a->num_taxa * ((size_t) range + 1));
	c->n = 0;
	unsigned char *keep = (unsigned char *) malloc((size_t) range + 1);
	recode_job job = { a, s, c->col, keep, start, end, 0xFFFFFFFFu };
	parallel_for((range + RECODE_TILE - 1) / RECODE_TILE, recode_tile, &job, s->threads, (size_t) RECODE_TILE * a->num_taxa * 4);
	if (job.bad_site != 0xFFFFFFFFu) {
		unsigned t = 0;
		while (base_set(a->chars[(size_t) t * a->seq_len + job.bad_site]) >= 0) ++t;
		free(keep);
		return xmph_fail("Invalid character '%c' in alignment, aborting.", a->chars[(size_t) t * a->seq_len + job.bad_site]);
	}
	for (unsigned i = 0; i < range; ++i) {
		if (!keep[i]) continue;
		if (c->n != i) memcpy(c->col + (size_t) c->n * a->num_taxa, c->col + (size_t) i * a->num_taxa, a->num_taxa);
		++c->n;
	}
	free(keep);
	c->wt = (unsigned *) malloc(((size_t) c->n + 1) * sizeof(unsigned));
	for (unsigned i = 0; i < c->n; ++i) c->wt[i] = 1;
	return 0;
}

static unsigned g_cmp_width;   /* qsort has no context argument in C99 */
typedef struct { unsigned wt; const unsigned char *col; } keyed;

static int by_bytes(const void *x, const void *y) {
	return memcmp(((const keyed *) x)->col, ((const keyed *) y)->col, g_cmp_width);
}

/* Heaviest first; equal weights in ascending byte order (a total order once columns are distinct). */
static int by_weight_then_bytes(const void *x, const void *y) {
	const keyed *a = (const keyed *) x, *b = (const keyed *) y;
	if (a->wt != b->wt) return a->wt > b->wt ? -1 : 1;
	return memcmp(a->col, b->col, g_cmp_width);
}

/* Large sets: runs sorted on all cores, then merged pairwise (also in parallel).  Both orders are total on distinct
 * columns (and equal columns are interchangeable: they are about to be merged), so the result does not depend on how
 * the work was cut. */
typedef struct { keyed *k, *tmp; size_t n; unsigned runs; int (*cmp)(const void *, const void *); size_t width; } sort_job;

static void sort_run(void *arg, unsigned r) {
	sort_job *j = (sort_job *) arg;
	const size_t lo = j->n * r / j->runs, hi = j->n * (r + 1) / j->runs;
	qsort(j->k + lo, hi - lo, sizeof(keyed), j->cmp);
}

static void merge_pair(void *arg, unsigned pair) {
	sort_job *j = (sort_job *) arg;
	const unsigned r = 2 * pair * (unsigned) j->width;
	const size_t lo = j->n * r / j->runs;
	const size_t mid = j->n * (r + j->width < j->runs ? r + j->width : j->runs) / j->runs;
	const size_t hi = j->n * (r + 2 * j->width < j->runs ? r + 2 * j->width : j->runs) / j->runs;
	size_t a = lo, b = mid, o = lo;
	while (a < mid && b < hi) j->tmp[o++] = j->cmp(&j->k[b], &j->k[a]) < 0 ? j->k[b++] : j->k[a++];
	while (a < mid) j->tmp[o++] = j->k[a++];
	while (b < hi) j->tmp[o++] = j->k[b++];
}

static void sort_keyed(keyed *k, size_t n, int (*cmp)(const void *, const void *), unsigned threads) {
	if (!threads) {
		long cores = sysconf(_SC_NPROCESSORS_ONLN);
		threads = cores > 0 ? (unsigned) cores : 1;
	}
	unsigned runs = 1;
	while (runs * 2 <= threads && runs < 32) runs *= 2;
	if (n < 200000 || runs == 1) {
		qsort(k, n, sizeof(keyed), cmp);
		return;
	}
	sort_job job = { k, (keyed *) malloc((n + 1) * sizeof(keyed)), n, runs, cmp, 1 };
	parallel_for(runs, sort_run, &job, threads, (size_t) 1 << 24);
	for (job.width = 1; job.width < runs; job.width *= 2) {
		parallel_for((runs + 2 * (unsigned) job.width - 1) / (2 * (unsigned) job.width), merge_pair, &job, threads, (size_t) 1 << 24);
		keyed *t = job.k;
		job.k = job.tmp;
		job.tmp = t;
	}
	if (job.k != k) {
		memcpy(k, job.k, n * sizeof(keyed));
		free(job.k);
	} else {
		free(job.tmp);
	}
}

static void sort_columns(columns *c, int (*cmp)(const void *, const void *), unsigned threads) {
	keyed *k = (keyed *) malloc(((size_t) c->n + 1) * sizeof(keyed));
	unsigned char *ncol = (unsigned char *) malloc((size_t) c->num_taxa * (c->n + 1));
	unsigned *nwt = (unsigned *) malloc(((size_t) c->n + 1) * sizeof(unsigned));
	for (unsigned i = 0; i < c->n; ++i) {
		k[i].wt = c->wt[i];
		k[i].col = c->col + (size_t) i * c->num_taxa;
	}
	g_cmp_width = c->num_taxa;
	sort_keyed(k, c->n, cmp, threads);
	for (unsigned i = 0; i < c->n; ++i) {
		nwt[i] = k[i].wt;
		memcpy(ncol + (size_t) i * c->num_taxa, k[i].col, c->num_taxa);
	}
	free(k);
	free(c->col);
	free(c->wt);
	c->col = ncol;
	c->wt = nwt;
}

/* Adjacent equal columns (after sorting by bytes) collapse into one whose weight is the sum. */
static void merge_equal_columns(columns *c) {
	unsigned j = 0;
	if (c->n == 0) return;
	for (unsigned i = 1; i < c->n; ++i) {
		unsigned char *ci = c->col + (size_t) i * c->num_taxa, *cj = c->col + (size_t) j * c->num_taxa;
		if (memcmp(cj, ci, c->num_taxa) == 0) {
			c->wt[j] += c->wt[i];
		} else {
			++j;
			memmove(c->col + (size_t) j * c->num_taxa, ci, c->num_taxa);
			c->wt[j] = c->wt[i];
		}
	}
	c->n = j + 1;
}

/* Length a column adds to every tree if it is parsimony-uninformative, else -1.  A column is
 * uninformative when, for some state r ("the rest"; the reference also tries the impossible fifth
 * state 16, kept here), the sets that do not contain r are pairwise disjoint: each of them then costs
 * exactly one step on any tree. */
static int uninformative_cost(const unsigned char *col, unsigned n) {
	for (unsigned r = 1; r <= 16; r <<= 1) {
		unsigned with_r = 0;
		int disjoint = 1;
		for (unsigned i = 0; i < n && disjoint; ++i) {
			if (col[i] & r) { ++with_r; continue; }
			for (unsigned j = i + 1; j < n; ++j) {
				if (!(col[j] & r) && (col[i] & col[j])) { disjoint = 0; break; }
			}
		}
		if (disjoint) return (int) (n - with_r);
	}
	return -1;
}

static void drop_uninformative(columns *c, xmph_problem *p) {
	unsigned j = 0;
	p->n_constant_sites = p->n_nonpi_sites = p->constant_weight = 0;
	for (unsigned i = 0; i < c->n; ++i) {
		const unsigned char *ci = c->col + (size_t) i * c->num_taxa;
		unsigned common = 15;
		for (unsigned t = 0; t < c->num_taxa && common; ++t) common &= ci[t];
		if (common) {
			p->n_constant_sites += c->wt[i];
			continue;
		}
		int cost = uninformative_cost(ci, c->num_taxa);
		if (cost >= 0) {
			p->constant_weight += (unsigned) cost * c->wt[i];
			p->n_nonpi_sites += c->wt[i];
			continue;
		}
		if (i != j) {
			memmove(c->col + (size_t) j * c->num_taxa, ci, c->num_taxa);
			c->wt[j] = c->wt[i];
		}
		++j;
	}
	c->n = j;
}

/* ---- unpacked (one set per byte) distances -------------------------------------------------- */
/* Both distances are written branch-free so that the compiler vectorises them: at 64 taxa x 2^20 sites
 * (BASELINE config 2 pushed through the CLI) the O(n^2) distances are the whole preprocessing time. */
static unsigned fitch_distance(const unsigned char *a, const unsigned char *b, const unsigned *w, unsigned n) {
	unsigned d = 0;
	for (unsigned i = 0; i < n; ++i) d += w[i] & (0u - (unsigned) ((a[i] & b[i]) == 0));
	return d;
}

static unsigned strict_distance(const unsigned char *a, const unsigned char *b, const unsigned *w, unsigned n) {
	unsigned d = 0;
	for (unsigned i = 0; i < n; ++i) d += w[i] & (0u - (unsigned) (a[i] != b[i]));
	return d;
}

/* All pairwise distances between `height` rows: D[i * height + j], symmetric, zero diagonal. */
typedef struct { const unsigned char *rows; const unsigned *w; unsigned *D; unsigned ns, height; int strict; } pair_job;

static void pair_item(void *x, unsigned item) {
	const pair_job *j = (const pair_job *) x;
	/* item -> (a, b), a < b, row-major over the strict upper triangle */
	unsigned a = 0, left = item;
	while (left >= j->height - 1 - a) { left -= j->height - 1 - a; ++a; }
	const unsigned b = a + 1 + left;
	const unsigned char *ra = j->rows + (size_t) a * j->ns, *rb = j->rows + (size_t) b * j->ns;
	unsigned d = j->strict ? strict_distance(ra, rb, j->w, j->ns) : fitch_distance(ra, rb, j->w, j->ns);
	j->D[(size_t) a * j->height + b] = j->D[(size_t) b * j->height + a] = d;
}

static unsigned *pair_matrix(const unsigned char *rows, unsigned ns, unsigned height, const unsigned *w, int strict, unsigned threads) {
	unsigned *D = (unsigned *) calloc((size_t) height * height + 1, sizeof(unsigned));
	pair_job j = { rows, w, D, ns, height, strict };
	if (D && height > 1) parallel_for(height * (height - 1) / 2, pair_item, &j, threads, ns);
	return D;
}

/* Weight of a minimum spanning tree of the complete graph with distance matrix D over `height` vertices (Prim). */
static unsigned mst_of_matrix(const unsigned *D, unsigned height) {
	unsigned *best = (unsigned *) malloc((height + 1) * sizeof(unsigned));
	unsigned char *in = (unsigned char *) calloc(height + 1, 1);
	unsigned total = 0;
	if (height < 2) { free(best); free(in); return 0; }
	in[0] = 1;
	for (unsigned j = 1; j < height; ++j) best[j] = D[j];
	for (unsigned k = 1; k < height; ++k) {
		unsigned pick = 0, d = 0xffffffffu;
		for (unsigned j = 1; j < height; ++j) {
			if (!in[j] && best[j] < d) { d = best[j]; pick = j; }
		}
		in[pick] = 1;
		total += d;
		for (unsigned j = 1; j < height; ++j) {
			if (!in[j] && D[(size_t) pick * height + j] < best[j]) best[j] = D[(size_t) pick * height + j];
		}
	}
	free(best);
	free(in);
	return total;
}

static unsigned mst_weight(const unsigned char *rows, unsigned n_sites, unsigned height, const unsigned *w, int strict, unsigned threads) {
	unsigned *D = pair_matrix(rows, n_sites, height, w, strict, threads);
	unsigned total = D ? mst_of_matrix(D, height) : 0;
	free(D);
	return total;
}

/* ---- taxon order ---------------------------------------------------------------------------- */
static int cmp_unsigned(const void *a, const void *b) {
	unsigned x = *(const unsigned *) a, y = *(const unsigned *) b;
	return x < y ? -1 : x > y;
}

/* Max-mini order: the two most distant taxa first (first such pair in (i, j) order); then repeatedly
 * the remaining taxon whose ascending-sorted vector of distances to the taxa already placed is
 * lexicographically greatest (first such taxon, in current row order, on ties).  The last taxon is
 * whatever is left.  Rows are exchanged exactly as the reference exchanges them (row order decides
 * ties), but on a permutation: the distances come from one matrix computed up front, and the rows
 * themselves move once, at the end. */
static void order_taxa(xmph_problem *p, unsigned threads) {
	const unsigned n = p->num_taxa, ns = p->num_sites;
	unsigned *D = pair_matrix(p->sites, ns, n, p->site_weights, 0, threads);
	unsigned *dist = (unsigned *) malloc(n * sizeof(unsigned));
	unsigned *best = (unsigned *) malloc(n * sizeof(unsigned));
	unsigned *at = p->taxon_map;                  /* at[k] = original row now at position k */
	unsigned fi = 0, fj = 1, fd = 0;

	for (unsigned i = 0; i < n; ++i) at[i] = i;
	for (unsigned i = 0; i + 1 < n; ++i) {
		for (unsigned j = i + 1; j < n; ++j) {
			unsigned d = D[(size_t) i * n + j];
			if (d > fd) { fi = i; fj = j; fd = d; }
		}
	}
	if (fi != 0) { at[0] = fi; at[fi] = 0; }
	if (fj != 1) { unsigned t = at[fj]; at[fj] = at[1]; at[1] = t; }
	for (unsigned k = 2; k + 1 < n; ++k) {
		unsigned pick = k;
		memset(best, 0, n * sizeof(unsigned));
		for (unsigned i = k; i < n; ++i) {
			for (unsigned j = 0; j < k; ++j) dist[j] = D[(size_t) at[i] * n + at[j]];
			qsort(dist, k, sizeof(unsigned), cmp_unsigned);
			for (unsigned j = 0; j < k; ++j) {
				if (dist[j] > best[j]) {
					pick = i;
					memcpy(best, dist, k * sizeof(unsigned));
					break;
				}
				if (dist[j] < best[j]) break;
			}
		}
		if (pick != k) { unsigned t = at[pick]; at[pick] = at[k]; at[k] = t; }
	}
	int moved = 0;
	for (unsigned i = 0; i < n; ++i) moved |= at[i] != i;
	if (moved) {
		unsigned char *sorted = (unsigned char *) malloc((size_t) n * ns + 1);
		for (unsigned i = 0; i < n; ++i) memcpy(sorted + (size_t) i * ns, p->sites + (size_t) at[i] * ns, ns);
		free(p->sites);
		p->sites = sorted;
	}
	free(best);
	free(dist);
	free(D);
}

/* ---- lower bounds --------------------------------------------------------------------------- */
/* Distinct-bases bound.  For every column made only of single states: once taxa 0..k are on the
 * tree, each state that occurs among the remaining taxa but not yet on the tree must cost one more
 * step.  rest_bound[k] = weighted count of such (column, state) pairs. */
/* Sites are independent: chunks of sites go to threads, each adds its share to a per-chunk row of totals. */
enum { DB_CHUNK = 16384 };
typedef struct { const xmph_problem *p; unsigned *totals; } db_job;      /* totals[chunk][taxon] */

static void distinct_bases_chunk(void *arg, unsigned chunk) {
	static const signed char slot[16] = { -1, 0, 1, -1, 2, -1, -1, -1, 3, -1, -1, -1, -1, -1, -1, -1 };
	db_job *j = (db_job *) arg;
	const xmph_problem *p = j->p;
	const unsigned n = p->num_taxa, ns = p->num_sites;
	const unsigned lo = chunk * DB_CHUNK, hi = lo + DB_CHUNK < ns ? lo + DB_CHUNK : ns, len = hi - lo;
	unsigned (*rest)[4] = (unsigned (*)[4]) calloc(len + 1, sizeof(unsigned[4]));
	unsigned char (*seen)[4] = (unsigned char (*)[4]) calloc(len + 1, 4);
	unsigned char *skip = (unsigned char *) calloc(len + 1, 1);
	unsigned *total = j->totals + (size_t) chunk * n;
	for (unsigned t = 0; t < n; ++t) {
		const unsigned char *row = p->sites + (size_t) t * ns + lo;
		for (unsigned i = 0; i < len; ++i) {
			int k = slot[row[i] & 15];
			if (k < 0) skip[i] = 1; else ++rest[i][k];
		}
	}
	for (unsigned t = 0; t < n; ++t) {
		const unsigned char *row = p->sites + (size_t) t * ns + lo;
		unsigned sum = 0;
		for (unsigned i = 0; i < len; ++i) {
			if (skip[i]) continue;
			int k = slot[row[i] & 15];
			seen[i][k] = 1;
			--rest[i][k];
			for (int b = 0; b < 4; ++b) {
				if (!seen[i][b] && rest[i][b] > 0) sum += p->site_weights[lo + i];
			}
		}
		total[t] = sum;
	}
	free(skip);
	free(seen);
	free(rest);
}

static void distinct_bases_bound(xmph_problem *p, unsigned threads) {
	const unsigned n = p->num_taxa, ns = p->num_sites, chunks = (ns + DB_CHUNK - 1) / DB_CHUNK;
	if (!chunks) return;
	db_job job = { p, (unsigned *) calloc((size_t) chunks * n + 1, sizeof(unsigned)) };
	parallel_for(chunks, distinct_bases_chunk, &job, threads, (size_t) DB_CHUNK * n * 8);
	for (unsigned t = 0; t < n; ++t) {
		unsigned total = 0;
		for (unsigned c = 0; c < chunks; ++c) total += job.totals[(size_t) c * n + t];
		if (total > p->rest_bound[t]) p->rest_bound[t] = total;
	}
	free(job.totals);
}

/* MST bound: half the MST weight (Fitch distances, so ambiguity never overestimates) over
 * { union of taxa 0..k, taxon k+1, ..., taxon n-1 }, rounded down. */
static void mst_bound(xmph_problem *p, unsigned threads) {
	const unsigned n = p->num_taxa, ns = p->num_sites;
	unsigned *D = pair_matrix(p->sites, ns, n, p->site_weights, 0, threads);   /* taxon-to-taxon distances never change */
	unsigned char *joined = (unsigned char *) malloc((size_t) ns + 1);          /* union of taxa 0..k */
	unsigned *sub = (unsigned *) malloc((size_t) n * n * sizeof(unsigned) + 4);
	memcpy(joined, p->sites, ns);
	for (unsigned k = 0; k + 1 < n; ++k) {
		/* vertices: the union, then taxa k+1 .. n-1 */
		const unsigned h = n - k;
		sub[0] = 0;
		for (unsigned j = 1; j < h; ++j) {
			sub[j] = sub[(size_t) j * h] = k == 0 ? D[k + j] : fitch_distance(joined, p->sites + (size_t) (k + j) * ns, p->site_weights, ns);
			for (unsigned i = 1; i < h; ++i) sub[(size_t) j * h + i] = D[(size_t) (k + j) * n + k + i];
		}
		unsigned b = mst_of_matrix(sub, h) / 2;
		if (b > p->rest_bound[k]) p->rest_bound[k] = b;
		const unsigned char *next = p->sites + (size_t) (k + 1) * ns;
		for (unsigned i = 0; i < ns; ++i) joined[i] |= next[i];
	}
	free(sub);
	free(joined);
	free(D);
}

/* Same text format as restBound.txt: a header line, then "<index> <bound>" lines. */
static int load_rest_bound(xmph_problem *p, const char *path) {
	char line[256];
	unsigned idx, val, line_no = 2;
	FILE *f = fopen(path, "rb");
	if (!f) return xmph_fail("Unable to open file '%s' to retrieve restBound[] array, aborting.", path);
	if (fgets(line, sizeof line, f)) {
		while (fgets(line, sizeof line, f)) {
			if (sscanf(line, "%u %u", &idx, &val) < 2) {
				fclose(f);
				return xmph_fail("Unable to parse line %u of '%s', aborting.", line_no, path);
			}
			if (idx >= p->num_taxa) {
				fclose(f);
				return xmph_fail("Line %u of '%s' names taxon index %u but there are only %u taxa, aborting.", line_no, path, idx, p->num_taxa);
			}
			p->rest_bound[idx] = val;
			++line_no;
		}
	}
	fclose(f);
	return 0;
}

int xmph_write_rest_bound(const xmph_problem *p, const char *path) {
	FILE *f = fopen(path, "wt");
	if (!f) return xmph_fail("Unable to open '%s' for output.", path);
	fprintf(f, "i\tMin weight added by taxa i+1, i+2, ..., n-1.\n");
	for (unsigned i = 2; i < p->num_taxa; ++i) fprintf(f, "%u\t%u\n", i, p->rest_bound[i]);
	fclose(f);
	return 0;
}

/* ---- packing -------------------------------------------------------------------------------- */
static void pack_rows(xmph_problem *p) {
	const unsigned ns = p->num_sites;
	/* No informative site at all: one block of padding, so that the search runs as usual and -- like the reference's,
	 * whose Fitch calls then see zero blocks -- finds every tree optimal at the constant weight. */
	p->n_blocks = ns ? (ns - 1) / XMPH_SITES_PER_BLOCK + 1 : 1;
	const size_t row_bytes = (size_t) p->n_blocks * XMPH_BLOCK_BYTES, padded = row_bytes * 2;
	if (posix_memalign((void **) &p->rows, 64, row_bytes * p->num_taxa + 64)) p->rows = NULL;
	p->weights = (unsigned *) malloc((padded + 1) * sizeof(unsigned));
	for (unsigned t = 0; t < p->num_taxa; ++t) {
		unsigned char *row = p->rows + t * row_bytes;
		for (size_t s = 0; s < padded; ++s) {
			unsigned set = s < ns ? p->sites[(size_t) t * ns + s] : 15u;   /* padding never scores */
			if (s & 1) row[s >> 1] |= (unsigned char) (set << 4); else row[s >> 1] = (unsigned char) set;
		}
	}
	for (size_t s = 0; s < padded; ++s) p->weights[s] = s < ns ? p->site_weights[s] : 1;
}

void xmph_free_problem(xmph_problem *p) {
	free(p->rows);
	free(p->weights);
	free(p->sites);
	free(p->site_weights);
	free(p->taxon_map);
	free(p->rest_bound);
	memset(p, 0, sizeof *p);
}

/* columns [site][taxon] -> rows [taxon][site], a tile of sites at a time */
enum { TRANSPOSE_TILE = 8192 };
typedef struct { const unsigned char *col; unsigned char *sites; unsigned n, num_taxa; } transpose_job;

static void transpose_tile(void *arg, unsigned tile) {
	transpose_job *j = (transpose_job *) arg;
	const unsigned lo = tile * TRANSPOSE_TILE, hi = lo + TRANSPOSE_TILE < j->n ? lo + TRANSPOSE_TILE : j->n;
	for (unsigned i = lo; i < hi; ++i) {
		const unsigned char *src = j->col + (size_t) i * j->num_taxa;
		for (unsigned t = 0; t < j->num_taxa; ++t) j->sites[(size_t) t * j->n + i] = src[t];
	}
}

int xmph_prepare(const xmph_alignment *a, const xmph_settings *s, xmph_problem *p) {
	columns c = { 0, 0, NULL, NULL };
	const unsigned unsupported = XMPH_OPT_USEINCOMPATBOUNDREST | XMPH_OPT_USEKICKASSBOUNDREST | XMPH_OPT_USEKA2BOUNDREST |
	                             XMPH_OPT_USEGREEDYHITTINGSETBOUND;
	memset(p, 0, sizeof *p);
	if (a->num_taxa < 4) return xmph_fail("You must have four or more taxa to perform this analysis.");
	if (a->num_taxa > XMPH_MAXTAXA) return xmph_fail("At most %d taxa are supported.", XMPH_MAXTAXA);
	if (s->options & unsupported) {
		return xmph_fail("Only the -Bd, -Bm, -Bp and -Bf lower bounds are built into this program; compute other bounds with the reference and load its restBound.txt with -Bf <file>.");
	}
	if (recode(a, s, &c)) { free(c.col); return -1; }
	sort_columns(&c, by_bytes, s->threads);
	merge_equal_columns(&c);
	sort_columns(&c, by_weight_then_bytes, s->threads);
	drop_uninformative(&c, p);

	p->num_taxa = a->num_taxa;
	p->num_sites = c.n;
	p->sites = (unsigned char *) malloc((size_t) c.n * a->num_taxa + 1);
	{
		transpose_job tj = { c.col, p->sites, c.n, a->num_taxa };
		parallel_for((c.n + TRANSPOSE_TILE - 1) / TRANSPOSE_TILE, transpose_tile, &tj, s->threads, (size_t) TRANSPOSE_TILE * a->num_taxa * 2);
	}
	p->site_weights = c.wt;
	free(c.col);
	p->taxon_map = (unsigned *) malloc(a->num_taxa * sizeof(unsigned));
	p->rest_bound = (unsigned *) calloc(a->num_taxa, sizeof(unsigned));
	order_taxa(p, s->threads);

	p->bound = s->bound;
	if (s->options & XMPH_OPT_IMPROVEINITUPPERBOUND) {
		p->mst_upper_bound = mst_weight(p->sites, p->num_sites, p->num_taxa, p->site_weights, 1, s->threads) + p->constant_weight;
		if (p->mst_upper_bound < p->bound) p->bound = p->mst_upper_bound;
	}
	if (s->options & XMPH_OPT_USEDISTBASESBOUNDREST) distinct_bases_bound(p, s->threads);
	if ((s->options & XMPH_OPT_USELOADRESTBOUND) && load_rest_bound(p, s->rest_bound_file)) {
		/* The reference has reported the site counts and the MST bound by the time it reads this file
		 * (src/common.c:1205-1209): leave those numbers (and nothing else) for the caller's report. */
		const xmph_problem keep = *p;
		xmph_free_problem(p);
		p->num_taxa = keep.num_taxa;
		p->num_sites = keep.num_sites;
		p->n_constant_sites = keep.n_constant_sites;
		p->n_nonpi_sites = keep.n_nonpi_sites;
		p->constant_weight = keep.constant_weight;
		p->mst_upper_bound = keep.mst_upper_bound;
		return -1;
	}
	if (s->options & XMPH_OPT_USEMSTBOUNDREST) mst_bound(p, s->threads);
	if ((s->options & XMPH_OPT_USEPARTBOUND) && xmph_part_bound(p, (int) s->part_strategy, s->part_max_stale, s->threads)) {
		xmph_free_problem(p);
		return -1;
	}
	pack_rows(p);
	return 0;
}
