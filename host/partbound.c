/* Partition lower bound ("-Bp"), the bound the reference's author benchmarks with.
 *
 * What it computes (reference src/partbound.c:1041-1164): for every tree size T = 4 .. n-1 a lower
 * bound on the length that taxa T .. n-1 must still add to any tree on taxa 0 .. T-1.  The site
 * patterns are split into parts; each part contributes max(additive, LB(part on all taxa) -
 * UB(part on the first T taxa)), where LB is the best of several spanning-tree bounds on the part's
 * distinct row sequences, UB is the per-site worst case and "additive" is the distinct-bases count.
 * Starting from one part per site, a greedy local search moves one site at a time between parts
 * while the total improves.  Weighted sites are handled in "chunks" of equal residual weight.
 *
 * restBound[] must equal the reference's exactly, so the local search visits part pairs, directions
 * and sites in the reference's order and breaks ties the same way (first strictly better move wins);
 * that order is restated here over a different representation: a part is an ordered list of site
 * indices into the column-major alignment (no per-part copies of the data), a part's bound is
 * evaluated on rows packed 16 sites per 64-bit word with the same SWAR empty-nibble count the
 * search uses, and the (tree size, chunk) sub-problems - which are independent - run on a thread
 * pool.  Known quirks of the reference that change values are kept and marked QUIRK.
 */
#include <limits.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include "xmp_host.h"

int xmph_fail(const char *fmt, ...);

enum { STRAT_FINAL = 0, STRAT_CLIPPED = 1, STRAT_LB = 2 };   /* src/partbound.h:5-9 */

typedef struct {
	int *cols;          /* site indices, in the reference's column order */
	int n, cap;
	int lb, ub;         /* LB on all taxa, UB on the first T taxa */
	int add;            /* additive score as the reference has it stored (QUIRK: 0 until the part first changes) */
	int add_true;       /* sum of the per-site additive scores */
	int touched;
} part;

typedef struct {
	unsigned n_taxa, n_sites;
	const unsigned char *col;     /* [site][taxon] */
	const unsigned *weights;
	const int *single;            /* bound of each one-site part (NULL while it is being computed) */
	int strategy, max_stale;
} shared;

typedef struct { int first_site, multiplier; } chunk;   /* sites first_site .. 0, each counted multiplier times */

typedef struct {
	const shared *sh;
	/* bound evaluation */
	uint64_t *rows;  size_t rows_cap;
	int *uniq;
	int *ew; unsigned char *ea, *eb;          /* edges among distinct rows */
	int *sw; unsigned char *sa, *sb;          /* the same, ordered by weight */
	int *bucket; size_t bucket_cap;
	int *parent, *comp_w, *compact, *compact_w;
	int *between;                             /* [c][c] least weight between two components */
	/* per tree size */
	int *site_ub, *site_add;
	/* partition */
	part *parts;
	int **blocks; size_t n_blocks, cap_blocks;  /* column lists are carved from these and freed per chunk */
	int *block_cur; size_t block_left;
} work;

/* ---- memory for column lists ------------------------------------------------------------------ */
static int *carve(work *w, size_t n) {
	if (n > w->block_left) {
		size_t sz = n > 4096 ? n : 4096;
		if (w->n_blocks == w->cap_blocks) {
			w->cap_blocks = w->cap_blocks ? w->cap_blocks * 2 : 16;
			w->blocks = (int **) realloc(w->blocks, w->cap_blocks * sizeof(int *));
		}
		w->block_cur = w->blocks[w->n_blocks++] = (int *) malloc(sz * sizeof(int));
		w->block_left = sz;
	}
	int *r = w->block_cur;
	w->block_cur += n;
	w->block_left -= n;
	return r;
}

static void release_blocks(work *w) {
	for (size_t i = 0; i < w->n_blocks; ++i) free(w->blocks[i]);
	w->n_blocks = 0;
	w->block_cur = NULL;
	w->block_left = 0;
}

static void reserve_one_more(work *w, part *p) {
	if (p->n + 1 <= p->cap) return;
	int cap = p->cap * 2 + 2;
	int *c = carve(w, (size_t) cap);
	memcpy(c, p->cols, (size_t) p->n * sizeof(int));
	p->cols = c;
	p->cap = cap;
}

/* ---- lower bound of one part -------------------------------------------------------------------- */
static int find_root(int *parent, int v) {
	int r = v;
	while (parent[r] != r) r = parent[r];
	while (parent[v] != r) { int next = parent[v]; parent[v] = r; v = next; }
	return r;
}

/* Sites whose two state sets do not intersect, over rows of `words` 64-bit words (padding nibbles are 0xF). */
static inline int row_distance(const uint64_t *a, const uint64_t *b, int words) {
	const uint64_t low3 = 0x7777777777777777ull, top = 0x8888888888888888ull;
	int d = 0;
	for (int k = 0; k < words; ++k) {
		uint64_t x = a[k] & b[k];
		uint64_t nonempty = (((x & low3) + low3) | x) & top;
		d += __builtin_popcountll(nonempty ^ top);
	}
	return d;
}

/* The best spanning-tree bound on `n_rows` packed rows of `n_sites` sites
 * (src/partbound.c:155-432; its special cases for 1 and 2 sites return the same values). */
static int bound_of_rows(work *w, const uint64_t *rows, int n_rows, int words, int n_sites) {
	int nu = 0, ne = 0, max_w = 0;

	/* distinct rows and all pairwise distances between them */
	for (int t = 0; t < n_rows; ++t) {
		const uint64_t *r = rows + (size_t) t * words;
		int dup = 0;
		for (int j = 0; j < nu; ++j) {
			const uint64_t *u = rows + (size_t) w->uniq[j] * words;
			if (!memcmp(r, u, (size_t) words * 8)) { dup = 1; break; }
			int d = row_distance(r, u, words);
			w->ew[ne + j] = d; w->ea[ne + j] = (unsigned char) j; w->eb[ne + j] = (unsigned char) nu;
		}
		if (dup) continue;
		for (int j = 0; j < nu; ++j) if (w->ew[ne + j] > max_w) max_w = w->ew[ne + j];
		ne += nu;
		w->uniq[nu++] = t;
	}
	if (nu < 2) return 0;

	/* order the edges by weight (counting sort) */
	if ((size_t) max_w + 2 > w->bucket_cap) {
		w->bucket_cap = (size_t) max_w * 2 + 64;
		free(w->bucket);
		w->bucket = (int *) malloc(w->bucket_cap * sizeof(int));
	}
	memset(w->bucket, 0, ((size_t) max_w + 2) * sizeof(int));
	for (int e = 0; e < ne; ++e) ++w->bucket[w->ew[e] + 1];
	for (int k = 1; k <= max_w; ++k) w->bucket[k] += w->bucket[k - 1];
	for (int e = 0; e < ne; ++e) {
		int pos = w->bucket[w->ew[e]]++;
		w->sw[pos] = w->ew[e]; w->sa[pos] = w->ea[e]; w->sb[pos] = w->eb[e];
	}

	for (int v = 0; v < nu; ++v) { w->parent[v] = v; w->comp_w[v] = 0; }
	int n_comp = nu, forest = 0, first_long = 0;
	int b_connected = 0, b_three = 0, b_mst = 0;

	for (int e = 0; n_comp > 1; ++e) {
		int ra = find_root(w->parent, w->sa[e]), rb = find_root(w->parent, w->sb[e]);
		if (ra == rb) continue;
		const int wt = w->sw[e];
		if (wt > 1 && !first_long) {
			/* the components joined by edges of weight <= 1 are complete: the forest so far plus this
			 * edge is a bound, and so is the cheapest way to join any three components */
			first_long = wt;
			b_connected = forest + wt;
			if (n_comp >= 3) {
				int c = 0;
				for (int v = 0; v < nu; ++v) if (w->parent[v] == v) w->compact[v] = c++;
				memset(w->between, 0, (size_t) c * c * sizeof(int));
				int wanted = c * (c - 1) / 2;
				for (int f = e; wanted; ++f) {
					int x = w->compact[find_root(w->parent, w->sa[f])], y = w->compact[find_root(w->parent, w->sb[f])];
					if (x == y) continue;
					if (x > y) { int s = x; x = y; y = s; }
					if (!w->between[x * c + y]) { w->between[x * c + y] = w->sw[f]; --wanted; }
				}
				int *cw = w->compact_w;
				for (int v = 0; v < nu; ++v) if (w->parent[v] == v) cw[w->compact[v]] = w->comp_w[v];
				for (int x = 0; x < c - 2; ++x) {
					for (int y = x + 1; y < c - 1; ++y) {
						const int dxy = w->between[x * c + y], wxy = cw[x] + cw[y];
						for (int z = y + 1; z < c; ++z) {
							const int dxz = w->between[x * c + z], dyz = w->between[y * c + z];
							const int sum = dxy + dxz + dyz;
							int longest = dxy > dxz ? dxy : dxz;
							if (dyz > longest) longest = dyz;
							int direct = sum - longest, steiner = (sum + 1) / 2;
							int cand = (direct < steiner ? direct : steiner) + wxy + cw[z];
							if (cand > b_three) b_three = cand;
						}
					}
				}
			}
		}
		w->parent[ra] = rb;
		w->comp_w[rb] += w->comp_w[ra] + wt;
		--n_comp;
		forest += wt;
	}
	if (!first_long) b_connected = forest;
	else b_mst = n_sites > 2 ? forest / 2 : forest;

	int best = b_connected;
	if (b_three > best) best = b_three;
	if (b_mst > best) best = b_mst;
	return best;
}

/* Answers for two-site parts without ambiguity, keyed by the set of (state, state) pairs present
 * (role of the reference's generated 64 KB table src/sitepaircosts.c; here filled on first use by the
 * general evaluation, which yields the same numbers: tests/test_host.py checks all 65,535). */
static unsigned char pair_memo[65536];
static unsigned char pair_known[65536];

static int pair_bound(work *w, unsigned index) {
	if (__atomic_load_n(&pair_known[index], __ATOMIC_ACQUIRE)) return pair_memo[index];
	uint64_t rows[16];
	int n = 0;
	for (unsigned i = 0; i < 16; ++i) {
		if (index >> i & 1) rows[n++] = ~0xffull | (1u << (i >> 2)) | (16u << (i & 3));
	}
	int v = bound_of_rows(w, rows, n, 1, 2);
	pair_memo[index] = (unsigned char) v;
	__atomic_store_n(&pair_known[index], 1, __ATOMIC_RELEASE);
	return v;
}

int xmph_site_pair_cost(unsigned index);   /* exported for the parity test against the reference's table */

static const signed char state_slot[16] = { -1, 0, 1, -1, 2, -1, -1, -1, 3, -1, -1, -1, -1, -1, -1, -1 };

static int part_bound(work *w, const int *cols, int ns) {
	const shared *sh = w->sh;
	const int n = (int) sh->n_taxa;
	if (!ns) return 0;
	if (ns == 1 && sh->single) return sh->single[cols[0]];
	if (ns == 2) {
		const unsigned char *c0 = sh->col + (size_t) cols[0] * n, *c1 = sh->col + (size_t) cols[1] * n;
		unsigned index = 0;
		int t = 0;
		for (; t < n; ++t) {
			int a = state_slot[c0[t] & 15], b = state_slot[c1[t] & 15];
			if ((a | b) < 0) break;
			index |= 1u << (a * 4 + b);
		}
		if (t == n) return pair_bound(w, index);
	}
	const int words = (ns + 15) >> 4;
	if ((size_t) words * n > w->rows_cap) {
		w->rows_cap = (size_t) words * n * 2;
		free(w->rows);
		w->rows = (uint64_t *) malloc(w->rows_cap * 8);
	}
	uint64_t *rows = w->rows;
	const uint64_t last = (ns & 15) ? ~0ull << (4 * (ns & 15)) : 0;
	for (int t = 0; t < n; ++t) {
		for (int k = 0; k < words - 1; ++k) rows[(size_t) t * words + k] = 0;
		rows[(size_t) t * words + words - 1] = last;
	}
	for (int s = 0; s < ns; ++s) {
		const unsigned char *c = sh->col + (size_t) cols[s] * n;
		uint64_t *dst = rows + (s >> 4);
		const int shift = (s & 15) * 4;
		for (int t = 0; t < n; ++t) dst[(size_t) t * words] |= (uint64_t) (c[t] & 15) << shift;
	}
	return bound_of_rows(w, rows, n, words, ns);
}

/* ---- scores ------------------------------------------------------------------------------------- */
static inline int final_score(int strategy, int lb, int ub, int additive) {   /* src/partbound.c:106-115 */
	if (strategy == STRAT_FINAL) return additive > lb - ub ? additive : lb - ub;
	if (strategy == STRAT_CLIPPED) return lb - ub > 0 ? lb - ub : 0;
	return lb;
}

/* Per-site tables for tree size T: the worst-case cost of the site on the first T taxa
 * (src/common.c:2735-2757) and the number of states first seen at a taxon >= T (src/partbound.c:472-490). */
static void site_tables(work *w, int T) {
	const shared *sh = w->sh;
	const int n = (int) sh->n_taxa;
	for (unsigned s = 0; s < sh->n_sites; ++s) {
		const unsigned char *c = sh->col + (size_t) s * n;
		int most = 0;
		for (int b = 1; b <= 8; b <<= 1) {
			int cnt = 0;
			for (int t = 0; t < T; ++t) cnt += (c[t] & b) != 0;
			if (cnt > most) most = cnt;
		}
		w->site_ub[s] = T - most;
		int seen = 0, fresh = 0;
		for (int t = 0; t < n; ++t) {
			if (t >= T && (c[t] & ~seen)) ++fresh;
			seen |= c[t];
		}
		w->site_add[s] = fresh;
	}
}

/* ---- the local search --------------------------------------------------------------------------- */
typedef struct { int from, to, k, lb[2], ub[2], add[2]; } move;

static void apply_move(work *w, part *parts, const move *m) {   /* src/partbound.c:562-588 */
	part *src = &parts[m->from], *dst = &parts[m->to];
	reserve_one_more(w, dst);
	dst->cols[dst->n++] = src->cols[m->k];
	if (m->k != src->n - 1) { int s = src->cols[m->k]; src->cols[m->k] = src->cols[src->n - 1]; src->cols[src->n - 1] = s; }
	--src->n;
	src->lb = m->lb[0]; src->ub = m->ub[0]; src->add = src->add_true = m->add[0];
	dst->lb = m->lb[1]; dst->ub = m->ub[1]; dst->add = dst->add_true = m->add[1];
	src->touched = dst->touched = 1;
}

/* One pass over all pairs (part >= first, any earlier part).  Returns the position of the first part
 * changed in this pass after the changed parts have been gathered at the end, or *n_parts if nothing
 * moved (src/partbound.c:600-914). */
static int one_pass(work *w, part *parts, int *n_parts, int first, int *total) {
	const int strategy = w->sh->strategy;
	int n_moves = 0, cur = first;
	int best_gain = 0, best_raw = INT_MIN, best_raw_gain = 0;
	move gain_move, raw_move;
	memset(&gain_move, 0, sizeof gain_move);
	memset(&raw_move, 0, sizeof raw_move);

	for (int i = first; i < *n_parts; ++i) {
		if (!parts[i].n) continue;
		if (cur != i) parts[cur] = parts[i];
		parts[cur].touched = 0;
		for (int other = 0; other < cur; ++other) {
			const int idx[2] = { cur, other };
			const int n_dir = (parts[cur].n == 1 && parts[other].n == 1) ? 1 : 2;
			for (int d = 0; d < n_dir; ++d) {
				part *src = &parts[idx[d]], *dst = &parts[idx[1 - d]];
				const int raw_s = src->lb - src->ub, raw_d = dst->lb - dst->ub;
				const int old_raw = raw_s > raw_d ? raw_s : raw_d;
				const int old_final = final_score(strategy, src->lb, src->ub, src->add) + final_score(strategy, dst->lb, dst->ub, dst->add);
				reserve_one_more(w, dst);
				const int left = src->n - 1;
				for (int k = 0; k <= left; ++k) {
					const int site = src->cols[k];
					src->cols[k] = src->cols[left];          /* src without site k, in the order the reference leaves it */
					src->cols[left] = site;
					dst->cols[dst->n] = site;
					const int lb0 = part_bound(w, src->cols, left);
					const int lb1 = part_bound(w, dst->cols, dst->n + 1);
					src->cols[left] = src->cols[k];
					src->cols[k] = site;
					const int ub0 = src->ub - w->site_ub[site], ub1 = dst->ub + w->site_ub[site];
					const int add0 = src->add_true - w->site_add[site], add1 = dst->add_true + w->site_add[site];
					int raw = lb1 - ub1;
					if (left && lb0 - ub0 > raw) raw = lb0 - ub0;    /* an emptied part has no raw score */
					const int fin = final_score(strategy, lb0, ub0, add0) + final_score(strategy, lb1, ub1, add1);
					move *rec = NULL;
					if (fin - old_final > best_gain) {
						best_gain = fin - old_final;
						rec = &gain_move;
					} else if (fin >= old_final && raw <= 0 && raw > old_raw && raw > best_raw) {
						best_raw = raw;
						best_raw_gain = raw - old_raw;
						rec = &raw_move;
					}
					if (rec) {
						rec->from = idx[d]; rec->to = idx[1 - d]; rec->k = k;
						rec->lb[0] = lb0; rec->lb[1] = lb1; rec->ub[0] = ub0; rec->ub[1] = ub1;
						rec->add[0] = add0; rec->add[1] = add1;
					}
				}
			}
			/* take the best move of this pair at once */
			int moved = 0;
			if (best_gain > 0) {
				apply_move(w, parts, &gain_move);
				*total += best_gain;
				moved = 1;
			} else if (best_raw_gain > 0) {
				apply_move(w, parts, &raw_move);
				moved = 1;
			}
			if (moved) {
				best_gain = 0; best_raw = INT_MIN; best_raw_gain = 0;
				++n_moves;
				if (parts[cur].n && !parts[other].n) { part s = parts[other]; parts[other] = parts[cur]; parts[cur] = s; }
				if (!parts[cur].n) { --cur; break; }     /* the slot is refilled by the next part */
			}
		}
		++cur;
	}
	*n_parts = cur;
	if (!n_moves) return cur;

	/* changed parts to the far end, keeping the reference's two-pointer order */
	int i = 0, j = cur - 1;
	for (;;) {
		while (!parts[i].touched && i != j) ++i;
		if (i == j) break;
		while (parts[j].touched && i != j) --j;
		if (i == j) break;
		part s = parts[i]; parts[i] = parts[j]; parts[j] = s;
	}
	if (!parts[i].touched) ++i;
	return i;
}

/* One chunk for one tree size: the trivial partition of sites first_site .. 0, improved until no pass
 * moves anything or max_stale passes in a row leave the total unchanged (src/partbound.c:920-943,1071-1137). */
static int chunk_bound(work *w, int T, chunk ck) {
	const shared *sh = w->sh;
	part *parts = w->parts;
	int n_parts = 0, total = 0;
	(void) T;
	for (int s = ck.first_site; s >= 0; --s) {
		part *p = &parts[n_parts++];
		p->cols = carve(w, 2);
		p->cols[0] = s;
		p->n = 1; p->cap = 2;
		p->lb = part_bound(w, p->cols, 1);
		p->ub = w->site_ub[s];
		p->add = 0;                       /* QUIRK: src/partbound.c:1090 passes (nTaxaOnTree, numTaxa) swapped, which counts nothing */
		p->add_true = w->site_add[s];
		p->touched = 1;
		total += final_score(sh->strategy, p->lb, p->ub, p->add);
	}
	int first = 0, stale = 0;
	while (first < n_parts) {
		const int before = total;
		first = one_pass(w, parts, &n_parts, first, &total);
		if (total == before) {
			if (sh->max_stale && ++stale == sh->max_stale) break;
		} else {
			stale = 0;
		}
	}
	release_blocks(w);
	return total * ck.multiplier;
}

/* ---- driver ------------------------------------------------------------------------------------- */
typedef struct {
	const shared *sh;
	const chunk *chunks; int n_chunks;
	int n_jobs; int next_job;     /* job = (T - 4) * n_chunks + chunk */
	int *job_value;
	int failed;
} pool;

static int work_init(work *w, const shared *sh) {
	const size_t n = sh->n_taxa, pairs = n * (n - 1) / 2 + 1;
	memset(w, 0, sizeof *w);
	w->sh = sh;
	w->uniq = (int *) malloc(n * sizeof(int));
	w->ew = (int *) malloc(pairs * sizeof(int)); w->ea = (unsigned char *) malloc(pairs); w->eb = (unsigned char *) malloc(pairs);
	w->sw = (int *) malloc(pairs * sizeof(int)); w->sa = (unsigned char *) malloc(pairs); w->sb = (unsigned char *) malloc(pairs);
	w->parent = (int *) malloc(n * sizeof(int));
	w->comp_w = (int *) malloc(n * sizeof(int));
	w->compact = (int *) malloc(n * sizeof(int));
	w->compact_w = (int *) malloc(n * sizeof(int));
	w->between = (int *) malloc(n * n * sizeof(int));
	w->site_ub = (int *) malloc((sh->n_sites + 1) * sizeof(int));
	w->site_add = (int *) malloc((sh->n_sites + 1) * sizeof(int));
	w->parts = (part *) malloc((sh->n_sites + 1) * sizeof(part));
	return w->uniq && w->ew && w->ea && w->eb && w->sw && w->sa && w->sb && w->parent && w->comp_w && w->compact && w->compact_w &&
	       w->between && w->site_ub && w->site_add && w->parts ? 0 : -1;
}

static void work_free(work *w) {
	release_blocks(w);
	free(w->blocks); free(w->rows); free(w->uniq); free(w->ew); free(w->ea); free(w->eb); free(w->sw); free(w->sa); free(w->sb);
	free(w->bucket); free(w->parent); free(w->comp_w); free(w->compact); free(w->compact_w); free(w->between); free(w->site_ub); free(w->site_add);
	free(w->parts);
}

static void *pool_thread(void *arg) {
	pool *pl = (pool *) arg;
	work w;
	if (work_init(&w, pl->sh)) { __atomic_store_n(&pl->failed, 1, __ATOMIC_RELAXED); work_free(&w); return NULL; }
	int tables_for = -1;
	for (;;) {
		int job = __atomic_fetch_add(&pl->next_job, 1, __ATOMIC_RELAXED);
		if (job >= pl->n_jobs) break;
		int T = 4 + job / pl->n_chunks, c = job % pl->n_chunks;
		if (T != tables_for) { site_tables(&w, T); tables_for = T; }
		pl->job_value[job] = chunk_bound(&w, T, pl->chunks[c]);
	}
	work_free(&w);
	return NULL;
}

int xmph_part_bound(xmph_problem *p, int strategy, unsigned max_stale_passes, unsigned n_threads) {
	const unsigned n = p->num_taxa, S = p->num_sites;
	if (n < 5 || !S) return 0;
	shared sh;
	sh.n_taxa = n; sh.n_sites = S; sh.weights = p->site_weights; sh.strategy = strategy; sh.max_stale = (int) max_stale_passes;
	unsigned char *col = (unsigned char *) malloc((size_t) n * S);
	if (!col) return xmph_fail("Out of memory in the partition bound.");
	for (unsigned t = 0; t < n; ++t) for (unsigned s = 0; s < S; ++s) col[(size_t) s * n + t] = p->sites[(size_t) t * S + s];
	sh.col = col;
	sh.single = NULL;
	int *single = (int *) malloc((size_t) S * sizeof(int));
	{
		work w;
		if (!single || work_init(&w, &sh)) { work_free(&w); free(single); free(col); return xmph_fail("Out of memory in the partition bound."); }
		for (unsigned s = 0; s < S; ++s) { int c = (int) s; single[s] = part_bound(&w, &c, 1); }
		work_free(&w);
	}
	sh.single = single;

	/* chunks: peel the smallest remaining weight off every site that still has some (src/partbound.c:1071-1105) */
	chunk *chunks = (chunk *) malloc((size_t) S * sizeof(chunk));
	int n_chunks = 0, peeled = 0, rightmost = (int) S - 1;
	while ((int) p->site_weights[0] > peeled) {
		int mult = (int) p->site_weights[rightmost] - peeled;
		peeled += mult;
		chunks[n_chunks].first_site = rightmost;
		chunks[n_chunks].multiplier = mult;
		++n_chunks;
		for (int i = rightmost; i >= 0; --i) {
			if (i < (int) S - 1 && (int) p->site_weights[i] > peeled && (int) p->site_weights[i + 1] == peeled) rightmost = i;
		}
	}

	pool pl;
	pl.sh = &sh; pl.chunks = chunks; pl.n_chunks = n_chunks;
	pl.n_jobs = (int) (n - 4) * n_chunks; pl.next_job = 0; pl.failed = 0;
	pl.job_value = (int *) calloc((size_t) pl.n_jobs + 1, sizeof(int));
	if (!n_threads) {
		long cores = sysconf(_SC_NPROCESSORS_ONLN);
		n_threads = cores > 0 ? (unsigned) cores : 1;
	}
	if ((int) n_threads > pl.n_jobs) n_threads = (unsigned) pl.n_jobs;
	if (n_threads <= 1) {
		pool_thread(&pl);
	} else {
		pthread_t *th = (pthread_t *) malloc(n_threads * sizeof(pthread_t));
		unsigned started = 0;
		for (; started < n_threads; ++started) if (pthread_create(&th[started], NULL, pool_thread, &pl)) break;
		if (!started) pool_thread(&pl);
		for (unsigned i = 0; i < started; ++i) pthread_join(th[i], NULL);
		free(th);
	}
	int rc = 0;
	if (pl.failed) {
		rc = xmph_fail("Out of memory in the partition bound.");
	} else {
		for (unsigned T = 4; T < n; ++T) {
			int b = 0;
			for (int c = 0; c < n_chunks; ++c) b += pl.job_value[(T - 4) * n_chunks + c];
			if (b > 0 && (unsigned) b > p->rest_bound[T - 1]) p->rest_bound[T - 1] = (unsigned) b;   /* src/partbound.c:1155-1163 */
		}
	}
	free(pl.job_value);
	free(chunks);
	free(single);
	free(col);
	return rc;
}

int xmph_site_pair_cost(unsigned index) {
	if (!index || index > 0xffffu) return -1;
	work w;
	shared sh;
	memset(&sh, 0, sizeof sh);
	sh.n_taxa = 16;
	if (work_init(&w, &sh)) { work_free(&w); return -1; }
	int v = pair_bound(&w, index);
	work_free(&w);
	return v;
}
