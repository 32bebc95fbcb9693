/* TEST INFRASTRUCTURE ONLY -- see oracle.h.  Partial-tree state sets and exact branch and bound.
 *
 * Deliberately the simplest correct formulation: every partial tree's UP / DOWN / MIDPOINT sets are
 * recomputed from scratch (one post-order and one pre-order pass) instead of the reference's
 * incremental UpdateTree (src/yanbader.c:314-495); the two are equivalent because the incremental
 * version only skips recomputation where the result is unchanged (src/yanbader.c:364-369).
 *
 * Definitions restated from the reference:
 *   tree shape    root node = taxon 0 with a single LEFT child (src/common.c:547-550)
 *   UP(leaf)      the taxon's row;  UP(v) = F(UP(left), UP(right))        (src/common.c:559-567)
 *   DOWN(c)       root's row for the root's child (src/common.c:569-570);
 *                 otherwise F(DOWN(parent), UP(sibling))                  (src/common.c:582-610)
 *   MIDPOINT(v)   F(UP(v), DOWN(v))                                       (src/common.c:572-580)
 *   insertion     new leaf LEFT, old subtree RIGHT                        (src/common.c:735-736)
 *   cost          FitchScore(MIDPOINT(edge), row(taxon)); accept iff
 *                 (int) cost <= (int) (bound - score - restBound[m])      (src/yanbader.c:152-176)
 *   edge order    pre-order from root->LEFT: own edge, LEFT, RIGHT        (src/yanbader.c:103,238-243)
 *   full tree     score < bound: lower the bound and forget saved trees;
 *                 then record the tree, counting past max_trees           (src/yanbader.c:195-213,260-294)
 */
#include <limits.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "oracle.h"

/* The search uses the 64-bit forms (rows are whole 16-byte blocks); tests/test_oracle.py checks them
 * against the nibble loops and against the compiled reference. */
static void bases(const uint8_t *a, const uint8_t *b, uint8_t *d, size_t nbytes) {
	xo_fitch_bases_swar64((const uint64_t *) a, (const uint64_t *) b, (uint64_t *) d, nbytes / 8);
}
static unsigned score_w(const uint8_t *a, const uint8_t *b, const unsigned *w, size_t nbytes) {
	return w ? xo_fitch_score_swar64((const uint64_t *) a, (const uint64_t *) b, w, nbytes / 8)
	         : xo_fitch_score_weight1_swar64((const uint64_t *) a, (const uint64_t *) b, nbytes / 8);
}

typedef struct xnode {
	int taxon;                    /* search-order taxon index for a leaf, -1 for an internal node */
	unsigned edge;                /* edge id in the shared numbering (oracle.h) */
	struct xnode *par, *kid[2];   /* kid[0] = LEFT, kid[1] = RIGHT */
	uint8_t *up, *down, *mid;
} xnode;

typedef struct {
	const uint8_t *rows;
	size_t nbytes;
	xnode *pool;
	unsigned used;
	uint8_t *setmem;
	int skip_cost;            /* the search tracks the tree length incrementally */
} forest;

static xnode *new_node(forest *f, int taxon, unsigned edge) {
	xnode *n = &f->pool[f->used];
	n->taxon = taxon;
	n->edge = edge;
	n->par = n->kid[0] = n->kid[1] = NULL;
	n->up = f->setmem + (size_t) f->used * 3 * f->nbytes;
	n->down = n->up + f->nbytes;
	n->mid = n->down + f->nbytes;
	if (taxon >= 0) memcpy(n->up, f->rows + (size_t) taxon * f->nbytes, f->nbytes);
	++f->used;
	return n;
}

static void forest_init(forest *f, const uint8_t *rows, size_t nbytes, unsigned max_taxa) {
	f->rows = rows;
	f->nbytes = nbytes;
	f->pool = (xnode *) calloc(2 * (size_t) max_taxa, sizeof(xnode));
	f->setmem = (uint8_t *) aligned_alloc(64, ((2 * (size_t) max_taxa * 3 * nbytes + 64) + 63) / 64 * 64);
	f->used = 0;
	f->skip_cost = 0;
	if (!f->pool || !f->setmem) {
		fprintf(stderr, "oracle: out of memory allocating %zu bytes of state sets\n", 2 * (size_t) max_taxa * 3 * nbytes);
		abort();
	}
}

static void forest_free(forest *f) {
	free(f->pool);
	free(f->setmem);
}

/* The 3-taxon start tree: root = taxon 0, internal node with LEFT = taxon 1, RIGHT = taxon 2. */
static xnode *base_tree(forest *f) {
	xnode *root = new_node(f, 0, UINT_MAX);
	xnode *l1 = new_node(f, 1, 0), *l2 = new_node(f, 2, 1), *in = new_node(f, -1, 2);
	root->kid[0] = in; in->par = root;
	in->kid[0] = l1; l1->par = in;
	in->kid[1] = l2; l2->par = in;
	return root;
}

/* Splice taxon t into the edge above v.  Returns the new internal node. */
static xnode *insert_above(forest *f, xnode *v, int t) {
	xnode *leaf = new_node(f, t, 2u * (unsigned) t - 3u);
	xnode *in = new_node(f, -1, 2u * (unsigned) t - 2u);
	xnode *p = v->par;
	int side = (p->kid[1] == v);
	in->par = p; p->kid[side] = in;
	in->kid[0] = leaf; leaf->par = in;
	in->kid[1] = v; v->par = in;
	return in;
}

/* Undo the most recent insert_above (nodes are released LIFO). */
static void remove_inserted(forest *f, xnode *in) {
	xnode *p = in->par, *v = in->kid[1];
	int side = (p->kid[1] == in);
	p->kid[side] = v; v->par = p;
	f->used -= 2;
}

static unsigned post_up(const forest *f, xnode *n, const unsigned *weights) {
	unsigned cost = 0;
	if (n->taxon >= 0) return 0;
	cost += post_up(f, n->kid[0], weights);
	cost += post_up(f, n->kid[1], weights);
	if (!f->skip_cost) cost += score_w(n->kid[0]->up, n->kid[1]->up, weights, f->nbytes);
	bases(n->kid[0]->up, n->kid[1]->up, n->up, f->nbytes);
	return cost;
}

static void pre_down(const forest *f, xnode *n) {
	xnode *p = n->par;
	if (p->par == NULL) {
		memcpy(n->down, p->up, f->nbytes);
	} else {
		xnode *sib = p->kid[p->kid[0] == n];
		bases(p->down, sib->up, n->down, f->nbytes);
	}
	bases(n->up, n->down, n->mid, f->nbytes);
	if (n->taxon < 0) {
		pre_down(f, n->kid[0]);
		pre_down(f, n->kid[1]);
	}
}

/* Recompute every set of the tree; returns its Fitch length. */
static unsigned recompute(const forest *f, xnode *root, const unsigned *weights) {
	xnode *top = root->kid[0];
	unsigned cost = post_up(f, top, weights);
	if (!f->skip_cost) cost += score_w(top->up, root->up, weights, f->nbytes);
	pre_down(f, top);
	return cost;
}

static xnode *find_edge(xnode *n, unsigned edge) {
	xnode *r;
	if (!n) return NULL;
	if (n->edge == edge) return n;
	if ((r = find_edge(n->kid[0], edge))) return r;
	return find_edge(n->kid[1], edge);
}

unsigned xo_tree_sets(const uint8_t *char_mat, size_t nbytes, const unsigned *weights, unsigned m,
                      const uint8_t *path, uint8_t *mid, uint8_t *up, uint8_t *down) {
	forest f;
	unsigned cost;
	forest_init(&f, char_mat, nbytes, m);
	xnode *root = base_tree(&f);
	for (unsigned t = 3; t < m; ++t) insert_above(&f, find_edge(root->kid[0], path[t]), (int) t);
	cost = recompute(&f, root, weights);
	for (unsigned i = 1; i < f.used; ++i) {
		xnode *n = &f.pool[i];
		if (mid) memcpy(mid + (size_t) n->edge * nbytes, n->mid, nbytes);
		if (up) memcpy(up + (size_t) n->edge * nbytes, n->up, nbytes);
		if (down) memcpy(down + (size_t) n->edge * nbytes, n->down, nbytes);
	}
	forest_free(&f);
	return cost;
}

/* ---- search -------------------------------------------------------------------------------- */
typedef struct {
	const xo_problem *p;
	xo_result *r;
	forest f;
	xnode *root;
	unsigned bound, score;
	size_t cap, pcap;
	uint8_t cur_path[256];
	uint8_t **level_mid;      /* per tree size: MIDPOINT of every edge in pre-order */
	xnode ***level_edge;      /* per tree size: the node below every edge in pre-order */
} search;

static void emit(search *s, const char *txt, size_t n) {
	xo_result *r = s->r;
	if (r->newick_len + n + 1 > s->cap) {
		s->cap = (r->newick_len + n + 1) * 2 + 4096;
		r->newick = (char *) realloc(r->newick, s->cap);
	}
	memcpy(r->newick + r->newick_len, txt, n);
	r->newick_len += n;
	r->newick[r->newick_len] = 0;
}

static void print_sub(search *s, const xnode *n) {
	char buf[16];
	if (n->taxon >= 0) {
		emit(s, buf, (size_t) sprintf(buf, "%u", s->p->taxon_map[n->taxon] + 1));
	} else {
		emit(s, "(", 1);
		print_sub(s, n->kid[0]);
		emit(s, ",", 1);
		print_sub(s, n->kid[1]);
		emit(s, ")", 1);
	}
}

static void record_tree(search *s) {
	xo_result *r = s->r;
	if (r->num_trees < s->p->max_trees) {
		const unsigned n = s->p->num_taxa;
		if (r->paths_len + n > s->pcap) {
			s->pcap = (r->paths_len + n) * 2 + 4096;
			r->paths = (uint8_t *) realloc(r->paths, s->pcap);
		}
		memcpy(r->paths + r->paths_len, s->cur_path, n);
		r->paths_len += n;
		char buf[16];
		emit(s, "(", 1);
		emit(s, buf, (size_t) sprintf(buf, "%u", s->p->taxon_map[0] + 1));
		emit(s, ",", 1);
		print_sub(s, s->root->kid[0]);
		emit(s, ");\n", 3);
	}
	++r->num_trees;
}

static unsigned collect_preorder(xnode *n, xnode **out, unsigned k) {
	out[k++] = n;
	if (n->taxon < 0) {
		k = collect_preorder(n->kid[0], out, k);
		k = collect_preorder(n->kid[1], out, k);
	}
	return k;
}

static void branch_and_bound(search *s, unsigned m) {
	const xo_problem *p = s->p;
	const size_t nb = p->nbytes;
	if (m >= p->num_taxa) return;
	++s->r->nodes;
	recompute(&s->f, s->root, NULL);
	xnode **edges = s->level_edge[m];
	uint8_t *mids = s->level_mid[m];
	unsigned ne = collect_preorder(s->root->kid[0], edges, 0);
	for (unsigned e = 0; e < ne; ++e) memcpy(mids + (size_t) e * nb, edges[e]->mid, nb);
	const uint8_t *row = p->char_mat + (size_t) m * nb;
	for (unsigned e = 0; e < ne; ++e) {
		int allowable = (int) (s->bound - s->score - p->rest_bound[m]);
		int inc = (int) score_w(mids + (size_t) e * nb, row, p->weights, nb);
		++s->r->score_calls;
		s->score += (unsigned) inc;
		if (inc <= allowable) {
			s->cur_path[m] = (uint8_t) edges[e]->edge;
			xnode *in = insert_above(&s->f, edges[e], (int) m);
			if (m + 1 == p->num_taxa) {
				if (s->score < s->bound) {
					s->bound = s->score;
					s->r->num_trees = 0;
					s->r->newick_len = 0;
					s->r->paths_len = 0;
				}
				record_tree(s);
			} else {
				branch_and_bound(s, m + 1);
			}
			remove_inserted(&s->f, in);
		}
		s->score -= (unsigned) inc;
	}
}

int xo_search_subtree(const xo_problem *p, const uint8_t *path, unsigned m, xo_result *r) {
	search s;
	memset(&s, 0, sizeof s);
	memset(r, 0, sizeof *r);
	if (p->num_taxa < 4 || p->num_taxa > 128 || m < 3 || m >= p->num_taxa) return -1;
	s.p = p;
	s.r = r;
	forest_init(&s.f, p->char_mat, p->nbytes, p->num_taxa);
	s.root = base_tree(&s.f);
	for (unsigned t = 3; t < m; ++t) {
		xnode *v = find_edge(s.root->kid[0], path[t]);
		if (!v) { forest_free(&s.f); return -1; }
		insert_above(&s.f, v, (int) t);
		s.cur_path[t] = path[t];
	}
	s.level_mid = (uint8_t **) calloc(p->num_taxa + 1, sizeof(uint8_t *));
	s.level_edge = (xnode ***) calloc(p->num_taxa + 1, sizeof(xnode **));
	for (unsigned k = m; k < p->num_taxa; ++k) {
		s.level_mid[k] = (uint8_t *) malloc((size_t) (2 * k - 3) * p->nbytes);
		s.level_edge[k] = (xnode **) malloc((size_t) (2 * k - 3) * sizeof(xnode *));
	}
	s.bound = p->bound;
	s.score = recompute(&s.f, s.root, p->weights) + p->constant_weight;
	s.f.skip_cost = 1;
	branch_and_bound(&s, m);
	r->score = s.bound;
	for (unsigned k = m; k < p->num_taxa; ++k) {
		free(s.level_mid[k]);
		free(s.level_edge[k]);
	}
	free(s.level_mid);
	free(s.level_edge);
	forest_free(&s.f);
	return 0;
}

int xo_search(const xo_problem *p, xo_result *r) {
	return xo_search_subtree(p, NULL, 3, r);
}

void xo_result_free(xo_result *r) {
	free(r->newick);
	free(r->paths);
	r->newick = NULL;
	r->paths = NULL;
	r->newick_len = r->paths_len = 0;
}
