// Host-side check of xmp_b200/csrc/walk_prog.h: a child's walk program derived from its parent's by
// child_word() must equal the program computed from scratch from the child's topology, for random
// insertion sequences.  Also replays the root-path walk the kernels do (ancestors via parent_pos).
// Built and run by tests/test_host.py; prints "OK <cases>" or a diagnostic and exits non-zero.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include "../xmp_b200/csrc/walk_prog.h"

using namespace xmp;

enum { MAXE = 200 };

struct Topo {
	unsigned m, top;
	uint8_t par[MAXE], k0[MAXE], k1[MAXE];
	void base() {
		m = 3;
		top = 2;
		memset(par, 0xFF, sizeof par);
		memset(k0, 0xFF, sizeof k0);
		memset(k1, 0xFF, sizeof k1);
		par[0] = par[1] = 2;
		k0[2] = 0;
		k1[2] = 1;
	}
	// taxon index t (0-based, = current number of taxa) goes on edge v: new leaf LEFT, v RIGHT
	void insert(unsigned v) {
		const unsigned a = 2 * m - 3, b = 2 * m - 2, pv = par[v];
		par[b] = (uint8_t) pv;
		if (pv == 0xFF) top = b;
		else if (k1[pv] == v) k1[pv] = (uint8_t) b;
		else k0[pv] = (uint8_t) b;
		k0[b] = (uint8_t) a;
		k1[b] = (uint8_t) v;
		par[a] = par[v] = (uint8_t) b;
		++m;
	}
	unsigned preorder(uint8_t *out) const {
		uint8_t st[MAXE];
		unsigned sp = 0, n = 0;
		st[sp++] = (uint8_t) top;
		while (sp) {
			unsigned e = st[--sp];
			out[n++] = (uint8_t) e;
			if (k0[e] != 0xFF) {
				st[sp++] = k1[e];
				st[sp++] = k0[e];
			}
		}
		return n;
	}
};

static uint64_t rng_state = 88172645463325252ull;
static unsigned rnd(unsigned n) {
	rng_state ^= rng_state << 13;
	rng_state ^= rng_state >> 7;
	rng_state ^= rng_state << 17;
	return (unsigned) ((rng_state >> 11) % n);
}

int main(int argc, char **argv) {
	const unsigned max_taxa = argc > 1 ? (unsigned) atoi(argv[1]) : 100;
	const unsigned reps = argc > 2 ? (unsigned) atoi(argv[2]) : 200;
	unsigned long cases = 0;
	for (unsigned rep = 0; rep < reps; ++rep) {
		Topo t;
		t.base();
		uint8_t pre[MAXE];
		unsigned char pos[MAXE], size[MAXE];
		unsigned prog[MAXE], child[MAXE], want[MAXE];
		unsigned E = t.preorder(pre);
		prog_from_topology(t.par, t.k0, t.k1, pre, E, prog, pos, size);
		while (t.m < max_taxa) {
			// caterpillar-heavy, balanced-heavy and uniform choices all get exercised
			unsigned P = rep % 3 == 0 ? rnd(E) : (rep % 3 == 1 ? E - 1 - rnd(E < 3 ? E : 3) : rnd(E < 4 ? E : 4));
			const Splice sp = make_splice(P, prog[P], E);
			// proper ancestors of P, found the way the kernels find them
			bool anc[MAXE];
			memset(anc, 0, sizeof anc);
			unsigned cur = P, other = sp.sv, depth = sp.dv, steps = 0;
			while (depth > 0) {
				if (other == XMP_NO_POS) { printf("rep %u: no sibling at depth %u\n", rep, depth); return 1; }
				const unsigned q = parent_pos(cur, other);
				if (!(q < P && q + prog_size(prog[q]) > P)) { printf("rep %u: parent_pos gave a non-ancestor\n", rep); return 1; }
				if (prog_depth(prog[q]) != depth - 1) { printf("rep %u: ancestor depth mismatch\n", rep); return 1; }
				anc[q] = true;
				other = prog_sib(prog[q]);
				cur = q;
				--depth;
				++steps;
			}
			if (cur != 0 || steps != sp.dv) { printf("rep %u: root path did not end at position 0\n", rep); return 1; }
			for (unsigned j = 0; j < E + 2; ++j) {
				unsigned i = parent_index(j, P);
				if (i >= E) i = E - 1;                      // the kernels clamp; only reached on the spliced steps
				child[j] = child_word(j, prog[i], anc[i] && i < P, sp);
			}
			t.insert(sp.xv);
			const unsigned E2 = t.preorder(pre);
			if (E2 != E + 2) { printf("rep %u: edge count\n", rep); return 1; }
			prog_from_topology(t.par, t.k0, t.k1, pre, E2, want, pos, size);
			for (unsigned j = 0; j < E2; ++j) {
				if (child[j] != want[j]) {
					printf("rep %u, %u taxa, insertion at position %u: step %u derived %08x, from scratch %08x\n", rep, t.m, P, j, child[j], want[j]);
					return 1;
				}
			}
			memcpy(prog, want, sizeof(unsigned) * E2);
			E = E2;
			++cases;
		}
	}
	printf("OK %lu\n", cases);
	return 0;
}
