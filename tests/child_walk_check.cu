// Host-side replay of the search kernels' child construction (xmp_b200/csrc/child_walk.cuh, the very code
// k_build_children_eager / k_build_children / k_expand_last_fused run per thread) against a from-scratch
// computation of the child tree's state sets.  TEST INFRASTRUCTURE: compiled with nvcc for the host only,
// run by tests/test_host.py without a GPU; prints "OK <children checked>" or a diagnostic and exits non-zero.
//
// For random alignments (single states and ambiguity sets, 1-3 blocks per set, weighted and unweighted) and
// random stepwise-addition trees, every insertion position P of a parent node laid out exactly as a frontier
// pool lays it out (tiled header with the walk program, tiled UP rows, eager or classic) is expanded with
// ChildWalk::init / word_for / step and the loop body of k_build_children_eager; per step the child's program
// word, UP set, MIDPOINT set and the cost of the next taxon must equal the values computed directly from the
// child's topology by the definitions of the reference (src/common.c:543-650).
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <cuda_runtime.h>
#include "../xmp_b200/csrc/child_walk.cuh"

using namespace xmp;

enum { MAXE = 2 * 100 };

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
	void insert(unsigned v) {          // taxon m goes on edge v: new leaf LEFT, v RIGHT (src/common.c:722-770)
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

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static unsigned rnd(unsigned n) {
	rng_state ^= rng_state << 13;
	rng_state ^= rng_state >> 7;
	rng_state ^= rng_state << 17;
	return (unsigned) ((rng_state >> 11) % n);
}

static bool same(const uint4 &a, const uint4 &b) { return a.x == b.x && a.y == b.y && a.z == b.z && a.w == b.w; }

// UP / DOWN / MIDPOINT of every edge of t, column `col`, straight from the definitions.
static void sets_from_scratch(const Topo &t, const std::vector<uint4> &rows, unsigned W, unsigned col, uint4 *up, uint4 *down, uint4 *mid) {
	uint8_t pre[MAXE];
	const unsigned E = t.preorder(pre);
	for (int i = (int) E - 1; i >= 0; --i) {
		const unsigned e = pre[i];
		up[e] = t.k0[e] == 0xFF ? rows[(size_t) edge_taxon(e) * W + col] : fitch_block(up[t.k0[e]], up[t.k1[e]]);
	}
	for (unsigned i = 0; i < E; ++i) {
		const unsigned e = pre[i], p = t.par[e];
		if (p == 0xFF) down[e] = rows[col];
		else down[e] = fitch_block(down[p], up[t.k0[p] == e ? t.k1[p] : t.k0[p]]);
		mid[e] = fitch_block(up[e], down[e]);
	}
}

// The SWAR primitives themselves (fitch_dev.cuh), against the definition nibble by nibble (src/seq_b2_w1_basic.c:155-244: the
// intersection if it is not empty, else the union and one unit of cost): every pair of state sets -- the empty set included,
// which only padding can produce -- at every nibble position, with random neighbours so that a carry or a borrow crossing a
// nibble boundary would show.
static unsigned check_primitives() {
	unsigned checked = 0;
	for (unsigned pos = 0; pos < 8; ++pos)
		for (unsigned sa = 0; sa < 16; ++sa)
			for (unsigned sb = 0; sb < 16; ++sb)
				for (unsigned rep = 0; rep < 64; ++rep) {
					unsigned a = 0, b = 0;
					for (unsigned s = 0; s < 8; ++s) {
						a |= (s == pos ? sa : rnd(16)) << (4 * s);
						b |= (s == pos ? sb : rnd(16)) << (4 * s);
					}
					unsigned want = 0, want_cost = 0, want_empty = 0;
					for (unsigned s = 0; s < 8; ++s) {
						const unsigned x = (a >> (4 * s)) & 15u, y = (b >> (4 * s)) & 15u;
						want |= ((x & y) ? (x & y) : (x | y)) << (4 * s);
						if (!(x & y)) {
							++want_cost;
							want_empty |= 8u << (4 * s);
						}
					}
					if (fitch_word(a, b) != want || cost_word_w1(a, b) != want_cost || empty_mask(a, b) != want_empty) {
						printf("primitive mismatch: a=%08x b=%08x merge %08x (want %08x) cost %u (want %u)\n", a, b, fitch_word(a, b), want, cost_word_w1(a, b), want_cost);
						exit(1);
					}
					++checked;
				}
	return checked;
}

// The planar form of a block (fitch_dev.cuh): the transposition bit by bit, then the two facts a kernel that keeps its sets
// planar relies on -- merging commutes with the transposition, costs do not change -- on random blocks (single states,
// ambiguity sets, padding nibbles 0xF, and empty sets for good measure), weight-1 and weighted, heavy and light blocks.
static unsigned check_planar() {
	unsigned checked = 0;
	for (unsigned rep = 0; rep < 60000; ++rep) {
		unsigned aw[4], bw[4];
		for (unsigned i = 0; i < 4; ++i) {
			aw[i] = bw[i] = 0;
			for (unsigned j = 0; j < 8; ++j) {
				const unsigned kind = rnd(10);
				aw[i] |= (kind < 6 ? 1u << rnd(4) : (kind < 9 ? rnd(16) : 15u)) << (4 * j);
				bw[i] |= (rnd(10) < 6 ? 1u << rnd(4) : rnd(16)) << (4 * j);
			}
		}
		const uint4 a = make_uint4(aw[0], aw[1], aw[2], aw[3]), b = make_uint4(bw[0], bw[1], bw[2], bw[3]);
		const uint4 pa = to_planar(a), pb = to_planar(b);
		const unsigned pw[4] = {pa.x, pa.y, pa.z, pa.w};
		for (unsigned i = 0; i < 4; ++i)
			for (unsigned j = 0; j < 8; ++j)
				for (unsigned k = 0; k < 4; ++k)
					if (((pw[k] >> (4 * j + i)) & 1u) != ((aw[i] >> (4 * j + k)) & 1u)) {
						printf("to_planar: word %u nibble %u base %u\n", i, j, k);
						exit(1);
					}
		if (!same(to_planar(fitch_block(a, b)), fitch_block_planar(pa, pb))) {
			printf("planar merge differs\n");
			exit(1);
		}
		const unsigned w1 = cost_word_w1(a.x, b.x) + cost_word_w1(a.y, b.y) + cost_word_w1(a.z, b.z) + cost_word_w1(a.w, b.w);
		if (cost_block_planar_w1(pa, pb) != w1) {
			printf("planar weight-1 cost differs\n");
			exit(1);
		}
		// weighted: 3 blocks of planes built the way the library builds them, the heavy part ending anywhere (also inside a block)
		const unsigned n_words = 12, n_planes = 3, heavy_words = rnd(n_words + 1), block = rnd(3);
		unsigned planes[3 * 12];
		for (unsigned x = 0; x < n_planes * n_words; ++x) planes[x] = 0;
		for (unsigned w = 0; w < heavy_words; ++w)
			for (unsigned nib = 0; nib < 8; ++nib) {
				const unsigned wt = 1 + rnd(7);
				for (unsigned k = 0; k < n_planes; ++k)
					if ((wt >> k) & 1u) planes[k * n_words + w] |= 8u << (4 * nib);
			}
		WeightPlanes wp{planes, n_words, heavy_words, n_planes};
		if (cost_block_planar(pa, pb, block, wp) != cost_block(a, b, block, wp)) {
			printf("planar weighted cost differs (block %u, heavy words %u)\n", block, heavy_words);
			exit(1);
		}
		++checked;
	}
	// the neutral element of the merge is all ones in both forms
	const uint4 ones = make_uint4(~0u, ~0u, ~0u, ~0u);
	if (!same(to_planar(ones), ones)) {
		printf("to_planar(ones)\n");
		exit(1);
	}
	return checked;
}

int main(int argc, char **argv) {
	check_primitives();
	check_planar();
	const unsigned max_taxa = argc > 1 ? (unsigned) atoi(argv[1]) : 24;
	const unsigned reps = argc > 2 ? (unsigned) atoi(argv[2]) : 40;
	unsigned long checked = 0;
	// second pass: the same replay with every state set in planar form (a search run with XMP_SEARCH_PLANAR=1: planar rows in,
	// the kernels' <PLANAR = true> instantiations) -- sets must be the transposed sets, programs and costs the same
	for (unsigned pass = 0; pass < 2; ++pass)
	for (unsigned rep = 0; rep < reps; ++rep) {
		const bool planar = pass == 1;
		auto X = [&](const uint4 &v) { return planar ? to_planar(v) : v; };
		auto MB = [&](const uint4 &a, const uint4 &b) { return planar ? merge_block<true>(a, b) : merge_block<false>(a, b); };
		const unsigned n = max_taxa, W = 1 + rep % 3;
		const bool eager = rep % 4 != 3, weighted = rep % 2 == 1;
		// alignment: mostly single states, some ambiguity sets; padding nibbles do not occur (every nibble is a set)
		std::vector<uint4> rows((size_t) n * W);
		for (auto &r : rows) {
			unsigned w[4];
			for (unsigned k = 0; k < 4; ++k) {
				w[k] = 0;
				for (unsigned s = 0; s < 8; ++s) w[k] |= (rnd(10) ? 1u << rnd(4) : 1u + rnd(15)) << (4 * s);
			}
			r = make_uint4(w[0], w[1], w[2], w[3]);
		}
		// weights: descending, bit-sliced planes as the library builds them (fitch_dev.cuh)
		const unsigned n_words = 4 * W, n_planes = 3;
		std::vector<unsigned> weights((size_t) 8 * n_words, 1), planes((size_t) n_planes * n_words, 0);
		unsigned heavy_words = 0;
		if (weighted) {
			const unsigned heavy_sites = 1 + rnd(8 * n_words - 1);
			for (unsigned s = 0; s < heavy_sites; ++s) weights[s] = 2 + rnd(6);
			for (unsigned s = 1; s < heavy_sites; ++s) if (weights[s] > weights[s - 1]) weights[s] = weights[s - 1];
			heavy_words = (heavy_sites + 7) / 8;
			for (unsigned s = 0; s < 8 * n_words; ++s)
				for (unsigned k = 0; k < n_planes; ++k)
					if (weights[s] >> k & 1u) planes[(size_t) k * n_words + s / 8] |= 8u << (4 * (s % 8));
		}
		WeightPlanes wp{planes.data(), n_words, heavy_words, n_planes};

		Layout L;
		L.n = n;
		L.W = W;
		L.off_prog = HDR_PATH_W + ((n + 3u) >> 2);
		const unsigned Emax = 2 * n - 3;
		L.off_cost = eager ? L.off_prog + Emax : 0u;
		L.NW = L.off_prog + Emax + (eager ? Emax : 0u);
		L.ts_shift = W == 1 ? 3u : (W == 2 ? 2u : (W <= 4 ? 1u : 0u));
		L.trec = 4 + ((n + 3u) & ~3u);
		const unsigned slot = rnd(32);                 // any slot of the first header tile

		Topo t;
		t.base();
		while (t.m + 1 < n) {
			const unsigned m = t.m, E = 2 * m - 3, Ec = E + 2;
			uint8_t pre[MAXE];
			unsigned char pos[MAXE], size[MAXE];
			unsigned prog[MAXE];
			t.preorder(pre);
			prog_from_topology(t.par, t.k0, t.k1, pre, E, prog, pos, size);
			// the parent node as a pool holds it
			const unsigned rows_per_node = eager ? E : 2 * E, TS = 1u << L.ts_shift;
			std::vector<unsigned> hdr((size_t) L.NW * 32, 0xDEADBEEFu);
			std::vector<uint4> sets((size_t) 32 / TS * rows_per_node * TS * W + 64);
			PoolView pool{hdr.data(), sets.data(), rows_per_node, 32};
			unsigned *ph = hdr_of(pool, L, slot);
			for (unsigned i = 0; i < E; ++i) ph[(L.off_prog + i) * 32u] = prog[i];
			std::vector<uint4> up(MAXE), down(MAXE), mid(MAXE);
			for (unsigned col = 0; col < W; ++col) {
				sets_from_scratch(t, rows, W, col, up.data(), down.data(), mid.data());
				// the same tree with every set in planar form, as k_score_trees_wide keeps them: rows transposed once, planar
				// merges throughout -- every set must be the transposed packed set, every insertion cost the same
				{
					std::vector<uint4> pup(MAXE), pdown(MAXE);
					for (int i = (int) E - 1; i >= 0; --i) {
						const unsigned e = pre[i];
						pup[e] = t.k0[e] == 0xFF ? to_planar(rows[(size_t) edge_taxon(e) * W + col]) : fitch_block_planar(pup[t.k0[e]], pup[t.k1[e]]);
					}
					const uint4 pnext = to_planar(rows[(size_t) m * W + col]);
					for (unsigned i = 0; i < E; ++i) {
						const unsigned e = pre[i], p = t.par[e];
						pdown[e] = p == 0xFF ? fitch_block_planar(to_planar(rows[col]), make_uint4(~0u, ~0u, ~0u, ~0u))
						                     : fitch_block_planar(pdown[p], pup[t.k0[p] == e ? t.k1[p] : t.k0[p]]);
						const uint4 pmid = fitch_block_planar(pup[e], pdown[e]);
						const bool w1 = 4u * col >= wp.heavy_words;
						const unsigned c = w1 ? cost_block_planar_w1(pmid, pnext) : cost_block_planar(pmid, pnext, col, wp);
						if (!same(pup[e], to_planar(up[e])) || !same(pdown[e], to_planar(down[e])) || !same(pmid, to_planar(mid[e])) ||
						    c != cost_block(mid[e], rows[(size_t) m * W + col], col, wp)) {
							printf("planar tree: m=%u edge %u column %u\n", m, e, col);
							return 1;
						}
					}
				}
				uint4 *base = sets_of(pool, L, slot) + col;
				const size_t rs = (size_t) W << L.ts_shift;
				for (unsigned i = 0; i < E; ++i) {
					base[(size_t) (up_row0(L, E) + i) * rs] = X(up[pre[i]]);
					if (!eager) base[(size_t) i * rs] = X(mid[pre[i]]);
				}
			}
			// expand at every position; continue the tree with one of them
			const unsigned keep = rnd(E);
			Topo next = t;
			for (unsigned P = 0; P < E; ++P) {
				Topo c = t;
				c.insert(prog_edge(prog[P]));
				uint8_t cpre[MAXE];
				unsigned want[MAXE];
				c.preorder(cpre);
				prog_from_topology(c.par, c.k0, c.k1, cpre, Ec, want, pos, size);
				std::vector<unsigned> cost(Ec, 0), want_cost(Ec, 0);
				for (unsigned col = 0; col < W; ++col) {
					std::vector<uint4> cup(MAXE), cdown(MAXE), cmid(MAXE);
					sets_from_scratch(c, rows, W, col, cup.data(), cdown.data(), cmid.data());
					// --- what one thread of k_build_children_eager does for (this child, this column) ---
					const unsigned T = 1, D = m + 1;
					std::vector<uint4> upath(D), dstack(D);
					const uint4 *PU = sets_of(pool, L, slot) + (size_t) up_row0(L, E) * ((size_t) W << L.ts_shift) + col;
					const uint4 r0 = X(rows[col]), rt = X(rows[(size_t) m * W + col]), rn = X(rows[(size_t) (m + 1) * W + col]);
					bool overflow = false;
					ChildWalk cw;
					// eager layout: the root-path UP sets live in the child's own rows (tiled like a pool's), as in the kernel
					std::vector<uint4> crows((size_t) Ec * ((size_t) W << L.ts_shift) + 64);
					uint4 *CU = eager ? crows.data() + col : nullptr;
					uint4 *up_arg = eager ? nullptr : upath.data();
					if (planar) cw.init<true>(ph, L, m, P, PU, rt, up_arg, T, D, overflow, CU);
					else cw.init<false>(ph, L, m, P, PU, rt, up_arg, T, D, overflow, CU);
					unsigned wj, pw1 = cw.word_for(1);
					uint4 ux, us;
					cw.step(0, cw.word_for(0), wj, ux, us, up_arg, T, r0);
					for (unsigned j = 0; j < Ec; ++j) {
						const unsigned pw2 = cw.word_for(j + 2);
						unsigned wn = wj;
						uint4 uxn = ux, usn = us;
						if (j + 1 < Ec) cw.step(j + 1, pw1, wn, uxn, usn, up_arg, T, r0);
						pw1 = pw2;
						const unsigned d = prog_depth(wj);
						const uint4 dn = d == 0 ? r0 : MB(dstack[d - 1], us);
						if (prog_size(wj) > 1u) {
							if (d < D) dstack[d] = dn; else overflow = true;
						}
						if (CU) CU[(size_t) j * ((size_t) W << L.ts_shift)] = ux;
						const uint4 m_j = MB(ux, dn);
						// the kernels' own selection: the weight-1 form for columns past the heavy words
						if (4u * col >= wp.heavy_words) cost[j] += planar ? score_block_w1<true>(m_j, rn) : score_block_w1<false>(m_j, rn);
						else cost[j] += planar ? score_block<true>(m_j, rn, col, wp) : score_block<false>(m_j, rn, col, wp);
						// --- checks ---
						const unsigned e = cpre[j];
						if (wj != want[j]) { printf("rep %u m %u P %u step %u: word %08x, from scratch %08x\n", rep, m, P, j, wj, want[j]); return 1; }
						if (!same(ux, X(cup[e]))) { printf("rep %u m %u P %u step %u col %u: UP set differs\n", rep, m, P, j, col); return 1; }
						if (!same(dn, X(cdown[e]))) { printf("rep %u m %u P %u step %u col %u: DOWN set differs\n", rep, m, P, j, col); return 1; }
						if (!same(m_j, X(cmid[e]))) { printf("rep %u m %u P %u step %u col %u: MIDPOINT set differs\n", rep, m, P, j, col); return 1; }
						// the reference's definition of the cost, site by site
						const uint4 rn_packed = rows[(size_t) (m + 1) * W + col];
						const unsigned mw[4] = {cmid[e].x, cmid[e].y, cmid[e].z, cmid[e].w}, tw[4] = {rn_packed.x, rn_packed.y, rn_packed.z, rn_packed.w};
						for (unsigned k = 0; k < 4; ++k)
							for (unsigned s = 0; s < 8; ++s)
								if (!((mw[k] >> (4 * s)) & (tw[k] >> (4 * s)) & 15u)) want_cost[j] += weights[(size_t) 8 * (4 * col + k) + s];
						wj = wn; ux = uxn; us = usn;
					}
					if (overflow) { printf("rep %u m %u P %u: stack overflow flagged with the worst-case depth\n", rep, m, P); return 1; }
					// --- the lean step of k_expand_last_fused (depth, internal flag and the two UP sets only; upath[] variant) ---
					if (!eager) {
						std::vector<uint4> upath2(D), dstack2(D);
						ChildWalk lw;
						bool of2 = false;
						if (planar) lw.init<true>(ph, L, m, P, PU, rt, upath2.data(), T, D, of2);
						else lw.init<false>(ph, L, m, P, PU, rt, upath2.data(), T, D, of2);
						for (unsigned j = 0; j < Ec; ++j) {
							unsigned depth = 0;
							bool internal = false;
							uint4 lux, lus;
							lw.step_lean(j, lw.word_for(j), depth, internal, lux, lus, upath2.data(), T, r0);
							const unsigned e = cpre[j];
							const uint4 dn = depth == 0 ? r0 : MB(dstack2[depth - 1], lus);
							if (internal) dstack2[depth] = dn;
							if (depth != prog_depth(want[j]) || internal != (prog_size(want[j]) > 1u) || lw.edge_at(j) != prog_edge(want[j])) {
								printf("rep %u m %u P %u step %u: lean step gives depth %u internal %d edge %u, from scratch %08x\n", rep, m, P, j, depth, (int) internal, lw.edge_at(j), want[j]);
								return 1;
							}
							if (!same(lux, X(cup[e])) || !same(dn, X(cdown[e]))) { printf("rep %u m %u P %u step %u col %u: lean step: UP or DOWN set differs\n", rep, m, P, j, col); return 1; }
						}
					}
				}
				for (unsigned j = 0; j < Ec; ++j) {
					if (cost[j] != want_cost[j]) { printf("rep %u m %u P %u step %u: cost %u, site by site %u\n", rep, m, P, j, cost[j], want_cost[j]); return 1; }
				}
				if (P == keep) next = c;
				++checked;
			}
			t = next;
		}
	}
	printf("OK %lu\n", checked);
	return 0;
}
