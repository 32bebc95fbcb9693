// Frontier record layout and the child walk shared by the search kernels (search.cu).
//
// Everything here is plain arithmetic on a parent node's record, written __host__ __device__ so that the
// same code that the kernels run can be replayed on the CPU against a from-scratch construction of the child
// (tests/child_walk_check.cu, a "not gpu" test).  The product path only ever calls it from device code.
#pragma once
#include "fitch_dev.cuh"
#include "walk_prog.h"

namespace xmp {

struct Layout {
	unsigned n;         // taxa in the problem
	unsigned W;         // 128-bit blocks per state set
	unsigned NW;        // header words per node: score | insertion path (packed bytes) | walk program
	unsigned off_prog;  // index of the first program word
	unsigned ts_shift;  // records are tiled: 1 << ts_shift consecutive slots interleave their rows
	unsigned trec;      // bytes per emitted tree record: score + path
	unsigned off_cost;  // eager scoring: index of the first cost word (one per pre-order position); 0 = classic layout
};
// Two record layouts.  Classic: rows [MIDPOINT | UP], the scoring pass reads the MIDPOINT rows.  Eager (off_cost != 0,
// state sets of up to 32 blocks): the kernel that builds a node also scores the NEXT taxon on each of its edges and
// leaves the costs in the header, one word per position; MIDPOINT sets are never stored (rows = UP only, half the
// bytes), and the scoring pass shrinks to reading 4 bytes per (tree, edge) pair and applying the current bound.
XMP_HD unsigned up_row0(const Layout &L, unsigned E) { return L.off_cost ? 0u : E; }
// Header words: [0] tree length so far | [1, 1 + ceil(n/4)) insertion path, one byte per taxon (byte 0 = the
// root's child) | [off_prog, off_prog + 2n-3) walk program (walk_prog.h), one word per pre-order position.
#define HDR_PATH_W 1u

// Both arrays of a pool are tiled so that threads working on CONSECUTIVE slots touch consecutive addresses
// when they access the same row -- and every row is indexed by pre-order position, which all lanes of a
// warp reach at the same step of their walks:
//   header word w of slot s:        hdr [((s / 32) * NW + w) * 32 + s % 32]                     (128-byte lines)
//   block c of row k of slot s:     sets[(((s / TS) * nsets + k) * TS + s % TS) * W + c]        TS = 1 << ts_shift
// with TS * W * 16 B = one 128-byte line for narrow sets (W = 1: 8 slots, W = 2: 4, W <= 4: 2) and TS = 1,
// i.e. plain records, from W = 5 up.  Rows of a node: MIDPOINT at [0, E), UP at [E, 2E); DOWN sets are transient.
struct PoolView {
	unsigned *hdr;
	uint4 *sets;
	unsigned nsets;           // rows per node: 2 * (2m-3), or 2m-3 with eager scoring
	unsigned long long cap;   // slots (a multiple of 32); pools are rings: slot = index modulo cap
};

XMP_HD unsigned *hdr_of(const PoolView &p, const Layout &L, unsigned long long slot) {
	return p.hdr + (slot >> 5) * L.NW * 32ull + (slot & 31ull);               // word w at [w * 32]
}
XMP_HD uint4 *sets_of(const PoolView &p, const Layout &L, unsigned long long slot) {
	const unsigned ts = L.ts_shift;
	return p.sets + ((((slot >> ts) * p.nsets) << ts) + (slot & ((1ull << ts) - 1ull))) * L.W;   // row k at [(k << ts) * W]
}
// ---- walking a child tree ----------------------------------------------------------------------------
// The child of parent node N obtained by inserting the next taxon on the edge at position P, visited in
// pre-order without ever materialising its topology: step j of the child is step j (before P) or j - 2 (after)
// of the parent's walk program with three spliced steps at P (walk_prog.h) -- selects, the same trip count
// for every lane of a level, and every lane is at the same position j at the same time.
//
// Per-lane state with dynamic indexing lives in shared memory, element k of thread t at [k * blockDim.x + t]
// (16-byte accesses are then bank-conflict free whatever k each lane uses; local memory would serialise
// into one wavefront per distinct index):
//   upath[k]   UP sets recomputed along the root path: k = 0 the new internal node, k = dv - d the ancestor at
//              depth d; every other UP set is the parent tree's, read from its record (L1 hits: the lanes of a
//              warp share a handful of parents).  A kernel that writes the child's UP rows anyway (the eager
//              builder) keeps these sets in the child's record instead (ChildWalk::gup) and has no upath[];
//   dstack[d]  DOWN set of the current ancestor at depth d.
struct ChildWalk {
	const unsigned *prog;         // parent's program: word i at prog[i * 32]
	const uint4 *PU;              // parent's UP rows for this lane's column: row i at PU[i * rs]
	unsigned rs, E;
	Splice sp;
	// "Is my sibling P or an ancestor of P?" without a table: in pre-order the ancestors a_0 .. a_{dv-1} of P come
	// first at their depths, so before the walk leaves P's subtree the latest node seen at depth e is an ancestor iff
	// e < (ancestors seen so far); afterwards the off-path siblings come back up with strictly decreasing depth.
	// lim = number of depths whose latest node is still an ancestor of P.  A node at depth d hangs off the root path
	// (its sibling is P or an ancestor, whose recomputed UP set sits in upath[dv - d]) iff it is no ancestor itself
	// and d <= lim; then lim = d (ancestor: lim = d + 1).
	unsigned lim;
	uint4 rt;
	// Where the UP sets recomputed along the root path live.  nullptr: the shared-memory stack upath[].  Otherwise the
	// child's own UP rows in its record (row = pre-order position, stride rs): init() writes the ancestors' rows there
	// (an ancestor keeps its position, the new internal node sits at P), step() reads them back by position with plain
	// loads -- a thread sees its own earlier stores -- and the per-thread stack, half of the walk's shared memory, goes.
	uint4 *gup;

	// Decodes the insertion at position P of the parent and recomputes UP along the root path.
	// PLANAR: the state sets (the parent's rows, row_new) are in planar form (fitch_dev.cuh); the walk itself never looks inside a set.
	template <bool PLANAR = false>
	XMP_HD unsigned init(const unsigned *ph, const Layout &L, unsigned m, unsigned P, const uint4 *parent_up,
	                                         const uint4 &row_new, uint4 *upath, unsigned T, unsigned D, bool &overflow,
	                                         uint4 *child_rows = nullptr) {
		prog = ph + L.off_prog * 32u;
		PU = parent_up;
		rs = L.W << L.ts_shift;
		E = 2 * m - 3;
		rt = row_new;
		sp = make_splice(P, prog[P * 32u], E);
		lim = 0;
		uint4 u = merge_block<PLANAR>(rt, XMP_LDG(PU + P * rs));
		gup = child_rows;
		if (gup) gup[P * rs] = u;
		else upath[0] = u;
		// Up the root path.  The set hanging off the path at each level is requested one level ahead: the chain
		// of dependent loads is then the program words alone (mostly L1 hits: a warp shares a few parents), not
		// program word -> row -> program word -> row.
		unsigned k = 1, cur = P, other = sp.sv;
		uint4 off = sp.dv > 0 ? XMP_LDG(PU + other * rs) : rt;       // `other` hangs off the path: unchanged
		for (unsigned depth = sp.dv; depth > 0; --depth) {
			cur = parent_pos(cur, other);
			other = prog_sib(prog[cur * 32u]);
			const uint4 off_next = depth > 1 ? XMP_LDG(PU + other * rs) : rt;
			u = merge_block<PLANAR>(u, off);
			if (gup) gup[cur * rs] = u;
			else if (k < D) upath[k * T] = u;
			else overflow = true;
			off = off_next;
			++k;
		}
		return k;           // merges done
	}
	// Step j of the child's pre-order: its program word, its UP set and its sibling's.
	// The parent's program word that step j of the child starts from (a load from the parent's header).
	XMP_HD unsigned word_for(unsigned j) const {
		unsigned i = parent_index(j, sp.P);
		if (i >= E) i = E - 1;
		return prog[i * 32u];
	}
	// w = word_for(j), loaded by the caller two steps ahead: the word and the rows it leads to are two dependent
	// loads, and with the word requested inside the step that chain sat on the critical path of every step.
	// Steps must be taken in order j = 0, 1, 2, ... (lim is running state).
	XMP_HD void step(unsigned j, unsigned w, unsigned &word, uint4 &ux, uint4 &us, const uint4 *upath, unsigned T, const uint4 &r0) {
		const unsigned P = sp.P;
		unsigned i = parent_index(j, P);
		if (i >= E) i = E - 1;
		const bool anc = i < P && i + prog_size(w) > P;          // proper ancestor of P: its subtree [i, i + size) holds P
		word = child_word(j, w, anc, sp);
		const unsigned d = prog_depth(w);
		// where the two UP sets come from: 0 = the parent's record (row), 1 = upath[k], 2 = the new taxon's row, 3 = none
		unsigned src_u = anc ? 1u : 0u, row_u = i, k_u = sp.dv - d;
		unsigned s = prog_sib(w);
		unsigned src_s = 0u, row_s = s, k_s = sp.dv - d;
		if (s == XMP_NO_POS) src_s = 3u;
		else if (!anc && d <= lim) src_s = 1u;                  // s == P included: d == dv there, k_s = 0
		lim = anc ? d + 1u : (lim < d ? lim : d);               // the spliced steps re-read positions P - 1 and P: no change
		if (j == P) {
			src_u = 1u; k_u = 0u;
			row_s = sp.sv; src_s = sp.sv == XMP_NO_POS ? 3u : 0u;
		}
		if (j == P + 1u) { src_u = 2u; src_s = 0u; row_s = P; }
		if (j == P + 2u) { src_u = 0u; row_u = P; src_s = 2u; }
		if (gup) {          // on-path sets by position: my own (an ancestor, or the new node at P: position j) / my sibling's (s <= P)
#ifdef XMP_WALK_UL
			// A/B variant (scripts/build_variant.sh): both rows come from global memory, so pick the ADDRESS and load
			// unconditionally (any valid row when the value is not used) -- selects instead of branch regions.
			const uint4 *pu = src_u == 1u ? gup + j * rs : PU + row_u * rs;
			const uint4 *ps = src_s == 1u ? gup + s * rs : PU + (src_s == 0u ? row_s : 0u) * rs;
			const uint4 a = *pu, b = *ps;
			ux = src_u == 2u ? rt : a;
			us = src_s < 2u ? b : (src_s == 2u ? rt : r0);
			return;
#else
			ux = src_u == 0u ? XMP_LDG(PU + row_u * rs) : (src_u == 1u ? gup[j * rs] : rt);
			us = src_s == 0u ? XMP_LDG(PU + row_s * rs) : (src_s == 1u ? gup[s * rs] : (src_s == 2u ? rt : r0));
			return;
#endif
		}
		ux = src_u == 0u ? XMP_LDG(PU + row_u * rs) : (src_u == 1u ? upath[k_u * T] : rt);
		us = src_s == 0u ? XMP_LDG(PU + row_s * rs) : (src_s == 1u ? upath[k_s * T] : (src_s == 2u ? rt : r0));
	}

	// The same step for a kernel that does not keep the child (k_expand_last_fused): instead of the child's whole
	// program word only what the walk itself consumes -- the depth of position j in the child and whether it is an
	// internal node -- and the two UP sets.  (The word costs ~50 instructions per step in selects; the fused kernel needs
	// the edge id of a position only for the few trees it emits, and takes it from edge_at().)  upath[] variant only.
	XMP_HD void step_lean(unsigned j, unsigned w, unsigned &depth, bool &internal, uint4 &ux, uint4 &us, const uint4 *upath, unsigned T,
	                      const uint4 &r0) {
		const unsigned P = sp.P, q = j - P;                      // q = 0, 1, 2: the spliced steps b, a, v
		const bool after = j > P + 2u;
		const unsigned i = after ? j - 2u : (j < E ? j : E - 1u); // parent position this step starts from (w = prog[parent_index(j)])
		const unsigned d = prog_depth(w), s = prog_sib(w), sz = prog_size(w);
		const bool anc = j < P && j + sz > P;                    // proper ancestor of P
		const bool on_path = !anc && d <= lim && s != XMP_NO_POS; // my sibling is P or an ancestor of P
		if (q >= 3u) lim = anc ? d + 1u : (lim < d ? lim : d);    // the spliced steps re-read positions P - 1 and P: no change
		// defaults: an ordinary node
		depth = d + ((after && i < P + sp.szv) ? 1u : 0u);
		internal = sz > 1u;
		unsigned src_u = anc ? 1u : 0u, row_u = i, src_s = s == XMP_NO_POS ? 3u : (on_path ? 1u : 0u), row_s = s, k = sp.dv - d;
		unsigned k_u = k;
		if (q == 0u) { src_u = 1u; k_u = 0u; row_s = sp.sv; src_s = sp.sv == XMP_NO_POS ? 3u : 0u; depth = sp.dv; internal = true; }
		if (q == 1u) { src_u = 2u; src_s = 0u; row_s = P; depth = sp.dv + 1u; internal = false; }
		if (q == 2u) { src_u = 0u; row_u = P; src_s = 2u; depth = sp.dv + 1u; internal = sp.szv > 1u; }
		ux = src_u == 0u ? XMP_LDG(PU + row_u * rs) : (src_u == 1u ? upath[k_u * T] : rt);
		us = src_s == 0u ? XMP_LDG(PU + row_s * rs) : (src_s == 1u ? upath[k * T] : (src_s == 2u ? rt : r0));
	}
	// Edge id of the child's position j (for the trees the fused kernel emits).
	XMP_HD unsigned edge_at(unsigned j) const {
		const unsigned P = sp.P;
		if (j == P) return sp.b;
		if (j == P + 1u) return sp.a;
		if (j == P + 2u) return sp.xv;
		return prog_edge(prog[parent_index(j, P) * 32u]);
	}
};

}  // namespace xmp
