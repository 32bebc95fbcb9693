// The "walk program" of a partial tree and how a child's program derives from its parent's.
// Plain integer logic, shared by the device kernels (search.cu) and a host-side check (tests/walk_check.cpp).
//
// A tree on m taxa has E = 2m-3 edges.  Position i (0 <= i < E) is the i-th edge in the reference's
// pre-order (src/yanbader.c:103,238-243: an edge, then its LEFT subtree, then its RIGHT subtree, starting at
// the root's only child).  One 32-bit word per position:
//
//     edge id | sibling's position << 8 | depth << 16 | subtree size << 24
//
// (edge ids are the append-only ids of include/xmp_cuda.h; the root's child has no sibling: 0xFF; subtree
// size counts edges, 1 for a leaf).  Because a node's first child directly follows it in pre-order, the
// parent of the node at position c whose sibling sits at position s is at min(c, s) - 1: no parent links,
// no edge-id-indexed tables are needed, and every per-node array of the frontier (state sets, program) is
// indexed by position -- which is what lets a warp of threads, each building a different child, write the
// same row at the same step.
//
// Inserting the next taxon on the edge at position P (edge id xv, sibling position sv, depth dv, subtree size
// szv) puts the new internal edge b at position P, the new leaf a at P + 1 (LEFT) and the old edge at
// P + 2 (RIGHT), shifts everything behind P by two, deepens xv's subtree by one and grows every proper
// ancestor of P by two (reference src/common.c:722-770).
#pragma once

#if defined(__CUDACC__)
#define XMP_HD __host__ __device__ __forceinline__
#else
#define XMP_HD inline
#endif

namespace xmp {

#define XMP_NO_POS 0xFFu

XMP_HD unsigned prog_word(unsigned edge, unsigned sib_pos, unsigned depth, unsigned size) {
	return edge | (sib_pos << 8) | (depth << 16) | (size << 24);
}
XMP_HD unsigned prog_edge(unsigned w) { return w & 0xFFu; }
XMP_HD unsigned prog_sib(unsigned w) { return (w >> 8) & 0xFFu; }
XMP_HD unsigned prog_depth(unsigned w) { return (w >> 16) & 0xFFu; }
XMP_HD unsigned prog_size(unsigned w) { return w >> 24; }

XMP_HD unsigned parent_pos(unsigned c, unsigned s) { return (c < s ? c : s) - 1u; }

struct Splice {
	unsigned P;        // position of the insertion edge in the parent tree
	unsigned xv, sv, dv, szv;   // that position's word, unpacked
	unsigned a, b;     // ids of the new leaf edge and the new internal edge (E and E + 1)
};

XMP_HD Splice make_splice(unsigned P, unsigned word_at_P, unsigned E) {
	Splice s;
	s.P = P;
	s.xv = prog_edge(word_at_P);
	s.sv = prog_sib(word_at_P);
	s.dv = prog_depth(word_at_P);
	s.szv = prog_size(word_at_P);
	s.a = E;
	s.b = E + 1;
	return s;
}

// Parent position that step j of the child corresponds to (meaningless for the three spliced steps).
XMP_HD unsigned parent_index(unsigned j, unsigned P) { return j <= P ? j : j - 2u; }

// Parent position -> child position for every node other than the insertion edge itself.
XMP_HD unsigned shift_pos(unsigned q, unsigned P) { return q < P ? q : q + 2u; }

// The child's word at step j.  w = the parent's word at parent_index(j, P); anc = that parent position is
// a proper ancestor of P (its subtree contains P).
XMP_HD unsigned child_word(unsigned j, unsigned w, bool anc, const Splice &sp) {
	const unsigned P = sp.P;
	unsigned x = prog_edge(w), s = prog_sib(w), d = prog_depth(w), sz = prog_size(w);
	const unsigned i = parent_index(j, P);
	if (j > P + 2u && i < P + sp.szv) d += 1u;                    // inside xv's subtree: one level deeper
	if (s != XMP_NO_POS) s = (s == P) ? P : shift_pos(s, P);      // xv's sibling now hangs next to b, which sits at P
	if (anc) sz += 2u;
	if (j == P) {
		x = sp.b;
		s = sp.sv == XMP_NO_POS ? XMP_NO_POS : shift_pos(sp.sv, P);
		d = sp.dv;
		sz = sp.szv + 2u;
	}
	if (j == P + 1u) { x = sp.a; s = P + 2u; d = sp.dv + 1u; sz = 1u; }
	if (j == P + 2u) { x = sp.xv; s = P + 1u; d = sp.dv + 1u; sz = sp.szv; }
	return prog_word(x, s, d, sz);
}

// Program of a tree given as arrays over edge ids (parent edge or 0xFF, the two child edges or 0xFF) and its
// pre-order list.  pos[] and size[] are scratch arrays of E bytes.
template <typename Byte>
XMP_HD void prog_from_topology(const Byte *par, const Byte *k0, const Byte *k1, const Byte *pre, unsigned E,
                               unsigned *prog, unsigned char *pos, unsigned char *size) {
	for (unsigned i = 0; i < E; ++i) {
		pos[pre[i]] = (unsigned char) i;
		size[pre[i]] = 1;
	}
	for (unsigned i = E; i-- > 0;) {                               // children precede nothing they depend on: bottom-up
		const unsigned e = pre[i], p = par[e];
		if (p != 0xFFu) size[p] = (unsigned char) (size[p] + size[e]);
	}
	for (unsigned i = 0; i < E; ++i) {
		const unsigned e = pre[i], p = par[e];
		unsigned d = 0, s = XMP_NO_POS;
		if (p != 0xFFu) {
			d = prog_depth(prog[pos[p]]) + 1u;                     // parents precede children in pre-order
			s = pos[k0[p] == e ? k1[p] : k0[p]];
		}
		prog[i] = prog_word(e, s, d, size[e]);
	}
}

}  // namespace xmp
