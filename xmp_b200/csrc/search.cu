// Exact maximum-parsimony branch and bound on the device: a batched frontier in place of the
// reference's recursive per-edge loop.
//
// Reference behaviour being replaced (all in src/yanbader.c): BranchAndBound (:21-109) visits one
// partial tree; BandBScanEdge (:113-244) scores the next taxon against each of its 2m-3 edges and, for
// every edge whose cost fits under the bound, InsertNode + UpdateTreeAfterInsertion (:450-495) build the
// child tree's UP / DOWN / MIDPOINT sets and recurse; complete trees lower the bound and are recorded
// (:195-213).  Here the same search tree is explored a batch of partial trees at a time:
//
//   k_build_children_eager  one thread per (survivor, 128-bit column): walks the child in pre-order through the
//                       parent's walk program with the insertion spliced in (child_walk.cuh), recomputes UP along
//                       the root path and DOWN / MIDPOINT everywhere, scores the NEXT taxon on every MIDPOINT
//                       (AND / empty-nibble mask / POPC) and writes the child's UP rows, header, program and costs.
//   k_select            every (partial tree, edge) pair of a batch, 8 per thread: cost and length from the headers,
//                       the reference's signed acceptance test against the current bound, one survivor-list
//                       reservation per CTA; on the last level survivors are complete trees: atomicMin on the
//                       bound, a tree record.
//   k_expand_last_fused the same walk for parents with n-2 taxa, but each MIDPOINT is scored against the
//                       last taxon on the spot: the deepest level is never stored.
//   k_expand_score / k_build_children   the classic pair for state sets wider than 32 blocks: MIDPOINT rows are
//                       stored by the builder and scored by a separate pass (group-shuffle reduction per pair).
//
// Frontier layout in HBM: one ring-buffer pool per tree size m.  A node is a header (length, insertion
// path, walk program, and -- eager layout -- the insertion cost of the next taxon per position) plus a record
// of state-set rows indexed by pre-order position: [UP] (eager) or [MIDPOINT | UP] (classic); DOWN sets are
// transient.  Headers and rows are tiled so that a warp whose lanes work on different nodes at the same position
// touches whole 128-byte lines.  The scheduler (run_sweep) expands every level that holds a worthwhile batch in
// one pass, deepest first, with a single host synchronisation per pass; pool capacities bound the memory.
//
// Results are independent of batch order: ties are kept (<=), the bound only decreases, every tree
// recorded carries its length and the final filter keeps those equal to the final bound.
#include <algorithm>
#include <cmath>
#include <numeric>
#include <thread>
#include <stdlib.h>
#include <string.h>
#include "ctx.h"
#include "walk_prog.h"
#include "child_walk.cuh"
#include "tree_book.h"
#include "exchange_proto.h"

namespace xmp {

__device__ __forceinline__ unsigned path_byte(const unsigned *h, unsigned t) {
	return (h[(HDR_PATH_W + (t >> 2)) * 32u] >> (8u * (t & 3u))) & 0xFFu;
}

// Writes a tree record: its length, then the insertion path of the node with header h (one byte per taxon,
// copied a word at a time) with bytes t0 and t0 + 1 replaced by e0 and e1 (pass t0 + 1 >= 4 * words to skip e1).
__device__ __forceinline__ void emit_tree(uint8_t *rec, const Layout &L, const unsigned *h, unsigned length,
                                          unsigned t0, unsigned e0, unsigned e1, bool two) {
	unsigned *out = (unsigned *) rec;
	out[0] = length;
	const unsigned words = (L.n + 3u) >> 2;
	for (unsigned q = 0; q < words; ++q) {
		unsigned w = h[(HDR_PATH_W + q) * 32u];
		if (q == (t0 >> 2)) w = (w & ~(0xFFu << (8u * (t0 & 3u)))) | (e0 << (8u * (t0 & 3u)));
		if (two && q == ((t0 + 1u) >> 2)) w = (w & ~(0xFFu << (8u * ((t0 + 1u) & 3u)))) | (e1 << (8u * ((t0 + 1u) & 3u)));
		out[1 + q] = w;
	}
}

// A better complete tree was found: lower this GPU's bound word and, if that really lowered it, every peer's
// (system-scope atomics on peer-mapped memory, exchange.cu).  Replaces the reference's MSG_NEWBOUND relay through
// the boss (src/yanbader.c:200-211, src/mpiboss.c:162-171, src/mpiworker.c:19-41).  Improvements are rare -- a handful
// per search -- so the peer traffic is a few words; what matters is that no host is in the path.
__device__ __forceinline__ void lower_bound(unsigned *bound_ptr, unsigned v, const XchPeerTable *pt) {
	if (atomicMin(bound_ptr, v) > v && pt) {
		const unsigned world = pt->world, me = pt->rank;
		for (unsigned r = 0; r < world; ++r)
			if (r != me) atomicMin_system(pt->blk[r], v);          // word 0 of a block is its bound
	}
}

__device__ __forceinline__ unsigned path_hash(const unsigned *h, unsigned m, unsigned e) {
	unsigned x = 2166136261u;
	for (unsigned t = 3; t < m; ++t) x = (x ^ path_byte(h, t)) * 16777619u;
	x = (x ^ e) * 16777619u;
	return x ^ (x >> 15);
}

// ---- scoring + pruning of a batch of partial trees --------------------------------------------------
// Pairs are enumerated tile by tile, position by position, slot-minor: a warp reads whole 128-byte lines of
// MIDPOINT rows, and the survivors of one tile (<= 8 parents) stay together in the survivor list, so that
// the child kernels find their parents' data in L1.
template <int G, bool PLANAR>
__global__ void __launch_bounds__(256)
k_expand_score(PoolView pool, Layout L, unsigned m, unsigned long long first, unsigned long long n_nodes,
               const uint4 *__restrict__ taxon_row, WeightPlanes wp, unsigned rest, unsigned *bound_ptr, int last,
               uint4 *__restrict__ surv, unsigned *n_surv, uint8_t *__restrict__ trees, unsigned *n_trees,
               unsigned trees_cap, unsigned shard_count, unsigned shard_index, const XchPeerTable *peers) {
	__shared__ unsigned s_bound, s_warp[9];
	if (threadIdx.x == 0) s_bound = *(volatile unsigned *) bound_ptr;   // one read of the hot word per CTA
	const unsigned E = 2 * m - 3, ts = L.ts_shift, tmask = (1u << ts) - 1u;
	const unsigned long long gthread = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x;
	const unsigned long long pair = gthread / G;
	const unsigned sub = (unsigned) (gthread % G);
	const unsigned tile_pairs = E << ts;
	const unsigned long long tile = pair / tile_pairs;
	const unsigned r = (unsigned) (pair - tile * tile_pairs);
	const unsigned pos = r >> ts;
	const long long lin = (long long) ((tile << ts) + (r & tmask)) - (long long) (first & tmask);   // index inside the batch
	const bool valid = lin >= 0 && (unsigned long long) lin < n_nodes;
	unsigned long long node = first + (valid ? (unsigned long long) lin : 0ull);
	if (node >= pool.cap) node -= pool.cap;              // the batch may wrap around the end of the ring
	unsigned s = 0;
	__syncthreads();
	if (valid) {
		const uint4 *src = sets_of(pool, L, node) + ((size_t) pos << ts) * L.W;
		for (unsigned b = sub; b < L.W; b += G) s += score_block<PLANAR>(src[b], __ldg(taxon_row + b), b, wp);
	}
#pragma unroll
	for (int o = G / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);

	bool keep = false;
	unsigned new_score = 0;
	const unsigned *hdr = hdr_of(pool, L, node);
	if (valid && sub == 0) {
		const unsigned score = hdr[0];
		keep = accept_insertion(s, s_bound, score, rest);
		new_score = score + s;
		if (keep && shard_count > 1) keep = (path_hash(hdr, m, prog_edge(hdr[(L.off_prog + pos) * 32u])) % shard_count) == shard_index;
	}
	if (last) {
		if (keep && new_score < s_bound) lower_bound(bound_ptr, new_score, peers);
		unsigned slot = block_append(keep, n_trees, s_warp);
		if (keep && slot < trees_cap) {
			emit_tree(trees + (size_t) slot * L.trec, L, hdr, new_score, m, prog_edge(hdr[(L.off_prog + pos) * 32u]), 0u, false);
		}
	} else {
		unsigned slot = block_append(keep, n_surv, s_warp);
		if (keep) surv[slot] = make_uint4((unsigned) node, pos, new_score, 0u);
	}
}

// Eager layout: the costs are already in the headers.  Pairs are enumerated by header tile (32 slots), position,
// slot, so that a warp reads whole 128-byte lines of cost words.  A CTA's life is a chain of latencies -- the
// bound word, the header words, the atomic that reserves its place in the survivor list -- and with one pair
// per thread a launch is limited by how many such chains the GPU can have in flight (measured: 68 us for 11 M
// pairs, the same as the kernel that also read and scored the MIDPOINT sets).  So every thread takes K pairs,
// 256 apart, loads them all at once, and the CTA reserves list space once.
template <int K>
__global__ void __launch_bounds__(256)
k_select(PoolView pool, Layout L, unsigned m, unsigned long long first, unsigned long long n_nodes, unsigned rest,
         unsigned *bound_ptr, int last, uint4 *__restrict__ surv, unsigned *n_surv, uint8_t *__restrict__ trees, unsigned *n_trees,
         unsigned trees_cap, unsigned shard_count, unsigned shard_index, const XchPeerTable *peers) {
	__shared__ unsigned s_bound, s_warp[8], s_base;
	if (threadIdx.x == 0) s_bound = *(volatile unsigned *) bound_ptr;   // one read of the hot word per CTA
	const unsigned E = 2 * m - 3, tile_pairs = E * 32u;
	const unsigned long long base = (unsigned long long) blockIdx.x * (256u * K);
	const unsigned long long tile0 = base / tile_pairs;
	const unsigned r0 = (unsigned) (base - tile0 * tile_pairs) + threadIdx.x;
	unsigned node[K], pos[K], score[K], cost[K];
	bool valid[K];
#pragma unroll
	for (int i = 0; i < K; ++i) {
		const unsigned rr = r0 + 256u * i, q = rr / tile_pairs, r = rr - q * tile_pairs;
		pos[i] = r >> 5;
		const long long lin = (long long) (((tile0 + q) << 5) + (r & 31u)) - (long long) (first & 31ull);   // index inside the batch
		valid[i] = lin >= 0 && (unsigned long long) lin < n_nodes;
		unsigned long long nd = first + (valid[i] ? (unsigned long long) lin : 0ull);
		if (nd >= pool.cap) nd -= pool.cap;              // the batch may wrap around the end of the ring
		node[i] = (unsigned) nd;
		const unsigned *hdr = hdr_of(pool, L, nd);
		score[i] = valid[i] ? hdr[0] : 0u;
		cost[i] = valid[i] ? hdr[(L.off_cost + pos[i]) * 32u] : 0u;
	}
	__syncthreads();
	const unsigned bound = s_bound;
	unsigned mask = 0, best = 0xffffffffu;
#pragma unroll
	for (int i = 0; i < K; ++i) {
		bool keep = valid[i] && accept_insertion(cost[i], bound, score[i], rest);
		if (keep && shard_count > 1) {
			const unsigned *hdr = hdr_of(pool, L, node[i]);
			keep = (path_hash(hdr, m, prog_edge(hdr[(L.off_prog + pos[i]) * 32u])) % shard_count) == shard_index;
		}
		if (keep) {
			mask |= 1u << i;
			best = min(best, score[i] + cost[i]);
		}
	}
	if (last && best < bound) lower_bound(bound_ptr, best, peers);
	// place in the list: exclusive scan of the per-thread counts over the CTA, one atomic per CTA
	const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, cnt = (unsigned) __popc(mask);
	unsigned incl = cnt;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		const unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
		if ((int) lane >= o) incl += v;
	}
	if (lane == 31) s_warp[warp] = incl;
	__syncthreads();
	if (threadIdx.x == 0) {
		unsigned total = 0;
		for (unsigned w = 0; w < 8; ++w) {
			const unsigned c = s_warp[w];
			s_warp[w] = total;
			total += c;
		}
		s_base = total ? atomicAdd(last ? n_trees : n_surv, total) : 0u;
	}
	__syncthreads();
	unsigned slot = s_base + s_warp[warp] + incl - cnt;
#pragma unroll
	for (int i = 0; i < K; ++i) {
		if (!((mask >> i) & 1u)) continue;
		if (last) {
			if (slot < trees_cap) {
				const unsigned *hdr = hdr_of(pool, L, node[i]);
				emit_tree(trees + (size_t) slot * L.trec, L, hdr, score[i] + cost[i], m, prog_edge(hdr[(L.off_prog + pos[i]) * 32u]), 0u, false);
			}
		} else {
			surv[slot] = make_uint4(node[i], pos[i], score[i] + cost[i], 0u);
		}
		++slot;
	}
}

// ---- child construction ------------------------------------------------------------------------------
// One thread per (survivor, 128-bit column): at step j the UP and MIDPOINT sets of the child's position j go
// to row j of its record, the column-0 thread adds the program word -- consecutive threads build consecutive
// slots, so with the tiled layout every such store fills whole 128-byte lines.
// The two per-thread stacks hold D entries each; D comes from the deepest tree the level has produced so far
// (depth_out collects it for the next level), not from the worst case m + 1, so that deep levels (20-36 taxa)
// still fit four CTAs per SM.  (Compiling for six CTAs -- 80 registers -- was measured and is slower:
// profiles/r01_search_tuning.md.)
template <bool PLANAR>
__global__ void __launch_bounds__(128)
k_build_children(PoolView parent, PoolView child, Layout L, unsigned m, const uint4 *__restrict__ surv,
                 const unsigned *n_surv_ptr, unsigned s_begin, unsigned s_end, unsigned long long child_first,
                 const uint4 *__restrict__ rows, unsigned long long *merges, unsigned D, unsigned *depth_out, unsigned *overflow_flag) {
	extern __shared__ uint4 smem[];
	const unsigned T = blockDim.x;
	uint4 *upath = smem + threadIdx.x;
	uint4 *dstack = smem + (size_t) D * T + threadIdx.x;
	// survivors s_begin .. min(*n_surv_ptr, s_end) - 1 become children child_first, child_first + 1, ...
	const unsigned n_surv = min(*n_surv_ptr, s_end);
	const unsigned W = L.W, E = 2 * m - 3, Ec = E + 2, rs = W << L.ts_shift;
	const unsigned long long total = n_surv > s_begin ? (unsigned long long) (n_surv - s_begin) * W : 0ull;
	unsigned my_merges = 0, my_depth = 0;
	bool overflow = false;
	for (unsigned long long idx = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x; idx < total;
	     idx += (unsigned long long) gridDim.x * blockDim.x) {
		const unsigned k = (unsigned) (idx / W), col = (unsigned) (idx % W);
		const unsigned s = s_begin + k;
		unsigned long long cslot = child_first + k;
		if (cslot >= child.cap) cslot -= child.cap;
		const uint4 sv = surv[s];
		const unsigned pnode = sv.x, P = sv.y;
		const unsigned *ph = hdr_of(parent, L, pnode);
		const uint4 *PU = sets_of(parent, L, pnode) + (size_t) E * rs + col;      // classic layout only (launch_children)
		uint4 *CM = sets_of(child, L, cslot) + col;
		uint4 *CU = CM + (size_t) Ec * rs;                  // DOWN sets are not kept: no later step reads them
		unsigned *ch = hdr_of(child, L, cslot);
		unsigned *cprog = ch + L.off_prog * 32u;
		const uint4 r0 = __ldg(rows + col);
		ChildWalk cw;
		const unsigned n_up = cw.init<PLANAR>(ph, L, m, P, PU, __ldg(rows + (size_t) m * W + col), upath, T, D, overflow);
		// software-pipelined by one step: the UP sets of step j + 1 are requested before step j's arithmetic
		// (two steps ahead was measured and gains nothing: profiles/r01_search_tuning.md)
		unsigned wj, first_edge = 0;
		uint4 ux, us;
		unsigned pw1 = cw.word_for(1);                      // Ec >= 5
		cw.step(0, cw.word_for(0), wj, ux, us, upath, T, r0);
		for (unsigned j = 0; j < Ec; ++j) {
			const unsigned pw2 = cw.word_for(j + 2);        // clamped past the end
			unsigned wn = wj;
			uint4 uxn = ux, usn = us;
			if (j + 1 < Ec) cw.step(j + 1, pw1, wn, uxn, usn, upath, T, r0);
			pw1 = pw2;
			const unsigned d = prog_depth(wj);
			const uint4 dn = d == 0 ? r0 : merge_block<PLANAR>(dstack[(size_t) (d - 1 < D ? d - 1 : 0u) * T], us);   // d - 1 >= D: overflow is already flagged
			if (prog_size(wj) > 1u) {
				if (d < D) dstack[(size_t) d * T] = dn; else overflow = true;
			}
			my_depth = max(my_depth, d);
			CU[(size_t) j * rs] = ux;
			CM[(size_t) j * rs] = merge_block<PLANAR>(ux, dn);
			if (col == 0) {
				cprog[j * 32u] = wj;
				if (j == 0) first_edge = prog_edge(wj);
			}
			wj = wn; ux = uxn; us = usn;
		}
		if (col == 0) {
			ch[0] = sv.z;
			const unsigned pw = (L.n + 3u) >> 2;
			for (unsigned q = 0; q < pw; ++q) {
				unsigned w = ph[(HDR_PATH_W + q) * 32u];
				if (q == (m >> 2)) w = (w & ~(0xFFu << (8u * (m & 3u)))) | (cw.sp.xv << (8u * (m & 3u)));
				if (q == 0) w = (w & ~0xFFu) | first_edge;
				ch[(HDR_PATH_W + q) * 32u] = w;
			}
			my_merges += n_up + 2 * Ec;
		}
	}
	// statistics: one atomic per warp that did any work, not one per child
	my_merges = warp_sum(my_merges);
	if ((threadIdx.x & 31u) == 0 && my_merges) atomicAdd(merges, (unsigned long long) my_merges);
	my_depth = __reduce_max_sync(0xffffffffu, my_depth);
	if ((threadIdx.x & 31u) == 0 && my_depth > *(volatile unsigned *) depth_out) atomicMax(depth_out, my_depth);
	if (overflow) atomicOr(overflow_flag, 1u);
}

// Eager variant: the child's MIDPOINT sets are not stored; each is scored against the taxon the child will
// receive next, and the per-position costs go to the child's header.  A child's W columns sit in consecutive
// threads of ONE CTA (whole children per CTA), their costs meet in shared memory.
template <bool PLANAR>
#ifdef XMP_EAGER_MIN_CTAS            // A/B knob (scripts/build_variant.sh): register target as CTAs of 128 threads per SM
__global__ void __launch_bounds__(128, XMP_EAGER_MIN_CTAS)
#else
__global__ void __launch_bounds__(128)
#endif
k_build_children_eager(PoolView parent, PoolView child, Layout L, unsigned m, const uint4 *__restrict__ surv,
                       const unsigned *n_surv_ptr, unsigned s_begin, unsigned s_end, unsigned long long child_first,
                       const uint4 *__restrict__ rows, WeightPlanes wp, unsigned long long *merges, unsigned D,
                       unsigned *depth_out, unsigned *overflow_flag) {
	extern __shared__ uint4 smem[];
	const unsigned T = blockDim.x;
	uint4 *upath = nullptr;                                           // the root-path UP sets go through the child's record
	// the DOWN stack by 32-bit shared address: entry d of this thread at ds0 + d * SB
	const unsigned ds0 = (unsigned) __cvta_generic_to_shared(smem + threadIdx.x), SB = 16u * T;
	unsigned *acc = (unsigned *) (smem + (size_t) D * T);            // [cpc][Ec], used when W > 1
	const unsigned W = L.W, E = 2 * m - 3, Ec = E + 2, rs = W << L.ts_shift;
	const unsigned cpc = T / W;                                        // whole children per CTA (W <= 32 <= T)
	const unsigned n_surv = min(*n_surv_ptr, s_end);
	const unsigned n_children = n_surv > s_begin ? n_surv - s_begin : 0u;
	const unsigned lc = threadIdx.x / W, col = threadIdx.x - lc * W;
	const uint4 r0 = __ldg(rows + col);
	const uint4 rt = __ldg(rows + (size_t) m * W + col);
	const uint4 rn = __ldg(rows + (size_t) (m + 1) * W + col);         // the taxon the child is scored against
	const bool w1 = 4u * col >= wp.heavy_words;                        // this column holds weight-1 sites only
	unsigned my_merges = 0, my_depth = 0;
	bool overflow = false;
	for (unsigned base = blockIdx.x * cpc; base < n_children; base += gridDim.x * cpc) {
		const unsigned k = base + lc;
		const bool active = lc < cpc && k < n_children;
		if (W > 1) {
			for (unsigned x = threadIdx.x; x < cpc * Ec; x += T) acc[x] = 0;
			__syncthreads();
		}
		if (active) {
			const unsigned s = s_begin + k;
			unsigned long long cslot = child_first + k;
			if (cslot >= child.cap) cslot -= child.cap;
			const uint4 sv = surv[s];
			const unsigned pnode = sv.x, P = sv.y;
			const unsigned *ph = hdr_of(parent, L, pnode);
			const uint4 *PU = sets_of(parent, L, pnode) + col;             // eager layout: rows are the UP sets
			uint4 *CU = sets_of(child, L, cslot) + col;
			unsigned *ch = hdr_of(child, L, cslot);
			unsigned *cprog = ch + L.off_prog * 32u, *ccost = ch + L.off_cost * 32u;
			ChildWalk cw;
			const unsigned n_up = cw.init<PLANAR>(ph, L, m, P, PU, rt, upath, T, D, overflow, CU);
			unsigned wj, first_edge = 0;
			uint4 ux, us;
			unsigned pw1 = cw.word_for(1);                  // Ec >= 5
			cw.step(0, cw.word_for(0), wj, ux, us, upath, T, r0);
			uint4 *cu_at = CU;                              // row j of the child, its program and cost words: running pointers
			unsigned *prog_at = cprog, *cost_at = ccost;
			for (unsigned j = 0; j < Ec; ++j, cu_at += rs, prog_at += 32, cost_at += 32) {
				const unsigned pw2 = cw.word_for(j + 2);    // clamped past the end
				unsigned wn = wj;
				uint4 uxn = ux, usn = us;
				if (j + 1 < Ec) cw.step(j + 1, pw1, wn, uxn, usn, upath, T, r0);
				pw1 = pw2;
				const unsigned d = prog_depth(wj);
				uint4 dn = r0;
				if (d) dn = merge_block<PLANAR>(lds128(ds0 + (d - 1 < D ? d - 1 : 0u) * SB), us);   // d - 1 >= D: overflow is already flagged
				if (prog_size(wj) > 1u) {
					if (d < D) sts128(ds0 + d * SB, dn); else overflow = true;
				}
				my_depth = max(my_depth, d);
				*cu_at = ux;
				const uint4 mid = merge_block<PLANAR>(ux, dn);
				unsigned c;
				if (w1) c = score_block_w1<PLANAR>(mid, rn);
				else c = score_block<PLANAR>(mid, rn, col, wp);
				if (W == 1) *cost_at = c;
				else if (c) atomicAdd(&acc[lc * Ec + j], c);
				if (col == 0) {
					*prog_at = wj;
					if (j == 0) first_edge = prog_edge(wj);
				}
				wj = wn; ux = uxn; us = usn;
			}
			if (col == 0) {
				ch[0] = sv.z;
				const unsigned pw = (L.n + 3u) >> 2;
				for (unsigned q = 0; q < pw; ++q) {
					unsigned w = ph[(HDR_PATH_W + q) * 32u];
					if (q == (m >> 2)) w = (w & ~(0xFFu << (8u * (m & 3u)))) | (cw.sp.xv << (8u * (m & 3u)));
					if (q == 0) w = (w & ~0xFFu) | first_edge;
					ch[(HDR_PATH_W + q) * 32u] = w;
				}
				my_merges += n_up + 2 * Ec;
			}
		}
		if (W > 1) {
			__syncthreads();
			for (unsigned x = threadIdx.x; x < cpc * Ec; x += T) {
				const unsigned c_l = x / Ec, j = x - c_l * Ec, kk = base + c_l;
				if (kk < n_children) {
					unsigned long long cslot = child_first + kk;
					if (cslot >= child.cap) cslot -= child.cap;
					hdr_of(child, L, cslot)[(L.off_cost + j) * 32u] = acc[x];
				}
			}
			__syncthreads();
		}
	}
	my_merges = warp_sum(my_merges);
	if ((threadIdx.x & 31u) == 0 && my_merges) atomicAdd(merges, (unsigned long long) my_merges);
	my_depth = __reduce_max_sync(0xffffffffu, my_depth);
	if ((threadIdx.x & 31u) == 0 && my_depth > *(volatile unsigned *) depth_out) atomicMax(depth_out, my_depth);
	if (overflow) atomicOr(overflow_flag, 1u);
}

// ---- last two levels fused ---------------------------------------------------------------------------
// Parents have n-2 taxa.  A surviving child (n-1 taxa) is only ever scored once more, against the last
// taxon, so it is never written to HBM: G lanes per child (one per 128-bit column) walk it, form each
// edge's MIDPOINT on the fly, score the last taxon on it (group-shuffle sum over the columns), apply the
// acceptance test and emit complete trees.  This replaces k_build_children for the deepest pool and the
// whole k_expand_score pass over it -- where almost all nodes of a search live.
template <int G, bool PLANAR>
__global__ void __launch_bounds__(128)
k_expand_last_fused(PoolView parent, Layout L, unsigned m, const uint4 *__restrict__ surv, const unsigned *n_surv_ptr,
                    unsigned s_begin, unsigned s_end, const uint4 *__restrict__ rows, WeightPlanes wp, unsigned rest_last,
                    unsigned *bound_ptr, uint8_t *__restrict__ trees, unsigned *n_trees, unsigned trees_cap,
                    unsigned long long *merges, unsigned D, unsigned *overflow_flag, const XchPeerTable *peers) {
	extern __shared__ uint4 smem[];
	const unsigned T = blockDim.x;
	uint4 *upath = smem + threadIdx.x;
	// the DOWN stack by 32-bit shared address: entry d of this thread at ds0 + d * SB
	const unsigned ds0 = (unsigned) __cvta_generic_to_shared(smem + (size_t) D * T + threadIdx.x), SB = 16u * T;
	const unsigned W = L.W, E = 2 * m - 3, Ec = E + 2, rs = W << L.ts_shift;
	const unsigned n_surv = min(*n_surv_ptr, s_end);
	if (n_surv <= s_begin) return;
	const unsigned lane = threadIdx.x & 31u, sub = lane % G;
	const unsigned groups_per_block = T / G;
	const bool col_ok = sub < W;
	const unsigned col = col_ok ? sub : 0;
	const uint4 r0 = __ldg(rows + col);
	const uint4 rt = __ldg(rows + (size_t) m * W + col);
	const uint4 rl = __ldg(rows + (size_t) (m + 1) * W + col);
	const bool w1 = 4u * col >= wp.heavy_words;                 // this column holds weight-1 sites only
	unsigned my_merges = 0;
	bool overflow = false;
	// Every lane of a warp makes the same number of trips (the cost of a child is summed over its G lanes with
	// full-mask shuffles); lanes past the end of the list walk the last survivor again and emit nothing.
	const unsigned long long stride = (unsigned long long) gridDim.x * groups_per_block;
	for (unsigned long long s0 = (unsigned long long) s_begin + (unsigned long long) blockIdx.x * groups_per_block + (threadIdx.x & ~31u) / G;
	     s0 < n_surv; s0 += stride) {
		const unsigned long long s = s0 + lane / G;
		const bool live = s < n_surv;
		const uint4 sv = surv[live ? s : (unsigned long long) n_surv - 1ull];
		const unsigned pnode = sv.x, P = sv.y, score = sv.z;
		const unsigned *ph = hdr_of(parent, L, pnode);
		const uint4 *PU = sets_of(parent, L, pnode) + (size_t) up_row0(L, E) * rs + col;
		ChildWalk cw;
		const unsigned n_up = cw.init<PLANAR>(ph, L, m, P, PU, rt, upath, T, D, overflow);
		// the bound is a hot word: one lane per warp reads it, the others get it by shuffle
		unsigned bound = 0;
		if (lane == 0) bound = *(volatile unsigned *) bound_ptr;
		bound = __shfl_sync(0xffffffffu, bound, 0);
		// software-pipelined by one step: the UP sets of step j + 1 are requested before step j's arithmetic
		unsigned depth, depth_n = 0;
		bool internal, internal_n = false;
		uint4 ux, us, uxn, usn;
		unsigned pw1 = cw.word_for(1);                      // Ec >= 5
		cw.step_lean(0, cw.word_for(0), depth, internal, ux, us, upath, T, r0);
		for (unsigned j = 0; j < Ec; ++j) {
			const unsigned pw2 = cw.word_for(j + 2);        // clamped past the end
			if (j + 1 < Ec) cw.step_lean(j + 1, pw1, depth_n, internal_n, uxn, usn, upath, T, r0);
			pw1 = pw2;
			uint4 dn = r0;
			if (depth) dn = merge_block<PLANAR>(lds128(ds0 + (depth - 1 < D ? depth - 1 : 0u) * SB), us);   // depth - 1 >= D: overflow is already flagged
			if (internal) {
				if (depth < D) sts128(ds0 + depth * SB, dn); else overflow = true;
			}
			const uint4 mid = merge_block<PLANAR>(ux, dn);
			unsigned c;
			if (w1) c = score_block_w1<PLANAR>(mid, rl);
			else c = score_block<PLANAR>(mid, rl, col, wp);
			if (!col_ok) c = 0;
#pragma unroll
			for (int o = G / 2; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
			if (live && sub == 0 && accept_insertion(c, bound, score, rest_last)) {
				if (score + c < bound) lower_bound(bound_ptr, score + c, peers);
				const unsigned slot = atomicAdd(n_trees, 1u);
				if (slot < trees_cap) {
					emit_tree(trees + (size_t) slot * L.trec, L, ph, score + c, m, cw.sp.xv, cw.edge_at(j), true);
				}
			}
			depth = depth_n; internal = internal_n; ux = uxn; us = usn;
		}
		if (live && sub == 0) my_merges += n_up + 2 * Ec;
	}
	my_merges = warp_sum(my_merges);
	if (lane == 0 && my_merges) atomicAdd(merges, (unsigned long long) my_merges);
	if (overflow) atomicOr(overflow_flag, 1u);
}

// Headers for nodes whose sets were built from topologies (the root job, imported jobs).
__global__ void k_write_headers(PoolView pool, Layout L, unsigned m, unsigned long long first, unsigned n_jobs,
                                const uint8_t *__restrict__ topo, unsigned topo_stride, const uint8_t *__restrict__ paths,
                                const unsigned *__restrict__ tree_len) {
	const unsigned j = blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= n_jobs) return;
	const unsigned E = 2 * m - 3;
	unsigned *h = hdr_of(pool, L, first + j);
	const uint8_t *t = topo + (size_t) j * 4 * topo_stride;
	const uint8_t *par = t, *k0 = t + topo_stride, *k1 = t + 2 * topo_stride, *pre = t + 3 * topo_stride;
	unsigned prog[2 * XMP_MAX_TAXA];
	unsigned char pos[2 * XMP_MAX_TAXA], size[2 * XMP_MAX_TAXA];
	prog_from_topology(par, k0, k1, pre, E, prog, pos, size);
	h[0] = tree_len[j];
	const unsigned pw = (L.n + 3u) >> 2;
	for (unsigned q = 0; q < pw; ++q) {
		unsigned w = 0;
		for (unsigned b = 0; b < 4; ++b) {
			const unsigned k = 4 * q + b;
			unsigned byte = k < L.n ? paths[(size_t) j * L.n + k] : 0u;
			if (k == 0) byte = pre[0];                // first edge in pre-order = the root's child
			w |= byte << (8u * b);
		}
		h[(HDR_PATH_W + q) * 32u] = w;
	}
	for (unsigned i = 0; i < E; ++i) h[(L.off_prog + i) * 32u] = prog[i];
}

// Sets built from topologies (k_build_sets: plain records indexed by edge id, MIDPOINT then UP) -> the pool's
// tiled, position-indexed rows.  One thread per (job, position, column).
__global__ void k_scatter_sets(PoolView pool, Layout L, unsigned m, unsigned long long first, unsigned n_jobs,
                               const uint8_t *__restrict__ topo, unsigned topo_stride, const uint4 *__restrict__ src) {
	const unsigned E = 2 * m - 3, W = L.W;
	const unsigned long long idx = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= (unsigned long long) n_jobs * E * W) return;
	const unsigned col = (unsigned) (idx % W);
	const unsigned i = (unsigned) ((idx / W) % E);
	const unsigned j = (unsigned) (idx / ((unsigned long long) W * E));
	const unsigned e = topo[(size_t) j * 4 * topo_stride + 3 * topo_stride + i];
	const uint4 *rec = src + (size_t) j * 2 * E * W;
	uint4 *dst = sets_of(pool, L, first + j) + col;
	const size_t rs = (size_t) W << L.ts_shift;
	if (L.off_cost) {
		dst[(size_t) i * rs] = rec[((size_t) E + e) * W + col];
	} else {
		dst[(size_t) i * rs] = rec[(size_t) e * W + col];
		dst[(size_t) (E + i) * rs] = rec[((size_t) E + e) * W + col];
	}
}

// Eager layout: cost of inserting taxon m at every position of the imported nodes, from their staged MIDPOINT sets.
template <bool PLANAR>
__global__ void k_stage_costs(PoolView pool, Layout L, unsigned m, unsigned long long first, unsigned n_jobs,
                              const uint8_t *__restrict__ topo, unsigned topo_stride, const uint4 *__restrict__ src,
                              const uint4 *__restrict__ taxon_row, WeightPlanes wp) {
	const unsigned E = 2 * m - 3, W = L.W;
	const unsigned long long idx = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= (unsigned long long) n_jobs * E) return;
	const unsigned i = (unsigned) (idx % E), j = (unsigned) (idx / E);
	const unsigned e = topo[(size_t) j * 4 * topo_stride + 3 * topo_stride + i];
	const uint4 *mid = src + ((size_t) j * 2 * E + e) * W;
	unsigned c = 0;
	for (unsigned b = 0; b < W; ++b) c += score_block<PLANAR>(mid[b], __ldg(taxon_row + b), b, wp);
	hdr_of(pool, L, first + j)[(L.off_cost + i) * 32u] = c;
}

// Insertion paths of nodes first .. first + n - 1 as plain byte records (job descriptors for export).
__global__ void k_read_paths(PoolView pool, Layout L, unsigned m, unsigned long long first, unsigned n, uint8_t *__restrict__ out) {
	const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const unsigned *h = hdr_of(pool, L, first + i);
	uint8_t *job = out + (size_t) i * L.n;
	job[0] = (uint8_t) m;                        // a job descriptor: byte 0 = tree size, byte t = path[t] for 3 <= t < m
	job[1] = job[2] = 0;
	for (unsigned t = 3; t < L.n; ++t) job[t] = (uint8_t) path_byte(h, t);
}

__global__ void k_filter_trees(const uint8_t *__restrict__ src, unsigned n_src, unsigned trec, const unsigned *bound_ptr,
                               uint8_t *__restrict__ dst, unsigned *n_dst) {
	const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
	bool keep = false;
	if (i < n_src) keep = *(const unsigned *) (src + (size_t) i * trec) <= *bound_ptr;
	unsigned slot = warp_append(keep, n_dst);
	if (keep) {
		const unsigned *s = (const unsigned *) (src + (size_t) i * trec);
		unsigned *d = (unsigned *) (dst + (size_t) slot * trec);
		for (unsigned k = 0; k < trec / 4; ++k) d[k] = s[k];
	}
}

__global__ void k_lower_bound(unsigned *bound_ptr, unsigned value) { atomicMin(bound_ptr, value); }

// The alignment rows in planar form (fitch_dev.cuh), for a search run with XMP_SEARCH_PLANAR=1.
__global__ void k_rows_to_planar(const uint4 *__restrict__ in, uint4 *__restrict__ out, unsigned long long n) {
	const unsigned long long i = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) out[i] = to_planar(in[i]);
}

// The beam of the greedy dive: keeps the `keep` cheapest entries of a survivor list (by the length of the child tree,
// surv[i].z), on the device -- one CTA, a histogram of lengths above the minimum, a threshold, a compaction.  Ties at
// the threshold are taken in arrival order.  (This used to be a device-to-host copy, a host sort and a copy back per
// tree size: most of the ~10 ms a search spent in the dive.)
__global__ void __launch_bounds__(1024)
k_beam_select(const uint4 *__restrict__ surv, const unsigned *n_ptr, unsigned keep, uint4 *__restrict__ out, unsigned *n_out) {
	__shared__ unsigned hist[2048];
	__shared__ unsigned s_min, s_T, s_quota, s_count, s_ties;
	const unsigned n = *n_ptr, tid = threadIdx.x;
	if (tid == 0) { s_min = 0xffffffffu; s_count = 0; s_ties = 0; }
	for (unsigned i = tid; i < 2048; i += 1024) hist[i] = 0;
	__syncthreads();
	unsigned lo = 0xffffffffu;
	for (unsigned i = tid; i < n; i += 1024) lo = min(lo, surv[i].z);
	lo = __reduce_min_sync(0xffffffffu, lo);
	if ((tid & 31u) == 0) atomicMin(&s_min, lo);
	__syncthreads();
	const unsigned smin = s_min;
	for (unsigned i = tid; i < n; i += 1024) atomicAdd(&hist[min(surv[i].z - smin, 2047u)], 1u);
	__syncthreads();
	if (tid == 0) {
		unsigned cum = 0, T = 2048;
		for (unsigned b = 0; b < 2048; ++b) {
			if (cum + hist[b] >= keep) { T = b; break; }
			cum += hist[b];
		}
		s_T = T;
		s_quota = keep - cum;                    // entries at the threshold that still fit
	}
	__syncthreads();
	const unsigned T = s_T, quota = s_quota;
	for (unsigned i = tid; i < n; i += 1024) {
		const uint4 e = surv[i];
		const unsigned bin = min(e.z - smin, 2047u);
		bool take = bin < T;
		if (bin == T) take = atomicAdd(&s_ties, 1u) < quota;
		if (take) out[atomicAdd(&s_count, 1u)] = e;
	}
	__syncthreads();
	if (tid == 0) *n_out = s_count;
}

// ---- host side ----------------------------------------------------------------------------------------
struct Pool {
	unsigned *hdr = nullptr;
	uint4 *sets = nullptr;
	unsigned long long cap = 0, count = 0, rec = 0;   // rec = uint4 per node
	unsigned long long head = 0;          // ring: open nodes occupy slots head .. head + count - 1 (modulo cap)
	unsigned E = 0, rows = 0;             // rows per node: 2E classic, E eager
	PoolView view() const { return PoolView{hdr, sets, rows, cap}; }
	unsigned long long slot(unsigned long long i) const { return (head + i) % cap; }
};

struct Search {
	xmp_search_params prm;
	Layout L;
	unsigned n = 0, split_level = 0;
	Pool pool[XMP_MAX_TAXA + 2];
	unsigned *d_bound = nullptr, *d_nsurv = nullptr, *d_ntrees = nullptr, *d_len = nullptr;
	const XchPeerTable *d_peers = nullptr;   // non-null when this search is one rank of a multi-GPU search (exchange.cu)
	unsigned steals_served = 0, steals_got = 0, open_published = 0xffffffffu;
	unsigned long long *d_merges = nullptr;
	uint4 *d_surv = nullptr;
	unsigned long long surv_cap = 0;
	uint4 *d_surv_lvl[XMP_MAX_TAXA + 2] = {nullptr};            // per-level survivor lists used by the sweep
	unsigned long long surv_cap_lvl[XMP_MAX_TAXA + 2] = {0};
	unsigned *d_nsurv_lvl = nullptr;                           // [XMP_MAX_TAXA + 1] counters, one per level
	bool sweep = true;
	unsigned sweep_fallback = 1;
	unsigned sweep_levels = 6;            // levels per sweep while connected to peers (XMP_SWEEP_LEVELS)
	// A level whose batch produced more children than the next pool could take keeps the batch at the head
	// of its ring and builds the remaining survivors in later sweeps.
	struct Pending { bool active = false; unsigned long long chunk = 0; unsigned head = 0, ns = 0; };
	Pending pend[XMP_MAX_TAXA + 2];
	double rate[XMP_MAX_TAXA + 2];        // survivors per scored pair seen so far, per level (sizes the batches)
	uint8_t *d_trees[2] = {nullptr, nullptr};
	int cur = 0;
	unsigned trees_cap = 0, ntrees_dev = 0;
	unsigned *h_pin = nullptr;            // pinned copy of the device scalars: [0] bound, [1] n_surv, [2] n_trees
	TreeBook book;                        // records (score + path) that left the device, each <= the bound at that time (tree_book.h)
	unsigned bound = 0x7fffffffu;
	bool done = false, finished = false;
	bool fused_last = false;      // the deepest pool is never materialised (k_expand_last_fused)
	bool eager = false;           // nodes carry their insertion costs instead of MIDPOINT sets (Layout::off_cost)
	unsigned min_batch = 32768, beam = 4096;
	unsigned maxdepth[XMP_MAX_TAXA + 2] = {0};   // deepest tree seen so far per tree size (sizes the walk kernels' stacks)
	unsigned *d_maxdepth = nullptr, *d_overflow = nullptr;
	xmp_search_stats st;
	Scratch topo, pathbuf, stage;
	// XMP_SEARCH_PLANAR=1 (an A/B switch, off by default): every state set of the search -- the rows below, the pools, the staged
	// sets of imported subtrees -- is kept in planar form (fitch_dev.cuh).  The form is closed under merging and preserves costs
	// (tests/child_walk_check.cu), and nothing but paths and lengths ever leaves a search, so results cannot depend on it.
	bool planar = false;
	const uint4 *rows = nullptr;          // the alignment rows the search's kernels read: ctx->d_rows, or their planar copy
	Scratch prows;
};

// Scheduler knobs (defaults chosen on B200, see profiles/r01_search_tuning.md; overridable for tuning):
// XMP_MIN_BATCH  open trees a level must offer to join a sweep, XMP_BEAM  width of the initial greedy dive,
// XMP_SWEEP=0  one level per synchronisation, XMP_SWEEP_FALLBACK=0  every level goes when none qualifies.
static unsigned env_unsigned(const char *name, unsigned dflt) {
	const char *v = getenv(name);
	if (!v || !*v) return dflt;
	char *end = nullptr;
	unsigned long x = strtoul(v, &end, 10);
	return end != v ? (unsigned) x : dflt;
}

// Counters of one expanded batch: `nodes` partial trees on m taxa, each scored on all of its 2m-3 edges
// (the reference's branchAndBoundCount / branchAndBoundScanEdgeCountByDepth[m], src/measure.c:8-10).
static void count_batch(xmp_search_stats &st, unsigned m, unsigned long long nodes) {
	st.nodes += nodes;
	st.score_calls += nodes * (2ull * m - 3);
	st.nodes_by_depth[m] += nodes;
	st.edge_evals_by_depth[m] += nodes * (2ull * m - 3);
	st.batches_by_depth[m] += 1;
}

static double odd_factorial(unsigned m) {      // number of distinct trees on m taxa: (2m-5)!!
	double t = 1;
	for (unsigned k = 4; k <= m; ++k) t *= (double) (2 * k - 5);
	return t;
}

void search_destroy(xmp_ctx *ctx) {
	Search *S = ctx->search;
	if (!S) return;
	cudaSetDevice(ctx->device);
	cudaStreamSynchronize(ctx->stream);
	// device buffers live in ctx->arena, the pinned scalars in ctx->h_pin: both outlive the search
	S->topo.release();
	S->pathbuf.release();
	S->stage.release();
	S->prows.release();
	delete S;
	ctx->search = nullptr;
}

static int alloc_search(xmp_ctx *ctx, Search *S, unsigned W) {
	const unsigned n = S->n, Emax = 2 * n - 3;
	Layout &L = S->L;
	L.n = n;
	L.W = W;
	L.off_prog = HDR_PATH_W + ((n + 3u) >> 2);
	S->eager = !(S->prm.flags & XMP_SEARCH_NO_EAGER) && W <= 32;
	L.off_cost = S->eager ? L.off_prog + Emax : 0u;
	L.NW = L.off_prog + Emax + (S->eager ? Emax : 0u);
	L.ts_shift = W == 1 ? 3u : (W == 2 ? 2u : (W <= 4 ? 1u : 0u));     // one 128-byte line per row of a tile
	L.trec = 4 + ((n + 3u) & ~3u);

	size_t free_b = 0, total_b = 0;
	XMP_CUDA(cudaMemGetInfo(&free_b, &total_b));
	// Default frontier budget: deep searches (many tree sizes in flight at once) profit from large pools --
	// batches stay big -- while allocating tens of GiB costs tens of milliseconds, which short searches on
	// few taxa would notice (profiles/r01_search_tuning.md).
	const size_t dflt = n >= 17 ? ((size_t) 48 << 30) : ((size_t) 12 << 30);
	// free_b does not count an arena this context already holds
	size_t budget = S->prm.mem_budget ? S->prm.mem_budget : std::min<size_t>((free_b + ctx->arena_bytes) / 2, dflt);
	S->trees_cap = budget >= ((size_t) 8 << 30) ? 1u << 22 : 1u << 20;      // records the device holds between drains
	if (budget < 2 * (size_t) S->trees_cap * L.trec + ((size_t) 128 << 20)) return set_error(XMP_ERR_NOMEM, "memory budget of %zu bytes is too small for the search", budget);

	// The deepest pool is skipped when the last two levels are fused (narrow state sets): it then only
	// ever holds imported jobs.
	S->fused_last = !(S->prm.flags & XMP_SEARCH_NO_FUSE) && W <= 32 && n >= 5;
	const double last_cap = S->fused_last ? 1024.0 : 1e300;
	// Everything the search keeps on the device is carved out of ONE allocation owned by the context
	// (ctx->arena): a search costs one cudaMalloc instead of ~2n, and none at all when the arena was
	// reserved beforehand (xmp_cuda_search_reserve) or is left over from an earlier search.
	// layout(C, base) places every buffer for per-level capacity C and returns the bytes used; with
	// base == nullptr it only measures.  A level never needs more than (2m-5)!! nodes.
	auto layout = [&](unsigned long long C, char *base) -> size_t {
		size_t off = 0;
		auto take = [&](size_t bytes) -> void * {
			off = (off + 255) & ~(size_t) 255;
			void *ptr = base ? base + off : nullptr;
			off += bytes;
			return ptr;
		};
		for (unsigned m = 3; m < n; ++m) {
			Pool &P = S->pool[m];
			P.E = 2 * m - 3;
			P.rows = S->eager ? P.E : 2 * P.E;
			P.rec = (unsigned long long) P.rows * W;       // classic: [MIDPOINT | UP][edge][block]; eager: [UP][edge][block]
			P.cap = (unsigned long long) std::min((double) C, odd_factorial(m));
			if (m + 1 == n) P.cap = (unsigned long long) std::min((double) P.cap, last_cap);
			P.cap = (P.cap + 31ull) & ~31ull;                                 // whole tiles (headers: 32 slots, sets: <= 8)
			P.count = 0;
			P.hdr = (unsigned *) take((size_t) P.cap * L.NW * 4);
			P.sets = (uint4 *) take((size_t) P.cap * P.rec * 16);
		}
		S->surv_cap = std::max<unsigned long long>(C, (unsigned long long) S->beam * Emax);
		S->d_surv = (uint4 *) take((size_t) S->surv_cap * 16);
		S->d_trees[0] = (uint8_t *) take((size_t) S->trees_cap * L.trec);
		S->d_trees[1] = (uint8_t *) take((size_t) S->trees_cap * L.trec);
		// one survivor list per level: as many entries as the pool its children go to (the fused level
		// keeps them only until the fused kernel has consumed them)
		for (unsigned m = 3; m + 1 < n; ++m) {
			S->surv_cap_lvl[m] = (S->fused_last && m + 2 == n) ? 4 * S->surv_cap : 4 * S->pool[m + 1].cap;
			S->d_surv_lvl[m] = (uint4 *) take((size_t) S->surv_cap_lvl[m] * 16);
		}
		S->d_len = (unsigned *) take(65536 * sizeof(unsigned));
		return off + 256;
	};
	unsigned long long lo = 256, hi = 16777216ull;
	if (layout(lo, nullptr) > budget) return set_error(XMP_ERR_NOMEM, "memory budget of %zu bytes cannot hold even a minimal frontier", budget);
	while (hi - lo > 1) {
		unsigned long long mid = (lo + hi) / 2;
		if (layout(mid, nullptr) <= budget) lo = mid; else hi = mid;
	}
	const size_t need = layout(lo, nullptr);
	if (ctx->arena_bytes < need) {
		if (ctx->arena) cudaFree(ctx->arena);
		ctx->arena = nullptr;
		ctx->arena_bytes = 0;
		XMP_CUDA(cudaMalloc(&ctx->arena, need));
		ctx->arena_bytes = need;
	}
	layout(lo, (char *) ctx->arena);
	// The counter block (bound, survivor and tree counters, per-level depths) lives in the context's exchange block,
	// where peer GPUs can reach the bound word.  Its protocol words (bound, steal request / reply, busy counter) belong
	// to the exchange once ranks are connected: peers may already have written them.
	int xrc = xch_ensure_block(ctx);
	if (xrc) return xrc;
	S->d_bound = ctx->xch;
	S->d_peers = ctx->xch_connected ? (const XchPeerTable *) ((const char *) ctx->xch + XCH_TABLE_OFF) : nullptr;
	S->d_nsurv = S->d_bound + 1;
	S->d_ntrees = S->d_bound + 2;
	S->d_merges = (unsigned long long *) (S->d_bound + 4);
	S->d_nsurv_lvl = S->d_bound + 8;
	S->d_overflow = S->d_bound + 6;
	S->d_maxdepth = S->d_bound + 128;          // [XMP_MAX_TAXA + 2]
	if (ctx->xch_connected) {
		XMP_CUDA(cudaMemsetAsync(S->d_bound + 1, 0, (XW_FIRST - 1) * sizeof(unsigned), ctx->stream));
		XMP_CUDA(cudaMemsetAsync(S->d_bound + XW_FIRST + XW_COUNT, 0, (256 - XW_FIRST - XW_COUNT) * sizeof(unsigned), ctx->stream));
	} else {
		XMP_CUDA(cudaMemsetAsync(S->d_bound, 0, 1024, ctx->stream));
	}
	if (!ctx->h_pin) XMP_CUDA(cudaMallocHost(&ctx->h_pin, 1024));
	S->h_pin = ctx->h_pin;
	return XMP_OK;
}

// Creates open nodes from insertion paths (job descriptors): the thief-side "fast-forward" of the
// reference (src/yanbader.c:95-99,136-143), done as one batched set construction per tree size.
// jobs: n_jobs records of L.n bytes, byte 0 = tree size m, byte t (3 <= t < m) = path[t].
static int import_nodes(xmp_ctx *ctx, Search *S, const uint8_t *jobs, unsigned n_jobs) {
	const Layout &L = S->L;
	for (unsigned m = 3; m < S->n; ++m) {
		std::vector<uint8_t> sel;
		for (unsigned j = 0; j < n_jobs; ++j) {
			if (jobs[(size_t) j * L.n] == m) sel.insert(sel.end(), jobs + (size_t) j * L.n, jobs + (size_t) (j + 1) * L.n);
		}
		unsigned k = (unsigned) (sel.size() / L.n);
		size_t done = 0;
		while (done < k) {
			Pool &P = S->pool[m];
			// at most 65536 nodes and 256 MiB of staging per round
			unsigned part = (unsigned) std::min<size_t>(k - done, std::max<size_t>(1, std::min<size_t>(65536, ((size_t) 256 << 20) / ((size_t) 2 * P.E * L.W * 16))));
			if (P.count + part > P.cap) return set_error(XMP_ERR_NOMEM, "frontier pool for %u taxa is full (%llu nodes)", m, P.cap);
			const unsigned long long tail = P.slot(P.count);
			part = (unsigned) std::min<unsigned long long>(part, P.cap - tail);      // up to the end of the ring, the rest next time
			const unsigned E = 2 * m - 3, stride = (E + 15u) & ~15u;
			std::vector<uint8_t> h((size_t) part * 4 * stride, 0xFF);
			HostTopo t;
			for (unsigned i = 0; i < part; ++i) {
				const uint8_t *path = sel.data() + (done + i) * L.n;
				if (!t.from_path(path, m)) return set_error(XMP_ERR_ARG, "job %u names an edge that does not exist", i);
				uint8_t *row = h.data() + (size_t) i * 4 * stride;
				memcpy(row, t.par, E);
				memcpy(row + stride, t.kid[0], E);
				memcpy(row + 2 * stride, t.kid[1], E);
				t.preorder(row + 3 * stride);
				// depth of this tree (parents precede children in pre-order): the walk kernels size their stacks by it
				uint8_t depth[2 * XMP_MAX_TAXA];
				for (unsigned q = 0; q < E; ++q) {
					const unsigned e = row[3 * stride + q], pe = t.par[e];
					depth[e] = pe == 0xFFu ? 0 : (uint8_t) (depth[pe] + 1);
					S->maxdepth[m] = std::max<unsigned>(S->maxdepth[m], depth[e]);
				}
			}
			int rc = S->topo.reserve(h.size());
			if (rc) return rc;
			rc = S->pathbuf.reserve((size_t) part * L.n);
			if (rc) return rc;
			XMP_CUDA(cudaMemcpyAsync(S->topo.ptr, h.data(), h.size(), cudaMemcpyHostToDevice, ctx->stream));
			XMP_CUDA(cudaMemcpyAsync(S->pathbuf.ptr, sel.data() + done * L.n, (size_t) part * L.n, cudaMemcpyHostToDevice, ctx->stream));
			rc = S->stage.reserve((size_t) part * 2 * P.E * L.W * 16);
			if (rc) return rc;
			rc = launch_build_sets(ctx, (const uint8_t *) S->topo.ptr, stride, part, m, (uint4 *) S->stage.ptr, (size_t) 2 * P.E * L.W, true, S->d_len,
			                       S->planar ? S->rows : nullptr);
			if (rc) return rc;
			{
				const unsigned long long threads = (unsigned long long) part * E * L.W;
				k_scatter_sets<<<(unsigned) ((threads + 255) / 256), 256, 0, ctx->stream>>>(P.view(), L, m, tail, part, (const uint8_t *) S->topo.ptr,
				                                                                          stride, (const uint4 *) S->stage.ptr);
				XMP_CUDA(cudaGetLastError());
			}
			k_write_headers<<<(part + 127) / 128, 128, 0, ctx->stream>>>(P.view(), L, m, tail, part, (const uint8_t *) S->topo.ptr, stride,
			                                                            (const uint8_t *) S->pathbuf.ptr, S->d_len);
			XMP_CUDA(cudaGetLastError());
			if (S->eager) {
				const unsigned long long threads = (unsigned long long) part * E;
				if (S->planar)
					k_stage_costs<true><<<(unsigned) ((threads + 255) / 256), 256, 0, ctx->stream>>>(P.view(), L, m, tail, part, (const uint8_t *) S->topo.ptr, stride,
					                                                                               (const uint4 *) S->stage.ptr, S->rows + (size_t) m * L.W, ctx->wp);
				else
					k_stage_costs<false><<<(unsigned) ((threads + 255) / 256), 256, 0, ctx->stream>>>(P.view(), L, m, tail, part, (const uint8_t *) S->topo.ptr, stride,
					                                                                                (const uint4 *) S->stage.ptr, S->rows + (size_t) m * L.W, ctx->wp);
				XMP_CUDA(cudaGetLastError());
			}
			XMP_CUDA(cudaStreamSynchronize(ctx->stream));
			S->st.launches += 3;
			P.count += part;
			done += part;
		}
	}
	return XMP_OK;
}

static int launch_expand(xmp_ctx *ctx, Search *S, unsigned m, unsigned long long first, unsigned long long chunk, bool shard,
                         uint4 *surv, unsigned *nsurv) {
	const Layout &L = S->L;
	Pool &P = S->pool[m];
	const int last = (m + 1 == S->n);
	const uint4 *row = S->rows + (size_t) m * L.W;
	const unsigned rest = ctx->rest_bound[m];
	const unsigned sc = shard ? S->prm.shard_count : 1u, si = shard ? S->prm.shard_index : 0u;
	// whole tiles of the pool's record layout are enumerated; slots outside the batch are masked in the kernel
	const unsigned long long tiles = ((first & ((1ull << L.ts_shift) - 1ull)) + chunk + (1ull << L.ts_shift) - 1ull) >> L.ts_shift;
	const unsigned long long pairs = (tiles << L.ts_shift) * P.E;
	uint8_t *trees = S->d_trees[S->cur];   // slots are absolute: d_ntrees already counts the records present
	const unsigned room = S->trees_cap;
	if (S->eager) {
		const unsigned long long tiles32 = ((first & 31ull) + chunk + 31ull) >> 5;
		const unsigned long long pairs32 = tiles32 * 32ull * P.E;
		k_select<8><<<(unsigned) ((pairs32 + 2047) / 2048), 256, 0, ctx->stream>>>(P.view(), L, m, first, chunk, rest, S->d_bound, last, surv, nsurv, trees,
		                                                                        S->d_ntrees, room, sc, si, S->d_peers);
		XMP_CUDA(cudaGetLastError());
		S->st.launches += 1;
		return XMP_OK;
	}
#define XMP_EXPAND_F(G, F)                                                                                                 \
	k_expand_score<G, F><<<(unsigned) ((pairs * G + 255) / 256), 256, 0, ctx->stream>>>(P.view(), L, m, first, chunk, row, ctx->wp, rest, \
	                                                                                 S->d_bound, last, surv, nsurv, trees,             \
	                                                                                 S->d_ntrees, room, sc, si, S->d_peers)
#define XMP_EXPAND(G)                    \
	do {                                 \
		if (S->planar) XMP_EXPAND_F(G, true); \
		else XMP_EXPAND_F(G, false);     \
	} while (0)
	const unsigned W = L.W;
	if (W > 16) XMP_EXPAND(32);
	else if (W > 8) XMP_EXPAND(16);
	else if (W > 4) XMP_EXPAND(8);
	else if (W > 2) XMP_EXPAND(4);
	else if (W > 1) XMP_EXPAND(2);
	else XMP_EXPAND(1);
#undef XMP_EXPAND
#undef XMP_EXPAND_F
	XMP_CUDA(cudaGetLastError());
	S->st.launches += 1;
	return XMP_OK;
}

// Shared memory of the child-walk kernels: two stacks of D blocks per thread (one for the eager builder, whose
// root-path UP sets live in the child's record).  A child of a tree of depth d is
// at most d + 1 deep; its stacks are indexed by depth, so D = (deepest parent so far) + 2, capped by the worst
// case m + 1 (a caterpillar).
static unsigned walk_threads(unsigned D, size_t *smem, unsigned stacks = 2) {
	unsigned threads = 128;
	while (threads > 32 && (size_t) stacks * D * 16 * threads > ((size_t) 96 << 10)) threads >>= 1;
	*smem = (size_t) stacks * D * 16 * threads;
	return threads;
}

static int launch_children(xmp_ctx *ctx, Search *S, unsigned m, const uint4 *surv, const unsigned *nsurv, unsigned long long child_first,
                           unsigned s_begin = 0, unsigned s_end = 0xffffffffu) {
	Pool &P = S->pool[m], &C = S->pool[m + 1];
	const unsigned D = std::min(m + 1, S->maxdepth[m] + 2);
	size_t smem = 0;
	const unsigned threads = walk_threads(D, &smem, S->eager ? 1 : 2);
	const unsigned blocks = (unsigned) ctx->sm_count * 16;
	if (S->eager) {
		if (S->L.W > 1) smem += (size_t) (threads / S->L.W) * (2 * m - 1) * sizeof(unsigned);      // per-position costs of the CTA's children
		if (S->planar) {
			XMP_CUDA(cudaFuncSetAttribute(k_build_children_eager<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
			k_build_children_eager<true><<<blocks, threads, smem, ctx->stream>>>(P.view(), C.view(), S->L, m, surv, nsurv, s_begin, s_end, child_first % C.cap,
			                                                                     S->rows, ctx->wp, S->d_merges, D, S->d_maxdepth + m + 1, S->d_overflow);
		} else {
			XMP_CUDA(cudaFuncSetAttribute(k_build_children_eager<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
			k_build_children_eager<false><<<blocks, threads, smem, ctx->stream>>>(P.view(), C.view(), S->L, m, surv, nsurv, s_begin, s_end, child_first % C.cap,
			                                                                      S->rows, ctx->wp, S->d_merges, D, S->d_maxdepth + m + 1, S->d_overflow);
		}
		XMP_CUDA(cudaGetLastError());
		S->st.launches += 1;
		return XMP_OK;
	}
	if (S->planar) {
		XMP_CUDA(cudaFuncSetAttribute(k_build_children<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
		k_build_children<true><<<blocks, threads, smem, ctx->stream>>>(P.view(), C.view(), S->L, m, surv, nsurv, s_begin, s_end, child_first % C.cap,
		                                                               S->rows, S->d_merges, D, S->d_maxdepth + m + 1, S->d_overflow);
	} else {
		XMP_CUDA(cudaFuncSetAttribute(k_build_children<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
		k_build_children<false><<<blocks, threads, smem, ctx->stream>>>(P.view(), C.view(), S->L, m, surv, nsurv, s_begin, s_end, child_first % C.cap,
		                                                                S->rows, S->d_merges, D, S->d_maxdepth + m + 1, S->d_overflow);
	}
	XMP_CUDA(cudaGetLastError());
	S->st.launches += 1;
	return XMP_OK;
}

static int launch_fused(xmp_ctx *ctx, Search *S, unsigned m, unsigned s_begin, unsigned s_end, const uint4 *surv, const unsigned *nsurv) {
	const Layout &L = S->L;
	Pool &P = S->pool[m];
	const unsigned D = std::min(m + 1, S->maxdepth[m] + 2);
	size_t smem = 0;
	const unsigned threads = walk_threads(D, &smem);
	const unsigned blocks = (unsigned) ctx->sm_count * 16;
#define XMP_FUSED_F(G, F)                                                                                                        \
	do {                                                                                                                         \
		XMP_CUDA(cudaFuncSetAttribute(k_expand_last_fused<G, F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));          \
		k_expand_last_fused<G, F><<<blocks, threads, smem, ctx->stream>>>(P.view(), L, m, surv, nsurv, s_begin, s_end,               \
		                                                               S->rows, ctx->wp, ctx->rest_bound[m + 1], S->d_bound,         \
		                                                               S->d_trees[S->cur], S->d_ntrees, S->trees_cap, S->d_merges, D, \
		                                                               S->d_overflow, S->d_peers);                                   \
	} while (0)
#define XMP_FUSED(G)                         \
	do {                                     \
		if (S->planar) XMP_FUSED_F(G, true); \
		else XMP_FUSED_F(G, false);          \
	} while (0)
	if (L.W == 1) XMP_FUSED(1);
	else if (L.W == 2) XMP_FUSED(2);
	else if (L.W <= 4) XMP_FUSED(4);
	else if (L.W <= 8) XMP_FUSED(8);
	else if (L.W <= 16) XMP_FUSED(16);
	else XMP_FUSED(32);
#undef XMP_FUSED
#undef XMP_FUSED_F
	XMP_CUDA(cudaGetLastError());
	S->st.launches += 1;
	return XMP_OK;
}

// Reads back n_surv, n_trees, bound and the merge counter (one small pinned copy, one sync).
static int fetch_counters(xmp_ctx *ctx, Search *S) {
	XMP_CUDA(cudaMemcpyAsync(S->h_pin, S->d_bound, 1024, cudaMemcpyDeviceToHost, ctx->stream));
	XMP_CUDA(cudaStreamSynchronize(ctx->stream));
	S->bound = S->h_pin[0];
	for (unsigned m = 3; m <= S->n; ++m) S->maxdepth[m] = std::max(S->maxdepth[m], S->h_pin[128 + m]);
	if (S->h_pin[6]) return set_error(XMP_ERR_NOMEM, "a walk stack overflowed: a tree was deeper than its level's recorded maximum (internal error)");
	return XMP_OK;
}

// ---- the -m cap in bounded memory: tree_book.h ------------------------------------------------------
// Host-side utility (no device needed): the first `keep` of n_paths insertion paths (num_taxa bytes each) in the
// reference's depth-first order are moved to the front, in that order.
extern "C" int xmp_cuda_paths_keep_first_dfs(uint8_t *paths, size_t n_paths, unsigned num_taxa, size_t keep) {
	if (!paths || num_taxa < 3 || num_taxa > XMP_MAX_TAXA) return set_error(XMP_ERR_ARG, "bad path array");
	keep_first_dfs(paths, n_paths, num_taxa, 0, num_taxa, keep);
	return XMP_OK;
}

// Moves every tree record off the device into the host-side book.
static int drain_trees(xmp_ctx *ctx, Search *S) {
	if (S->ntrees_dev) {
		const size_t bytes = (size_t) S->ntrees_dev * S->L.trec, old = S->book.recs.size();
		S->book.recs.resize(old + bytes);
		XMP_CUDA(cudaMemcpyAsync(S->book.recs.data() + old, S->d_trees[S->cur], bytes, cudaMemcpyDeviceToHost, ctx->stream));
		XMP_CUDA(cudaStreamSynchronize(ctx->stream));
		S->book.absorb(old, S->bound);
	}
	S->ntrees_dev = 0;
	XMP_CUDA(cudaMemsetAsync(S->d_ntrees, 0, sizeof(unsigned), ctx->stream));
	return XMP_OK;
}

// Keeps the device tree buffer at most half full: first by discarding records above the bound.
static int tidy_trees(xmp_ctx *ctx, Search *S) {
	if (S->ntrees_dev <= S->trees_cap / 4) return XMP_OK;
	const int other = S->cur ^ 1;
	XMP_CUDA(cudaMemsetAsync(S->d_ntrees, 0, sizeof(unsigned), ctx->stream));
	k_filter_trees<<<(S->ntrees_dev + 255) / 256, 256, 0, ctx->stream>>>(S->d_trees[S->cur], S->ntrees_dev, S->L.trec, S->d_bound,
	                                                                   S->d_trees[other], S->d_ntrees);
	XMP_CUDA(cudaGetLastError());
	S->st.launches += 1;
	S->cur = other;
	int rc = fetch_counters(ctx, S);
	if (rc) return rc;
	S->ntrees_dev = S->h_pin[2];
	if (S->ntrees_dev > S->trees_cap / 2) return drain_trees(ctx, S);
	return XMP_OK;
}

// The tree buffer overflowed inside the fused kernel (records past the end were dropped, the counter kept
// counting): restore the counter, make room and redo this level's survivors piece by piece.  The counter
// says how many trees the whole list yields, so the list is cut into pieces that should each fill about half
// the buffer; a piece that still overflows is halved.  Emission is idempotent: records carry their own length.
static int redo_fused(xmp_ctx *ctx, Search *S, unsigned m, unsigned ns, const uint4 *surv, const unsigned *nsurv, unsigned before) {
	struct Range { unsigned lo, hi; };
	std::vector<Range> todo;
	const unsigned long long produced = S->h_pin[2] - before, half = S->trees_cap / 2;
	unsigned pieces = (unsigned) std::min<unsigned long long>(ns, std::max<unsigned long long>(2, (produced + half - 1) / half));
	if (pieces < 1) pieces = 1;
	for (unsigned p = pieces; p-- > 0;) {
		const unsigned lo = (unsigned) ((unsigned long long) ns * p / pieces), hi = (unsigned) ((unsigned long long) ns * (p + 1) / pieces);
		if (hi > lo) todo.push_back(Range{lo, hi});
	}
	S->ntrees_dev = before;
	while (!todo.empty()) {
		Range r = todo.back();
		todo.pop_back();
		int rc = drain_trees(ctx, S);
		if (rc) return rc;
		rc = launch_fused(ctx, S, m, r.lo, r.hi, surv, nsurv);
		if (rc) return rc;
		rc = fetch_counters(ctx, S);
		if (rc) return rc;
		if (S->h_pin[2] > S->trees_cap) {
			if (r.hi - r.lo < 2) return set_error(XMP_ERR_NOMEM, "tree buffer too small for a single partial tree's completions");
			const unsigned mid = r.lo + (r.hi - r.lo) / 2;
			todo.push_back(Range{mid, r.hi});
			todo.push_back(Range{r.lo, mid});
			S->ntrees_dev = 0;
			XMP_CUDA(cudaMemsetAsync(S->d_ntrees, 0, sizeof(unsigned), ctx->stream));
			continue;
		}
		S->st.full_trees += S->h_pin[2];
		S->ntrees_dev = S->h_pin[2];
	}
	return XMP_OK;
}

// One frontier batch at level m.  keep_best > 0 (the dive) keeps only that many cheapest children.
static int run_batch(xmp_ctx *ctx, Search *S, unsigned m, unsigned long long chunk, unsigned keep_best) {
	Pool &P = S->pool[m];
	const bool last = (m + 1 == S->n);
	const bool shard = S->prm.shard_count > 1 && m + 1 == S->split_level && keep_best == 0;
	const unsigned long long first = P.slot(P.count - chunk);      // most recent nodes first (stack order)
	if (!last) XMP_CUDA(cudaMemsetAsync(S->d_nsurv, 0, sizeof(unsigned), ctx->stream));
	int rc = launch_expand(ctx, S, m, first, chunk, shard, S->d_surv, S->d_nsurv);
	if (rc) return rc;
	count_batch(S->st, m, chunk);
	S->st.iterations += 1;
	if (last) {
		rc = fetch_counters(ctx, S);
		if (rc) return rc;
		S->st.full_trees += S->h_pin[2] - S->ntrees_dev;
		S->ntrees_dev = S->h_pin[2];
		P.count -= chunk;
		return tidy_trees(ctx, S);
	}
	const uint4 *surv = S->d_surv;
	const unsigned *nsurv = S->d_nsurv;
	unsigned ns_word = 1;                       // which word of the counter block holds the length of the list used
	if (keep_best) {
		// never more children than the next pool can take (a small memory budget on many taxa: capacity < beam)
		if (!(S->fused_last && m + 2 == S->n)) keep_best = (unsigned) std::min<unsigned long long>(keep_best, S->pool[m + 1].cap - S->pool[m + 1].count);
		if (keep_best == 0) return set_error(XMP_ERR_NOMEM, "frontier pool for %u taxa is full (%llu nodes)", m + 1, S->pool[m + 1].cap);
		// the beam, cut on the device into the tail of the survivor buffer (the list itself is at most beam x (2m-3)
		// entries, the buffer beam x (2n-3)); no host round between scoring and building
		uint4 *beam_out = S->d_surv + (S->surv_cap - keep_best);
		k_beam_select<<<1, 1024, 0, ctx->stream>>>(S->d_surv, S->d_nsurv, keep_best, beam_out, S->d_bound + 7);
		XMP_CUDA(cudaGetLastError());
		S->st.launches += 1;
		surv = beam_out;
		nsurv = S->d_bound + 7;
		ns_word = 7;
	}
	if (S->fused_last && m + 2 == S->n) {
		// Children of this level are scored against the last taxon without ever being stored.
		const unsigned before = S->ntrees_dev;
		rc = launch_fused(ctx, S, m, 0, 0xffffffffu, surv, nsurv);
		if (rc) return rc;
		rc = fetch_counters(ctx, S);
		if (rc) return rc;
		const unsigned ns = S->h_pin[ns_word];
		if (S->h_pin[2] > S->trees_cap) {
			rc = redo_fused(ctx, S, m, ns, surv, nsurv, before);
			if (rc) return rc;
		} else {
			S->st.full_trees += S->h_pin[2] - before;
			S->ntrees_dev = S->h_pin[2];
		}
		S->st.children += ns;
		count_batch(S->st, m + 1, ns);
		P.count -= chunk;
		return tidy_trees(ctx, S);
	}
	rc = launch_children(ctx, S, m, surv, nsurv, S->pool[m + 1].slot(S->pool[m + 1].count));
	if (rc) return rc;
	rc = fetch_counters(ctx, S);
	if (rc) return rc;
	const unsigned ns = S->h_pin[ns_word];
	S->pool[m + 1].count += ns;
	S->st.children += ns;
	P.count -= chunk;
	return XMP_OK;
}

// One sweep over the frontier: every level that holds open trees is expanded in the same pass, deepest
// first, on one stream and with ONE host synchronisation at the end (per-level survivor lists and counters).
// Going deepest-first lets level m's children land exactly where level m + 1's batch was just consumed, so
// the pools stay plain stacks.  Compared with one level per synchronisation this cuts the number of host
// round trips on deep, narrow searches (20 taxa: 17 levels) by an order of magnitude.
static int run_sweep(xmp_ctx *ctx, Search *S, bool *any) {
	const unsigned n = S->n;
	*any = false;
	// Open trees on the last level only exist when the last two levels are not fused (or were imported):
	// they emit complete trees into the shared buffer, so they go alone, as a classic single batch.
	if (S->pool[n - 1].count) {
		Pool &P = S->pool[n - 1];
		unsigned long long room = (S->trees_cap - S->ntrees_dev) / P.E;
		if (room == 0) {
			int rc = drain_trees(ctx, S);
			if (rc) return rc;
			room = S->trees_cap / P.E;
		}
		*any = true;
		return run_batch(ctx, S, n - 1, std::min(P.count, room), 0);
	}
	struct Item { unsigned m; unsigned long long chunk, room; bool fused, cont; };
	Item items[XMP_MAX_TAXA + 2];
	unsigned n_items = 0;
	const unsigned before = S->ntrees_dev;
	// Batch of level m: as many of its oldest open trees as the next pool can be EXPECTED to absorb, going by
	// the survival rate seen so far (a worst-case bound of 2m-3 children per tree would keep batches an order
	// of magnitude smaller).  If more children survive than fit, the batch stays put and the rest of its
	// survivor list is built in later sweeps (Pending).
	auto plan = [&](unsigned m, unsigned long long room_next) -> unsigned long long {
		Pool &P = S->pool[m];
		const bool fused = S->fused_last && m + 2 == n;
		if (P.count == 0 || S->pend[m].active) return 0;
		const double per_tree = fused ? (double) P.E : std::max(1.0, P.E * std::min(1.0, S->rate[m] * 1.25));
		unsigned long long chunk = fused ? P.count : (unsigned long long) ((double) room_next / per_tree);
		chunk = std::min(chunk, S->surv_cap_lvl[m] / P.E);        // the survivor list itself must never overflow
		return std::min(chunk, P.count);
	};
	// A level joins the sweep only if it brings a worthwhile batch: every launch costs tens of microseconds
	// whatever its size.  If nothing qualifies (the start and the tail of a search) only the fullest level
	// goes, which grows the frontier (XMP_SWEEP_FALLBACK=0: every level goes).
	unsigned long long best_chunk = 0;
	unsigned best_level = 0;
	bool worthwhile = false;
	for (unsigned m = n - 2; m >= 3; --m) {
		const bool fused = S->fused_last && m + 2 == n;
		const unsigned long long room_next = fused ? 0 : S->pool[m + 1].cap - S->pool[m + 1].count;
		if (S->pend[m].active) {
			if (room_next) worthwhile = true;      // it can move on
			continue;
		}
		const unsigned long long chunk = plan(m, room_next);
		if (chunk >= std::min<unsigned long long>(S->min_batch, S->pool[m].cap)) worthwhile = true;
		if (chunk > best_chunk) { best_chunk = chunk; best_level = m; }
	}
	// A sparse frontier (fewer open trees in all than one worthwhile batch: the tail of a search, or a thief that has
	// just been handed a few subtrees) is a chain of latencies whatever is done; then every level goes in the same sweep,
	// so that the chain is as long as the remaining subtrees are deep, not as their levels are many.
	unsigned long long open_total = 0;
	for (unsigned m = 3; m + 1 < n; ++m) open_total += S->pool[m].count;
	const bool sparse = open_total < S->min_batch;
	// pass 0 applies the selection above; pass 1 (only if pass 0 launched nothing) lets every level go that can
	// One rank of a multi-GPU search answers steal requests between two sweeps: it keeps its sweeps short (the deepest
	// few levels that qualify) so that an idle peer waits milliseconds, not the tens of milliseconds that a pass over
	// twenty well-filled levels takes (its36 under -Bd on 2 GPUs: the thief spent 0.73 of 3.4 s waiting for replies).
	const unsigned level_limit = S->d_peers ? S->sweep_levels : 0xffffffffu;
	for (int pass = 0; pass < 2 && n_items == 0; ++pass) {
		for (unsigned m = n - 2; m >= 3 && n_items < level_limit; --m) {
			Pool &P = S->pool[m];
			const bool fused = S->fused_last && m + 2 == n;
			const unsigned long long room_next = fused ? 0 : S->pool[m + 1].cap - S->pool[m + 1].count;
			if (!fused && S->pend[m].active) {
				Search::Pending &pd = S->pend[m];
				const unsigned long long take = std::min<unsigned long long>(pd.ns - pd.head, room_next);
				if (take == 0) continue;
				// the level's counter still holds the length of the list written by its last scoring pass
				int rc = launch_children(ctx, S, m, S->d_surv_lvl[m], S->d_nsurv_lvl + m, S->pool[m + 1].slot(S->pool[m + 1].count),
				                         pd.head, pd.head + (unsigned) take);
				if (rc) return rc;
				items[n_items++] = Item{m, take, 0, false, true};
				continue;
			}
			const unsigned long long chunk = plan(m, room_next);
			if (chunk == 0) continue;
			if (pass == 0) {
				if (worthwhile) {
					if (chunk < std::min<unsigned long long>(S->min_batch, P.cap)) continue;
				} else if (S->sweep_fallback == 1 && !sparse && m != best_level) {
					continue;
				}
			}
			const bool shard = S->prm.shard_count > 1 && m + 1 == S->split_level;
			XMP_CUDA(cudaMemsetAsync(S->d_nsurv_lvl + m, 0, sizeof(unsigned), ctx->stream));
			int rc = launch_expand(ctx, S, m, P.slot(0), chunk, shard, S->d_surv_lvl[m], S->d_nsurv_lvl + m);
			if (rc) return rc;
			const unsigned s_end = (unsigned) std::min<unsigned long long>(room_next, 0xffffffffull);
			rc = fused ? launch_fused(ctx, S, m, 0, 0xffffffffu, S->d_surv_lvl[m], S->d_nsurv_lvl + m)
			           : launch_children(ctx, S, m, S->d_surv_lvl[m], S->d_nsurv_lvl + m, S->pool[m + 1].slot(S->pool[m + 1].count), 0, s_end);
			if (rc) return rc;
			items[n_items++] = Item{m, chunk, room_next, fused, false};
		}
	}
	if (n_items == 0) return XMP_OK;
	*any = true;
	int rc = fetch_counters(ctx, S);
	if (rc) return rc;
	for (unsigned i = 0; i < n_items; ++i) {
		const unsigned m = items[i].m;
		Pool &P = S->pool[m];
		if (items[i].cont) {
			Search::Pending &pd = S->pend[m];
			pd.head += (unsigned) items[i].chunk;
			S->pool[m + 1].count += items[i].chunk;
			S->st.children += items[i].chunk;
			if (pd.head == pd.ns) {
				P.head = P.slot(pd.chunk);
				P.count -= pd.chunk;
				pd.active = false;
			}
			continue;
		}
		const unsigned ns = S->h_pin[8 + m];
		count_batch(S->st, m, items[i].chunk);
		S->rate[m] = std::max((double) ns / ((double) items[i].chunk * P.E), 0.5 * S->rate[m]);
		if (items[i].fused) {
			P.head = P.slot(items[i].chunk);
			P.count -= items[i].chunk;
			S->st.children += ns;
			count_batch(S->st, m + 1, ns);
			if (S->h_pin[2] > S->trees_cap) {
				rc = redo_fused(ctx, S, m, ns, S->d_surv_lvl[m], S->d_nsurv_lvl + m, before);
				if (rc) return rc;
			} else {
				S->st.full_trees += S->h_pin[2] - before;
				S->ntrees_dev = S->h_pin[2];
			}
			continue;
		}
		const unsigned long long built = std::min<unsigned long long>(ns, items[i].room);
		S->pool[m + 1].count += built;
		S->st.children += built;
		if (ns > built) {
			S->pend[m].active = true;
			S->pend[m].chunk = items[i].chunk;
			S->pend[m].head = (unsigned) built;
			S->pend[m].ns = ns;
		} else {
			P.head = P.slot(items[i].chunk);
			P.count -= items[i].chunk;
		}
	}
	S->st.iterations += 1;
	return tidy_trees(ctx, S);
}

static int push_root(xmp_ctx *ctx, Search *S) {
	std::vector<uint8_t> job(S->L.n, 0);
	job[0] = 3;
	return import_nodes(ctx, S, job.data(), 1);
}

// Greedy beam descent from the 3-taxon tree: the same role as the reference's greedy stepwise-addition
// tree (src/greedytree.c:51-105) -- a quick complete tree whose length tightens the bound.  It changes
// how many nodes the search visits, never its result.
static int dive(xmp_ctx *ctx, Search *S) {
	int rc = push_root(ctx, S);
	if (rc) return rc;
	for (unsigned m = 3; m < S->n; ++m) {
		if (S->pool[m].count == 0) break;
		rc = run_batch(ctx, S, m, S->pool[m].count, S->beam);
		if (rc) return rc;
	}
	for (unsigned m = 3; m < S->n; ++m) S->pool[m].count = S->pool[m].head = 0;
	// Trees found by the dive will be found again by the search proper.
	S->ntrees_dev = 0;
	S->book.reset();
	XMP_CUDA(cudaMemsetAsync(S->d_ntrees, 0, sizeof(unsigned), ctx->stream));
	return fetch_counters(ctx, S);
}

static void init_knobs(xmp_ctx *ctx, Search *S) {
	S->n = ctx->num_taxa;
	S->min_batch = std::max(env_unsigned("XMP_MIN_BATCH", 32768), 1u);
	// the dive costs about a millisecond whatever its width now that the beam is cut on the device; a wider beam finds
	// the optimum (or a length next to it) more often: mh1 55 M -> 28 M nodes, h4 ~260 M -> 217 M (profiles/r02_beam_ab.md)
	S->beam = std::max(std::min(env_unsigned("XMP_BEAM", 4096), 65536u), 1u);
	S->sweep = env_unsigned("XMP_SWEEP", 1) != 0;
	S->sweep_fallback = env_unsigned("XMP_SWEEP_FALLBACK", 1);
	S->sweep_levels = std::max(env_unsigned("XMP_SWEEP_LEVELS", 6), 1u);
	S->planar = env_unsigned("XMP_SEARCH_PLANAR", 0) != 0;
	for (unsigned m = 0; m < XMP_MAX_TAXA + 2; ++m) S->rate[m] = 1.0;
	memset(&S->st, 0, sizeof S->st);
}

// How much work a tree on m + 1 taxa stands for, relative to one on m taxa, when a donor decides what to give away and
// a thief whom to ask.  Unpruned it would be 1 / (2m - 3); under the bound the search trees measured here grow by a
// factor of 1 to 2.5 per tree size in their busy middle (profiles/r02_probe_long_workloads_raw.txt), so most of the
// remaining work sits deep, and a share of it means a share of the deeper levels too.
static const double STEAL_LEVEL_WEIGHT = 0.5;

// Detaches open partial trees and stages their job descriptors (insertion paths, byte 0 = tree size) in S->pathbuf on
// the device.  The reference's thief takes "the last unprocessed edge at the victim's shallowest open level"
// (src/mpiworker.c:169-181); here
//   thieves == 0  (xmp_cuda_search_export_jobs): up to max_nodes trees of the shallowest open level;
//   thieves == k  (a donor serving k idle ranks at once): the share k / (k + 1) of the WORK behind the open trees
//                 (STEAL_LEVEL_WEIGHT per tree size), taken from the shallowest levels down -- never the only open tree,
//                 at least one tree per thief, at most XCH_MAX_JOBS.  Handing out one shallow tree per request was
//                 measured on 8 GPUs: the rank with the largest share of its36 served 256 requests, each good for
//                 ~1.5 ms of work, and was still the last to finish.
static int detach_jobs(xmp_ctx *ctx, Search *S, unsigned max_nodes, unsigned thieves, unsigned *n_out) {
	const Layout &L = S->L;
	*n_out = 0;
	// A sharded search deals the frontier out by hash when it expands the level below split_level; a node from a
	// shallower level would be dealt again -- with the thief's index -- on the other side.  So only subtrees that are
	// past the deal ever change hands (the shallow prefix is a handful of nodes, gone after the first sweeps).
	const unsigned m0 = S->prm.shard_count > 1 ? std::max(3u, S->split_level) : 3u;
	unsigned long long open = 0;
	double work = 0, unit = 1;
	for (unsigned m = m0; m < S->n; ++m) {
		if (!S->pend[m].active) {
			open += S->pool[m].count;
			work += (double) S->pool[m].count * unit;
		}
		unit *= STEAL_LEVEL_WEIGHT;
	}
	if (thieves && open < 2) return XMP_OK;
	const double target = work * (double) thieves / (double) (thieves + 1);
	int rc = S->pathbuf.reserve((size_t) std::max<unsigned>(max_nodes, 256) * L.n);
	if (rc) return rc;
	unsigned total = 0;
	double got = 0;
	unit = 1;
	for (unsigned m = m0; m < S->n && total < max_nodes; unit *= STEAL_LEVEL_WEIGHT, ++m) {
		Pool &P = S->pool[m];
		if (P.count == 0 || S->pend[m].active) continue;      // pending: its oldest nodes are half expanded; newer ones sit behind them in the ring
		unsigned long long want = P.count;
		if (thieves) {
			if (got >= target && total >= thieves) break;
			const double need = std::max((target - got) / unit, (double) (total < thieves ? thieves - total : 0));
			want = (unsigned long long) std::min((double) P.count, std::ceil(need));
			if (want >= open) want = open - 1;                                  // keep one tree to work on
		}
		unsigned k = (unsigned) std::min<unsigned long long>(want, max_nodes - total);
		const unsigned long long tail = P.slot(P.count);         // one past the newest node
		if (tail != 0) k = (unsigned) std::min<unsigned long long>(k, tail);   // stay on this side of the wrap
		if (k == 0) continue;
		const unsigned long long from = (tail == 0 ? P.cap : tail) - k;
		k_read_paths<<<(k + 127) / 128, 128, 0, ctx->stream>>>(P.view(), L, m, from, k, (uint8_t *) S->pathbuf.ptr + (size_t) total * L.n);
		XMP_CUDA(cudaGetLastError());
		S->st.launches += 1;
		P.count -= k;
		open -= k;
		total += k;
		got += (double) k * unit;
		if (!thieves) break;
	}
	*n_out = total;
	return XMP_OK;
}

// ---- the engine side of the exchange protocol (exchange_proto.h, exchange.cu) -------------------------
// What this rank could give away right now, published as a hint for idle ranks: 0 = nothing, 1 = a single open tree
// (not given away), otherwise 2 + an estimate of the work behind its open trees (STEAL_LEVEL_WEIGHT per tree size),
// which ranks a rank holding shallow trees above one holding as many deep ones -- that is where thieves should go.
static int publish_open(xmp_ctx *ctx, Search *S) {
	const unsigned m0 = S->prm.shard_count > 1 ? std::max(3u, S->split_level) : 3u;
	unsigned long long open = 0;
	double work = 0, unit = (double) (1u << 20);
	for (unsigned m = m0; m < S->n; ++m) {
		if (!S->pend[m].active) {
			open += S->pool[m].count;
			work += (double) S->pool[m].count * unit;
		}
		unit *= STEAL_LEVEL_WEIGHT;
	}
	const unsigned v = open < 2 ? (unsigned) open : 2u + (unsigned) std::min(work, 2.0e9);
	if (v == S->open_published) return XMP_OK;
	S->open_published = v;
	ctx->h_xres[8] = v;           // pinned; a later value overtaking this copy is just as good a hint
	XMP_CUDA(cudaMemcpyAsync(S->d_bound + XW_OPEN, ctx->h_xres + 8, sizeof(unsigned), cudaMemcpyHostToDevice, ctx->stream));
	return XMP_OK;
}

int search_export_for_steal(xmp_ctx *ctx, unsigned k, unsigned *n) {
	Search *S = ctx->search;
	int rc = detach_jobs(ctx, S, XCH_MAX_JOBS, k, n);
	if (rc == XMP_OK && *n) {
		S->steals_served += 1;
		rc = publish_open(ctx, S);
	}
	return rc;
}

// Donor -> thief: staged descriptors i, i + k, i + 2k, ... (a mix of shallow and deeper subtrees for every thief) go
// straight from this GPU's memory into the thief's mailbox -- a strided peer-to-peer copy over NVLink; the addresses
// are peer-mapped, cudaMemcpyDefault resolves them.
int search_send_staged(xmp_ctx *ctx, unsigned thief, unsigned i, unsigned k, unsigned n) {
	Search *S = ctx->search;
	const size_t rec = S->L.n, share = (n - i + k - 1) / k;
	XMP_CUDA(cudaMemcpy2DAsync((char *) ctx->xch_peer[thief] + XCH_MAILBOX_BYTE_OFF, rec, (const char *) S->pathbuf.ptr + (size_t) i * rec, (size_t) k * rec,
	                           rec, share, cudaMemcpyDefault, ctx->stream));
	return XMP_OK;
}

// Thief: the descriptors in this rank's mailbox become open subtrees (the reference's fast-forward,
// src/yanbader.c:95-99,136-143, as one batched set construction).
int search_import_mailbox(xmp_ctx *ctx, unsigned n) {
	Search *S = ctx->search;
	if (n > XCH_MAX_JOBS) return set_error(XMP_ERR_ARG, "a reply announces %u jobs, the mailbox holds %u", n, (unsigned) XCH_MAX_JOBS);
	std::vector<uint8_t> jobs((size_t) n * S->L.n);
	XMP_CUDA(cudaMemcpyAsync(jobs.data(), (const char *) ctx->xch + XCH_MAILBOX_BYTE_OFF, jobs.size(), cudaMemcpyDeviceToHost, ctx->stream));
	XMP_CUDA(cudaStreamSynchronize(ctx->stream));
	S->done = false;
	S->finished = false;
	S->steals_got += 1;
	int rc = import_nodes(ctx, S, jobs.data(), n);
	if (rc) return rc;
	return publish_open(ctx, S);
}

// After xmp_cuda_search_steal: either the search goes on, or it is over everywhere -- then the bound word holds the
// minimum over all ranks (every peer's atomicMin landed before that peer could go idle).
int search_after_steal(xmp_ctx *ctx, bool finished) {
	Search *S = ctx->search;
	if (!finished) return XMP_OK;
	int rc = drain_trees(ctx, S);
	if (rc == XMP_OK) rc = fetch_counters(ctx, S);
	if (rc == XMP_OK) S->done = S->finished = true;
	return rc;
}

int xch_service_requests(xmp_ctx *ctx, const unsigned *words);
static bool requests_pending(const Search *S, const xmp_ctx *ctx) {
	for (unsigned r = 0; r < ctx->xch_world; ++r)
		if (S->h_pin[XW_REQUEST + r]) return true;
	return false;
}

// The two halves of xmp_cuda_search_begin that talk to the device (split out so that every failure takes the same
// tear-down path in the caller).
static int begin_bound(xmp_ctx *ctx, Search *S) {
	if (ctx->xch_connected) {
		// the word was set to "no bound" by xmp_cuda_exchange_reset before the ranks' barrier; a peer that started
		// earlier may already have lowered it
		k_lower_bound<<<1, 1, 0, ctx->stream>>>(S->d_bound, S->bound);
		XMP_CUDA(cudaGetLastError());
		return fetch_counters(ctx, S);
	}
	XMP_CUDA(cudaMemcpyAsync(S->d_bound, &S->bound, sizeof(unsigned), cudaMemcpyHostToDevice, ctx->stream));
	XMP_CUDA(cudaStreamSynchronize(ctx->stream));
	return XMP_OK;
}

static int begin_timed(xmp_ctx *ctx, Search *S) {
	XMP_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
	if (!(S->prm.flags & XMP_SEARCH_NO_DIVE)) {
		int rc = dive(ctx, S);
		if (rc) return rc;
	}
	int rc = push_root(ctx, S);
	if (rc) return rc;
	XMP_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
	XMP_CUDA(cudaEventSynchronize(ctx->ev1));
	float ms = 0;
	XMP_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
	S->st.seconds += ms * 1e-3;
	S->st.begin_seconds = ms * 1e-3;
	return XMP_OK;
}

}  // namespace xmp

using namespace xmp;

// Sizes and allocates the search's device memory for a problem of the given shape ahead of time, the
// way the reference sets up its arena in InitTreeData before the branch-and-bound timer starts
// (src/common.c:1113-1224, src/arena.h): a following xmp_cuda_search_begin on such a problem with the
// same budget allocates nothing.
extern "C" int xmp_cuda_search_reserve(xmp_ctx *ctx, unsigned num_taxa, unsigned n_blocks, size_t mem_budget) {
	if (!ctx) return set_error(XMP_ERR_ARG, "null context");
	if (num_taxa < 4 || num_taxa > XMP_MAX_TAXA || n_blocks == 0) return set_error(XMP_ERR_ARG, "cannot reserve a search for %u taxa x %u blocks", num_taxa, n_blocks);
	XMP_CUDA(cudaSetDevice(ctx->device));
	search_destroy(ctx);
	Search *S = new Search();
	ctx->search = S;
	memset(&S->prm, 0, sizeof S->prm);
	S->prm.mem_budget = mem_budget;
	S->prm.shard_count = 1;
	init_knobs(ctx, S);
	S->n = num_taxa;
	int rc = alloc_search(ctx, S, n_blocks);
	search_destroy(ctx);
	return rc;
}

extern "C" int xmp_cuda_search_begin(xmp_ctx *ctx, const xmp_search_params *params) {
	if (!ctx || !ctx->d_rows || !params) return set_error(XMP_ERR_ARG, "no problem loaded");
	if (ctx->num_taxa < 4) return set_error(XMP_ERR_ARG, "the search needs at least four taxa");
	const unsigned shard_count = params->shard_count ? params->shard_count : 1u;
	if (params->shard_index >= shard_count) return set_error(XMP_ERR_ARG, "shard_index %u >= shard_count %u", params->shard_index, shard_count);
	if (ctx->xch_connected && (shard_count != ctx->xch_world || params->shard_index != ctx->xch_rank)) {
		return set_error(XMP_ERR_ARG, "this context is connected as rank %u of %u: its searches are shard %u of %u (got %u of %u); "
		                              "xmp_cuda_exchange_disconnect first for a search of its own",
		                 ctx->xch_rank, ctx->xch_world, ctx->xch_rank, ctx->xch_world, params->shard_index, shard_count);
	}
	XMP_CUDA(cudaSetDevice(ctx->device));
	search_destroy(ctx);
	Search *S = new Search();
	ctx->search = S;
	S->prm = *params;
	S->prm.shard_count = shard_count;
	init_knobs(ctx, S);
	// a search that fails to start leaves no half-built state behind: later calls then report "no search in progress"
	auto fail = [&](int rc) { search_destroy(ctx); return rc; };
	int rc = alloc_search(ctx, S, ctx->n_blocks);
	if (rc) return fail(rc);
	S->rows = ctx->d_rows;
	if (S->planar) {
		const unsigned long long nb = (unsigned long long) ctx->num_taxa * ctx->n_blocks;
		rc = S->prows.reserve((size_t) nb * sizeof(uint4));
		if (rc) return fail(rc);
		k_rows_to_planar<<<(unsigned) ((nb + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_rows, (uint4 *) S->prows.ptr, nb);
		const cudaError_t e = cudaGetLastError();
		if (e != cudaSuccess) return fail(cuda_failed(e, "k_rows_to_planar", __FILE__, __LINE__));
		S->rows = (const uint4 *) S->prows.ptr;
	}
	S->book.trec = S->L.trec;
	S->book.n = S->L.n;
	S->book.max_trees = S->prm.max_trees;
	S->bound = params->bound;
	rc = begin_bound(ctx, S);
	if (rc) return fail(rc);

	// Where the frontier is dealt out to shards: the first tree size with >= 65536 subtrees per shard.  The deal is the
	// load balance: subtree sizes are heavy-tailed, and only tens of thousands of them per shard even out.  Measured on
	// 8 GPUs (profiles/r02_scaling.md): with 64 per shard its36 under -Bd left one GPU with twice the work of the others,
	// and stealing cannot repair that cheaply -- a thief only ever gets what is open between two sweeps; with 4096 the
	// slowest GPU still ran 40 % longer than the fastest; with 65536 (tree size 10 on 8 GPUs) all finish within 15 %.
	// The shallow prefix that every shard expands redundantly is then ~150 K trees, a fraction of a millisecond; with
	// 2^20 per shard it starts to show (h4 +7 % nodes).  XMP_SPLIT_PER_SHARD overrides.
	S->split_level = 0;
	if (S->prm.shard_count > 1) {
		unsigned s = S->prm.split_level;
		if (s == 0) {
			const double per_shard = (double) std::max(env_unsigned("XMP_SPLIT_PER_SHARD", 65536), 1u);
			s = 4;
			while (s + 2 < S->n && odd_factorial(s) < per_shard * S->prm.shard_count) ++s;
		}
		if (s < 4) s = 4;
		if (s > S->n - 1) s = S->n - 1;
		S->split_level = s;
	}
	rc = begin_timed(ctx, S);
	if (rc) return fail(rc);
	return XMP_OK;
}

extern "C" int xmp_cuda_search_run(xmp_ctx *ctx, unsigned max_batches, int *done) {
	if (!ctx || !ctx->search) return set_error(XMP_ERR_ARG, "xmp_cuda_search_begin has not been called");
	Search *S = ctx->search;
	XMP_CUDA(cudaSetDevice(ctx->device));
	XMP_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
	unsigned batches = 0;
	int rc = XMP_OK;
	while (S->sweep && !S->done && (max_batches == 0 || batches < max_batches)) {
		bool any = false;
		rc = run_sweep(ctx, S, &any);
		if (rc) break;
		if (!any) {
			bool open = false;
			for (unsigned m = 3; m < S->n; ++m) open |= S->pool[m].count != 0;
			if (open) {
				// cannot happen: the deepest open level can always be expanded
				rc = set_error(XMP_ERR_NOMEM, "the frontier is blocked: no level can be expanded (internal error)");
				break;
			}
			S->done = true;
			break;
		}
		++batches;
		if (S->d_peers) {
			// one rank of a multi-GPU search: publish what could be given away, answer a waiting thief
			// (the request word came back with this sweep's counters)
			rc = publish_open(ctx, S);
			if (rc == XMP_OK && requests_pending(S, ctx)) rc = xch_service_requests(ctx, S->h_pin + XW_FIRST);
			if (rc) break;
		}
	}
	while (!S->sweep && !S->done && (max_batches == 0 || batches < max_batches)) {
		// XMP_SWEEP=0 (kept for A/B measurements, profiles/r01_search_tuning.md): one level per host
		// synchronisation, worst-case batch sizing; the deepest level holding a worthwhile batch, otherwise
		// the fullest one.
		int pick = -1, fallback = -1;
		unsigned long long pick_chunk = 0, fallback_chunk = 0;
		for (unsigned m = S->n - 1; m >= 3; --m) {
			Pool &P = S->pool[m];
			if (P.count == 0) continue;
			unsigned long long room;
			if (m + 1 == S->n) room = (S->trees_cap - S->ntrees_dev) / P.E;
			else if (S->fused_last && m + 2 == S->n) room = S->surv_cap / P.E;     // survivors are consumed in place
			else room = (S->pool[m + 1].cap - S->pool[m + 1].count) / P.E;
			unsigned long long chunk = std::min(P.count, room);
			if (chunk == 0) continue;
			if (chunk > fallback_chunk) {
				fallback = (int) m;
				fallback_chunk = chunk;
			}
			if (chunk >= S->min_batch) { pick = (int) m; pick_chunk = chunk; break; }
		}
		if (pick < 0) { pick = fallback; pick_chunk = fallback_chunk; }
		if (pick < 0) {
			bool any = false;
			for (unsigned m = 3; m < S->n; ++m) any |= S->pool[m].count != 0;
			if (any) {
				// Only possible when the tree buffer is full: make room and retry.
				rc = drain_trees(ctx, S);
				if (rc) break;
				continue;
			}
			S->done = true;
			break;
		}
		rc = run_batch(ctx, S, (unsigned) pick, pick_chunk, 0);
		if (rc) break;
		++batches;
		if (S->d_peers) {
			rc = publish_open(ctx, S);
			if (rc == XMP_OK && requests_pending(S, ctx)) rc = xch_service_requests(ctx, S->h_pin + XW_FIRST);
			if (rc) break;
		}
	}
	if (rc == XMP_OK && !S->done) {
		// A slice that stopped on max_batches may have emptied the frontier with its last sweep: report it
		// now, so that "nothing pending" and "done" never disagree (the multi-GPU hosts terminate on the former
		// and fetch results on the latter).
		bool open = false;
		for (unsigned m = 3; m < S->n; ++m) open |= S->pool[m].count != 0;
		if (!open) S->done = true;
	}
	if (rc == XMP_OK && S->d_peers) rc = publish_open(ctx, S);      // an empty frontier is published too: nothing to steal here
	if (rc == XMP_OK && S->done && !S->finished) {
		rc = drain_trees(ctx, S);
		if (rc == XMP_OK) rc = fetch_counters(ctx, S);
		if (rc == XMP_OK) S->finished = true;
	}
	cudaEventRecord(ctx->ev1, ctx->stream);
	cudaEventSynchronize(ctx->ev1);
	float ms = 0;
	cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
	S->st.seconds += ms * 1e-3;
	if (done) *done = S->done ? 1 : 0;
	return rc;
}

extern "C" int xmp_cuda_search_get_bound(xmp_ctx *ctx, unsigned *bound) {
	if (!ctx || !ctx->search || !bound) return set_error(XMP_ERR_ARG, "no search in progress");
	XMP_CUDA(cudaSetDevice(ctx->device));
	int rc = fetch_counters(ctx, ctx->search);
	*bound = ctx->search->bound;
	return rc;
}

extern "C" int xmp_cuda_search_set_bound(xmp_ctx *ctx, unsigned bound) {
	if (!ctx || !ctx->search) return set_error(XMP_ERR_ARG, "no search in progress");
	Search *S = ctx->search;
	XMP_CUDA(cudaSetDevice(ctx->device));
	k_lower_bound<<<1, 1, 0, ctx->stream>>>(S->d_bound, bound);
	XMP_CUDA(cudaGetLastError());
	return fetch_counters(ctx, S);   // trees above a bound lowered by a peer are dropped in xmp_cuda_search_result
}

extern "C" int xmp_cuda_search_pending(xmp_ctx *ctx, unsigned long long *per_level) {
	if (!ctx || !ctx->search || !per_level) return set_error(XMP_ERR_ARG, "no search in progress");
	for (unsigned m = 0; m <= XMP_MAX_TAXA; ++m) per_level[m] = ctx->search->pool[m].count;
	return XMP_OK;
}

extern "C" int xmp_cuda_search_export_jobs(xmp_ctx *ctx, unsigned max_nodes, uint8_t *jobs, unsigned *n_jobs) {
	if (!ctx || !ctx->search || !jobs || !n_jobs) return set_error(XMP_ERR_ARG, "no search in progress");
	Search *S = ctx->search;
	XMP_CUDA(cudaSetDevice(ctx->device));
	*n_jobs = 0;
	unsigned k = 0;
	int rc = detach_jobs(ctx, S, max_nodes, 0, &k);
	if (rc || k == 0) return rc;
	XMP_CUDA(cudaMemcpyAsync(jobs, S->pathbuf.ptr, (size_t) k * S->L.n, cudaMemcpyDeviceToHost, ctx->stream));
	XMP_CUDA(cudaStreamSynchronize(ctx->stream));
	*n_jobs = k;
	return XMP_OK;
}

extern "C" int xmp_cuda_search_import_jobs(xmp_ctx *ctx, const uint8_t *jobs, unsigned n_jobs) {
	if (!ctx || !ctx->search || (!jobs && n_jobs)) return set_error(XMP_ERR_ARG, "no search in progress");
	Search *S = ctx->search;
	XMP_CUDA(cudaSetDevice(ctx->device));
	if (n_jobs == 0) return XMP_OK;
	for (unsigned j = 0; j < n_jobs; ++j) {
		unsigned m = jobs[(size_t) j * S->L.n];
		if (m < 3 || m >= S->n) return set_error(XMP_ERR_ARG, "job %u has tree size %u, expected 3..%u", j, m, S->n - 1);
		if (S->prm.shard_count > 1 && m < S->split_level) return set_error(XMP_ERR_ARG, "job %u has tree size %u, below the split level %u of this sharded search", j, m, S->split_level);
	}
	S->done = false;       // an idle shard that receives stolen work is searching again
	S->finished = false;
	return import_nodes(ctx, S, jobs, n_jobs);
}

extern "C" int xmp_cuda_search_result(xmp_ctx *ctx, unsigned *score, unsigned long long *n_trees, uint8_t *paths, size_t cap) {
	if (!ctx || !ctx->search) return set_error(XMP_ERR_ARG, "no search in progress");
	Search *S = ctx->search;
	if (!S->finished) return set_error(XMP_ERR_ARG, "the search has not finished");
	// Filter here, not when the frontier ran dry: a peer shard may have lowered the bound since.
	const Layout &L = S->L;
	unsigned long long count = 0;
	for (size_t r = 0; r < S->book.recs.size(); r += L.trec) {
		if (S->book.length_of(r) != S->bound) continue;
		if (paths && count < cap) memcpy(paths + (size_t) count * L.n, S->book.recs.data() + r + 4, L.n);
		++count;
	}
	if (score) *score = S->bound;
	if (n_trees) *n_trees = S->book.optimal(S->bound, nullptr);
	return XMP_OK;
}

extern "C" int xmp_cuda_search_stored(xmp_ctx *ctx, unsigned long long *n_stored) {
	if (!ctx || !ctx->search || !n_stored) return set_error(XMP_ERR_ARG, "no search in progress");
	Search *S = ctx->search;
	if (!S->finished) return set_error(XMP_ERR_ARG, "the search has not finished");
	S->book.optimal(S->bound, n_stored);
	return XMP_OK;
}

extern "C" int xmp_cuda_search_stats(xmp_ctx *ctx, xmp_search_stats *out) {
	if (!ctx || !ctx->search || !out) return set_error(XMP_ERR_ARG, "no search in progress");
	Search *S = ctx->search;
	unsigned long long merges = 0;
	XMP_CUDA(cudaSetDevice(ctx->device));
	XMP_CUDA(cudaMemcpyAsync(&merges, S->d_merges, sizeof merges, cudaMemcpyDeviceToHost, ctx->stream));
	XMP_CUDA(cudaStreamSynchronize(ctx->stream));
	*out = S->st;
	out->steals_served = S->steals_served;
	out->steals_received = S->steals_got;
	out->site_merges = (S->st.score_calls + merges) * ctx->real_sites;
	return XMP_OK;
}
