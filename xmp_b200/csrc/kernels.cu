// Batched Fitch kernels for sm_100a: insertion scoring of (partial tree, edge) pairs, pruning with
// ballot compaction, and construction of edge-indexed state-set arrays from tree topologies.
//
// What they replace in the reference: the per-pair CHOSEN_FitchScore call and acceptance test inside
// BandBScanEdge (src/yanbader.c:152-176), and the FitchBases passes of ConstructBaseTree /
// UpdateTreeAfterInsertion (src/common.c:559-620, src/yanbader.c:450-495).
//
// This is HBM-bound bitwise integer work: 0.5 B per site-merge for scoring.
// No tensor cores; the design points are coalesced 128-bit loads, enough independent loads in flight
// per SM to cover HBM latency, and one reduction per (pair, tile).
#include <stdlib.h>
#include <algorithm>
#include "ctx.h"

namespace xmp {

// ---------------------------------------------------------------------------------------------
// Wide state sets (n_blocks >= 32): one warp per (pair, tile); a tile is 32*K consecutive blocks.
// The taxon's tile lives in registers for the whole CTA lifetime and is reused for every pair the
// warp visits, so the only streamed operand is the edge state set: 16 B per lane per load,
// 512 B per warp per load, 2*K independent loads in flight per lane.
// Grid: x = tile, y = group of 8*pairs_per_warp pairs.
// ---------------------------------------------------------------------------------------------
template <int K>
__global__ void __launch_bounds__(256, 2)
k_score_wide(const uint4 *__restrict__ sets, unsigned long long tree_stride, unsigned pairs_per_tree,
             const uint4 *__restrict__ taxon_row, WeightPlanes wp, unsigned n_blocks,
             unsigned long long n_pairs, unsigned pairs_per_warp, unsigned *__restrict__ cost) {
	const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
	const unsigned col0 = blockIdx.x * (32u * K) + lane;
	const uint4 ones = make_uint4(~0u, ~0u, ~0u, ~0u);
	uint4 t[K];
#pragma unroll
	for (int k = 0; k < K; ++k) t[k] = (col0 + 32u * k < n_blocks) ? __ldg(taxon_row + col0 + 32u * k) : ones;
	const bool heavy = 4u * blockIdx.x * (32u * K) < wp.heavy_words;   // uniform over the CTA

	unsigned long long p = ((unsigned long long) blockIdx.y * 8u + warp) * pairs_per_warp;
	const unsigned long long p_end = min(p + pairs_per_warp, n_pairs);

	auto base_of = [&](unsigned long long pair) {
		return sets + (pair / pairs_per_tree) * tree_stride + (pair % pairs_per_tree) * (unsigned long long) n_blocks + col0;
	};
	auto tile_cost = [&](const uint4 (&v)[K]) {
		unsigned s = 0;
		if (!heavy) {
#pragma unroll
			for (int k = 0; k < K; ++k) {
				s += __popc(empty_mask(v[k].x, t[k].x)) + __popc(empty_mask(v[k].y, t[k].y)) +
				     __popc(empty_mask(v[k].z, t[k].z)) + __popc(empty_mask(v[k].w, t[k].w));
			}
		} else {
#pragma unroll
			for (int k = 0; k < K; ++k) s += cost_block(v[k], t[k], col0 + 32u * k, wp);
		}
		return warp_sum(s);
	};

	for (; p + 1 < p_end; p += 2) {
		const uint4 *s0 = base_of(p), *s1 = base_of(p + 1);
		uint4 v0[K], v1[K];
#pragma unroll
		for (int k = 0; k < K; ++k) v0[k] = (col0 + 32u * k < n_blocks) ? ld_stream(s0 + 32u * k) : ones;
#pragma unroll
		for (int k = 0; k < K; ++k) v1[k] = (col0 + 32u * k < n_blocks) ? ld_stream(s1 + 32u * k) : ones;
		unsigned c0 = tile_cost(v0), c1 = tile_cost(v1);
		if (lane == 0) {
			atomicAdd(cost + p, c0);
			atomicAdd(cost + p + 1, c1);
		}
	}
	if (p < p_end) {
		const uint4 *s0 = base_of(p);
		uint4 v0[K];
#pragma unroll
		for (int k = 0; k < K; ++k) v0[k] = (col0 + 32u * k < n_blocks) ? ld_stream(s0 + 32u * k) : ones;
		unsigned c0 = tile_cost(v0);
		if (lane == 0) atomicAdd(cost + p, c0);
	}
}

// Acceptance test + ballot compaction over finished costs (wide path only).
__global__ void __launch_bounds__(256)
k_prune(const unsigned *__restrict__ cost, unsigned long long n_pairs, unsigned pairs_per_tree,
        const unsigned *__restrict__ base_score, unsigned bound, unsigned rest,
        unsigned *__restrict__ survivors, unsigned *n_survivors) {
	unsigned long long p = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x;
	bool keep = false;
	if (p < n_pairs) {
		unsigned base = base_score ? __ldg(base_score + p / pairs_per_tree) : 0u;
		keep = accept_insertion(cost[p], bound, base, rest);
	}
	unsigned slot = warp_append(keep, n_survivors);
	if (keep) survivors[slot] = (unsigned) p;
}

// ---------------------------------------------------------------------------------------------
// Narrow state sets: G lanes per pair (G = 1, 2, 4, 8, 16, 32), each lane takes blocks sub, sub+G, ...
// The cost is complete inside the group, so scoring, the acceptance test and the ballot compaction
// are one kernel.
// ---------------------------------------------------------------------------------------------
template <int G>
__global__ void __launch_bounds__(256)
k_score_group(const uint4 *__restrict__ sets, unsigned long long tree_stride, unsigned pairs_per_tree,
              const uint4 *__restrict__ taxon_row, WeightPlanes wp, unsigned n_blocks, unsigned long long n_pairs,
              const unsigned *__restrict__ base_score, unsigned bound, unsigned rest,
              unsigned *__restrict__ cost, unsigned *__restrict__ survivors, unsigned *n_survivors) {
	__shared__ unsigned s_warp[9];
	const unsigned long long gthread = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x;
	const unsigned long long pair = gthread / G;
	const unsigned sub = (unsigned) (gthread % G);
	const bool valid = pair < n_pairs;
	unsigned s = 0;
	if (valid) {
		const uint4 *src = sets + (pair / pairs_per_tree) * tree_stride + (pair % pairs_per_tree) * (unsigned long long) n_blocks;
		for (unsigned b = sub; b < n_blocks; b += G) s += cost_block(ld_stream(src + b), __ldg(taxon_row + b), b, wp);
	}
#pragma unroll
	for (int o = G / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
	bool keep = false;
	if (valid && sub == 0) {
		cost[pair] = s;
		if (survivors) {
			unsigned base = base_score ? __ldg(base_score + pair / pairs_per_tree) : 0u;
			keep = accept_insertion(s, bound, base, rest);
		}
	}
	if (survivors) {      // uniform over the grid
		unsigned slot = block_append(keep, n_survivors, s_warp);
		if (keep) survivors[slot] = (unsigned) pair;
	}
}

template <int G>
static void launch_group(xmp_ctx *ctx, const uint4 *d_sets, unsigned long long tree_stride, unsigned ppt, const uint4 *row,
                         unsigned long long n_pairs, const unsigned *d_base, unsigned bound, unsigned rest, unsigned *d_cost,
                         unsigned *d_surv, unsigned *d_nsurv) {
	unsigned long long threads = n_pairs * G;
	unsigned blocks = (unsigned) ((threads + 255) / 256);
	k_score_group<G><<<blocks, 256, 0, ctx->stream>>>(d_sets, tree_stride, ppt, row, ctx->wp, ctx->n_blocks, n_pairs, d_base,
	                                                  bound, rest, d_cost, d_surv, d_nsurv);
}

template <int K>
static void launch_wide(xmp_ctx *ctx, const uint4 *d_sets, unsigned long long tree_stride, unsigned ppt, const uint4 *row,
                        unsigned long long n_pairs, unsigned *d_cost) {
	const unsigned W = ctx->n_blocks;
	const unsigned tiles = (W + 32 * K - 1) / (32 * K);
	unsigned ppw = 32;
	// Keep at least ~8 CTAs per SM in the grid so that the tail wave is short.
	while (ppw > 1 && (unsigned long long) tiles * ((n_pairs + 8ull * ppw - 1) / (8ull * ppw)) < 8ull * ctx->sm_count) ppw >>= 1;
	unsigned long long groups = (n_pairs + 8ull * ppw - 1) / (8ull * ppw);
	if (groups > 65535ull) {
		ppw = (unsigned) ((n_pairs + 8ull * 65535ull - 1) / (8ull * 65535ull));
		groups = (n_pairs + 8ull * ppw - 1) / (8ull * ppw);
	}
	dim3 grid(tiles, (unsigned) groups);
	k_score_wide<K><<<grid, 256, 0, ctx->stream>>>(d_sets, tree_stride, ppt, row, ctx->wp, W, n_pairs, ppw, d_cost);
}


// ---------------------------------------------------------------------------------------------
// The same scorer with the edge sets staged through shared memory by TMA bulk copies
// (cp.async.bulk.shared::cluster.global + mbarrier complete_tx), no register staging.  Every warp runs its
// own S-deep ring of 32*K-block tiles: lane 0 arms a stage's mbarrier with the byte count and issues the
// copy of the warp's pair i + S - 1 while the warp scores pair i out of the stage that has just landed.
// Selected with XMP_SCORE_TMA=1; profiles/r01_tma_vs_ldg.md has the measured comparison that decides the
// default.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned) __cvta_generic_to_shared(p); }

template <int K, int S>
__global__ void __launch_bounds__(256, S <= 3 ? 2 : 1)
k_score_wide_tma(const uint4 *__restrict__ sets, unsigned long long tree_stride, unsigned pairs_per_tree,
                 const uint4 *__restrict__ taxon_row, WeightPlanes wp, unsigned n_blocks,
                 unsigned long long n_pairs, unsigned pairs_per_warp, unsigned *__restrict__ cost) {
	extern __shared__ __align__(128) unsigned char dyn[];
	const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
	constexpr unsigned TILE = 32u * K;                         // blocks per tile
	uint4 *ring = (uint4 *) dyn + (size_t) warp * S * TILE;     // this warp's stages
	unsigned long long *bars = (unsigned long long *) ((uint4 *) dyn + (size_t) 8 * S * TILE) + warp * S;
	const unsigned tile0 = blockIdx.x * TILE;
	const unsigned tile_blocks = min(TILE, n_blocks - tile0);
	const unsigned tile_bytes = tile_blocks * 16u;
	const unsigned col0 = tile0 + lane;
	const uint4 ones = make_uint4(~0u, ~0u, ~0u, ~0u);
	uint4 t[K];
#pragma unroll
	for (int k = 0; k < K; ++k) t[k] = (col0 + 32u * k < n_blocks) ? __ldg(taxon_row + col0 + 32u * k) : ones;
	const bool heavy = 4u * tile0 < wp.heavy_words;

	const unsigned long long p0 = ((unsigned long long) blockIdx.y * 8u + warp) * pairs_per_warp;
	const unsigned long long p_end = min(p0 + pairs_per_warp, n_pairs);
	const unsigned count = p0 < p_end ? (unsigned) (p_end - p0) : 0u;
	auto src_of = [&](unsigned long long pair) {
		return sets + (pair / pairs_per_tree) * tree_stride + (pair % pairs_per_tree) * (unsigned long long) n_blocks + tile0;
	};
	if (lane == 0) {
		for (int s = 0; s < S; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bars + s)));
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncwarp();
	auto issue = [&](unsigned i) {      // lane 0 only: start the copy of this warp's i-th pair into stage i % S
		const unsigned s = i % S;
		const unsigned bar = smem_u32(bars + s);
		asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(tile_bytes) : "memory");
		asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
		             ::"r"(smem_u32(ring + (size_t) s * TILE)), "l"(src_of(p0 + i)), "r"(tile_bytes), "r"(bar) : "memory");
	};
	if (lane == 0) {
		for (unsigned i = 0; i < (unsigned) (S - 1) && i < count; ++i) issue(i);
	}
	for (unsigned i = 0; i < count; ++i) {
		if (lane == 0 && i + S - 1 < count) issue(i + S - 1);      // its stage was drained in iteration i - 1
		const unsigned s = i % S, parity = (i / S) & 1u;
		const unsigned bar = smem_u32(bars + s);
		unsigned done = 0;
		while (!done) {
			asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
			             : "=r"(done) : "r"(bar), "r"(parity) : "memory");
		}
		const uint4 *st = ring + (size_t) s * TILE + lane;
		unsigned c = 0;
#pragma unroll
		for (int k = 0; k < K; ++k) {
			const uint4 v = (32u * k + lane < tile_blocks) ? st[32 * k] : ones;
			if (!heavy) {
				c += __popc(empty_mask(v.x, t[k].x)) + __popc(empty_mask(v.y, t[k].y)) + __popc(empty_mask(v.z, t[k].z)) + __popc(empty_mask(v.w, t[k].w));
			} else {
				c += cost_block(v, t[k], col0 + 32u * k, wp);
			}
		}
		c = warp_sum(c);
		if (lane == 0) atomicAdd(cost + p0 + i, c);
		__syncwarp();        // every lane has read the stage before lane 0 refills it
	}
}

template <int K, int S>
static int launch_wide_tma(xmp_ctx *ctx, const uint4 *d_sets, unsigned long long tree_stride, unsigned ppt, const uint4 *row,
                           unsigned long long n_pairs, unsigned *d_cost) {
	const unsigned W = ctx->n_blocks;
	const unsigned tiles = (W + 32 * K - 1) / (32 * K);
	unsigned ppw = 64;
	while (ppw > 4 && (unsigned long long) tiles * ((n_pairs + 8ull * ppw - 1) / (8ull * ppw)) < 4ull * ctx->sm_count) ppw >>= 1;
	unsigned long long groups = (n_pairs + 8ull * ppw - 1) / (8ull * ppw);
	if (groups > 65535ull) {
		ppw = (unsigned) ((n_pairs + 8ull * 65535ull - 1) / (8ull * 65535ull));
		groups = (n_pairs + 8ull * ppw - 1) / (8ull * ppw);
	}
	const size_t smem = (size_t) 8 * S * 32 * K * 16 + 8 * S * sizeof(unsigned long long);
	XMP_CUDA(cudaFuncSetAttribute(k_score_wide_tma<K, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
	dim3 grid(tiles, (unsigned) groups);
	k_score_wide_tma<K, S><<<grid, 256, smem, ctx->stream>>>(d_sets, tree_stride, ppt, row, ctx->wp, W, n_pairs, ppw, d_cost);
	return XMP_OK;
}

int launch_score_sets(xmp_ctx *ctx, const uint4 *d_sets, unsigned long long tree_stride, size_t n_pairs, unsigned ppt,
                      unsigned taxon, const unsigned *d_base, unsigned bound, unsigned *d_cost, unsigned *d_surv,
                      unsigned *d_nsurv) {
	if (n_pairs == 0) return XMP_OK;
	const uint4 *row = ctx->d_rows + (size_t) taxon * ctx->n_blocks;
	const unsigned rest = ctx->rest_bound[taxon];
	const unsigned W = ctx->n_blocks;
	if (d_nsurv) XMP_CUDA(cudaMemsetAsync(d_nsurv, 0, sizeof(unsigned), ctx->stream));
	if (W >= 64) {
		XMP_CUDA(cudaMemsetAsync(d_cost, 0, n_pairs * sizeof(unsigned), ctx->stream));
		static const bool use_tma = getenv("XMP_SCORE_TMA") && atoi(getenv("XMP_SCORE_TMA")) != 0;
		static const int tma_variant = getenv("XMP_SCORE_TMA") ? atoi(getenv("XMP_SCORE_TMA")) : 0;
		if (W >= 256 && use_tma) {
			// 1: three stages, two CTAs per SM (the best TMA configuration measured); 2: four stages, one CTA per SM
			int rc = tma_variant == 2 ? launch_wide_tma<8, 4>(ctx, d_sets, tree_stride, ppt, row, n_pairs, d_cost)
			                          : launch_wide_tma<8, 3>(ctx, d_sets, tree_stride, ppt, row, n_pairs, d_cost);
			if (rc) return rc;
		} else if (W >= 256) launch_wide<8>(ctx, d_sets, tree_stride, ppt, row, n_pairs, d_cost);
		else if (W >= 128) launch_wide<4>(ctx, d_sets, tree_stride, ppt, row, n_pairs, d_cost);
		else launch_wide<2>(ctx, d_sets, tree_stride, ppt, row, n_pairs, d_cost);
		if (d_surv) {
			unsigned blocks = (unsigned) ((n_pairs + 255) / 256);
			k_prune<<<blocks, 256, 0, ctx->stream>>>(d_cost, n_pairs, ppt, d_base, bound, rest, d_surv, d_nsurv);
		}
	} else if (W > 16) {
		launch_group<32>(ctx, d_sets, tree_stride, ppt, row, n_pairs, d_base, bound, rest, d_cost, d_surv, d_nsurv);
	} else if (W > 8) {
		launch_group<16>(ctx, d_sets, tree_stride, ppt, row, n_pairs, d_base, bound, rest, d_cost, d_surv, d_nsurv);
	} else if (W > 4) {
		launch_group<8>(ctx, d_sets, tree_stride, ppt, row, n_pairs, d_base, bound, rest, d_cost, d_surv, d_nsurv);
	} else if (W > 2) {
		launch_group<4>(ctx, d_sets, tree_stride, ppt, row, n_pairs, d_base, bound, rest, d_cost, d_surv, d_nsurv);
	} else if (W > 1) {
		launch_group<2>(ctx, d_sets, tree_stride, ppt, row, n_pairs, d_base, bound, rest, d_cost, d_surv, d_nsurv);
	} else {
		launch_group<1>(ctx, d_sets, tree_stride, ppt, row, n_pairs, d_base, bound, rest, d_cost, d_surv, d_nsurv);
	}
	XMP_CUDA(cudaGetLastError());
	return XMP_OK;
}

int launch_score_batch(xmp_ctx *ctx, const uint4 *d_sets, size_t n_pairs, unsigned pairs_per_tree, unsigned taxon,
                       const unsigned *d_base_score, unsigned bound, unsigned *d_cost, unsigned *d_survivors,
                       unsigned *d_n_survivors) {
	if (pairs_per_tree == 0) pairs_per_tree = 1;
	return launch_score_sets(ctx, d_sets, (unsigned long long) pairs_per_tree * ctx->n_blocks, n_pairs, pairs_per_tree, taxon,
	                         d_base_score, bound, d_cost, d_survivors, d_n_survivors);
}

// ---------------------------------------------------------------------------------------------
// State sets of whole trees from their topologies: one thread per (tree, 128-bit column).
// UP sets in reverse pre-order, then DOWN and MIDPOINT in pre-order.
// ---------------------------------------------------------------------------------------------
// PLANAR: rows come in and sets go out in planar form (fitch_dev.cuh) -- a search run with XMP_SEARCH_PLANAR=1 builds its root and
// its imported subtrees this way; the public entry points (xmp_cuda_build_midpoints) always use the packed form.
template <int MAXE, bool PLANAR>
__global__ void __launch_bounds__(128)
k_build_sets(const uint4 *__restrict__ rows, unsigned n_blocks, WeightPlanes wp, const uint8_t *__restrict__ topo,
             unsigned topo_stride, unsigned n_trees, unsigned m, uint4 *__restrict__ out, unsigned long long rec_blocks,
             int store_all, unsigned *tree_len, unsigned constant_weight) {
	const unsigned long long idx = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x;
	const unsigned long long tree = idx / n_blocks;
	const unsigned col = (unsigned) (idx % n_blocks);
	const bool valid = tree < n_trees;
	const unsigned E = 2 * m - 3;
	unsigned len = 0;
	if (valid) {
		const uint8_t *par = topo + tree * 4ull * topo_stride, *k0 = par + topo_stride, *k1 = k0 + topo_stride, *pre = k1 + topo_stride;
		uint4 up[MAXE], down[MAXE];
		for (int i = (int) E - 1; i >= 0; --i) {
			unsigned e = pre[i];
			if (edge_is_leaf(e)) {
				up[e] = __ldg(rows + (size_t) edge_taxon(e) * n_blocks + col);
			} else {
				uint4 a = up[k0[e]], b = up[k1[e]];
				len += score_block<PLANAR>(a, b, col, wp);
				up[e] = merge_block<PLANAR>(a, b);
			}
		}
		const uint4 r0 = __ldg(rows + col);
		len += score_block<PLANAR>(up[pre[0]], r0, col, wp);
		uint4 *rec = out + tree * rec_blocks;
		for (unsigned i = 0; i < E; ++i) {
			unsigned e = pre[i], p = par[e];
			uint4 d;
			if (p == 0xFFu) {
				d = r0;
			} else {
				unsigned sib = (k0[p] == e) ? k1[p] : k0[p];
				d = merge_block<PLANAR>(down[p], up[sib]);
			}
			down[e] = d;
			rec[(size_t) e * n_blocks + col] = merge_block<PLANAR>(up[e], d);
			if (store_all) rec[((size_t) E + e) * n_blocks + col] = up[e];
		}
	}
	if (tree_len) {
		const unsigned long long t0 = __shfl_sync(0xffffffffu, tree, 0);
		const bool uniform = __all_sync(0xffffffffu, tree == t0);
		if (uniform) {
			len = warp_sum(len);
			if ((threadIdx.x & 31u) == 0 && valid) atomicAdd(tree_len + tree, len + (col == 0 ? constant_weight : 0u));
		} else if (valid) {
			atomicAdd(tree_len + tree, len + (col == 0 ? constant_weight : 0u));
		}
	}
}

// ---------------------------------------------------------------------------------------------
// Topologies -> insertion costs in one pass, for wide state sets: the MIDPOINT sets are formed on the fly and scored
// at once, never written to HBM (saves 1.5 B written + 0.5 B read per site-merge against k_build_sets followed by
// k_score_wide).  A CTA owns a tile of blockDim.x columns and walks a chunk of trees over it; every lane follows the
// same topology, every branch is uniform.
//
// All per-thread state with dynamic indexing lives in SHARED memory as "slots" of one 128-bit block per thread (slot s
// of thread t at sm[s * blockDim.x + t]: conflict-free):
//     [0, m]        the rows of taxa 0 .. m for this column tile -- loaded once per CTA, shared by all its trees
//     m + 1         all ones (the neutral element: F(x, ones) = x), so that the root's child needs no special case
//     then m - 2    UP sets of the internal edges, m - 2 DOWN sets of the internal edges, one slot nobody reads
// Per tree the first 2m-3 threads turn the topology into two small step tables of slot numbers (shared memory,
// read back as broadcasts), so the walk itself is branch-free: per pre-order position three slot loads, one store,
// two merges, one cost, one warp reduction (REDUX).  History (profiles/r02_ncu_k_score_trees_wide.md): the first
// version kept `uint4 up[2m-3], down[2m-3]` as local arrays -- ncu showed 55 GB of DRAM writes per launch for an
// algorithm that needs 5 MB, because the local-memory footprint of the resident threads exceeds L2; moving only the
// internal sets to shared memory removed the traffic but left 3,300 instructions per (tree, column) in branches and
// 64-bit address arithmetic around 1,300 of Fitch arithmetic.
// ---------------------------------------------------------------------------------------------
// Step tables (shared memory, read back as 16-byte broadcasts); all fields are BYTE offsets of slots, i.e. slot * 16 * blockDim.x.
//   down step of pre-order position i:  DOWN(e) = F(src, sib);  MIDPOINT = F(upe, DOWN);  DOWN is kept at dst
//   up step (internal edges only, children before parents):  slot dst = F(a, b)
// PLANAR = false is the packed form the kernel used until the end of round 2 (every GPU measurement of that round); kept
// selectable (XMP_E2E_PACKED=1) so that bench.py can time both forms in one run.
template <bool PLANAR>
__global__ void __launch_bounds__(128)
k_score_trees_wide(const uint4 *__restrict__ rows, unsigned n_blocks, WeightPlanes wp, const uint8_t *__restrict__ topo,
                   unsigned topo_stride, unsigned n_trees, unsigned trees_per_cta, unsigned m, unsigned taxon,
                   unsigned *__restrict__ cost) {
	extern __shared__ uint4 sm[];                              // [3m - 1 slots][blockDim.x], then the step tables
	const unsigned T = blockDim.x, E = 2 * m - 3, I = m - 2, tid = threadIdx.x, n_warps = T >> 5;
	const unsigned n_steps = trees_per_cta * E;
	uint4 *t_dn = sm + (size_t) (3 * m - 1) * T;               // [trees_per_cta][E]
	uint4 *t_up = t_dn + n_steps;                              // [trees_per_cta][I]
	unsigned *accw = (unsigned *) (t_up + trees_per_cta * I);  // [n_warps][n_steps]: per warp, written exactly once each
	uint8_t *s_edge = (uint8_t *) (accw + n_warps * n_steps);
	const unsigned col = blockIdx.x * T + tid;
	const bool ok = col < n_blocks;
	const unsigned ONES = m + 1, UP0 = m + 2, DN0 = UP0 + I, DUMMY = DN0 + I, SB = 16u * T;   // SB: bytes per slot
	const uint4 ones = make_uint4(~0u, ~0u, ~0u, ~0u);
	const unsigned t_begin = blockIdx.y * trees_per_cta, t_count = min(n_trees, t_begin + trees_per_cta) - t_begin;
	auto up_slot = [&](unsigned y) -> unsigned { return edge_is_leaf(y) ? edge_taxon(y) : UP0 + (y / 2 - 1); };
	// Down steps of all this CTA's trees, one thread per (tree, pre-order position): topology bytes -> byte offsets.
	for (unsigned x = tid; x < t_count * E; x += T) {
		const unsigned lt = x / E, i = x - lt * E;
		const uint8_t *tp = topo + (size_t) (t_begin + lt) * 4 * topo_stride;
		const unsigned e = tp[3 * topo_stride + i], p = tp[e];
		uint4 d;
		if (p == 0xFFu) {
			d.x = 0;                                         // the root is taxon 0: DOWN of its only child is that row ...
			d.y = ONES * SB;                                 // ... merged with the neutral element
		} else {
			const unsigned pk0 = tp[topo_stride + p], pk1 = tp[2 * topo_stride + p];
			d.x = (DN0 + (p / 2 - 1)) * SB;
			d.y = up_slot(pk0 == e ? pk1 : pk0) * SB;
		}
		d.z = up_slot(e) * SB;
		d.w = (edge_is_leaf(e) ? DUMMY : DN0 + (e / 2 - 1)) * SB;
		t_dn[x] = d;
		s_edge[x] = (uint8_t) e;
	}
	// Up steps: one thread per tree lists its internal edges in reverse pre-order.
	if (tid < t_count) {
		const uint8_t *tp = topo + (size_t) (t_begin + tid) * 4 * topo_stride;
		unsigned k = 0;
		for (int i = (int) E - 1; i >= 0; --i) {
			const unsigned e = tp[3 * topo_stride + i];
			if (edge_is_leaf(e)) continue;
			t_up[tid * I + k++] = make_uint4((UP0 + (e / 2 - 1)) * SB, up_slot(tp[topo_stride + e]) * SB, up_slot(tp[2 * topo_stride + e]) * SB, 0u);
		}
	}
	// Rows of this column tile: the scored taxon is row `taxon` (= m for the search's use), rows 0 .. m-1 are the leaves.
	// Columns past the end hold all ones in every row: they merge to all ones and cost nothing.
	// Every slot holds its block in PLANAR form (fitch_dev.cuh): the rows are transposed here, once per CTA, and from then on a merge
	// is 8 LOP3 instead of 28 instructions and a weight-1 cost 4 LOP3 + POPC instead of 20; only the costs leave the kernel.
	for (unsigned t = 0; t <= m; ++t) {
		const uint4 r = ok ? __ldg(rows + (size_t) (t == m ? taxon : t) * n_blocks + col) : ones;
		sm[t * T + tid] = PLANAR ? to_planar(r) : r;                 // to_planar(ones) == ones
	}
	sm[ONES * T + tid] = ones;
	const uint4 rx = sm[m * T + tid];
	const bool w1 = 4u * (ok ? col : 0u) >= wp.heavy_words;     // this column holds weight-1 sites only
	const unsigned me = (unsigned) __cvta_generic_to_shared(sm) + 16u * tid;           // this thread's column of slot 0
	const unsigned dn0 = (unsigned) __cvta_generic_to_shared(t_dn), up0 = (unsigned) __cvta_generic_to_shared(t_up);
	unsigned acc_at = (unsigned) __cvta_generic_to_shared(accw) + 4u * (tid >> 5) * n_steps;
	const bool lane0 = (tid & 31u) == 0;
	__syncthreads();
	for (unsigned lt = 0; lt < t_count; ++lt) {
		unsigned at = up0 + 16u * lt * I;
		for (unsigned k = 0; k < I; ++k, at += 16u) {           // UP sets of the internal edges, children before parents
			const uint4 u = lds128(at);
			const uint4 a = lds128(me + u.y), b = lds128(me + u.z);
			sts128(me + u.x, PLANAR ? fitch_block_planar(a, b) : fitch_block(a, b));
		}
		at = dn0 + 16u * lt * E;
		for (unsigned i = 0; i < E; ++i, at += 16u, acc_at += 4u) {   // DOWN and MIDPOINT in pre-order, scored on the spot
			const uint4 w = lds128(at);
			const uint4 a = lds128(me + w.x), b = lds128(me + w.y);
			const uint4 d = PLANAR ? fitch_block_planar(a, b) : fitch_block(a, b);
			if (w.w != DUMMY * SB) sts128(me + w.w, d);               // a leaf's DOWN set is read by nobody: 8 of 13 stores at m = 8 (uniform per warp)
			const uint4 ue = lds128(me + w.z);
			const uint4 mid = PLANAR ? fitch_block_planar(ue, d) : fitch_block(ue, d);
			unsigned c;
			if (PLANAR) {
				if (w1) c = cost_block_planar_w1(mid, rx);
				else c = cost_block_planar(mid, rx, col, wp);
			} else {
				if (w1) c = cost_word_w1(mid.x, rx.x) + cost_word_w1(mid.y, rx.y) + cost_word_w1(mid.z, rx.z) + cost_word_w1(mid.w, rx.w);
				else c = cost_block(mid, rx, col, wp);
			}
			c = __reduce_add_sync(0xffffffffu, c);
			if (lane0) sts32(acc_at, c);
		}
	}
	__syncthreads();
	for (unsigned x = tid; x < t_count * E; x += T) {
		unsigned c = 0;
		for (unsigned w = 0; w < n_warps; ++w) c += accw[w * n_steps + x];
		if (c) atomicAdd(cost + (size_t) (t_begin + x / E) * E + s_edge[x], c);
	}
}

// cost[tree][edge] of inserting `taxon` on every edge of trees whose topologies are in d_topo.
int launch_score_trees_wide(xmp_ctx *ctx, const uint8_t *d_topo, unsigned topo_stride, unsigned n_trees, unsigned m,
                            unsigned taxon, unsigned *d_cost) {
	const unsigned E = 2 * m - 3, slots = 3 * m - 1;
	// threads per CTA: as many as keep two CTAs per SM (3m - 1 slots of 16 bytes per thread)
	unsigned threads = 128;
	while (threads > 32 && (size_t) slots * 16 * threads > ((size_t) 100 << 10)) threads >>= 1;
	// trees per CTA: enough to amortise loading the rows and building the tables, few enough to fill the GPU several times over
	const unsigned col_tiles = (ctx->n_blocks + threads - 1) / threads;
	unsigned per_cta = 16;
	while (per_cta > 1 && (unsigned long long) col_tiles * ((n_trees + per_cta - 1) / per_cta) < 16ull * ctx->sm_count) per_cta >>= 1;
	// shared memory: the slots, then per tree E down steps + (m - 2) up steps of 16 bytes, one cost word per warp and step, E edge ids
	auto bytes = [&](unsigned trees) { return (size_t) slots * 16 * threads + (size_t) trees * ((size_t) E * (16 + 4 * (threads / 32) + 1) + (size_t) (m - 2) * 16) + 16; };
	while (per_cta > 1 && bytes(per_cta) > bytes(1) + ((size_t) 9 << 10)) per_cta >>= 1;      // tables: at most 9 KiB on top of the slots
	const size_t smem = bytes(per_cta);
	if (smem > ((size_t) 220 << 10)) return set_error(XMP_ERR_ARG, "trees on %u taxa need %zu bytes of shared memory per CTA", m, smem);
	const char *env = getenv("XMP_E2E_PACKED");
	const bool packed = env && env[0] == '1';                      // A/B switch for measurements: the packed-form kernel
	XMP_CUDA(cudaFuncSetAttribute(k_score_trees_wide<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
	XMP_CUDA(cudaFuncSetAttribute(k_score_trees_wide<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
	XMP_CUDA(cudaMemsetAsync(d_cost, 0, (size_t) n_trees * E * sizeof(unsigned), ctx->stream));
	for (unsigned t0 = 0; t0 < n_trees; t0 += 65535u * per_cta) {
		const unsigned nt = std::min(n_trees - t0, 65535u * per_cta);
		dim3 grid(col_tiles, (nt + per_cta - 1) / per_cta);
		if (packed)
			k_score_trees_wide<false><<<grid, threads, smem, ctx->stream>>>(ctx->d_rows, ctx->n_blocks, ctx->wp, d_topo + (size_t) t0 * 4 * topo_stride, topo_stride,
			                                                               nt, per_cta, m, taxon, d_cost + (size_t) t0 * E);
		else
			k_score_trees_wide<true><<<grid, threads, smem, ctx->stream>>>(ctx->d_rows, ctx->n_blocks, ctx->wp, d_topo + (size_t) t0 * 4 * topo_stride, topo_stride,
			                                                              nt, per_cta, m, taxon, d_cost + (size_t) t0 * E);
	}
	XMP_CUDA(cudaGetLastError());
	return XMP_OK;
}

int launch_build_sets(xmp_ctx *ctx, const uint8_t *d_topo, unsigned topo_stride, unsigned n_trees, unsigned m, uint4 *d_out,
                      size_t rec_blocks, bool store_all, unsigned *d_tree_len, const uint4 *planar_rows) {
	if (n_trees == 0) return XMP_OK;
	const unsigned E = 2 * m - 3;
	const unsigned long long threads = (unsigned long long) n_trees * ctx->n_blocks;
	const unsigned blocks = (unsigned) ((threads + 127) / 128);
	if (d_tree_len) XMP_CUDA(cudaMemsetAsync(d_tree_len, 0, (size_t) n_trees * sizeof(unsigned), ctx->stream));
#define XMP_BUILD(MAXE)                                                                                                                  \
	do {                                                                                                                                 \
		if (planar_rows)                                                                                                                 \
			k_build_sets<MAXE, true><<<blocks, 128, 0, ctx->stream>>>(planar_rows, ctx->n_blocks, ctx->wp, d_topo, topo_stride, n_trees, m,  \
			                                                          d_out, rec_blocks, store_all ? 1 : 0, d_tree_len, ctx->constant_weight); \
		else                                                                                                                             \
			k_build_sets<MAXE, false><<<blocks, 128, 0, ctx->stream>>>(ctx->d_rows, ctx->n_blocks, ctx->wp, d_topo, topo_stride, n_trees, m, \
			                                                           d_out, rec_blocks, store_all ? 1 : 0, d_tree_len, ctx->constant_weight); \
	} while (0)
	if (E <= 16) XMP_BUILD(16);
	else if (E <= 32) XMP_BUILD(32);
	else if (E <= 72) XMP_BUILD(72);
	else if (E <= 128) XMP_BUILD(128);
	else XMP_BUILD(200);
#undef XMP_BUILD
	XMP_CUDA(cudaGetLastError());
	return XMP_OK;
}

}  // namespace xmp
