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
template <int MAXE>
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
				len += cost_block(a, b, col, wp);
				up[e] = fitch_block(a, b);
			}
		}
		const uint4 r0 = __ldg(rows + col);
		len += cost_block(up[pre[0]], r0, col, wp);
		uint4 *rec = out + tree * rec_blocks;
		for (unsigned i = 0; i < E; ++i) {
			unsigned e = pre[i], p = par[e];
			uint4 d;
			if (p == 0xFFu) {
				d = r0;
			} else {
				unsigned sib = (k0[p] == e) ? k1[p] : k0[p];
				d = fitch_block(down[p], up[sib]);
			}
			down[e] = d;
			rec[(size_t) e * n_blocks + col] = fitch_block(up[e], d);
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
// Topologies -> insertion costs in one pass, for wide state sets: the MIDPOINT sets are formed in
// registers / local arrays and scored at once, never written to HBM (saves 1.5 B written + 0.5 B read
// per site-merge against k_build_sets followed by k_score_wide).  One CTA = one tree x 256 columns, so
// every lane of a warp follows the same topology (coalesced local arrays); per-edge costs are summed by
// warp shuffle, then shared-memory atomics, then one global RED per (tree, edge, tile).
// ---------------------------------------------------------------------------------------------
template <int MAXE>
__global__ void __launch_bounds__(256)
k_score_trees_wide(const uint4 *__restrict__ rows, unsigned n_blocks, WeightPlanes wp, const uint8_t *__restrict__ topo,
                   unsigned topo_stride, unsigned m, unsigned taxon, unsigned *__restrict__ cost) {
	__shared__ unsigned acc[MAXE];
	__shared__ uint8_t s_par[MAXE], s_k0[MAXE], s_k1[MAXE], s_pre[MAXE];
	const unsigned tree = blockIdx.y, E = 2 * m - 3;
	const unsigned col = blockIdx.x * 256u + threadIdx.x;
	const bool ok = col < n_blocks;
	const uint8_t *t = topo + (size_t) tree * 4 * topo_stride;
	for (unsigned i = threadIdx.x; i < E; i += 256) {
		s_par[i] = t[i];
		s_k0[i] = t[topo_stride + i];
		s_k1[i] = t[2 * topo_stride + i];
		s_pre[i] = t[3 * topo_stride + i];
		acc[i] = 0;
	}
	__syncthreads();
	const uint4 ones = make_uint4(~0u, ~0u, ~0u, ~0u);
	uint4 up[MAXE], down[MAXE];
	for (int i = (int) E - 1; i >= 0; --i) {
		const unsigned e = s_pre[i];
		if (edge_is_leaf(e)) up[e] = ok ? __ldg(rows + (size_t) edge_taxon(e) * n_blocks + col) : ones;
		else up[e] = fitch_block(up[s_k0[e]], up[s_k1[e]]);
	}
	const uint4 r0 = ok ? __ldg(rows + col) : ones;
	const uint4 rx = ok ? __ldg(rows + (size_t) taxon * n_blocks + col) : ones;
	for (unsigned i = 0; i < E; ++i) {
		const unsigned e = s_pre[i], p = s_par[e];
		uint4 d;
		if (p == 0xFFu) {
			d = r0;
		} else {
			const unsigned sib = (s_k0[p] == e) ? s_k1[p] : s_k0[p];
			d = fitch_block(down[p], up[sib]);
		}
		down[e] = d;
		unsigned c = cost_block(fitch_block(up[e], d), rx, ok ? col : 0u, wp);
		c = warp_sum(ok ? c : 0u);
		if ((threadIdx.x & 31u) == 0 && c) atomicAdd(&acc[e], c);
	}
	__syncthreads();
	for (unsigned e = threadIdx.x; e < E; e += 256) {
		if (acc[e]) atomicAdd(cost + (size_t) tree * E + e, acc[e]);
	}
}

// cost[tree][edge] of inserting `taxon` on every edge of trees whose topologies are in d_topo.
int launch_score_trees_wide(xmp_ctx *ctx, const uint8_t *d_topo, unsigned topo_stride, unsigned n_trees, unsigned m,
                            unsigned taxon, unsigned *d_cost) {
	const unsigned E = 2 * m - 3;
	XMP_CUDA(cudaMemsetAsync(d_cost, 0, (size_t) n_trees * E * sizeof(unsigned), ctx->stream));
	for (unsigned t0 = 0; t0 < n_trees; t0 += 65535u) {
		const unsigned nt = n_trees - t0 < 65535u ? n_trees - t0 : 65535u;
		dim3 grid((ctx->n_blocks + 255) / 256, nt);
		const uint8_t *topo = d_topo + (size_t) t0 * 4 * topo_stride;
		unsigned *cost = d_cost + (size_t) t0 * E;
		if (E <= 16) k_score_trees_wide<16><<<grid, 256, 0, ctx->stream>>>(ctx->d_rows, ctx->n_blocks, ctx->wp, topo, topo_stride, m, taxon, cost);
		else if (E <= 32) k_score_trees_wide<32><<<grid, 256, 0, ctx->stream>>>(ctx->d_rows, ctx->n_blocks, ctx->wp, topo, topo_stride, m, taxon, cost);
		else if (E <= 72) k_score_trees_wide<72><<<grid, 256, 0, ctx->stream>>>(ctx->d_rows, ctx->n_blocks, ctx->wp, topo, topo_stride, m, taxon, cost);
		else k_score_trees_wide<200><<<grid, 256, 0, ctx->stream>>>(ctx->d_rows, ctx->n_blocks, ctx->wp, topo, topo_stride, m, taxon, cost);
	}
	XMP_CUDA(cudaGetLastError());
	return XMP_OK;
}

int launch_build_sets(xmp_ctx *ctx, const uint8_t *d_topo, unsigned topo_stride, unsigned n_trees, unsigned m, uint4 *d_out,
                      size_t rec_blocks, bool store_all, unsigned *d_tree_len) {
	if (n_trees == 0) return XMP_OK;
	const unsigned E = 2 * m - 3;
	const unsigned long long threads = (unsigned long long) n_trees * ctx->n_blocks;
	const unsigned blocks = (unsigned) ((threads + 127) / 128);
	if (d_tree_len) XMP_CUDA(cudaMemsetAsync(d_tree_len, 0, (size_t) n_trees * sizeof(unsigned), ctx->stream));
#define XMP_BUILD(MAXE)                                                                                             \
	k_build_sets<MAXE><<<blocks, 128, 0, ctx->stream>>>(ctx->d_rows, ctx->n_blocks, ctx->wp, d_topo, topo_stride, n_trees, m, \
	                                                    d_out, rec_blocks, store_all ? 1 : 0, d_tree_len, ctx->constant_weight)
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
