// Device-side Fitch primitives on 32-bit words (8 sites) and 128-bit blocks (32 sites), sm_100a.
//
// Semantics: reference src/seq_b2_w1_basic.c:155-244 (per site: intersection if non-empty, else union
// and charge the site's weight).  The SWAR formulation differs from the reference's 64-bit one
// (src/seq_b2_w8_fastc64.c:11-23): the "empty" marker lives on bit 3 of every nibble so that it feeds
// POPC directly, and weighted scoring uses bit-sliced weights (one POPC per weight bit) instead of
// sixteen masked adds per word (src/fitchtemplate_b2_w8_fastc64.inc:118-152).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

// The Fitch arithmetic below is also compiled for the host, where a "not gpu" test replays the kernels' child
// walk (tests/child_walk_check.cu); loads through the read-only path and POPC have host stand-ins there.
#if defined(__CUDA_ARCH__)
#define XMP_LDG(p) __ldg(p)
#define XMP_POPC(x) __popc(x)
#else
#define XMP_LDG(p) (*(p))
#define XMP_POPC(x) ((unsigned) __builtin_popcount(x))
#endif
#define XMP_FHD __host__ __device__ __forceinline__

namespace xmp {

// Site weights, bit-sliced: plane k holds bit 3 of nibble s set iff bit k of weights[site s] is set.
// Words >= heavy_words carry only weight-1 sites (weights are sorted descending, padding weighs 1),
// so their cost is a plain population count -- the reference's FASTWEIGHT1FITCH hand-over
// (src/fitchtemplate_b2_w8_fastc64.inc:105-113), at 8-site instead of 16-site granularity.
struct WeightPlanes {
	const unsigned *planes;   // [n_planes][n_words]
	unsigned n_words;         // 32-bit words per state set = 4 * n_blocks
	unsigned heavy_words;
	unsigned n_planes;
};

// Bit 3 of each nibble set iff that site's sets do not intersect.
XMP_FHD unsigned empty_mask(unsigned a, unsigned b) {
	unsigned x = a & b;
	return ~(((x & 0x77777777u) + 0x77777777u) | x) & 0x88888888u;
}

// Preliminary Fitch set of a node from its two neighbours.
XMP_FHD unsigned fitch_word(unsigned a, unsigned b) {
	unsigned e = empty_mask(a, b);
	unsigned m = (e >> 3) * 15u;              // 0xF where the intersection is empty; a multiply so that it runs on the FMA pipe
	return (a & b) | ((a | b) & m);           //   (IMAD) beside the ALU pipe that LOP3 / SHF / IADD share: 6 + 1 instead of 7 + 1
}

XMP_FHD uint4 fitch_block(const uint4 &a, const uint4 &b) {
	return make_uint4(fitch_word(a.x, b.x), fitch_word(a.y, b.y), fitch_word(a.z, b.z), fitch_word(a.w, b.w));
}

XMP_FHD unsigned cost_word_w1(unsigned a, unsigned b) {
	return XMP_POPC(empty_mask(a, b));
}

XMP_FHD unsigned cost_word(unsigned a, unsigned b, unsigned word, const WeightPlanes &wp) {
	unsigned e = empty_mask(a, b);
	if (word >= wp.heavy_words) return XMP_POPC(e);
	unsigned s = 0;
	for (unsigned k = 0; k < wp.n_planes; ++k) s += XMP_POPC(e & XMP_LDG(wp.planes + (size_t) k * wp.n_words + word)) << k;
	return s;
}

// Cost of one 128-bit block; `block` is its index inside the state set.
XMP_FHD unsigned cost_block(const uint4 &a, const uint4 &b, unsigned block, const WeightPlanes &wp) {
	unsigned w = 4 * block;
	if (w >= wp.heavy_words) {
		return XMP_POPC(empty_mask(a.x, b.x)) + XMP_POPC(empty_mask(a.y, b.y)) + XMP_POPC(empty_mask(a.z, b.z)) + XMP_POPC(empty_mask(a.w, b.w));
	}
	return cost_word(a.x, b.x, w, wp) + cost_word(a.y, b.y, w + 1, wp) + cost_word(a.z, b.z, w + 2, wp) + cost_word(a.w, b.w, w + 3, wp);
}

// ---- planar blocks -------------------------------------------------------------------------------------------
// The same 128 bits per 32 sites, transposed inside the block: word k holds bit k (base k) of every site, the site in
// nibble j of word i sitting at bit 4j + i.  In this form the Fitch merge is pure three-input logic -- the sites whose
// sets intersect are the OR over the four planes of (a AND b), and every plane of the result is one LOP3 of its two
// inputs and that word -- 8 LOP3 per 32 sites where the packed form takes 4 x (4 LOP3 + SHF + add + IMAD), and a
// weight-1 cost is 4 LOP3 and one POPC where the packed form takes 4 x (3 LOP3 + add + POPC).  Used by kernels that
// keep their state sets to themselves (k_score_trees_wide: rows are transposed when a CTA loads them, costs come out
// the same); everything that crosses the C ABI stays in the reference's packed form (src/common.h:75-78).
// The transposition is a homomorphism: to_planar(fitch_block(a, b)) == fitch_block_planar(to_planar(a), to_planar(b)),
// and the costs agree; tests/child_walk_check.cu checks both on the host.
XMP_FHD unsigned planar_pick(unsigned c0, unsigned c1, unsigned c2, unsigned c3) {
	return (c0 & 0x11111111u) | (c1 & 0x22222222u) | (c2 & 0x44444444u) | (c3 & 0x88888888u);
}
XMP_FHD uint4 to_planar(const uint4 &v) {          // plane k, bit 4j + i  <-  bit 4j + k of word i: word i shifted by i - k
	return make_uint4(planar_pick(v.x, v.y << 1, v.z << 2, v.w << 3),
	                  planar_pick(v.x >> 1, v.y, v.z << 1, v.w << 2),
	                  planar_pick(v.x >> 2, v.y >> 1, v.z, v.w << 1),
	                  planar_pick(v.x >> 3, v.y >> 2, v.z >> 1, v.w));
}
// Bit s set iff the sets of site s intersect.
XMP_FHD unsigned planar_meet(const uint4 &a, const uint4 &b) { return (a.x & b.x) | (a.y & b.y) | (a.z & b.z) | (a.w & b.w); }
XMP_FHD uint4 fitch_block_planar(const uint4 &a, const uint4 &b) {
	const unsigned ne = planar_meet(a, b);         // where it is set the result is a AND b, elsewhere a OR b
	return make_uint4((a.x & b.x) | ((a.x | b.x) & ~ne), (a.y & b.y) | ((a.y | b.y) & ~ne),
	                  (a.z & b.z) | ((a.z | b.z) & ~ne), (a.w & b.w) | ((a.w | b.w) & ~ne));
}
XMP_FHD unsigned cost_block_planar_w1(const uint4 &a, const uint4 &b) { return XMP_POPC(~planar_meet(a, b)); }
// Weighted: the flags of word i of the packed block are bits i, 4 + i, 8 + i, ... of the planar flag word; put back on
// bit 3 of their nibbles they meet the bit-sliced weight planes as in cost_word().
XMP_FHD unsigned cost_block_planar(const uint4 &a, const uint4 &b, unsigned block, const WeightPlanes &wp) {
	const unsigned n = ~planar_meet(a, b), w = 4 * block;
	if (w >= wp.heavy_words) return XMP_POPC(n);
	unsigned s = 0;
	for (unsigned i = 0; i < 4; ++i) {
		const unsigned e = ((n >> i) & 0x11111111u) << 3;
		if (w + i >= wp.heavy_words) s += XMP_POPC(e);
		else
			for (unsigned k = 0; k < wp.n_planes; ++k) s += XMP_POPC(e & XMP_LDG(wp.planes + (size_t) k * wp.n_words + w + i)) << k;
	}
	return s;
}

// The same three operations in whichever form a kernel's template parameter selects.
template <bool PLANAR>
XMP_FHD uint4 merge_block(const uint4 &a, const uint4 &b) { return PLANAR ? fitch_block_planar(a, b) : fitch_block(a, b); }
template <bool PLANAR>
XMP_FHD unsigned score_block(const uint4 &a, const uint4 &b, unsigned block, const WeightPlanes &wp) {
	return PLANAR ? cost_block_planar(a, b, block, wp) : cost_block(a, b, block, wp);
}
template <bool PLANAR>
XMP_FHD unsigned score_block_w1(const uint4 &a, const uint4 &b) {
	return PLANAR ? cost_block_planar_w1(a, b) : cost_word_w1(a.x, b.x) + cost_word_w1(a.y, b.y) + cost_word_w1(a.z, b.z) + cost_word_w1(a.w, b.w);
}

// Streaming 128-bit load: read once, do not keep in L1.
__device__ __forceinline__ uint4 ld_stream(const uint4 *p) {
	uint4 r;
	asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
	return r;
}

// Shared memory by 32-bit address (__cvta_generic_to_shared): a walk computes every address as (this thread's base) + (a byte
// offset), one add per access, instead of letting the compiler rebuild generic pointers to dynamic shared memory inside its loop.
__device__ __forceinline__ uint4 lds128(unsigned addr) {
	uint4 v;
	asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
	return v;
}
__device__ __forceinline__ void sts128(unsigned addr, const uint4 &v) {
	asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void sts32(unsigned addr, unsigned v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }

__device__ __forceinline__ unsigned warp_sum(unsigned v) {
	v += __shfl_xor_sync(0xffffffffu, v, 16);
	v += __shfl_xor_sync(0xffffffffu, v, 8);
	v += __shfl_xor_sync(0xffffffffu, v, 4);
	v += __shfl_xor_sync(0xffffffffu, v, 2);
	v += __shfl_xor_sync(0xffffffffu, v, 1);
	return v;
}

// The acceptance test of BandBScanEdge, in the reference's own signed arithmetic
// (src/yanbader.c:115-116,152,176): ties are kept.
__device__ __forceinline__ bool accept_insertion(unsigned cost, unsigned bound, unsigned score, unsigned rest) {
	int allowable = (int) (bound - score - rest);
	return (int) cost <= allowable;
}

// Warp-aggregated append: every lane with keep == true gets a distinct slot after *counter.
__device__ __forceinline__ unsigned warp_append(bool keep, unsigned *counter) {
	unsigned ballot = __ballot_sync(0xffffffffu, keep);
	unsigned lane = threadIdx.x & 31u;
	unsigned base = 0;
	if (ballot) {
		int leader = __ffs(ballot) - 1;
		if ((int) lane == leader) base = atomicAdd(counter, (unsigned) __popc(ballot));
		base = __shfl_sync(0xffffffffu, base, leader);
	}
	return base + __popc(ballot & ((1u << lane) - 1u));
}

// Block-aggregated append for 256-thread CTAs: ONE atomic on *counter per CTA instead of one per warp.
// A counter that every warp of a large grid bumps serialises at a single L2 slice (~0.5 ns per atomic):
// at 64 K warps per launch that alone was most of the scoring kernel's time on small batches.
// Must be called by all threads of the block.
__device__ __forceinline__ unsigned block_append(bool keep, unsigned *counter, unsigned *s_warp /*[9]*/) {
	const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, n_warps = (blockDim.x + 31u) >> 5;
	const unsigned ballot = __ballot_sync(0xffffffffu, keep);
	if (lane == 0) s_warp[warp] = (unsigned) __popc(ballot);
	__syncthreads();
	if (threadIdx.x == 0) {
		unsigned total = 0;
		for (unsigned w = 0; w < n_warps; ++w) {
			unsigned c = s_warp[w];
			s_warp[w] = total;
			total += c;
		}
		s_warp[8] = total ? atomicAdd(counter, total) : 0u;
	}
	__syncthreads();
	return s_warp[8] + s_warp[warp] + __popc(ballot & ((1u << lane) - 1u));
}

// Edge-id helpers (include/xmp_cuda.h "Edge ids").
__device__ __host__ __forceinline__ bool edge_is_leaf(unsigned e) { return e == 0 || (e & 1u); }
__device__ __host__ __forceinline__ unsigned edge_taxon(unsigned e) { return (e + 3u) >> 1; }

}  // namespace xmp
