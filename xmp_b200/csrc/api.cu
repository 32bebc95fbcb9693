// C-ABI entry points of include/xmp_cuda.h except the search (search.cu): context, problem upload,
// the per-pair `_b2_w16_cuda` Fitch family, and the batched scoring wrappers.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <mutex>
#include "ctx.h"

namespace xmp {

static thread_local char g_err[512] = "";

int set_error(int code, const char *fmt, ...) {
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(g_err, sizeof g_err, fmt, ap);
	va_end(ap);
	return code;
}

int cuda_failed(cudaError_t e, const char *what, const char *file, int line) {
	return set_error(XMP_ERR_CUDA, "CUDA error %d (%s) at %s:%d in %s", (int) e, cudaGetErrorString(e), file, line, what);
}

int Scratch::reserve(size_t need) {
	if (need <= bytes) return XMP_OK;
	if (ptr) cudaFree(ptr);
	ptr = nullptr;
	bytes = 0;
	size_t want = need + need / 4 + 256;
	cudaError_t e = cudaMalloc(&ptr, want);
	if (e != cudaSuccess) {
		ptr = nullptr;
		return set_error(XMP_ERR_NOMEM, "cudaMalloc of %zu scratch bytes failed: %s", want, cudaGetErrorString(e));
	}
	bytes = want;
	return XMP_OK;
}

void Scratch::release() {
	if (ptr) cudaFree(ptr);
	ptr = nullptr;
	bytes = 0;
}

void HostTopo::base() {
	m = 3;
	top = 2;
	memset(par, 0xFF, sizeof par);
	memset(kid, 0xFF, sizeof kid);
	par[0] = par[1] = 2;
	kid[0][2] = 0;
	kid[1][2] = 1;
}

void HostTopo::insert(unsigned taxon, unsigned v) {
	const unsigned a = 2 * taxon - 3, b = 2 * taxon - 2, pv = par[v];
	par[b] = (uint8_t) pv;
	if (pv == 0xFFu) top = b;
	else kid[kid[1][pv] == v][pv] = (uint8_t) b;
	kid[0][b] = (uint8_t) a;
	kid[1][b] = (uint8_t) v;
	par[a] = par[v] = (uint8_t) b;
	m = taxon + 1;
}

unsigned HostTopo::preorder(uint8_t *out) const {
	uint8_t stack[2 * XMP_MAX_TAXA];
	unsigned sp = 0, n = 0;
	stack[sp++] = (uint8_t) top;
	while (sp) {
		unsigned e = stack[--sp];
		out[n++] = (uint8_t) e;
		if (!edge_is_leaf(e)) {
			stack[sp++] = kid[1][e];
			stack[sp++] = kid[0][e];
		}
	}
	return n;
}

bool HostTopo::from_path(const uint8_t *path, unsigned m_) {
	base();
	for (unsigned t = 3; t < m_; ++t) {
		if (path[t] >= 2 * t - 3) return false;
		insert(t, path[t]);
	}
	return true;
}

}  // namespace xmp

using namespace xmp;

extern "C" const char *xmp_cuda_last_error(void) { return g_err; }

extern "C" int xmp_cuda_device_count(void) {
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
	return n;
}

extern "C" int xmp_cuda_init(int device, xmp_ctx **out) {
	int n = 0;
	cudaError_t e = cudaGetDeviceCount(&n);
	if (e != cudaSuccess || n == 0) {
		return set_error(XMP_ERR_NODEVICE, "no CUDA device available (%s); libxmp_b200 has no CPU fallback", cudaGetErrorString(e));
	}
	if (device < 0 || device >= n) return set_error(XMP_ERR_ARG, "device %d out of range (0..%d)", device, n - 1);
	cudaDeviceProp prop;
	XMP_CUDA(cudaGetDeviceProperties(&prop, device));
	if (prop.major != 10) {
		return set_error(XMP_ERR_NODEVICE, "device %d is sm_%d%d; this library contains sm_100a code only", device, prop.major, prop.minor);
	}
	XMP_CUDA(cudaSetDevice(device));
	xmp_ctx *ctx = new xmp_ctx();
	ctx->device = device;
	ctx->sm_count = prop.multiProcessorCount;
	cudaError_t e2 = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking);
	if (e2 == cudaSuccess) e2 = cudaEventCreate(&ctx->ev0);
	if (e2 == cudaSuccess) e2 = cudaEventCreate(&ctx->ev1);
	if (e2 != cudaSuccess) {
		delete ctx;
		return cuda_failed(e2, "context creation", __FILE__, __LINE__);
	}
	ctx->stream = ctx->own_stream;
	*out = ctx;
	return XMP_OK;
}

// keep_rows: leave the owned row buffer in place so that a reload of the same shape reuses it.
static void free_problem(xmp_ctx *ctx, bool keep_rows = false) {
	if (ctx->d_rows_owned && !keep_rows) {
		cudaFree(ctx->d_rows_owned);
		ctx->d_rows_owned = nullptr;
		ctx->rows_owned_bytes = 0;
	}
	if (ctx->d_planes) cudaFree(ctx->d_planes);
	if (ctx->d_rest_bound) cudaFree(ctx->d_rest_bound);
	ctx->d_rows = nullptr;
	ctx->d_planes = nullptr;
	ctx->d_rest_bound = nullptr;
	ctx->wp = WeightPlanes{nullptr, 0, 0, 0};
}

extern "C" void xmp_cuda_destroy(xmp_ctx *ctx) {
	if (!ctx) return;
	cudaSetDevice(ctx->device);
	cudaStreamSynchronize(ctx->stream);
	search_destroy(ctx);
	free_problem(ctx);
	xch_release(ctx);
	if (ctx->arena) cudaFree(ctx->arena);
	if (ctx->h_pin) cudaFreeHost(ctx->h_pin);
	ctx->scratch.release();
	ctx->scratch2.release();
	if (ctx->ev0) cudaEventDestroy(ctx->ev0);
	if (ctx->ev1) cudaEventDestroy(ctx->ev1);
	if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
	delete ctx;
}

extern "C" int xmp_cuda_set_stream(xmp_ctx *ctx, void *stream) {
	if (!ctx) return set_error(XMP_ERR_ARG, "null context");
	ctx->stream = stream ? (cudaStream_t) stream : ctx->own_stream;
	return XMP_OK;
}

extern "C" int xmp_cuda_synchronize(xmp_ctx *ctx) {
	if (!ctx) return set_error(XMP_ERR_ARG, "null context");
	XMP_CUDA(cudaStreamSynchronize(ctx->stream));
	return XMP_OK;
}

// Bit-slices the weights (fitch_dev.cuh WeightPlanes) and uploads everything but the rows.
static int load_common(xmp_ctx *ctx, const unsigned *weights, unsigned num_taxa, unsigned n_blocks,
                       const unsigned *rest_bound, unsigned constant_weight) {
	if (num_taxa < 3 || num_taxa > XMP_MAX_TAXA) return set_error(XMP_ERR_ARG, "num_taxa %u out of range (3..%d)", num_taxa, XMP_MAX_TAXA);
	if (n_blocks == 0) return set_error(XMP_ERR_ARG, "n_blocks must be positive (an alignment with no informative site has nothing to search)");
	ctx->num_taxa = num_taxa;
	ctx->n_blocks = n_blocks;
	ctx->constant_weight = constant_weight;
	const unsigned n_words = 4 * n_blocks;
	unsigned heavy_words = 0, n_planes = 0;
	if (weights) {
		unsigned wmax = 1;
		for (size_t s = 0; s < (size_t) n_words * 8; ++s) {
			if (weights[s] != 1) heavy_words = (unsigned) (s / 8) + 1;
			if (weights[s] > wmax) wmax = weights[s];
		}
		while (n_planes < 32 && (wmax >> n_planes)) ++n_planes;
	}
	if (heavy_words) {
		std::vector<unsigned> planes((size_t) n_planes * n_words, 0u);
		for (unsigned w = 0; w < heavy_words; ++w) {
			for (unsigned nib = 0; nib < 8; ++nib) {
				unsigned wt = weights[(size_t) w * 8 + nib];
				for (unsigned k = 0; k < n_planes; ++k) {
					if ((wt >> k) & 1u) planes[(size_t) k * n_words + w] |= 8u << (4 * nib);
				}
			}
		}
		XMP_CUDA(cudaMalloc(&ctx->d_planes, planes.size() * sizeof(unsigned)));
		XMP_CUDA(cudaMemcpyAsync(ctx->d_planes, planes.data(), planes.size() * sizeof(unsigned), cudaMemcpyHostToDevice, ctx->stream));
		XMP_CUDA(cudaStreamSynchronize(ctx->stream));
	}
	ctx->wp = WeightPlanes{ctx->d_planes, n_words, heavy_words, n_planes};
	memset(ctx->rest_bound, 0, sizeof ctx->rest_bound);
	if (rest_bound) memcpy(ctx->rest_bound, rest_bound, num_taxa * sizeof(unsigned));
	ctx->real_sites = (unsigned long long) n_blocks * 32;
	return XMP_OK;
}

extern "C" int xmp_cuda_load(xmp_ctx *ctx, const void *rows, const unsigned *weights, unsigned num_taxa, unsigned n_blocks,
                             const unsigned *rest_bound, unsigned constant_weight) {
	if (!ctx || !rows) return set_error(XMP_ERR_ARG, "null argument");
	XMP_CUDA(cudaSetDevice(ctx->device));
	search_destroy(ctx);
	const size_t bytes = (size_t) num_taxa * n_blocks * 16;
	free_problem(ctx, ctx->rows_owned_bytes >= bytes);
	int rc = load_common(ctx, weights, num_taxa, n_blocks, rest_bound, constant_weight);
	if (rc) return rc;
	if (!ctx->d_rows_owned) {
		XMP_CUDA(cudaMalloc(&ctx->d_rows_owned, bytes));
		ctx->rows_owned_bytes = bytes;
	}
	XMP_CUDA(cudaMemcpyAsync(ctx->d_rows_owned, rows, bytes, cudaMemcpyHostToDevice, ctx->stream));
	XMP_CUDA(cudaStreamSynchronize(ctx->stream));
	ctx->d_rows = ctx->d_rows_owned;
	// Padding = trailing sites that are fully ambiguous in every row (a real such site is constant
	// and never reaches the search); they are excluded from the site-merge accounting.
	const uint8_t *r = (const uint8_t *) rows;
	unsigned long long sites = (unsigned long long) n_blocks * 32;
	while (sites > 0) {
		size_t s = (size_t) sites - 1;
		bool pad = true;
		for (unsigned t = 0; t < num_taxa && pad; ++t) {
			uint8_t byte = r[(size_t) t * n_blocks * 16 + s / 2];
			pad = (((s & 1) ? byte >> 4 : byte & 15) == 15);
		}
		if (!pad) break;
		--sites;
	}
	ctx->real_sites = sites;
	return XMP_OK;
}

extern "C" int xmp_cuda_load_device_rows(xmp_ctx *ctx, const void *d_rows, const unsigned *weights, unsigned num_taxa,
                                         unsigned n_blocks, const unsigned *rest_bound, unsigned constant_weight) {
	if (!ctx || !d_rows) return set_error(XMP_ERR_ARG, "null argument");
	XMP_CUDA(cudaSetDevice(ctx->device));
	search_destroy(ctx);
	free_problem(ctx);
	int rc = load_common(ctx, weights, num_taxa, n_blocks, rest_bound, constant_weight);
	if (rc) return rc;
	ctx->d_rows = (const uint4 *) d_rows;
	return XMP_OK;
}

// ---- batched wrappers ----------------------------------------------------------------------------
static int upload_topologies(xmp_ctx *ctx, const uint8_t *paths, size_t path_stride, unsigned n_trees, unsigned m,
                             Scratch &dst, unsigned *stride_out) {
	const unsigned E = 2 * m - 3, stride = (E + 15u) & ~15u;
	std::vector<uint8_t> h((size_t) n_trees * 4 * stride, 0xFF);
	HostTopo t;
	for (unsigned i = 0; i < n_trees; ++i) {
		if (!t.from_path(paths + i * path_stride, m)) return set_error(XMP_ERR_ARG, "tree %u: insertion path names an edge that does not exist", i);
		uint8_t *row = h.data() + (size_t) i * 4 * stride;
		memcpy(row, t.par, E);
		memcpy(row + stride, t.kid[0], E);
		memcpy(row + 2 * stride, t.kid[1], E);
		t.preorder(row + 3 * stride);
	}
	int rc = dst.reserve(h.size());
	if (rc) return rc;
	XMP_CUDA(cudaMemcpyAsync(dst.ptr, h.data(), h.size(), cudaMemcpyHostToDevice, ctx->stream));
	XMP_CUDA(cudaStreamSynchronize(ctx->stream));   // h goes out of scope
	*stride_out = stride;
	return XMP_OK;
}

extern "C" int xmp_cuda_build_midpoints(xmp_ctx *ctx, const uint8_t *paths, size_t path_stride, unsigned n_trees, unsigned m,
                                        void *d_mid, unsigned *d_tree_len) {
	if (!ctx || !ctx->d_rows) return set_error(XMP_ERR_ARG, "no problem loaded");
	if (m < 3 || m > ctx->num_taxa) return set_error(XMP_ERR_ARG, "tree size %u out of range (3..%u)", m, ctx->num_taxa);
	if (n_trees == 0) return XMP_OK;
	if (!paths || !d_mid) return set_error(XMP_ERR_ARG, "null argument");
	XMP_CUDA(cudaSetDevice(ctx->device));
	unsigned stride = 0;
	int rc = upload_topologies(ctx, paths, path_stride, n_trees, m, ctx->scratch, &stride);
	if (rc) return rc;
	return launch_build_sets(ctx, (const uint8_t *) ctx->scratch.ptr, stride, n_trees, m, (uint4 *) d_mid,
	                         (size_t) (2 * m - 3) * ctx->n_blocks, false, d_tree_len);
}

extern "C" int xmp_cuda_score_batch(xmp_ctx *ctx, const void *d_sets, size_t n_pairs, unsigned pairs_per_tree, unsigned taxon,
                                    const unsigned *d_base_score, unsigned bound, unsigned *d_cost, unsigned *d_survivors,
                                    unsigned *d_n_survivors) {
	if (!ctx || !ctx->d_rows) return set_error(XMP_ERR_ARG, "no problem loaded");
	if (taxon >= ctx->num_taxa) return set_error(XMP_ERR_ARG, "taxon %u out of range", taxon);
	if (n_pairs == 0) {
		if (d_n_survivors) XMP_CUDA(cudaMemsetAsync(d_n_survivors, 0, sizeof(unsigned), ctx->stream));
		return XMP_OK;
	}
	if (!d_sets || !d_cost || (d_survivors && !d_n_survivors)) return set_error(XMP_ERR_ARG, "null argument");
	if (n_pairs > 0xFFFFFFFFull) return set_error(XMP_ERR_ARG, "at most 2^32-1 pairs per launch");
	XMP_CUDA(cudaSetDevice(ctx->device));
	return launch_score_batch(ctx, (const uint4 *) d_sets, n_pairs, pairs_per_tree, taxon, d_base_score, bound, d_cost,
	                          d_survivors, d_n_survivors);
}

extern "C" int xmp_cuda_score_insertions_host(xmp_ctx *ctx, const uint8_t *paths, size_t path_stride, unsigned n_trees,
                                              unsigned m, unsigned *cost_out) {
	if (!ctx || !ctx->d_rows) return set_error(XMP_ERR_ARG, "no problem loaded");
	if (m < 3 || m >= ctx->num_taxa) return set_error(XMP_ERR_ARG, "tree size %u out of range (3..%u)", m, ctx->num_taxa - 1);
	if (n_trees == 0) return XMP_OK;
	XMP_CUDA(cudaSetDevice(ctx->device));
	const unsigned E = 2 * m - 3;
	const size_t n_pairs = (size_t) n_trees * E;
	if (ctx->n_blocks >= 256) {
		// Wide state sets: topologies -> costs in one fused pass, nothing but the costs touches HBM.
		unsigned stride = 0;
		int rc = upload_topologies(ctx, paths, path_stride, n_trees, m, ctx->scratch, &stride);
		if (rc) return rc;
		rc = ctx->scratch2.reserve(n_pairs * sizeof(unsigned) + 256);
		if (rc) return rc;
		unsigned *d_cost = (unsigned *) ctx->scratch2.ptr;
		rc = launch_score_trees_wide(ctx, (const uint8_t *) ctx->scratch.ptr, stride, n_trees, m, m, d_cost);
		if (rc) return rc;
		XMP_CUDA(cudaMemcpyAsync(cost_out, d_cost, n_pairs * sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream));
		XMP_CUDA(cudaStreamSynchronize(ctx->stream));
		return XMP_OK;
	}
	const size_t set_bytes = n_pairs * ctx->n_blocks * 16;
	int rc = ctx->scratch2.reserve(set_bytes + n_pairs * sizeof(unsigned) + 256);
	if (rc) return rc;
	uint4 *d_mid = (uint4 *) ctx->scratch2.ptr;
	unsigned *d_cost = (unsigned *) ((char *) ctx->scratch2.ptr + ((set_bytes + 255) & ~(size_t) 255));
	rc = xmp_cuda_build_midpoints(ctx, paths, path_stride, n_trees, m, d_mid, nullptr);
	if (rc) return rc;
	rc = launch_score_batch(ctx, d_mid, n_pairs, E, m, nullptr, 0x7fffffffu, d_cost, nullptr, nullptr);
	if (rc) return rc;
	XMP_CUDA(cudaMemcpyAsync(cost_out, d_cost, n_pairs * sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream));
	XMP_CUDA(cudaStreamSynchronize(ctx->stream));
	return XMP_OK;
}

// ---- per-pair family -------------------------------------------------------------------------------
// One CTA streams the two state sets; results come back through a small device record.
struct PairOut {
	unsigned score;
	unsigned changed;
};

__global__ void __launch_bounds__(256)
k_pair(const uint4 *__restrict__ s1, const uint4 *__restrict__ s2, uint4 *__restrict__ dest, const uint4 *__restrict__ prev,
       const unsigned *__restrict__ weights, unsigned n_blocks, unsigned w1_from_block, int mode, PairOut *out) {
	// mode 0: bases (+compare when prev); 1: weight-1 score; 2: weighted score (weight-1 from w1_from_block on)
	unsigned score = 0, changed = 0;
	for (unsigned b = threadIdx.x; b < n_blocks; b += blockDim.x) {
		const uint4 a = s1[b], c = s2[b];
		if (mode == 0) {
			uint4 d = fitch_block(a, c);
			dest[b] = d;
			if (prev) {
				uint4 p = prev[b];
				changed |= (d.x ^ p.x) | (d.y ^ p.y) | (d.z ^ p.z) | (d.w ^ p.w);
			}
		} else if (mode == 1 || b >= w1_from_block) {
			score += cost_word_w1(a.x, c.x) + cost_word_w1(a.y, c.y) + cost_word_w1(a.z, c.z) + cost_word_w1(a.w, c.w);
		} else {
			const unsigned av[4] = {a.x, a.y, a.z, a.w}, cv[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
			for (int w = 0; w < 4; ++w) {
				unsigned e = empty_mask(av[w], cv[w]);
#pragma unroll
				for (int nib = 0; nib < 8; ++nib) {
					if (e & (8u << (4 * nib))) score += weights[(size_t) b * 32 + w * 8 + nib];
				}
			}
		}
	}
	score = warp_sum(score);
	changed = __any_sync(0xffffffffu, changed != 0);
	if ((threadIdx.x & 31u) == 0) {
		if (score) atomicAdd(&out->score, score);
		if (changed) atomicOr(&out->changed, 1u);
	}
}

namespace {
struct PairEngine {
	std::mutex mu;
	bool ready = false;
	cudaStream_t stream = nullptr;
	char *d_buf = nullptr;
	size_t cap = 0;
	PairOut *d_out = nullptr;

	[[noreturn]] static void die(const char *what, cudaError_t e) {
		fprintf(stderr, "libxmp_b200: %s failed: %s.  There is no CPU fallback, aborting.\n", what, cudaGetErrorString(e));
		exit(1);
	}
	void check(cudaError_t e, const char *what) {
		if (e != cudaSuccess) die(what, e);
	}
	void init() {
		if (ready) return;
		int n = 0;
		cudaError_t e = cudaGetDeviceCount(&n);
		if (e != cudaSuccess || n == 0) die("finding a CUDA device", e == cudaSuccess ? cudaErrorNoDevice : e);
		check(cudaSetDevice(0), "cudaSetDevice");
		check(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking), "cudaStreamCreate");
		check(cudaMalloc(&d_out, sizeof(PairOut)), "cudaMalloc");
		ready = true;
	}
	// Layout of d_buf: s1 | s2 | dest | prev | weights, each n_bytes (weights: 8x).
	unsigned run(int mode, const void *s1, const void *s2, void *dest, const void *prev, const unsigned *weights,
	             unsigned n_blocks, unsigned w1_from_block, unsigned *changed_out) {
		std::lock_guard<std::mutex> lock(mu);
		if (n_blocks == 0) {
			if (changed_out) *changed_out = 0;
			return 0;
		}
		init();
		check(cudaSetDevice(0), "cudaSetDevice");
		const size_t nb = (size_t) n_blocks * 16, need = nb * 4 + nb * 8;
		if (need > cap) {
			if (d_buf) cudaFree(d_buf);
			check(cudaMalloc(&d_buf, need), "cudaMalloc");
			cap = need;
		}
		char *d1 = d_buf, *d2 = d_buf + nb, *dd = d_buf + 2 * nb, *dp = d_buf + 3 * nb, *dw = d_buf + 4 * nb;
		check(cudaMemcpyAsync(d1, s1, nb, cudaMemcpyHostToDevice, stream), "H2D");
		check(cudaMemcpyAsync(d2, s2, nb, cudaMemcpyHostToDevice, stream), "H2D");
		if (prev) check(cudaMemcpyAsync(dp, prev, nb, cudaMemcpyHostToDevice, stream), "H2D");
		if (weights) check(cudaMemcpyAsync(dw, weights, nb * 8, cudaMemcpyHostToDevice, stream), "H2D");
		check(cudaMemsetAsync(d_out, 0, sizeof(PairOut), stream), "memset");
		k_pair<<<1, 256, 0, stream>>>((const uint4 *) d1, (const uint4 *) d2, (uint4 *) dd, prev ? (const uint4 *) dp : nullptr,
		                              (const unsigned *) dw, n_blocks, w1_from_block, mode, d_out);
		check(cudaGetLastError(), "kernel launch");
		PairOut h;
		if (dest) check(cudaMemcpyAsync(dest, dd, nb, cudaMemcpyDeviceToHost, stream), "D2H");
		check(cudaMemcpyAsync(&h, d_out, sizeof h, cudaMemcpyDeviceToHost, stream), "D2H");
		check(cudaStreamSynchronize(stream), "synchronize");
		if (changed_out) *changed_out = h.changed;
		return h.score;
	}
};
PairEngine g_pair;

// The reference's FASTWEIGHT1FITCH hand-over point (src/fitchtemplate_b2_w8_fastc64.inc:105-113) at this
// family's block size: the first 16-byte block whose first weight is 1.
unsigned first_weight1_block(const unsigned *weights, unsigned n_blocks) {
	for (unsigned b = 0; b < n_blocks; ++b) {
		if (weights[(size_t) b * 32] == 1) return b;
	}
	return n_blocks;
}
}  // namespace

extern "C" void FitchBases_b2_w16_cuda(void *s1, void *s2, void *dest, unsigned nBlocks) {
	g_pair.run(0, s1, s2, dest, nullptr, nullptr, nBlocks, 0, nullptr);
}

extern "C" unsigned FitchBasesAndCompare_b2_w16_cuda(void *s1, void *s2, void *dest, void *prev, unsigned nBlocks) {
	unsigned changed = 0;
	g_pair.run(0, s1, s2, dest, prev, nullptr, nBlocks, 0, &changed);
	return changed;
}

extern "C" unsigned FitchScoreWeight1_b2_w16_cuda(void *s1, void *s2, unsigned nBlocks) {
	return g_pair.run(1, s1, s2, nullptr, nullptr, nullptr, nBlocks, 0, nullptr);
}

extern "C" unsigned FitchScoreWeight1_stopearly_b2_w16_cuda(void *s1, void *s2, unsigned nBlocks, unsigned maxScore) {
	(void) maxScore;   // the exact score satisfies the stop-early contract
	return g_pair.run(1, s1, s2, nullptr, nullptr, nullptr, nBlocks, 0, nullptr);
}

extern "C" unsigned FitchScore_b2_w16_cuda(void *s1, void *s2, unsigned *weights, unsigned nBlocks) {
	return g_pair.run(2, s1, s2, nullptr, nullptr, weights, nBlocks, nBlocks, nullptr);
}

extern "C" unsigned FitchScore_stopearly_b2_w16_cuda(void *s1, void *s2, unsigned *weights, unsigned nBlocks, unsigned maxScore) {
	(void) maxScore;
	return g_pair.run(2, s1, s2, nullptr, nullptr, weights, nBlocks, nBlocks, nullptr);
}

extern "C" unsigned FitchScore_fw1_b2_w16_cuda(void *s1, void *s2, unsigned *weights, unsigned nBlocks) {
	return g_pair.run(2, s1, s2, nullptr, nullptr, weights, nBlocks, first_weight1_block(weights, nBlocks), nullptr);
}

extern "C" unsigned FitchScore_fw1_stopearly_b2_w16_cuda(void *s1, void *s2, unsigned *weights, unsigned nBlocks, unsigned maxScore) {
	(void) maxScore;
	return g_pair.run(2, s1, s2, nullptr, nullptr, weights, nBlocks, first_weight1_block(weights, nBlocks), nullptr);
}
