// Internal: the context behind the opaque xmp_ctx handle of include/xmp_cuda.h.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>
#include <vector>
#include "../../include/xmp_cuda.h"
#include "fitch_dev.cuh"

namespace xmp {

int set_error(int code, const char *fmt, ...);
int cuda_failed(cudaError_t e, const char *what, const char *file, int line);

#define XMP_CUDA(call)                                                              \
	do {                                                                            \
		cudaError_t e_ = (call);                                                    \
		if (e_ != cudaSuccess) return xmp::cuda_failed(e_, #call, __FILE__, __LINE__); \
	} while (0)

struct Search;

// The exchange block of a context: 2 MiB of device memory that peers can map (exchange.cu, exchange_proto.h).
//   [0, 1024)      the search's counter block; word 0 is the upper bound, words 112-119 the protocol's control words
//   [1024, 1280)   XchPeerTable: every rank's block as THIS device addresses it (read by the kernels and by k_xch_gather)
//   [2048, ...)    mailbox: job descriptors (insertion paths) written by a donor over NVLink
struct XchPeerTable {
	unsigned *blk[16];
	unsigned world, rank;
};
#define XCH_BLOCK_BYTES ((size_t) 2 << 20)
#define XCH_TABLE_OFF 1024
#define XCH_MAILBOX_BYTE_OFF 2048

// Grow-only device scratch buffer.
struct Scratch {
	void *ptr = nullptr;
	size_t bytes = 0;
	int reserve(size_t need);
	void release();
};

// Per-tree topology in edge-id space, as the kernels consume it (one byte per entry).
struct HostTopo {
	unsigned m = 0;                 // taxa on the tree
	unsigned top = 2;               // the root's only child
	uint8_t par[2 * XMP_MAX_TAXA];
	uint8_t kid[2][2 * XMP_MAX_TAXA];
	void base();
	void insert(unsigned taxon, unsigned v);
	unsigned preorder(uint8_t *out) const;
	// Builds from an insertion path (path[t] for 3 <= t < m); returns false on an invalid edge id.
	bool from_path(const uint8_t *path, unsigned m_);
};

}  // namespace xmp

struct xmp_ctx {
	int device = 0;
	int sm_count = 148;
	cudaStream_t stream = nullptr;
	cudaStream_t own_stream = nullptr;
	cudaEvent_t ev0 = nullptr, ev1 = nullptr;

	// loaded problem
	unsigned num_taxa = 0, n_blocks = 0, num_sites_padded = 0;
	const uint4 *d_rows = nullptr;      // [num_taxa][n_blocks]
	uint4 *d_rows_owned = nullptr;
	size_t rows_owned_bytes = 0;
	unsigned *d_planes = nullptr;
	xmp::WeightPlanes wp{nullptr, 0, 0, 0};
	unsigned *d_rest_bound = nullptr;
	unsigned rest_bound[XMP_MAX_TAXA + 1] = {0};
	unsigned constant_weight = 0;
	unsigned long long real_sites = 0;   // sites that are not padding (for site-merge accounting)

	xmp::Scratch scratch, scratch2;
	xmp::Search *search = nullptr;
	void *arena = nullptr;         // the search's device memory (frontier pools, survivor lists, tree buffers), kept between searches
	size_t arena_bytes = 0;
	unsigned *h_pin = nullptr;     // pinned scalars the search polls

	// multi-GPU exchange (exchange.cu)
	unsigned *xch = nullptr;               // this context's exchange block (device memory, cudaMalloc'ed on first use)
	unsigned xch_rank = 0, xch_world = 1;
	unsigned *xch_peer[16] = {nullptr};    // every rank's block as this device sees it (own included)
	bool xch_ipc[16] = {false};            // mapped with cudaIpcOpenMemHandle (closed on disconnect)
	bool xch_connected = false;
	unsigned *h_xres = nullptr;            // pinned: results of control operations, the published hint
};

namespace xmp {
// kernels.cu
int launch_score_batch(xmp_ctx *ctx, const uint4 *d_sets, size_t n_pairs, unsigned pairs_per_tree, unsigned taxon,
                       const unsigned *d_base_score, unsigned bound, unsigned *d_cost, unsigned *d_survivors,
                       unsigned *d_n_survivors);
// Same with an explicit record stride: pair p reads d_sets + (p / ppt) * tree_stride + (p % ppt) * n_blocks.
int launch_score_sets(xmp_ctx *ctx, const uint4 *d_sets, unsigned long long tree_stride, size_t n_pairs, unsigned ppt,
                      unsigned taxon, const unsigned *d_base, unsigned bound, unsigned *d_cost, unsigned *d_surv,
                      unsigned *d_nsurv);
// Builds state sets for n_trees trees on m taxa whose topologies were uploaded to d_topo
// ([tree][4][stride] bytes: par, kid0, kid1, preorder; byte `stride-1` of the par row = top).
// rec_blocks = uint4 per tree record; MIDPOINT sets at 0 and, when store_all, UP sets at E*W.
// planar_rows != nullptr: read those rows (the alignment in planar form) and write planar sets.
int launch_build_sets(xmp_ctx *ctx, const uint8_t *d_topo, unsigned topo_stride, unsigned n_trees, unsigned m,
                      uint4 *d_out, size_t rec_blocks, bool store_all, unsigned *d_tree_len, const uint4 *planar_rows = nullptr);
int launch_score_trees_wide(xmp_ctx *ctx, const uint8_t *d_topo, unsigned topo_stride, unsigned n_trees, unsigned m,
                            unsigned taxon, unsigned *d_cost);
void search_destroy(xmp_ctx *ctx);
// exchange.cu
int xch_ensure_block(xmp_ctx *ctx);
void xch_release(xmp_ctx *ctx);
}  // namespace xmp
