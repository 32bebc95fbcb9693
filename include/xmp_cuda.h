/* xmp_cuda.h -- C ABI of libxmp_b200.so: the B200 (sm_100a) Fitch / branch-and-bound hot path.
 *
 * Plain C, no CUDA or torch types in any signature: pointers, sizes and opaque handles only, so the
 * library can be bound from C (host/fastdnamp.c), from ctypes (xmp_b200/cuda.py) or from any FFI.
 * There is NO CPU fallback behind these entry points: without a usable sm_100 device every call that
 * needs one fails with a negative code and a message in xmp_cuda_last_error().
 *
 * Two surfaces:
 *
 *  (1) The per-pair Fitch family `_b2_w16_cuda` -- the eight symbols every Fitch family of the
 *      reference exports (reference src/seq_b2_w8_fastc64.h:7-21, src/seq_b2_sse2asm.h:13-32, chosen
 *      through the macro dispatch of src/seq.h:29-66).  Same argument meaning; BYTESPERBLOCK = 16.
 *      They exist for drop-in linkage and parity tests: one launch per call is never the fast path.
 *
 *  (2) The batched surface, which replaces the reference's per-edge loop
 *      (BranchAndBound / BandBScanEdge / UpdateTreeAfterInsertion, reference src/yanbader.c:21-495):
 *      load a packed problem once, then score every (partial tree, insertion edge) pair of a batch in
 *      one launch, or run the whole exact search on the device.
 *
 * Data format (reference src/common.h:75-78, src/common.c:1034-1060): a state set is 4 bits per site,
 * two sites per byte, even site in the low nibble; rows are padded to 16-byte blocks with 0xF nibbles
 * whose weight is 1.  `n_blocks` always counts 16-byte blocks (32 sites).
 *
 * Edge ids.  A partial tree on taxa 0..m-1 (search order) has 2m-3 edges, each named after the node
 * below it: edge 0 / 1 = leaves 1 / 2, edge 2 = the first internal node (LEFT = leaf 1, RIGHT = leaf 2)
 * hanging off the root, which is taxon 0 (reference src/common.c:543-550).  Inserting taxon t on edge v
 * appends leaf edge 2t-3 and internal edge 2t-2 (LEFT = new leaf, RIGHT = v; src/common.c:735-736).
 * Ids never change once assigned, so a tree IS its insertion path: path[t] = v for 3 <= t < m.
 */
#ifndef XMP_CUDA_H
#define XMP_CUDA_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XMP_MAX_TAXA 100
#define XMP_OK 0
#define XMP_ERR_CUDA (-1)      /* a CUDA runtime call failed; see xmp_cuda_last_error() */
#define XMP_ERR_ARG (-2)       /* invalid argument / call order */
#define XMP_ERR_NOMEM (-3)     /* device memory budget too small for the request */
#define XMP_ERR_NODEVICE (-4)  /* no CUDA device, or not compute capability 10.x */

typedef struct xmp_ctx xmp_ctx;

const char *xmp_cuda_last_error(void);
int xmp_cuda_device_count(void);

/* One context per (host thread, device).  `device` is a CUDA ordinal. */
int  xmp_cuda_init(int device, xmp_ctx **out);
void xmp_cuda_destroy(xmp_ctx *ctx);
/* Launch on a caller-owned stream (a cudaStream_t passed as void *); NULL = the context's own. */
int  xmp_cuda_set_stream(xmp_ctx *ctx, void *stream);
int  xmp_cuda_synchronize(xmp_ctx *ctx);

/* ------------------------------------------------------------------------------------------------
 * (1) per-pair family, HOST buffers, 16-byte aligned, nBlocks 16-byte blocks (may be 0).
 *     Replaces: FitchBases_b2_w8_fastc64 etc. (reference src/seq_b2_w8_fastc64.c:11-44,
 *     src/fitchtemplate_b2_w8_fastc64.inc:43-163).  Stop-early variants return the exact score,
 *     which the contract allows (any value > maxScore when the true score exceeds it).
 *     They run on device 0 through an internal context; a CUDA failure aborts the process the way
 *     the reference treats fatal errors (message on stderr, exit(1)), because these signatures,
 *     fixed by the reference, have no error channel.
 * ---------------------------------------------------------------------------------------------- */
void     FitchBases_b2_w16_cuda(void *s1, void *s2, void *dest, unsigned nBlocks);
unsigned FitchBasesAndCompare_b2_w16_cuda(void *s1, void *s2, void *dest, void *prev, unsigned nBlocks);
unsigned FitchScoreWeight1_b2_w16_cuda(void *s1, void *s2, unsigned nBlocks);
unsigned FitchScoreWeight1_stopearly_b2_w16_cuda(void *s1, void *s2, unsigned nBlocks, unsigned maxScore);
unsigned FitchScore_b2_w16_cuda(void *s1, void *s2, unsigned *weights, unsigned nBlocks);
unsigned FitchScore_stopearly_b2_w16_cuda(void *s1, void *s2, unsigned *weights, unsigned nBlocks, unsigned maxScore);
unsigned FitchScore_fw1_b2_w16_cuda(void *s1, void *s2, unsigned *weights, unsigned nBlocks);
unsigned FitchScore_fw1_stopearly_b2_w16_cuda(void *s1, void *s2, unsigned *weights, unsigned nBlocks, unsigned maxScore);

/* ------------------------------------------------------------------------------------------------
 * (2a) problem upload.  Replaces the frozen part of struct TreeData (reference src/common.h:152-237):
 *      charMat, weights, seqLenBlocks, restBound, constantWeight.
 *      rows      HOST [num_taxa][n_blocks*16] packed rows, taxa in search order
 *      weights   HOST [n_blocks*32] site weights, sorted descending (NULL = all 1)
 *      rest_bound HOST [num_taxa] (NULL = all 0)
 * ---------------------------------------------------------------------------------------------- */
int xmp_cuda_load(xmp_ctx *ctx, const void *rows, const unsigned *weights, unsigned num_taxa, unsigned n_blocks,
                  const unsigned *rest_bound, unsigned constant_weight);
/* Same, but the rows already live on the device (e.g. a synthetic alignment generated there). */
int xmp_cuda_load_device_rows(xmp_ctx *ctx, const void *d_rows, const unsigned *weights, unsigned num_taxa,
                              unsigned n_blocks, const unsigned *rest_bound, unsigned constant_weight);

/* ------------------------------------------------------------------------------------------------
 * (2b) batched insertion scoring: "all edges of T partial trees in one launch".
 *      Replaces one BandBScanEdge sweep per tree (reference src/yanbader.c:113-244, the
 *      CHOSEN_FitchScore call at :160-166 and the acceptance test at :152,176).
 * ---------------------------------------------------------------------------------------------- */
/* Builds the edge-indexed MIDPOINT arrays of n_trees partial trees on taxa 0..m-1 from their
 * insertion paths (HOST, n_trees x path_stride bytes, path[t] at byte t).  d_mid is DEVICE memory of
 * n_trees * (2m-3) * n_blocks * 16 bytes, laid out [tree][edge][block].  tree_len (DEVICE, n_trees
 * unsigned, may be NULL) receives each tree's Fitch length + constant_weight. */
int xmp_cuda_build_midpoints(xmp_ctx *ctx, const uint8_t *paths, size_t path_stride, unsigned n_trees, unsigned m,
                             void *d_mid, unsigned *d_tree_len);

/* Scores taxon `taxon` against n_pairs state sets (DEVICE, [pair][n_blocks] 16-byte blocks) and prunes.
 *   d_base_score  DEVICE [n_pairs / pairs_per_tree] length of the tree each pair belongs to, or NULL (= 0)
 *   bound         current upper bound; a pair survives iff
 *                 (int) cost <= (int) (bound - base_score - rest_bound[taxon])   (src/yanbader.c:152,176)
 *   d_cost        DEVICE [n_pairs]  insertion cost of every pair (exact, never truncated)
 *   d_survivors   DEVICE [n_pairs]  indices of surviving pairs, ballot-compacted (may be NULL)
 *   d_n_survivors DEVICE [1]        their count (may be NULL iff d_survivors is)
 */
int xmp_cuda_score_batch(xmp_ctx *ctx, const void *d_sets, size_t n_pairs, unsigned pairs_per_tree, unsigned taxon,
                         const unsigned *d_base_score, unsigned bound, unsigned *d_cost, unsigned *d_survivors,
                         unsigned *d_n_survivors);

/* Convenience for callers holding HOST data (the end-to-end path of bench.py): uploads the paths,
 * builds the midpoints of n_trees trees on m taxa, scores taxon m against every edge, and returns the
 * costs ([n_trees][2m-3]) in HOST memory.  Scratch is taken from the context. */
int xmp_cuda_score_insertions_host(xmp_ctx *ctx, const uint8_t *paths, size_t path_stride, unsigned n_trees, unsigned m,
                                   unsigned *cost_out);

/* ------------------------------------------------------------------------------------------------
 * (2c) exact search on the device.  Replaces BranchAndBound(root, 3, td) (reference
 *      src/fastdnamp.c:127, src/yanbader.c:21-109) and, with jobs, the MPI boss/worker pair
 *      (src/mpiboss.c, src/mpiworker.c).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
	unsigned bound;             /* initial upper bound (0x7fffffff = none) */
	unsigned max_trees;         /* keep at most this many optimal trees (0x7fffffff = all) */
	size_t   mem_budget;        /* bytes of device memory the frontier may use (0 = default) */
	unsigned shard_index;       /* this context searches the subtrees whose index at the split level  */
	unsigned shard_count;       /*   is congruent to shard_index modulo shard_count (1 = everything)  */
	unsigned split_level;       /* tree size at which the frontier is dealt out (0 = automatic)      */
	unsigned flags;             /* XMP_SEARCH_* */
} xmp_search_params;
#define XMP_SEARCH_NO_DIVE 1u   /* skip the greedy beam dive that tightens the bound first */
#define XMP_SEARCH_NO_FUSE 2u   /* materialise the deepest pool instead of fusing the last two levels (testing) */
#define XMP_SEARCH_NO_EAGER 4u  /* keep MIDPOINT sets in the frontier and score them in a separate pass (the layout for > 32 blocks) */

typedef struct {
	unsigned long long nodes;          /* partial trees expanded (= BranchAndBound() invocations) */
	unsigned long long score_calls;    /* (tree, edge) pairs scored (= FitchScore calls) */
	unsigned long long children;       /* partial trees built (= accepted, non-final insertions) */
	unsigned long long full_trees;     /* complete trees that passed the bound test */
	unsigned long long site_merges;    /* 4-bit site operations performed, padding excluded */
	unsigned long long launches;       /* kernels launched by the search */
	unsigned long long iterations;     /* frontier batches processed */
	double   seconds;                  /* device time spent inside xmp_cuda_search_run */
	/* per tree size m (taxa on the tree being expanded), the reference's per-depth counters
	 * (src/measure.c:10 branchAndBoundScanEdgeCountByDepth[MAXTAXA], re-implemented as 64-bit): */
	unsigned long long nodes_by_depth[XMP_MAX_TAXA + 1];       /* partial trees on m taxa expanded */
	unsigned long long edge_evals_by_depth[XMP_MAX_TAXA + 1];  /* (tree, edge) pairs scored on trees of m taxa */
	unsigned long long batches_by_depth[XMP_MAX_TAXA + 1];     /* frontier batches (launch groups) at tree size m */
	unsigned long long steals_served;  /* steal requests of other ranks answered with work (reference: stealCount) */
	unsigned long long steals_received; /* times this rank ran dry and obtained work from a peer */
	double   begin_seconds;            /* part of `seconds` spent in xmp_cuda_search_begin (the greedy dive, the root) */
} xmp_search_stats;

/* Optional: allocate the search's device memory for a problem of this shape now (mem_budget as in
 * xmp_search_params, 0 = default), so that xmp_cuda_search_begin finds it in place -- the reference
 * likewise sets up its arena in InitTreeData, before its branch-and-bound timer starts
 * (src/common.c:1113-1224).  The memory stays with the context and is reused by later searches. */
int xmp_cuda_search_reserve(xmp_ctx *ctx, unsigned num_taxa, unsigned n_blocks, size_t mem_budget);
int xmp_cuda_search_begin(xmp_ctx *ctx, const xmp_search_params *params);
/* Runs at most max_batches frontier batches (0 = until done).  *done = 1 once the frontier is empty. */
int xmp_cuda_search_run(xmp_ctx *ctx, unsigned max_batches, int *done);
/* Upper-bound exchange between shards: read ours, lower ours to a peer's (purges trees above it). */
int xmp_cuda_search_get_bound(xmp_ctx *ctx, unsigned *bound);
int xmp_cuda_search_set_bound(xmp_ctx *ctx, unsigned bound);
/* Pending work, for load balancing: number of open partial trees per tree size, [XMP_MAX_TAXA + 1]. */
int xmp_cuda_search_pending(xmp_ctx *ctx, unsigned long long *per_level);
/* Work stealing: detach up to max_nodes open partial trees of the shallowest non-empty level as job
 * descriptors (insertion paths, num_taxa bytes each, byte 0 = tree size), and the converse. */
int xmp_cuda_search_export_jobs(xmp_ctx *ctx, unsigned max_nodes, uint8_t *jobs, unsigned *n_jobs);
int xmp_cuda_search_import_jobs(xmp_ctx *ctx, const uint8_t *jobs, unsigned n_jobs);
/* Result: optimal length, number of optimal trees found by this shard, and up to cap of their paths
 * (num_taxa bytes each, unordered).  Valid once done. */
int xmp_cuda_search_result(xmp_ctx *ctx, unsigned *score, unsigned long long *n_trees, uint8_t *paths, size_t cap);
/* The -m cap (reference src/yanbader.c:272-293: keep the first maxTrees trees of the depth-first search, go on
 * counting).  With a finite max_trees a shard that has collected more than 4 x max_trees (and a million) optimal
 * trees keeps only the max_trees that come first in the reference's depth-first order; n_trees above still
 * counts all of them, *n_stored says for how many the paths are held (= n_trees when nothing was cut). */
int xmp_cuda_search_stored(xmp_ctx *ctx, unsigned long long *n_stored);
/* Host-side utility (no device needed): moves the first `keep` of n_paths insertion paths (num_taxa bytes each)
 * in the reference's depth-first order (src/yanbader.c:103,238-243) to the front, in that order. */
int xmp_cuda_paths_keep_first_dfs(uint8_t *paths, size_t n_paths, unsigned num_taxa, size_t keep);
int xmp_cuda_search_stats(xmp_ctx *ctx, xmp_search_stats *out);

/* ------------------------------------------------------------------------------------------------
 * (2d) several GPUs of one node on one search.  Replaces the reference's MPI boss / worker pair
 *      (src/mpiboss.c:30-429, src/mpiworker.c:19-41,117-181,214-380) by words in peer-mapped DEVICE memory:
 *      every context owns an exchange block that the other ranks' GPUs address over NVLink.
 *        - the upper bound: the search kernels lower the bound word of every peer with a system-scope atomicMin
 *          in the launch that finds a better tree (the reference relays MSG_NEWBOUND through the boss);
 *        - work stealing: a rank whose frontier ran dry asks ONE peer (a compare-and-swap on that peer's request
 *          word); the donor answers between two sweeps of its own search by copying job descriptors (insertion
 *          paths) into the thief's mailbox, device to device.  No collective, no lock step;
 *        - termination: a counter of ranks that hold work; whoever brings it to zero raises every rank's flag.
 *      Use: shard the search with shard_index / shard_count = rank / world as before, and
 *          xmp_cuda_exchange_handle   on every rank; exchange the 64-byte handles (any side channel),
 *          xmp_cuda_exchange_connect  (ranks are processes: CUDA IPC) or _connect_local (ranks are threads),
 *          xmp_cuda_exchange_reset    before every search, then a barrier of your own, then xmp_cuda_search_begin;
 *          loop { xmp_cuda_search_run(ctx, 0, &done); xmp_cuda_search_steal(ctx, &state); } until state is
 *          XMP_STEAL_FINISHED; then xmp_cuda_search_result on every rank (each tree is held by exactly one).
 *      A context that is not connected behaves as before (one rank).
 * ---------------------------------------------------------------------------------------------- */
#define XMP_EXCHANGE_HANDLE_BYTES 64
#define XMP_STEAL_GOT_WORK 1
#define XMP_STEAL_FINISHED 2
int xmp_cuda_exchange_handle(xmp_ctx *ctx, unsigned char *handle /* [XMP_EXCHANGE_HANDLE_BYTES] */);
int xmp_cuda_exchange_connect(xmp_ctx *ctx, unsigned rank, unsigned world, const unsigned char *handles /* [world][64] */);
int xmp_cuda_exchange_connect_local(xmp_ctx *ctx, unsigned rank, unsigned world, xmp_ctx *const *peers /* [world] */);
int xmp_cuda_exchange_disconnect(xmp_ctx *ctx);
int xmp_cuda_exchange_reset(xmp_ctx *ctx);
/* Call when xmp_cuda_search_run reported done: waits until a peer hands over open subtrees (*state =
 * XMP_STEAL_GOT_WORK: call xmp_cuda_search_run again) or until no rank holds work (XMP_STEAL_FINISHED). */
int xmp_cuda_search_steal(xmp_ctx *ctx, int *state);

/* ------------------------------------------------------------------------------------------------
 * Measurement aid: what the integer pipe sustains on the Fitch instruction mixes with operands in
 * registers (no memory traffic), in G site-operations per second: out[0] weight-1 cost, out[1] merge
 * (FitchBases), out[2] merge + cost.  The denominator of the integer roofline.
 * ---------------------------------------------------------------------------------------------- */
int xmp_cuda_int_pipe_peak(xmp_ctx *ctx, double *out);

#ifdef __cplusplus
}
#endif
#endif
