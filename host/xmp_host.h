/* Host side of xmp_b200: everything fastdnamp does before and after the search, in C.
 *
 * This mirrors, behaviour for behaviour, the parts of the reference that surround the hot path
 * (reference file:line given at each function): the PHYLIP reader, column recoding / pattern
 * compression / removal of uninformative sites, the max-mini taxon order, the MST upper bound, the
 * distinct-bases, MST and partition lower bounds, packing into 4-bit state sets, and the Newick / NEXUS writer.
 * Nothing here computes Fitch sets for the search: that is libxmp_b200.so (include/xmp_cuda.h).
 *
 * All functions return 0 on success and a negative value on failure, with a message retrievable
 * from xmph_last_error(); the CLI prints it and exits 1, as the reference does in place.
 */
#ifndef XMP_HOST_H
#define XMP_HOST_H
#include <stddef.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Option bits: same values as the reference's OPT_* (src/common.h:51-68). */
#define XMPH_OPT_NEXUSFMT              1u
#define XMPH_OPT_DISCARDMISSAMBIG      2u
#define XMPH_OPT_DISPLAYNEWBOUNDS      4u
#define XMPH_OPT_DISPLAYNEWTREES       8u
#define XMPH_OPT_QUIET                 16u
#define XMPH_OPT_USELOADRESTBOUND      32u
#define XMPH_OPT_USEDISTBASESBOUNDREST 64u
#define XMPH_OPT_USEINCOMPATBOUNDREST  128u
#define XMPH_OPT_USEMSTBOUNDREST       256u
#define XMPH_OPT_USEKICKASSBOUNDREST   512u
#define XMPH_OPT_USEKA2BOUNDREST       1024u
#define XMPH_OPT_USEGREEDYHITTINGSETBOUND 2048u
#define XMPH_OPT_IMPROVEINITUPPERBOUND 4096u
#define XMPH_OPT_USEPARTBOUND          8192u
#define XMPH_OPT_ONLYCOMPUTELOWERBOUND 16384u
#define XMPH_OPT_KEEPTEMPFILES         32768u
#define XMPH_OPT_SAVEREORDEREDDATASET  65536u
#define XMPH_OPT_TBR                   131072u

#define XMPH_NO_LIMIT 0x7fffffffu      /* INT_MAX: "no bound" / "no tree limit" (src/common.c:59-60) */
#define XMPH_MAXTAXA 100               /* src/common.h:43 */
#define XMPH_BLOCK_BYTES 16            /* BYTESPERBLOCK of the _b2_w16 family */
#define XMPH_SITES_PER_BLOCK 32

typedef struct {
	unsigned num_taxa, seq_len;
	unsigned char *chars;   /* [num_taxa][seq_len], upper-case IUPAC / '-' / '?' */
	char **labels;          /* [num_taxa] */
} xmph_alignment;

typedef struct {
	unsigned options;       /* XMPH_OPT_* */
	unsigned bound;         /* initial upper bound, XMPH_NO_LIMIT = none */
	unsigned max_trees;     /* XMPH_NO_LIMIT = keep all */
	unsigned start_col, end_col;
	const char *rest_bound_file;
	unsigned part_strategy;   /* -Bp: 0 final scores (default), 1 clipped only (-BpC), 2 lower bounds only (-BpL); src/partbound.h:5-9 */
	unsigned part_max_stale;  /* -Bp: stop after this many passes without improvement (--pbmaxiters, default 2; 0 = never) */
	unsigned threads;         /* host threads for bound precomputation, 0 = all cores */
} xmph_settings;

typedef struct {
	unsigned num_taxa;
	unsigned num_sites;          /* parsimony-informative site patterns kept */
	unsigned n_blocks;           /* 16-byte blocks per packed row */
	unsigned char *rows;         /* [num_taxa][n_blocks*16] packed state sets, taxa in search order */
	unsigned *weights;           /* [n_blocks*32], descending, padding = 1 */
	unsigned char *sites;        /* [num_taxa][num_sites] one state set per byte, search order */
	unsigned *site_weights;      /* [num_sites] */
	unsigned *taxon_map;         /* taxon_map[i] = original 0-based row of search-order taxon i */
	unsigned *rest_bound;        /* [num_taxa] */
	unsigned constant_weight;    /* length contributed by the non-informative sites on any tree */
	unsigned n_constant_sites, n_nonpi_sites;
	unsigned bound;              /* settings bound, lowered to the MST bound when enabled */
	unsigned mst_upper_bound;    /* MST weight + constant_weight (0 if not computed) */
} xmph_problem;

const char *xmph_last_error(void);

void xmph_default_settings(xmph_settings *s);                                  /* src/common.c:51-71 */
int  xmph_read_phylip(FILE *in, xmph_alignment *a);                            /* src/common.c:2850-2948 */
void xmph_free_alignment(xmph_alignment *a);
int  xmph_prepare(const xmph_alignment *a, const xmph_settings *s, xmph_problem *p); /* src/common.c:1113-1224 */
void xmph_free_problem(xmph_problem *p);
int  xmph_write_rest_bound(const xmph_problem *p, const char *path);           /* src/bound.c:47-56 */
/* Partition lower bound, raising p->rest_bound[3 .. num_taxa-2] (src/partbound.c:1144-1164); called by xmph_prepare for -Bp. */
int  xmph_part_bound(xmph_problem *p, int strategy, unsigned max_stale_passes, unsigned n_threads);
/* Exact length of a pair of unambiguous sites holding the (state, state) pairs in the bit set `index`
 * (bit 4a+b); the entry the reference keeps in its generated table src/sitepaircosts.c. */
int  xmph_site_pair_cost(unsigned index);

/* ---- trees ------------------------------------------------------------------------------- */
/* A tree is its insertion path: path[t] (3 <= t < num_taxa) = edge id taxon t was inserted on, in the
 * append-only edge numbering of include/xmp_cuda.h.  Paths are stored num_taxa bytes apart. */

/* Writes the reference's Newick text "(r,(...))" without ';' (src/common.c:933-974).  Returns length. */
size_t xmph_path_to_newick(const unsigned char *path, unsigned num_taxa, const unsigned *taxon_map, char *out);

/* Converts a path to the reference's job-descriptor coordinates: dfs[t] = pre-order index of the
 * edge taxon t was inserted on, in the tree of t taxa (src/yanbader.c:103,238-243). */
void xmph_path_to_dfs(const unsigned char *path, unsigned num_taxa, unsigned char *dfs);
/* And back (used to replay reference-style job descriptors). */
void xmph_dfs_to_path(const unsigned char *dfs, unsigned num_taxa, unsigned char *path);

/* Sorts paths into the order in which the reference's depth-first search would have found them. */
void xmph_sort_paths_dfs(unsigned char *paths, size_t n, unsigned num_taxa);

/* Writes the result file exactly as OutputResults does (src/common.c:371-454).
 * nexus != 0: NEXUS tree block with TRANSLATE table; else one "<newick>;" per line. */
int xmph_write_trees(FILE *out, int nexus, unsigned score, const xmph_alignment *a, const unsigned *taxon_map,
                     const unsigned char *paths, size_t n_trees);

#ifdef __cplusplus
}
#endif
#endif
