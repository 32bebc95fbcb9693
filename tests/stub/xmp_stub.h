/* TEST INFRASTRUCTURE ONLY -- shared between the two halves of the CPU test double (xmp_stub.c, xmp_stub_exchange.cpp). */
#ifndef XMP_STUB_H
#define XMP_STUB_H
#include <stddef.h>
#include <stdint.h>
#include "xmp_cuda.h"

#ifdef __cplusplus
extern "C" {
#endif

struct xmp_ctx {
	int device;
	unsigned n, n_blocks, constant_weight;
	uint8_t *rows;
	unsigned *weights, *rest_bound, *taxon_map;
	unsigned bound;
	uint8_t *jobs;            /* n bytes each, byte 0 = tree size */
	size_t n_jobs, cap_jobs, head;
	uint8_t *trees;           /* n bytes each, all of length `bound` */
	size_t n_trees, cap_trees;
	xmp_search_stats st;
	int begun;
	/* exchange (xmp_stub_exchange.cpp) */
	int xch_connected;
	unsigned xch_rank, xch_world;
	uint8_t *staged;          /* job descriptors detached for a thief */
	unsigned n_staged;
	unsigned long long steals_served, steals_received;
};


int stub_fail(int code, const char *msg);
void stub_push_job(xmp_ctx *c, const uint8_t *path, unsigned m);
/* exchange hooks called by the search loop of xmp_stub.c */
unsigned stub_xch_bound_in(xmp_ctx *c);
void stub_xch_bound_local(xmp_ctx *c, unsigned v);     /* search_begin: fold the initial bound into this rank's word */              /* the bound word of this rank's block (peers lower it) */
void stub_xch_bound_out(xmp_ctx *c, unsigned v);     /* a better tree was found here: lower every rank's word */
int stub_xch_after_job(xmp_ctx *c);                  /* publish the open count, answer a waiting thief */

#ifdef __cplusplus
}
#endif
#endif
