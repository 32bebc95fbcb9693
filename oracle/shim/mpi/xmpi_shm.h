/* TEST INFRASTRUCTURE ONLY.  Layout of the shared segment of the one-node MPI (xmpi.c, xmpirun.c). */
#ifndef XMPI_SHM_H
#define XMPI_SHM_H
#include <stdatomic.h>
#include <stddef.h>
#include <stdint.h>

#define XMPI_MAX_RANKS   256
#define XMPI_RING_BYTES  ((size_t) 4 << 20)    /* a power of two; the largest message is the reference's 2 MiB tree buffer */
#define XMPI_HEADER_BYTES ((size_t) 64 * (XMPI_MAX_RANKS + 1))

struct xmpi_ring {
	_Atomic uint64_t head;                     /* bytes consumed (written by the receiver) */
	char pad0[56];
	_Atomic uint64_t tail;                     /* bytes produced (written by the sender) */
	char pad1[56];
	unsigned char data[XMPI_RING_BYTES];
};

struct xmpi_shm {
	uint32_t n_ranks, ring_bytes;
	char pad[56];
	struct { _Atomic uint32_t word; char pad[60]; } doorbell[XMPI_MAX_RANKS];   /* bumped by every send to that rank */
	/* struct xmpi_ring rings[n_ranks][n_ranks] follows at XMPI_HEADER_BYTES: ring (source, destination) */
};

static inline size_t xmpi_shm_bytes(int n_ranks) {
	return XMPI_HEADER_BYTES + (size_t) n_ranks * (size_t) n_ranks * sizeof(struct xmpi_ring);
}
#endif
