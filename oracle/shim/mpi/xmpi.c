/* TEST INFRASTRUCTURE ONLY -- never linked into the product.
 *
 * The one-node MPI of oracle/shim/mpi/mpi.h: ranks are processes started by xmpirun (xmpirun.c), which hands
 * every rank a shared-memory segment holding one ring per (source, destination) pair.  A ring has a single
 * producer and a single consumer, so it needs no lock: 64-bit byte counters `tail` (written by the sender) and
 * `head` (written by the receiver), records of {tag, length, payload} padded to 8 bytes.  A receiver drains its
 * rings into private per-source queues whenever it enters the library, so senders practically never wait and the
 * matching rules of MPI (per source in arrival order, any tag / any source wildcards, receives in posting order)
 * are applied on private memory.  Blocking calls sleep on a futex word per rank that every sender bumps.
 *
 * Not the reference's code and no substitute for a real MPI: it exists so that the reference's own boss / worker
 * program runs here as the multi-core CPU baseline (oracle/Makefile, target ref_mpi). */
#define _GNU_SOURCE
#include "mpi.h"
#include <errno.h>
#include <linux/futex.h>
#include <stdatomic.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/syscall.h>
#include <time.h>
#include <unistd.h>
#include "xmpi_shm.h"

static struct xmpi_shm *g_shm;
static int g_rank = -1, g_size = 0;

struct msg {
	struct msg *next;
	int tag;
	uint32_t len;
	unsigned char data[];
};
static struct msg **g_q_head, **g_q_tail;      /* unexpected messages, one FIFO per source */

struct req {
	int active, complete, is_recv;
	void *buf;
	size_t cap;
	int src, tag;
	MPI_Status st;
};
static struct req *g_req;                      /* handle h refers to g_req[h - 1] */
static int g_req_cap;
static int *g_posted, g_n_posted, g_posted_cap; /* receive requests not matched yet, in posting order */

static void die(const char *what) {
	fprintf(stderr, "xmpi (rank %d): %s\n", g_rank, what);
	_exit(86);
}

static struct xmpi_ring *ring_of(int src, int dst) {
	return (struct xmpi_ring *) ((char *) g_shm + XMPI_HEADER_BYTES + ((size_t) src * g_size + dst) * sizeof(struct xmpi_ring));
}

static void futex_wait(_Atomic uint32_t *word, uint32_t seen) {
	struct timespec ts = {0, 2000000};         /* 2 ms: a lost wake-up costs a short delay, never a hang */
	syscall(SYS_futex, (uint32_t *) word, FUTEX_WAIT, seen, &ts, NULL, 0);
}
static void futex_wake(_Atomic uint32_t *word) { syscall(SYS_futex, (uint32_t *) word, FUTEX_WAKE, 1, NULL, NULL, 0); }

static void ring_read(struct xmpi_ring *r, uint64_t at, void *dst, size_t n) {
	const size_t off = (size_t) (at & (XMPI_RING_BYTES - 1)), first = XMPI_RING_BYTES - off < n ? XMPI_RING_BYTES - off : n;
	memcpy(dst, r->data + off, first);
	if (first < n) memcpy((char *) dst + first, r->data, n - first);
}
static void ring_write(struct xmpi_ring *r, uint64_t at, const void *src, size_t n) {
	const size_t off = (size_t) (at & (XMPI_RING_BYTES - 1)), first = XMPI_RING_BYTES - off < n ? XMPI_RING_BYTES - off : n;
	memcpy(r->data + off, src, first);
	if (first < n) memcpy(r->data, (const char *) src + first, n - first);
}

/* Moves everything that has arrived into the private queues. */
static void progress(void) {
	for (int s = 0; s < g_size; ++s) {
		if (s == g_rank) continue;
		struct xmpi_ring *r = ring_of(s, g_rank);
		uint64_t head = atomic_load_explicit(&r->head, memory_order_relaxed);
		const uint64_t tail = atomic_load_explicit(&r->tail, memory_order_acquire);
		while (head < tail) {
			uint32_t hdr[2];
			ring_read(r, head, hdr, 8);
			struct msg *m = (struct msg *) malloc(sizeof *m + hdr[1]);
			if (!m) die("out of memory");
			m->next = NULL;
			m->tag = (int) hdr[0];
			m->len = hdr[1];
			ring_read(r, head + 8, m->data, hdr[1]);
			head += 8 + (((uint64_t) hdr[1] + 7) & ~(uint64_t) 7);
			if (g_q_tail[s]) g_q_tail[s]->next = m; else g_q_head[s] = m;
			g_q_tail[s] = m;
		}
		atomic_store_explicit(&r->head, head, memory_order_release);
	}
}

/* First queued message from source s that carries `tag` (or any); unlinks and returns it. */
static struct msg *take_from(int s, int tag) {
	struct msg *prev = NULL;
	for (struct msg *m = g_q_head[s]; m; prev = m, m = m->next) {
		if (tag != MPI_ANY_TAG && m->tag != tag) continue;
		if (prev) prev->next = m->next; else g_q_head[s] = m->next;
		if (g_q_tail[s] == m) g_q_tail[s] = prev;
		return m;
	}
	return NULL;
}

/* Completes posted receives, in posting order, against the queued messages. */
static void match(void) {
	int w = 0;
	for (int i = 0; i < g_n_posted; ++i) {
		struct req *q = &g_req[g_posted[i] - 1];
		struct msg *m = NULL;
		int src = q->src;
		if (src != MPI_ANY_SOURCE) m = take_from(src, q->tag);
		else
			for (src = 0; src < g_size && !m; ++src)
				if (src != g_rank && (m = take_from(src, q->tag)) != NULL) break;
		if (!m) {
			g_posted[w++] = g_posted[i];
			continue;
		}
		if (m->len > q->cap) die("message longer than the receive buffer");
		if (m->len) memcpy(q->buf, m->data, m->len);
		q->st.MPI_SOURCE = src;
		q->st.MPI_TAG = m->tag;
		q->st.MPI_ERROR = MPI_SUCCESS;
		q->st.xmpi_bytes = (int) m->len;
		q->complete = 1;
		free(m);
	}
	g_n_posted = w;
}

static int new_req(void) {
	for (int i = 0; i < g_req_cap; ++i)
		if (!g_req[i].active) {
			memset(&g_req[i], 0, sizeof g_req[i]);
			g_req[i].active = 1;
			return i + 1;
		}
	const int old = g_req_cap;
	g_req_cap = old ? 2 * old : 64;
	g_req = (struct req *) realloc(g_req, (size_t) g_req_cap * sizeof *g_req);
	if (!g_req) die("out of memory");
	memset(g_req + old, 0, (size_t) (g_req_cap - old) * sizeof *g_req);
	g_req[old].active = 1;
	return old + 1;
}
/* A handle the program may pass: null, or one of ours that is in use.  (The reference's boss waits on request
 * arrays parts of which it never wrote, src/mpiboss.c:42,237: anything that is not a live handle counts as null.) */
static struct req *live(MPI_Request h) {
	if (h <= 0 || h > g_req_cap || !g_req[h - 1].active) return NULL;
	return &g_req[h - 1];
}
static void finish(MPI_Request *h, MPI_Status *st) {
	struct req *q = &g_req[*h - 1];
	if (st) {
		if (q->is_recv) *st = q->st;
		else { st->MPI_SOURCE = g_rank; st->MPI_TAG = q->tag; st->MPI_ERROR = MPI_SUCCESS; st->xmpi_bytes = 0; }
	}
	q->active = 0;
	*h = MPI_REQUEST_NULL;
}

static void do_send(const void *buf, int count, MPI_Datatype dt, int dest, int tag) {
	if (dest < 0 || dest >= g_size || dest == g_rank) die("bad destination rank");
	const size_t bytes = (size_t) count * (size_t) dt, need = 8 + ((bytes + 7) & ~(size_t) 7);
	if (need > XMPI_RING_BYTES) die("message larger than a ring");
	struct xmpi_ring *r = ring_of(g_rank, dest);
	const uint64_t tail = atomic_load_explicit(&r->tail, memory_order_relaxed);
	while (XMPI_RING_BYTES - (tail - atomic_load_explicit(&r->head, memory_order_acquire)) < need) {
		progress();                            /* keep draining our own rings: two ranks sending to each other must not deadlock */
		struct timespec ts = {0, 50000};
		nanosleep(&ts, NULL);
	}
	const uint32_t hdr[2] = {(uint32_t) tag, (uint32_t) bytes};
	ring_write(r, tail, hdr, 8);
	if (bytes) ring_write(r, tail + 8, buf, bytes);
	atomic_store_explicit(&r->tail, tail + need, memory_order_release);
	atomic_fetch_add_explicit(&g_shm->doorbell[dest].word, 1, memory_order_release);
	futex_wake(&g_shm->doorbell[dest].word);
}

int MPI_Init(int *argc, char ***argv) {
	(void) argc; (void) argv;
	const char *fd = getenv("XMPI_FD"), *rk = getenv("XMPI_RANK"), *sz = getenv("XMPI_SIZE");
	if (!fd || !rk || !sz) {
		fprintf(stderr, "xmpi: this program has to be started through xmpirun (oracle/_ref/xmpirun -n N program ...)\n");
		exit(86);
	}
	g_rank = atoi(rk);
	g_size = atoi(sz);
	const size_t bytes = xmpi_shm_bytes(g_size);
	g_shm = (struct xmpi_shm *) mmap(NULL, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, atoi(fd), 0);
	if (g_shm == MAP_FAILED) die("cannot map the shared segment");
	if (g_shm->n_ranks != (uint32_t) g_size || g_shm->ring_bytes != XMPI_RING_BYTES) die("shared segment does not match");
	g_q_head = (struct msg **) calloc((size_t) g_size, sizeof *g_q_head);
	g_q_tail = (struct msg **) calloc((size_t) g_size, sizeof *g_q_tail);
	g_posted_cap = 64;
	g_posted = (int *) malloc((size_t) g_posted_cap * sizeof *g_posted);
	return MPI_SUCCESS;
}
int MPI_Finalize(void) { return MPI_SUCCESS; }
int MPI_Comm_rank(MPI_Comm comm, int *rank) { (void) comm; *rank = g_rank; return MPI_SUCCESS; }
int MPI_Comm_size(MPI_Comm comm, int *size) { (void) comm; *size = g_size; return MPI_SUCCESS; }

int MPI_Send(const void *buf, int count, MPI_Datatype dt, int dest, int tag, MPI_Comm comm) {
	(void) comm;
	do_send(buf, count, dt, dest, tag);
	return MPI_SUCCESS;
}
int MPI_Ssend(const void *buf, int count, MPI_Datatype dt, int dest, int tag, MPI_Comm comm) { return MPI_Send(buf, count, dt, dest, tag, comm); }
int MPI_Isend(const void *buf, int count, MPI_Datatype dt, int dest, int tag, MPI_Comm comm, MPI_Request *request) {
	(void) comm;
	do_send(buf, count, dt, dest, tag);
	const int h = new_req();
	g_req[h - 1].complete = 1;
	g_req[h - 1].tag = tag;
	*request = h;
	return MPI_SUCCESS;
}
int MPI_Issend(const void *buf, int count, MPI_Datatype dt, int dest, int tag, MPI_Comm comm, MPI_Request *request) {
	return MPI_Isend(buf, count, dt, dest, tag, comm, request);
}
int MPI_Irecv(void *buf, int count, MPI_Datatype dt, int source, int tag, MPI_Comm comm, MPI_Request *request) {
	(void) comm;
	if (source != MPI_ANY_SOURCE && (source < 0 || source >= g_size || source == g_rank)) die("bad source rank");
	const int h = new_req();
	struct req *q = &g_req[h - 1];
	q->is_recv = 1;
	q->buf = buf;
	q->cap = (size_t) count * (size_t) dt;
	q->src = source;
	q->tag = tag;
	if (g_n_posted == g_posted_cap) {
		g_posted_cap *= 2;
		g_posted = (int *) realloc(g_posted, (size_t) g_posted_cap * sizeof *g_posted);
		if (!g_posted) die("out of memory");
	}
	g_posted[g_n_posted++] = h;
	*request = h;
	match();                                   /* the message may be queued already */
	return MPI_SUCCESS;
}

/* Index of the first complete request, -1 if none is complete, -2 if none is live. */
static int first_complete(int count, MPI_Request *requests) {
	int any = 0;
	for (int i = 0; i < count; ++i) {
		struct req *q = live(requests[i]);
		if (!q) continue;
		any = 1;
		if (q->complete) return i;
	}
	return any ? -1 : -2;
}

int MPI_Waitany(int count, MPI_Request *requests, int *index, MPI_Status *status) {
	for (;;) {
		const uint32_t seen = atomic_load_explicit(&g_shm->doorbell[g_rank].word, memory_order_acquire);
		progress();
		match();
		const int i = first_complete(count, requests);
		if (i == -2) { *index = MPI_UNDEFINED; return MPI_SUCCESS; }
		if (i >= 0) { *index = i; finish(&requests[i], status); return MPI_SUCCESS; }
		futex_wait(&g_shm->doorbell[g_rank].word, seen);
	}
}
int MPI_Wait(MPI_Request *request, MPI_Status *status) {
	int idx;
	return MPI_Waitany(1, request, &idx, status);
}
int MPI_Waitall(int count, MPI_Request *requests, MPI_Status *statuses) {
	for (int i = 0; i < count; ++i)
		if (live(requests[i])) MPI_Wait(&requests[i], statuses ? &statuses[i] : NULL);
		else requests[i] = MPI_REQUEST_NULL;
	return MPI_SUCCESS;
}
int MPI_Waitsome(int incount, MPI_Request *requests, int *outcount, int *indices, MPI_Status *statuses) {
	for (;;) {
		const uint32_t seen = atomic_load_explicit(&g_shm->doorbell[g_rank].word, memory_order_acquire);
		progress();
		match();
		int n = 0, any = 0;
		for (int i = 0; i < incount; ++i) {
			struct req *q = live(requests[i]);
			if (!q) continue;
			any = 1;
			if (!q->complete) continue;
			indices[n] = i;
			finish(&requests[i], statuses ? &statuses[n] : NULL);
			++n;
		}
		if (!any) { *outcount = MPI_UNDEFINED; return MPI_SUCCESS; }
		if (n) { *outcount = n; return MPI_SUCCESS; }
		futex_wait(&g_shm->doorbell[g_rank].word, seen);
	}
}
/* The reference polls this once per branch-and-bound node (src/yanbader.c:78-94): with nothing new it costs one
 * pass over the (few dozen) handles and one load of the doorbell word. */
int MPI_Testany(int count, MPI_Request *requests, int *index, int *flag, MPI_Status *status) {
	static uint32_t last_seen;
	int i = first_complete(count, requests);
	if (i == -1) {
		const uint32_t now = atomic_load_explicit(&g_shm->doorbell[g_rank].word, memory_order_acquire);
		if (now != last_seen) {
			last_seen = now;
			progress();
			match();
			i = first_complete(count, requests);
		}
	}
	if (i == -2) { *flag = 1; *index = MPI_UNDEFINED; return MPI_SUCCESS; }
	if (i < 0) { *flag = 0; *index = MPI_UNDEFINED; return MPI_SUCCESS; }
	*flag = 1;
	*index = i;
	finish(&requests[i], status);
	return MPI_SUCCESS;
}
int MPI_Recv(void *buf, int count, MPI_Datatype dt, int source, int tag, MPI_Comm comm, MPI_Status *status) {
	MPI_Request h;
	MPI_Irecv(buf, count, dt, source, tag, comm, &h);
	return MPI_Wait(&h, status);
}
int MPI_Get_count(const MPI_Status *status, MPI_Datatype dt, int *count) {
	*count = status->xmpi_bytes / dt;
	return MPI_SUCCESS;
}
