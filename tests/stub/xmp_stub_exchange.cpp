// TEST INFRASTRUCTURE ONLY -- the exchange entry points of include/xmp_cuda.h (2d) for the CPU test double.
//
// Runs the product's protocol source (xmp_b200/csrc/exchange_proto.h: steal requests and replies, the busy counter,
// termination) over HOST memory instead of peer-mapped device memory: a static region when the ranks are threads of
// one process (fastdnamp_b200 --gpus=N over the stub, also under ThreadSanitizer), a file in /dev/shm named by
// XMP_STUB_SHM when they are processes (tests/test_distributed.py over gloo).  The words are std::atomic, which is what
// the system-scope atomics of exchange.cu are on the device.
#include <atomic>
#include <chrono>
#include <thread>
#include <fcntl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <unistd.h>
#include "xmp_stub.h"
#include "../../xmp_b200/csrc/exchange_proto.h"

using namespace xmp;

namespace {

struct Region {
	std::atomic<unsigned> word[XCH_MAX_RANKS][256];
	uint8_t mail[XCH_MAX_RANKS][XCH_MAX_JOBS * XMP_MAX_TAXA];
};
static_assert(std::atomic<unsigned>::is_always_lock_free, "words must be plain memory");

Region *g_region = nullptr;
Region g_static;       // threads of one process

Region *region() {
	if (g_region) return g_region;
	const char *name = getenv("XMP_STUB_SHM");
	if (!name || !*name) return g_region = &g_static;
	const int fd = open(name, O_RDWR | O_CREAT, 0600);
	if (fd < 0 || ftruncate(fd, sizeof(Region)) != 0) return nullptr;
	void *p = mmap(nullptr, sizeof(Region), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
	close(fd);
	return g_region = (p == MAP_FAILED ? nullptr : (Region *) p);
}

struct HostOps {
	xmp_ctx *c;
	int snapshot(unsigned *out) {
		for (unsigned i = 0; i < XW_COUNT; ++i) out[i] = region()->word[c->xch_rank][XW_FIRST + i].load();
		return 0;
	}
	int store(unsigned r, unsigned w, unsigned v) { region()->word[r][w].store(v); return 0; }
	int add(unsigned r, unsigned w, unsigned delta, unsigned *old) {
		const unsigned o = region()->word[r][w].fetch_add(delta);
		if (old) *old = o;
		return 0;
	}
	int gather(unsigned w, unsigned *out) {
		for (unsigned r = 0; r < c->xch_world; ++r) out[r] = region()->word[r][w].load();
		return 0;
	}
	void pause() { std::this_thread::sleep_for(std::chrono::microseconds(50)); }
};

struct HostEng {
	xmp_ctx *c;
	// the share t / (t + 1) of the open subtrees for t thieves, oldest (shallowest) first, never the only one
	int export_for_steal(unsigned thieves, unsigned *n) {
		const size_t open = c->n_jobs - c->head;
		unsigned k = 0;
		if (open >= 2) {
			k = (unsigned) ((open * thieves + thieves) / (thieves + 1));
			if (k >= open) k = (unsigned) open - 1;
		}
		if (k > XCH_MAX_JOBS) k = XCH_MAX_JOBS;
		c->staged = (uint8_t *) realloc(c->staged, (size_t) (k ? k : 1) * c->n);
		for (unsigned i = 0; i < k; ++i) memcpy(c->staged + (size_t) i * c->n, c->jobs + c->head++ * c->n, c->n);
		c->n_staged = k;
		if (k) c->steals_served++;
		*n = k;
		return 0;
	}
	int send_staged(unsigned thief, unsigned i, unsigned k, unsigned n) {
		unsigned out = 0;
		for (unsigned j = i; j < n; j += k) memcpy(region()->mail[thief] + (size_t) out++ * c->n, c->staged + (size_t) j * c->n, c->n);
		return 0;
	}
	int import_mailbox(unsigned n) {
		const uint8_t *jobs = region()->mail[c->xch_rank];
		if (getenv("XMP_STUB_LOG")) fprintf(stderr, "stub: device %d imports %u jobs\n", c->device, n);
		for (unsigned i = 0; i < n; ++i) stub_push_job(c, jobs + (size_t) i * c->n, jobs[(size_t) i * c->n]);
		c->steals_received++;
		return 0;
	}
};

void atomic_min(std::atomic<unsigned> &w, unsigned v) {
	unsigned cur = w.load();
	while (v < cur && !w.compare_exchange_weak(cur, v)) {}
}

}  // namespace

extern "C" {

int xmp_cuda_exchange_handle(xmp_ctx *c, unsigned char *handle) {
	if (!c || !handle) return stub_fail(XMP_ERR_ARG, "null argument");
	memset(handle, 0, XMP_EXCHANGE_HANDLE_BYTES);
	handle[0] = (unsigned char) c->device;
	return region() ? XMP_OK : stub_fail(XMP_ERR_NODEVICE, "cannot map the shared region");
}

static int connect_common(xmp_ctx *c, unsigned rank, unsigned world) {
	if (!c || rank >= world || world > XCH_MAX_RANKS || !region()) return stub_fail(XMP_ERR_ARG, "bad exchange arguments");
	// XMP_STUB_FAIL_CONNECT=<rank>: that rank cannot map its peers (what a node without peer access looks like to the driver)
	const char *bad = getenv("XMP_STUB_FAIL_CONNECT");
	if (bad && *bad && (unsigned) atoi(bad) == rank) return stub_fail(XMP_ERR_NODEVICE, "stub: peers cannot be mapped");
	c->xch_rank = rank;
	c->xch_world = world;
	c->xch_connected = world > 1;
	return XMP_OK;
}
int xmp_cuda_exchange_connect(xmp_ctx *c, unsigned rank, unsigned world, const unsigned char *handles) {
	(void) handles;
	return connect_common(c, rank, world);
}
int xmp_cuda_exchange_connect_local(xmp_ctx *c, unsigned rank, unsigned world, xmp_ctx *const *peers) {
	(void) peers;
	return connect_common(c, rank, world);
}
int xmp_cuda_exchange_disconnect(xmp_ctx *c) {
	if (c) c->xch_connected = 0;
	return XMP_OK;
}
int xmp_cuda_exchange_reset(xmp_ctx *c) {
	if (!c || !region()) return stub_fail(XMP_ERR_ARG, "no exchange");
	for (unsigned w = 0; w < 256; ++w) region()->word[c->xch_rank][w].store(0);
	region()->word[c->xch_rank][XW_BOUND].store(0x7fffffffu);
	region()->word[c->xch_rank][XW_BUSY].store(c->xch_world);
	return XMP_OK;
}

unsigned stub_xch_bound_in(xmp_ctx *c) { return region()->word[c->xch_rank][XW_BOUND].load(); }
void stub_xch_bound_local(xmp_ctx *c, unsigned v) { atomic_min(region()->word[c->xch_rank][XW_BOUND], v); }
void stub_xch_bound_out(xmp_ctx *c, unsigned v) {
	for (unsigned r = 0; r < c->xch_world; ++r) atomic_min(region()->word[r][XW_BOUND], v);
}
int stub_xch_after_job(xmp_ctx *c) {
	HostOps ops{c};
	HostEng eng{c};
	const size_t open = c->n_jobs - c->head;
	region()->word[c->xch_rank][XW_OPEN].store((unsigned) open);
	unsigned w[XW_COUNT];
	ops.snapshot(w);
	const int rc = xch_service(ops, eng, c->xch_rank, c->xch_world, w, false);
	region()->word[c->xch_rank][XW_OPEN].store((unsigned) (c->n_jobs - c->head));
	return rc;
}

int xmp_cuda_search_steal(xmp_ctx *c, int *state) {
	if (!c || !state) return stub_fail(XMP_ERR_ARG, "null argument");
	if (!c->xch_connected) { *state = XMP_STEAL_FINISHED; return XMP_OK; }
	HostOps ops{c};
	HostEng eng{c};
	region()->word[c->xch_rank][XW_OPEN].store(0);
	int result = 0;
	const int rc = xch_steal(ops, eng, c->xch_rank, c->xch_world, &result);
	if (rc) return stub_fail(XMP_ERR_ARG, "steal failed");
	*state = result == XCH_GOT_WORK ? XMP_STEAL_GOT_WORK : XMP_STEAL_FINISHED;
	const unsigned b = stub_xch_bound_in(c);
	if (b < c->bound) { c->bound = b; c->n_trees = 0; }
	return XMP_OK;
}

}  // extern "C"
