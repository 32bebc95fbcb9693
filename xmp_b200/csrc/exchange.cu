// The multi-GPU exchange over peer-mapped device memory (NVLink): replaces the reference's MPI transport
// (src/mpiboss.c, src/mpiworker.c) for the three things ranks tell each other -- a better upper bound, "give me
// work", "everybody is done".  exchange_proto.h holds the protocol; this file is its device side:
//
//   * every context owns one exchange block in device memory (ctx.h).  Between processes (torchrun, one rank per
//     GPU) the blocks are shared as CUDA IPC handles; between threads of one process (fastdnamp_b200 --gpus=N)
//     by enabling peer access.  Either way a rank ends up with a table of pointers, one per rank, that its own
//     kernels can dereference: loads, stores and atomics on a peer's block travel over NVLink / NVSwitch.
//   * the bound: search kernels call lower_bound() (search.cu) -- atomicMin on the local word and, only when that
//     actually lowered it, a system-scope atomicMin on every peer's word.  Each launch reads the local word once
//     per CTA, so a bound found anywhere prunes everywhere from the next launch on, without any host round.
//   * steal requests, replies, the busy counter and the finished flag are single words updated by one-thread
//     kernels (k_xch_rmw) with system-scope atomics; job descriptors go donor -> thief as a device-to-device copy
//     into the thief's mailbox.
#include <chrono>
#include <thread>
#include <string.h>
#include "ctx.h"
#include "exchange_proto.h"

namespace xmp {

__global__ void k_xch_rmw(unsigned *word, int op, unsigned a, unsigned b, unsigned *result) {
	unsigned old = 0;
	switch (op) {
	case 0:                                  // release store: whatever this rank wrote before (job descriptors) is visible first
		__threadfence_system();
		old = atomicExch_system(word, a);
		break;
	case 1: old = atomicAdd_system(word, a); break;
	default: old = atomicMin_system(word, a + 0u * b); break;
	}
	__threadfence_system();
	*result = old;                           // pinned host memory
}

__global__ void k_xch_gather(const XchPeerTable *pt, unsigned word, unsigned *out) {
	const unsigned r = threadIdx.x;
	if (r < pt->world) out[r] = *(volatile unsigned *) (pt->blk[r] + word);
}

int xch_ensure_block(xmp_ctx *ctx) {
	if (ctx->xch) return XMP_OK;
	XMP_CUDA(cudaSetDevice(ctx->device));
	XMP_CUDA(cudaMalloc(&ctx->xch, XCH_BLOCK_BYTES));
	XMP_CUDA(cudaMemset(ctx->xch, 0, XCH_BLOCK_BYTES));
	XMP_CUDA(cudaMallocHost(&ctx->h_xres, 256));
	memset(ctx->h_xres, 0, 256);
	ctx->xch_peer[0] = ctx->xch;
	return XMP_OK;
}

static void xch_disconnect(xmp_ctx *ctx) {
	for (unsigned r = 0; r < XCH_MAX_RANKS; ++r) {
		if (ctx->xch_ipc[r] && ctx->xch_peer[r]) cudaIpcCloseMemHandle(ctx->xch_peer[r]);
		ctx->xch_ipc[r] = false;
		ctx->xch_peer[r] = nullptr;
	}
	ctx->xch_connected = false;
	ctx->xch_rank = 0;
	ctx->xch_world = 1;
	ctx->xch_peer[0] = ctx->xch;
}

void xch_release(xmp_ctx *ctx) {
	xch_disconnect(ctx);
	if (ctx->xch) cudaFree(ctx->xch);
	if (ctx->h_xres) cudaFreeHost(ctx->h_xres);
	ctx->xch = nullptr;
	ctx->h_xres = nullptr;
}

// Uploads the pointer table and marks the context connected.
static int xch_finish_connect(xmp_ctx *ctx, unsigned rank, unsigned world) {
	XchPeerTable t;
	memset(&t, 0, sizeof t);
	for (unsigned r = 0; r < world; ++r) t.blk[r] = ctx->xch_peer[r];
	t.world = world;
	t.rank = rank;
	XMP_CUDA(cudaMemcpy((char *) ctx->xch + XCH_TABLE_OFF, &t, sizeof t, cudaMemcpyHostToDevice));
	ctx->xch_rank = rank;
	ctx->xch_world = world;
	ctx->xch_connected = world > 1;
	// first launches load the kernels' code (lazy module loading, milliseconds): not inside somebody's search
	k_xch_rmw<<<1, 1, 0, ctx->stream>>>(ctx->xch + 7, 1, 0u, 0u, ctx->h_xres);
	k_xch_gather<<<1, 32, 0, ctx->stream>>>((const XchPeerTable *) ((char *) ctx->xch + XCH_TABLE_OFF), XW_OPEN, ctx->h_xres + 40);
	XMP_CUDA(cudaGetLastError());
	XMP_CUDA(cudaStreamSynchronize(ctx->stream));
	return XMP_OK;
}

// One control operation on word w of rank r's block; the previous value comes back through pinned memory.
struct DevOps {
	xmp_ctx *ctx;
	int rmw(unsigned r, unsigned w, int op, unsigned a, unsigned b, unsigned *old) {
		k_xch_rmw<<<1, 1, 0, ctx->stream>>>(ctx->xch_peer[r] + w, op, a, b, ctx->h_xres);
		XMP_CUDA(cudaGetLastError());
		XMP_CUDA(cudaStreamSynchronize(ctx->stream));
		if (old) *old = ctx->h_xres[0];
		return XMP_OK;
	}
	int store(unsigned r, unsigned w, unsigned v) { return rmw(r, w, 0, v, 0, nullptr); }
	int add(unsigned r, unsigned w, unsigned delta, unsigned *old) { return rmw(r, w, 1, delta, 0, old); }
	int snapshot(unsigned *out) {
		XMP_CUDA(cudaMemcpyAsync(ctx->h_xres + 16, ctx->xch + XW_FIRST, XW_COUNT * sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream));
		XMP_CUDA(cudaStreamSynchronize(ctx->stream));
		memcpy(out, ctx->h_xres + 16, XW_COUNT * sizeof(unsigned));           // h_xres: [0] rmw result, [8] hint, [16, 40) snapshot, [40, 56) gather
		return XMP_OK;
	}
	int gather(unsigned w, unsigned *out) {
		k_xch_gather<<<1, 32, 0, ctx->stream>>>((const XchPeerTable *) ((char *) ctx->xch + XCH_TABLE_OFF), w, ctx->h_xres + 40);
		XMP_CUDA(cudaGetLastError());
		XMP_CUDA(cudaStreamSynchronize(ctx->stream));
		memcpy(out, ctx->h_xres + 40, ctx->xch_world * sizeof(unsigned));
		return XMP_OK;
	}
	void pause() { std::this_thread::sleep_for(std::chrono::microseconds(20)); }
};

// search.cu: the engine side of the protocol
int search_export_for_steal(xmp_ctx *ctx, unsigned k, unsigned *n);
int search_send_staged(xmp_ctx *ctx, unsigned thief, unsigned i, unsigned k, unsigned n);
int search_import_mailbox(xmp_ctx *ctx, unsigned n);
int search_after_steal(xmp_ctx *ctx, bool finished);

struct DevEng {
	xmp_ctx *ctx;
	int export_for_steal(unsigned k, unsigned *n) { return search_export_for_steal(ctx, k, n); }
	int send_staged(unsigned thief, unsigned i, unsigned k, unsigned n) { return search_send_staged(ctx, thief, i, k, n); }
	int import_mailbox(unsigned n) { return search_import_mailbox(ctx, n); }
};

// Called by the search between sweeps: answers the pending requests out of the open frontier.
// words = this rank's protocol words (XW_FIRST ...) as they came back with the sweep's counters.
int xch_service_requests(xmp_ctx *ctx, const unsigned *words) {
	DevOps ops{ctx};
	DevEng eng{ctx};
	return xch_service(ops, eng, ctx->xch_rank, ctx->xch_world, words, false);
}

}  // namespace xmp

using namespace xmp;

extern "C" int xmp_cuda_exchange_handle(xmp_ctx *ctx, unsigned char *handle) {
	if (!ctx || !handle) return set_error(XMP_ERR_ARG, "null argument");
	static_assert(sizeof(cudaIpcMemHandle_t) == XMP_EXCHANGE_HANDLE_BYTES, "handle size");
	int rc = xch_ensure_block(ctx);
	if (rc) return rc;
	cudaIpcMemHandle_t h;
	XMP_CUDA(cudaIpcGetMemHandle(&h, ctx->xch));
	memcpy(handle, &h, sizeof h);
	return XMP_OK;
}

extern "C" int xmp_cuda_exchange_connect(xmp_ctx *ctx, unsigned rank, unsigned world, const unsigned char *handles) {
	if (!ctx || !handles) return set_error(XMP_ERR_ARG, "null argument");
	if (world < 1 || world > XCH_MAX_RANKS || rank >= world) return set_error(XMP_ERR_ARG, "rank %u of %u: at most %u ranks", rank, world, (unsigned) XCH_MAX_RANKS);
	int rc = xch_ensure_block(ctx);
	if (rc) return rc;
	XMP_CUDA(cudaSetDevice(ctx->device));
	xch_disconnect(ctx);
	for (unsigned r = 0; r < world; ++r) {
		if (r == rank) { ctx->xch_peer[r] = ctx->xch; continue; }
		cudaIpcMemHandle_t h;
		memcpy(&h, handles + (size_t) r * XMP_EXCHANGE_HANDLE_BYTES, sizeof h);
		void *p = nullptr;
		cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
		if (e != cudaSuccess) {
			xch_disconnect(ctx);
			return cuda_failed(e, "cudaIpcOpenMemHandle (are the GPUs NVLink peers, and the ranks on one node?)", __FILE__, __LINE__);
		}
		ctx->xch_peer[r] = (unsigned *) p;
		ctx->xch_ipc[r] = true;
	}
	return xch_finish_connect(ctx, rank, world);
}

extern "C" int xmp_cuda_exchange_connect_local(xmp_ctx *ctx, unsigned rank, unsigned world, xmp_ctx *const *peers) {
	if (!ctx || !peers) return set_error(XMP_ERR_ARG, "null argument");
	if (world < 1 || world > XCH_MAX_RANKS || rank >= world || peers[rank] != ctx) return set_error(XMP_ERR_ARG, "rank %u of %u: bad peer list", rank, world);
	XMP_CUDA(cudaSetDevice(ctx->device));
	xch_disconnect(ctx);
	for (unsigned r = 0; r < world; ++r) {
		if (!peers[r] || !peers[r]->xch) return set_error(XMP_ERR_ARG, "peer %u has no exchange block yet (call xmp_cuda_exchange_handle on every context first)", r);
		if (r != rank && peers[r]->device != ctx->device) {
			int can = 0;
			XMP_CUDA(cudaDeviceCanAccessPeer(&can, ctx->device, peers[r]->device));
			if (!can) return set_error(XMP_ERR_NODEVICE, "device %d cannot address device %d: the multi-GPU search needs peer access (NVLink)", ctx->device, peers[r]->device);
			cudaError_t e = cudaDeviceEnablePeerAccess(peers[r]->device, 0);
			if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
			else if (e != cudaSuccess) return cuda_failed(e, "cudaDeviceEnablePeerAccess", __FILE__, __LINE__);
		}
		ctx->xch_peer[r] = peers[r]->xch;
	}
	return xch_finish_connect(ctx, rank, world);
}

extern "C" int xmp_cuda_exchange_disconnect(xmp_ctx *ctx) {
	if (!ctx) return set_error(XMP_ERR_ARG, "null context");
	cudaSetDevice(ctx->device);
	cudaStreamSynchronize(ctx->stream);
	xch_disconnect(ctx);
	return XMP_OK;
}

// Before every search, on every rank, followed by a barrier of the caller's (only then may peers write here again).
extern "C" int xmp_cuda_exchange_reset(xmp_ctx *ctx) {
	if (!ctx) return set_error(XMP_ERR_ARG, "null context");
	int rc = xch_ensure_block(ctx);
	if (rc) return rc;
	XMP_CUDA(cudaSetDevice(ctx->device));
	unsigned words[256];
	memset(words, 0, sizeof words);
	words[XW_BOUND] = 0x7fffffffu;
	words[XW_BUSY] = ctx->xch_world;               // meaningful on rank 0: every rank starts out holding work
	XMP_CUDA(cudaMemcpyAsync(ctx->xch, words, sizeof words, cudaMemcpyHostToDevice, ctx->stream));
	XMP_CUDA(cudaStreamSynchronize(ctx->stream));
	return XMP_OK;
}

extern "C" int xmp_cuda_search_steal(xmp_ctx *ctx, int *state) {
	if (!ctx || !ctx->search || !state) return set_error(XMP_ERR_ARG, "no search in progress");
	if (!ctx->xch_connected) {                     // a single rank: running dry is the end
		*state = XMP_STEAL_FINISHED;
		return search_after_steal(ctx, true);
	}
	XMP_CUDA(cudaSetDevice(ctx->device));
	DevOps ops{ctx};
	DevEng eng{ctx};
	int result = 0;
	int rc = xch_steal(ops, eng, ctx->xch_rank, ctx->xch_world, &result);
	if (rc) return rc;
	*state = result == XCH_GOT_WORK ? XMP_STEAL_GOT_WORK : XMP_STEAL_FINISHED;
	return search_after_steal(ctx, result != XCH_GOT_WORK);
}
