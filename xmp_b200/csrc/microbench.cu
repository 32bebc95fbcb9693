// Register-resident microbenchmark of the integer pipe on the exact instruction mix of the Fitch
// kernels (the INT32 / LOP3 / POPC rates are not in MEASURED_PEAKS.json and must be
// measured before any fraction of an integer roofline is quoted).  No memory traffic inside the loop:
// every thread keeps 8 independent words in registers and applies either the weight-1 cost
// (AND, AND+ADD, LOP3, POPC, ADD per 8 sites) or the state-set merge (fitch_word) to them.
#include "ctx.h"

namespace xmp {

template <int MODE>   // 0 = weight-1 cost, 1 = merge (FitchBases), 2 = merge followed by cost (one search step)
__global__ void __launch_bounds__(256)
k_int_peak(unsigned iters, unsigned seed, unsigned *sink) {
	unsigned a[8], b[8], acc = 0;
#pragma unroll
	for (int i = 0; i < 8; ++i) {
		a[i] = (threadIdx.x * 2654435761u + i * 40503u + seed) | 0x11111111u;
		b[i] = (blockIdx.x * 2246822519u + i * 7919u + seed) | 0x11111111u;
	}
	for (unsigned it = 0; it < iters; ++it) {
#pragma unroll
		for (int i = 0; i < 8; ++i) {
			if (MODE == 0) {
				acc += cost_word_w1(a[i], b[i]);
				a[i] += acc;                       // keep the chain live without adding Fitch work
			} else if (MODE == 1) {
				a[i] = fitch_word(a[i], b[i]) + it;
			} else {
				unsigned d = fitch_word(a[i], b[i]);
				acc += cost_word_w1(d, b[(i + 1) & 7]);
				a[i] = d + it;
			}
		}
	}
#pragma unroll
	for (int i = 0; i < 8; ++i) acc += a[i];
	if (acc == 0x12345678u) *sink = acc;        // never true in practice; defeats dead-code elimination
}

}  // namespace xmp

using namespace xmp;

// Site operations per second the integer pipe sustains on the three instruction mixes, in G/s.
// out[0] = weight-1 cost, out[1] = merge, out[2] = merge + cost.  Each word is 8 sites.
extern "C" int xmp_cuda_int_pipe_peak(xmp_ctx *ctx, double *out) {
	if (!ctx || !out) return set_error(XMP_ERR_ARG, "null argument");
	XMP_CUDA(cudaSetDevice(ctx->device));
	int rc = ctx->scratch.reserve(256);
	if (rc) return rc;
	const unsigned iters = 4096, blocks = (unsigned) ctx->sm_count * 8;
	for (int mode = 0; mode < 3; ++mode) {
		float best = 1e30f;
		for (int rep = 0; rep < 4; ++rep) {
			XMP_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
			if (mode == 0) k_int_peak<0><<<blocks, 256, 0, ctx->stream>>>(iters, (unsigned) rep, (unsigned *) ctx->scratch.ptr);
			else if (mode == 1) k_int_peak<1><<<blocks, 256, 0, ctx->stream>>>(iters, (unsigned) rep, (unsigned *) ctx->scratch.ptr);
			else k_int_peak<2><<<blocks, 256, 0, ctx->stream>>>(iters, (unsigned) rep, (unsigned *) ctx->scratch.ptr);
			XMP_CUDA(cudaGetLastError());
			XMP_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
			XMP_CUDA(cudaEventSynchronize(ctx->ev1));
			float ms = 0;
			XMP_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
			if (rep > 0 && ms < best) best = ms;
		}
		const double sites = (double) blocks * 256.0 * iters * 8.0 * 8.0;
		out[mode] = sites / (best * 1e-3) / 1e9;
	}
	return XMP_OK;
}
