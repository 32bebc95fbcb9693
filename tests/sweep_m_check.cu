// the synthetic microbenchmark swept over the partial-tree size m in {4, 8, 16, 63}.
//
// A stand-alone GPU check program (no Python, no torch: it starts in a second): builds the BASELINE config-2
// alignment (64 taxa x 2^20 i.i.d. single-state sites, xorshift64*, seed 1) on the host, and for every m
//   * builds the MIDPOINT sets of the partial trees on the device (xmp_cuda_build_midpoints),
//   * times xmp_cuda_score_batch (taxon m against every edge of every tree) with CUDA events on its stream,
//   * checks the costs of the first two trees of every chunk against the CPU oracle (oracle/liboracle.so; this
//     file is test infrastructure, the oracle is only the checker) and the whole chunk against the fused
//     host-buffer path (xmp_cuda_score_insertions_host), which shares no kernel with the resident path.
// 4096 trees per m; a chunk is as many trees as fit 62 GiB of edge state (m = 63: 4 chunks of 1032 trees --
// "in tree-chunks on one GPU").  One JSON line per m on stdout.
//
//   nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -Iinclude -o tests/sweep_m_check tests/sweep_m_check.cu \
//        -Lxmp_b200 -lxmp_b200 -Loracle -loracle -Xlinker -rpath,'$ORIGIN/../xmp_b200' -Xlinker -rpath,'$ORIGIN/../oracle'
//   tests/sweep_m_check [peak_GBps] [m ...]          exit 0 = all equal, 1 = a parity check failed, 2 / 3 = CUDA / library error
#include <cuda_runtime.h>
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "xmp_cuda.h"

extern "C" {
void xo_fitch_bases_swar64(const uint64_t *s1, const uint64_t *s2, uint64_t *dest, size_t nwords);
unsigned xo_fitch_score_weight1_swar64(const uint64_t *s1, const uint64_t *s2, size_t nwords);
}

static const unsigned NUM_TAXA = 64, NUM_SITES = 1u << 20, ROW_BYTES = NUM_SITES / 2, N_BLOCKS = ROW_BYTES / 16;

#define CK(x)                                                                                        \
	do {                                                                                             \
		cudaError_t e_ = (x);                                                                        \
		if (e_ != cudaSuccess) {                                                                     \
			fprintf(stderr, "%s:%d: %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_));      \
			exit(2);                                                                                 \
		}                                                                                            \
	} while (0)
#define XK(x)                                                                                        \
	do {                                                                                             \
		int r_ = (x);                                                                                \
		if (r_ != XMP_OK) {                                                                          \
			fprintf(stderr, "%s:%d: %s -> %d: %s\n", __FILE__, __LINE__, #x, r_, xmp_cuda_last_error()); \
			exit(3);                                                                                 \
		}                                                                                            \
	} while (0)

static uint64_t rng_state;
static inline uint64_t xorshift64star() {
	rng_state ^= rng_state >> 12;
	rng_state ^= rng_state << 25;
	rng_state ^= rng_state >> 27;
	return rng_state * 0x2545F4914F6CDD1DULL;
}
static double now_s() {
	return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// Insertion costs of taxon m on every edge of the tree given by `path`, from scratch on the CPU with the oracle.
// Edge ids as in include/xmp_cuda.h: edges 0 / 1 = leaves of taxa 1 / 2, edge 2 = the first internal node; taxon t
// on edge v appends leaf edge 2t-3 and internal edge 2t-2 (LEFT = new leaf, RIGHT = v).
static void oracle_costs(const uint8_t *rows, const uint8_t *path, unsigned m, unsigned *cost) {
	const unsigned E = 2 * m - 3, NONE = 0xFFFFu;
	const size_t nw = ROW_BYTES / 8;
	std::vector<unsigned> par(E, NONE), k0(E, NONE), k1(E, NONE);
	par[0] = par[1] = 2;
	k0[2] = 0;
	k1[2] = 1;
	unsigned top = 2;
	for (unsigned t = 3; t < m; ++t) {
		const unsigned v = path[t], a = 2 * t - 3, b = 2 * t - 2, pv = par[v];
		par[b] = pv;
		if (pv == NONE) top = b;
		else if (k0[pv] == v) k0[pv] = b;
		else k1[pv] = b;
		k0[b] = a;
		k1[b] = v;
		par[a] = par[v] = b;
	}
	std::vector<unsigned> pre, stack{top};
	while (!stack.empty()) {
		const unsigned e = stack.back();
		stack.pop_back();
		pre.push_back(e);
		if (k0[e] != NONE) {
			stack.push_back(k1[e]);
			stack.push_back(k0[e]);
		}
	}
	std::vector<uint64_t> up((size_t) E * nw), down((size_t) E * nw), mid(nw);
	auto row = [&](unsigned taxon) { return (const uint64_t *) (rows + (size_t) taxon * ROW_BYTES); };
	auto upp = [&](unsigned e) -> const uint64_t * { return k0[e] == NONE ? row((e + 3) / 2) : &up[(size_t) e * nw]; };
	for (size_t i = pre.size(); i-- > 0;) {
		const unsigned e = pre[i];
		if (k0[e] != NONE) xo_fitch_bases_swar64(upp(k0[e]), upp(k1[e]), &up[(size_t) e * nw], nw);
	}
	for (unsigned e : pre) {
		const uint64_t *dn;
		if (par[e] == NONE) dn = row(0);
		else {
			const unsigned p = par[e], sib = k0[p] == e ? k1[p] : k0[p];
			const uint64_t *dp = par[p] == NONE ? row(0) : &down[(size_t) p * nw];
			xo_fitch_bases_swar64(dp, upp(sib), &down[(size_t) e * nw], nw);
			dn = &down[(size_t) e * nw];
		}
		xo_fitch_bases_swar64(upp(e), dn, mid.data(), nw);
		cost[e] = xo_fitch_score_weight1_swar64(mid.data(), row(m), nw);
	}
}

int main(int argc, char **argv) {
	const double peak = argc > 1 ? atof(argv[1]) : 6554.2;      // measured HBM copy peak, GB/s (MEASURED_PEAKS.json)
	std::vector<unsigned> ms;
	for (int i = 2; i < argc; ++i) ms.push_back((unsigned) atoi(argv[i]));
	if (ms.empty()) ms = {4, 8, 16, 63};
	// defaults = the BASELINE workload; tests/test_gpu.py runs a reduced copy (SWEEP_TREES, SWEEP_STATE_GIB)
	const unsigned N_TREES = getenv("SWEEP_TREES") ? (unsigned) atoi(getenv("SWEEP_TREES")) : 4096;
	const size_t STATE_BUDGET = (size_t) (getenv("SWEEP_STATE_GIB") ? atoi(getenv("SWEEP_STATE_GIB")) : 62) << 30;
	if (N_TREES == 0 || STATE_BUDGET < ((size_t) 1 << 30)) {
		fprintf(stderr, "SWEEP_TREES / SWEEP_STATE_GIB out of range\n");
		return 2;
	}

	double t0 = now_s();
	uint8_t *rows;
	CK(cudaMallocHost(&rows, (size_t) NUM_TAXA * ROW_BYTES));
	rng_state = 1;
	for (size_t i = 0; i < (size_t) NUM_TAXA * ROW_BYTES; i += 4) {      // 8 sites (16 random bits) per draw
		const uint64_t r = xorshift64star() >> 16;
		for (unsigned k = 0; k < 4; ++k) {
			const unsigned lo = (r >> (4 * k)) & 3u, hi = (r >> (4 * k + 2)) & 3u;
			rows[i + k] = (uint8_t) ((1u << lo) | (1u << (hi + 4)));
		}
	}
	fprintf(stderr, "alignment generated in %.2f s\n", now_s() - t0);

	xmp_ctx *ctx;
	XK(xmp_cuda_init(0, &ctx));
	cudaStream_t stream;
	CK(cudaStreamCreate(&stream));
	XK(xmp_cuda_set_stream(ctx, stream));
	XK(xmp_cuda_load(ctx, rows, nullptr, NUM_TAXA, N_BLOCKS, nullptr, 0));
	void *d_mid;
	unsigned *d_cost;
	CK(cudaMalloc(&d_mid, STATE_BUDGET));
	CK(cudaMalloc(&d_cost, (size_t) N_TREES * 123 * sizeof(unsigned)));
	cudaEvent_t e0, e1;
	CK(cudaEventCreate(&e0));
	CK(cudaEventCreate(&e1));
	fprintf(stderr, "device ready at %.2f s\n", now_s() - t0);

	int failures = 0;
	for (unsigned m : ms) {
		if (m < 3 || m >= NUM_TAXA) continue;
		const unsigned E = 2 * m - 3;
		const size_t pair_bytes = ROW_BYTES;
		unsigned chunk = (unsigned) std::min<size_t>(N_TREES, STATE_BUDGET / (pair_bytes * E));
		std::vector<uint8_t> paths((size_t) N_TREES * NUM_TAXA, 0);
		rng_state = 2 + m;
		for (unsigned i = 0; i < N_TREES; ++i)
			for (unsigned t = 3; t < m; ++t) paths[(size_t) i * NUM_TAXA + t] = (uint8_t) (xorshift64star() % (2 * t - 3));
		std::vector<unsigned> cost((size_t) N_TREES * E), cost_fused((size_t) chunk * E), want(E);
		double score_ms_total = 0, build_ms_total = 0, fused_ms_total = 0;
		std::vector<double> launch_ms;
		unsigned checked_trees = 0, n_chunks = 0;
		bool ok = true, fused_ok = true;
		for (unsigned first = 0; first < N_TREES; first += chunk, ++n_chunks) {
			const unsigned nt = std::min(chunk, N_TREES - first);
			const size_t n_pairs = (size_t) nt * E;
			const uint8_t *pp = paths.data() + (size_t) first * NUM_TAXA;
			CK(cudaStreamSynchronize(stream));
			double tb = now_s();
			XK(xmp_cuda_build_midpoints(ctx, pp, NUM_TAXA, nt, m, d_mid, nullptr));
			CK(cudaStreamSynchronize(stream));
			build_ms_total += (now_s() - tb) * 1e3;
			for (int rep = 0; rep < 2; ++rep)       // warm-up
				XK(xmp_cuda_score_batch(ctx, d_mid, n_pairs, E, m, nullptr, 0x7fffffffu, d_cost, nullptr, nullptr));
			const int REPS = 5;                      // every launch streams n_pairs x 512 KiB >> the 126 MB L2
			for (int rep = 0; rep < REPS; ++rep) {
				CK(cudaEventRecord(e0, stream));
				XK(xmp_cuda_score_batch(ctx, d_mid, n_pairs, E, m, nullptr, 0x7fffffffu, d_cost, nullptr, nullptr));
				CK(cudaEventRecord(e1, stream));
				CK(cudaEventSynchronize(e1));
				float ms;
				CK(cudaEventElapsedTime(&ms, e0, e1));
				launch_ms.push_back(ms * (double) chunk / nt);          // normalised to a full chunk
			}
			CK(cudaMemcpyAsync(cost.data() + (size_t) first * E, d_cost, n_pairs * sizeof(unsigned), cudaMemcpyDeviceToHost, stream));
			CK(cudaStreamSynchronize(stream));
			{
				std::vector<double> v(launch_ms.end() - REPS, launch_ms.end());
				std::sort(v.begin(), v.end());
				score_ms_total += v[REPS / 2] * nt / chunk;
			}
			for (unsigned i = 0; i < std::min(2u, nt); ++i) {           // oracle on the first two trees of the chunk
				oracle_costs(rows, pp + (size_t) i * NUM_TAXA, m, want.data());
				if (memcmp(want.data(), cost.data() + (size_t) (first + i) * E, E * sizeof(unsigned)) != 0) ok = false;
				++checked_trees;
			}
			if (n_chunks == 0) {                                         // the fused host-buffer path on the first chunk
				double tf = now_s();
				XK(xmp_cuda_score_insertions_host(ctx, pp, NUM_TAXA, nt, m, cost_fused.data()));
				fused_ms_total = (now_s() - tf) * 1e3;
				if (memcmp(cost_fused.data(), cost.data() + (size_t) first * E, n_pairs * sizeof(unsigned)) != 0) fused_ok = false;
			}
		}
		unsigned long long checksum = 0;
		for (unsigned c : cost) checksum += c;
		const double pairs = (double) N_TREES * E, bytes = pairs * pair_bytes;
		const double gbps = bytes / (score_ms_total * 1e-3) / 1e9;
		printf("{\"m\": %u, \"edges\": %u, \"trees\": %u, \"pairs\": %.0f, \"state_GiB\": %.2f, \"chunks\": %u, \"trees_per_chunk\": %u, "
		       "\"score_ms\": %.4f, \"GBps\": %.1f, \"frac_of_peak\": %.4f, \"peak_GBps\": %.1f, \"Tsite_merges_per_s\": %.3f, "
		       "\"build_ms\": %.2f, \"fused_host_path_ms_first_chunk\": %.2f, \"oracle_trees_checked\": %u, \"oracle_ok\": %s, "
		       "\"fused_path_equal\": %s, \"cost_checksum\": %llu}\n",
		       m, E, N_TREES, pairs, bytes / (double) (1ull << 30), n_chunks, chunk, score_ms_total, gbps, gbps / peak, peak,
		       pairs * NUM_SITES / (score_ms_total * 1e-3) / 1e12, build_ms_total, fused_ms_total, checked_trees,
		       ok ? "true" : "false", fused_ok ? "true" : "false", checksum);
		fflush(stdout);
		if (!ok || !fused_ok) ++failures;
	}
	fprintf(stderr, "done at %.2f s\n", now_s() - t0);
	xmp_cuda_destroy(ctx);
	return failures ? 1 : 0;
}
