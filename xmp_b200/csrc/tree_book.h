// Host-side bookkeeping of the tree records that leave the device: filtering on a falling bound, the -m cut in
// bounded memory, counting what was cut.  Plain C++ (no CUDA), so that the very code the search runs can be
// replayed on the CPU by a test-only library (tests/stub/selftest_tree_cap.cpp); the product library exports
// only the search and xmp_cuda_paths_keep_first_dfs.
//
// The reference keeps the first maxTrees optimal trees its depth-first search meets and goes on counting
// (src/yanbader.c:272-293).  "First" is lexicographic order of the job coordinates: at every tree size the
// pre-order position of the edge that received the next taxon (src/yanbader.c:103,238-243).  A path in the
// library's append-only edge ids becomes those coordinates by replaying the insertions on a position table:
// inserting at position P puts the new internal edge at P, the new leaf at P + 1, and moves everything that
// sat at P or behind it two places back.
#pragma once
#include <algorithm>
#include <numeric>
#include <thread>
#include <vector>
#include <stdint.h>
#include <string.h>
#include "../../include/xmp_cuda.h"

namespace xmp {

inline void dfs_key(const uint8_t *path, unsigned n, uint8_t *key) {
	uint8_t pos[2 * XMP_MAX_TAXA];
	pos[2] = 0;
	pos[0] = 1;
	pos[1] = 2;
	key[0] = key[1] = key[2] = 0;
	for (unsigned t = 3; t < n; ++t) {
		const unsigned E = 2 * t - 3, P = pos[path[t]];
		for (unsigned e = 0; e < E; ++e)
			if (pos[e] >= P) pos[e] = (uint8_t) (pos[e] + 2);
		pos[2 * t - 3] = (uint8_t) (P + 1);
		pos[2 * t - 2] = (uint8_t) P;
		key[t] = (uint8_t) P;
	}
}

// key(path) < thr in that order?  Stops at the first tree size where the two differ: a tree that does not come before
// the threshold is usually told apart within the first few insertions, long before its whole key is known.
inline bool dfs_key_before(const uint8_t *path, unsigned n, const uint8_t *thr) {
	uint8_t pos[2 * XMP_MAX_TAXA];
	pos[2] = 0;
	pos[0] = 1;
	pos[1] = 2;
	for (unsigned t = 3; t < n; ++t) {
		const unsigned E = 2 * t - 3, P = pos[path[t]];
		if (P != thr[t]) return P < thr[t];
		for (unsigned e = 0; e < E; ++e)
			if (pos[e] >= P) pos[e] = (uint8_t) (pos[e] + 2);
		pos[2 * t - 3] = (uint8_t) (P + 1);
		pos[2 * t - 2] = (uint8_t) P;
	}
	return false;
}

// Moves the `keep` records that come first in the reference's depth-first order to the front of recs (in that
// order); the others follow in no particular order.  path_off = offset of the path inside a record.
inline void keep_first_dfs(uint8_t *recs, size_t n_rec, size_t stride, size_t path_off, unsigned n, size_t keep) {
	if (n_rec == 0 || keep == 0) return;
	keep = std::min(keep, n_rec);
	std::vector<uint8_t> keys(n_rec * n);
	unsigned n_thr = std::max(1u, std::min(std::thread::hardware_concurrency(), 32u));
	if (n_rec < 65536) n_thr = 1;
	auto work = [&](size_t lo, size_t hi) {
		for (size_t i = lo; i < hi; ++i) dfs_key(recs + i * stride + path_off, n, keys.data() + i * n);
	};
	if (n_thr == 1) work(0, n_rec);
	else {
		std::vector<std::thread> th;
		for (unsigned k = 0; k < n_thr; ++k) th.emplace_back(work, n_rec * k / n_thr, n_rec * (k + 1) / n_thr);
		for (auto &x : th) x.join();
	}
	std::vector<size_t> idx(n_rec);
	std::iota(idx.begin(), idx.end(), (size_t) 0);
	auto less = [&](size_t a, size_t b) {
		const int c = memcmp(keys.data() + a * n, keys.data() + b * n, n);
		return c != 0 ? c < 0 : a < b;
	};
	if (keep < n_rec) std::nth_element(idx.begin(), idx.begin() + keep, idx.end(), less);
	std::sort(idx.begin(), idx.begin() + keep, less);
	std::vector<uint8_t> out(n_rec * stride);
	for (size_t i = 0; i < n_rec; ++i) memcpy(out.data() + i * stride, recs + idx[i] * stride, stride);
	memcpy(recs, out.data(), n_rec * stride);
}

// Records (length, path) of trees that passed the bound test at the time they were found.
struct TreeBook {
	std::vector<uint8_t> recs;            // trec bytes each: 4-byte length, then the insertion path
	unsigned trec = 0, n = 0;             // record size, taxa
	unsigned max_trees = 0x7fffffffu;     // the -m cap (0x7fffffff = none)
	unsigned host_bound = 0xffffffffu;    // the bound recs was last filtered against
	unsigned long long dropped = 0;       // optimal trees counted but no longer held (the -m cut)
	unsigned dropped_score = 0;           //   ... and the length they had: they only count while that is still the bound
	// After a cut the held records are the first max_trees seen so far; a later tree whose key does not come before the
	// last of them can never be among the first max_trees overall: it is counted and not even stored.  (53humans.29 under
	// -Bp -m 1000 has 88 M optimal trees: without this every one of them was stored, keyed in full and selected against.)
	bool have_thr = false;
	uint8_t thr[XMP_MAX_TAXA + 4];        // key of the last record kept by the latest cut (valid while dropped_score is the bound)

	void reset() {
		recs.clear();
		dropped = 0;
		dropped_score = 0;
		host_bound = 0xffffffffu;
		have_thr = false;
	}
	unsigned length_of(size_t r) const {
		unsigned sc;
		memcpy(&sc, recs.data() + r, 4);
		return sc;
	}
	// With a finite max_trees the records held are cut back to the max_trees that come first whenever they outgrow
	// 4 x max_trees (and a million): a search with billions of optimal trees then needs the memory of a few million
	// records instead of all of them.  What is cut is still counted as long as its length stays the bound.  Every
	// shard keeps its own first max_trees; the first max_trees overall are among them.
	void cap(unsigned bound) {
		const size_t n_rec = recs.size() / trec, keep = max_trees;
		if (max_trees >= 0x7fffffffu || keep == 0) return;
		if (n_rec <= std::max<size_t>(4 * keep, (size_t) 1 << 20)) return;
		for (size_t r = 0; r < recs.size(); r += trec)       // all at the bound? (they are, once filtered)
			if (length_of(r) != bound) return;
		if (dropped_score != bound) dropped = 0;
		keep_first_dfs(recs.data(), n_rec, trec, 4, n, keep);
		recs.resize(keep * trec);
		dropped += n_rec - keep;
		dropped_score = bound;
		dfs_key(recs.data() + (keep - 1) * trec + 4, n, thr);     // the kept records are in order: the last one is the threshold
		have_thr = true;
	}
	// Records [old, end) have just arrived.  Drop what the current bound already rules out -- only the new records,
	// unless the bound has moved since the older ones were checked -- and, after a cut, what cannot be among the first
	// max_trees any more (counted, not stored); then apply the -m cut.
	void absorb(size_t old, unsigned bound) {
		const size_t first = bound < host_bound ? 0 : old;
		host_bound = bound;
#ifdef XMP_TREEBOOK_NO_PREFILTER                 // A/B switch of scripts/cpu_tree_book_bench.py only: the round-2 code before the pre-filter
		const bool use_thr = false;
#else
		const bool use_thr = have_thr && dropped_score == bound && dropped > 0;
#endif
		const size_t n_new = (recs.size() - first) / trec;
		std::vector<uint8_t> keepf(n_new);
		unsigned n_thr = std::max(1u, std::min(std::thread::hardware_concurrency(), 32u));
		if (n_new < 65536) n_thr = 1;
		std::vector<unsigned long long> counted(n_thr, 0);
		auto work = [&](unsigned k, size_t lo, size_t hi) {
			for (size_t i = lo; i < hi; ++i) {
				const size_t r = first + i * trec;
				const unsigned len = length_of(r);
				bool keep = len <= bound;
				if (keep && use_thr && len == bound && !dfs_key_before(recs.data() + r + 4, n, thr)) {
					keep = false;
					++counted[k];
				}
				keepf[i] = keep;
			}
		};
		if (n_thr == 1) work(0, 0, n_new);
		else {
			std::vector<std::thread> th;
			for (unsigned k = 0; k < n_thr; ++k) th.emplace_back(work, k, n_new * k / n_thr, n_new * (k + 1) / n_thr);
			for (auto &x : th) x.join();
		}
		size_t w = first;
		for (size_t i = 0; i < n_new; ++i) {
			if (!keepf[i]) continue;
			const size_t r = first + i * trec;
			if (w != r) memmove(recs.data() + w, recs.data() + r, trec);
			w += trec;
		}
		recs.resize(w);
		for (unsigned k = 0; k < n_thr; ++k) dropped += counted[k];
		cap(bound);
	}
	// Optimal trees known: how many there are, and for how many of them the path is still held.
	unsigned long long optimal(unsigned bound, unsigned long long *stored) const {
		unsigned long long count = 0;
		for (size_t r = 0; r < recs.size(); r += trec) count += length_of(r) == bound;
		if (stored) *stored = count;
		return count + (dropped_score == bound ? dropped : 0);
	}
};

}  // namespace xmp
