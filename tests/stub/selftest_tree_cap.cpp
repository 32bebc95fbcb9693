// TEST INFRASTRUCTURE ONLY -- replays the search's host-side tree bookkeeping (xmp_b200/csrc/tree_book.h: filtering on a
// falling bound, the -m cut, counting what was cut) on the CPU.  Built by tests/test_abi.py into a scratch directory;
// the product library does not contain this entry point.
//
// Batch b brings batch_records[b] records of rec_bytes = 4 + ((num_taxa + 3) & ~3) bytes (length, path) while the bound
// stands at batch_bounds[b].  Returns what xmp_cuda_search_result / _stored would report after the last batch.
#include "../../xmp_b200/csrc/tree_book.h"

extern "C" int xmp_selftest_tree_cap(const uint8_t *records, const unsigned *batch_records, const unsigned *batch_bounds,
                                     unsigned n_batches, unsigned num_taxa, unsigned max_trees, unsigned long long *n_total,
                                     unsigned long long *n_stored, uint8_t *paths, size_t cap) {
	if (!records || num_taxa < 4 || num_taxa > XMP_MAX_TAXA || !n_total || !n_stored) return XMP_ERR_ARG;
	xmp::TreeBook book;
	book.n = num_taxa;
	book.trec = 4 + ((num_taxa + 3u) & ~3u);
	book.max_trees = max_trees;
	size_t at = 0;
	unsigned bound = 0xffffffffu;
	for (unsigned b = 0; b < n_batches; ++b) {
		const size_t bytes = (size_t) batch_records[b] * book.trec, old = book.recs.size();
		bound = batch_bounds[b];
		book.recs.resize(old + bytes);
		memcpy(book.recs.data() + old, records + at, bytes);
		at += bytes;
		book.absorb(old, bound);
	}
	*n_total = book.optimal(bound, n_stored);
	size_t k = 0;
	for (size_t r = 0; r < book.recs.size() && paths && k < cap; r += book.trec) {
		if (book.length_of(r) == bound) memcpy(paths + k++ * num_taxa, book.recs.data() + r + 4, num_taxa);
	}
	return XMP_OK;
}
