// The multi-GPU exchange protocol: one shared upper bound, asynchronous work stealing, termination.
//
// Replaces the reference's MPI boss/worker traffic (src/mpiboss.c:30-429, src/mpiworker.c:19-41,117-181,214-380:
// MSG_NEWBOUND relayed through the boss, MSG_STEALJOB / MSG_NEWJOB between workers, the boss counting idle workers)
// by words in memory that every rank can reach.  In the product that memory is DEVICE memory, one block per GPU,
// mapped into the peers over NVLink (exchange.cu: CUDA IPC handles between processes, peer access between threads);
// the search kernels themselves lower the bound word of every peer with a system-scope atomicMin the moment they
// find a better tree, so a bound travels in the launch that found it, with no host in between.  This header is the
// host-side part -- who asks whom for work, and when everybody may stop -- written against an abstract `Ops` so that
// the same code runs over host memory in the test double (tests/stub/xmp_stub_exchange.cpp) under ThreadSanitizer
// and across processes.  No collectives, no lock-step rounds: a steal involves the idle rank and one donor.
//
// Words of a rank's block (indices into its 1 KiB counter block, see search.cu for the others):
//   XW_BOUND       the upper bound.  Written by local kernels (atomicMin) and by peers (system-scope atomicMin).
//   XW_REQUEST + r 1 = rank r asks this rank for work (one slot per thief: a plain store, and a donor can serve every
//                  waiting thief between two sweeps).
//   XW_REPLY       0 = waiting, 1 + k = the donor put k jobs into this rank's mailbox (k = 0: it had none to give).
//   XW_FINISHED    1 = the search is over everywhere (written into every block by the rank that went idle last).
//   XW_BUSY        rank 0's block only: ranks that hold work, plus jobs travelling.  A donor adds the thief BEFORE the
//                  jobs leave, a rank subtracts itself when its frontier runs dry; 0 is reached exactly once.
//   XW_OPEN        hint, published by the owner: 0 = nothing to give, 1 = a single open subtree (kept), larger = an
//                  estimate of the work behind what it could give away (thieves ask the largest).
// Invariants (the ones the reference model-checked for its MPI protocol, spin/fastdnamp.mpiprom): every job is held
// by exactly one rank or is in exactly one mailbox whose owner is waiting for it; a rank only stops after XW_BUSY
// reached 0, i.e. when no rank holds work and nothing travels; a thief has at most one request out and waits for its
// reply; no request is left unanswered while the search runs (idle ranks answer "nothing" from their waiting loop).
#pragma once
#include <stddef.h>
#include <stdint.h>

namespace xmp {

enum : unsigned { XCH_MAX_RANKS = 16, XCH_MAX_JOBS = 16384 };
enum : unsigned { XW_BOUND = 0, XW_FIRST = 232, XW_REQUEST = 232, XW_REPLY = 248, XW_FINISHED = 249, XW_BUSY = 250, XW_OPEN = 251, XW_COUNT = 24 };
enum : int { XCH_GOT_WORK = 1, XCH_FINISHED = 2 };

// Ops (all return 0 or an error code; `r` is a rank, `w` a word index):
//   snapshot(unsigned out[XW_COUNT])            this rank's words XW_FIRST .. XW_FIRST + XW_COUNT - 1
//   store(r, w, v)                              release store (everything this rank wrote before is visible first)
//   add(r, w, delta, unsigned *old)             atomic add, returns the previous value
//   gather(w, unsigned out[world])              word w of every rank's block
//   pause()                                     a short wait between polls
// Eng:
//   export_for_steal(k, unsigned *n)            detaches open subtrees for k thieves into a staging buffer (n = 0: none to
//                                               spare); thief i of k is meant to get staged jobs i, i + k, i + 2k, ...
//   send_staged(thief, i, k, n)                 copies that share of the n staged descriptors into the thief's mailbox
//   import_mailbox(n)                           turns n job descriptors in this rank's mailbox into open subtrees

// Answers every pending steal request (words = this rank's snapshot).  `idle`: this rank has nothing to give (it is
// waiting for work itself).
template <class Ops, class Eng>
int xch_service(Ops &ops, Eng &eng, unsigned rank, unsigned world, const unsigned *words, bool idle) {
	unsigned thieves[XCH_MAX_RANKS], k = 0;
	for (unsigned r = 0; r < world; ++r)
		if (r != rank && words[XW_REQUEST - XW_FIRST + r]) thieves[k++] = r;
	if (k == 0) return 0;
	unsigned n = 0;
	int rc = idle ? 0 : eng.export_for_steal(k, &n);
	if (rc) return rc;
	for (unsigned i = 0; i < k; ++i) {
		const unsigned thief = thieves[i];
		const unsigned share = n > i ? (n - i + k - 1) / k : 0;            // jobs i, i + k, ... of n
		if ((rc = ops.store(rank, XW_REQUEST + thief, 0)) != 0) return rc; // the slot is free again
		if (share) {
			if ((rc = ops.add(0, XW_BUSY, 1u, nullptr)) != 0) return rc;    // the thief counts as busy before the jobs travel
			if ((rc = eng.send_staged(thief, i, k, n)) != 0) return rc;
		}
		if ((rc = ops.store(thief, XW_REPLY, 1u + share)) != 0) return rc;
	}
	return 0;
}

// Called when this rank's frontier has run dry.  Returns with *result = XCH_GOT_WORK (stolen subtrees are open in
// the engine: search on) or XCH_FINISHED (no rank holds work any more).
template <class Ops, class Eng>
int xch_steal(Ops &ops, Eng &eng, unsigned rank, unsigned world, int *result) {
	unsigned w[XW_COUNT], open[XCH_MAX_RANKS], old = 0;
	int rc = ops.add(0, XW_BUSY, 0u - 1u, &old);
	if (rc) return rc;
	if (old == 1) {                                                        // I was the last one holding work
		for (unsigned r = 0; r < world; ++r)
			if ((rc = ops.store(r, XW_FINISHED, 1u)) != 0) return rc;
		*result = XCH_FINISHED;
		return 0;
	}
	unsigned first = (rank + 1) % world;                                   // ties between donors: rotate
	for (;;) {
		if ((rc = ops.snapshot(w)) != 0) return rc;
		if (w[XW_FINISHED - XW_FIRST]) break;
		if ((rc = xch_service(ops, eng, rank, world, w, true)) != 0) return rc;
		if ((rc = ops.gather(XW_OPEN, open)) != 0) return rc;
		unsigned victim = world, best = 1;
		for (unsigned k = 0; k < world; ++k) {
			const unsigned r = (first + k) % world;
			if (r != rank && open[r] > best) { best = open[r]; victim = r; }
		}
		if (victim == world) { ops.pause(); continue; }                    // nobody has anything to spare right now
		if ((rc = ops.store(rank, XW_REPLY, 0u)) != 0) return rc;
		if ((rc = ops.store(victim, XW_REQUEST + rank, 1u)) != 0) return rc;
		for (;;) {                                                         // one request, one reply
			if ((rc = ops.snapshot(w)) != 0) return rc;
			if (w[XW_REPLY - XW_FIRST] || w[XW_FINISHED - XW_FIRST]) break;
			if ((rc = xch_service(ops, eng, rank, world, w, true)) != 0) return rc;
			ops.pause();
		}
		const unsigned reply = w[XW_REPLY - XW_FIRST];
		if (reply > 1) {
			if ((rc = eng.import_mailbox(reply - 1)) != 0) return rc;
			*result = XCH_GOT_WORK;
			return 0;
		}
		if (w[XW_FINISHED - XW_FIRST]) break;
		first = (victim + 1) % world;                                      // it had nothing to give: look elsewhere
	}
	*result = XCH_FINISHED;
	return 0;
}

}  // namespace xmp
