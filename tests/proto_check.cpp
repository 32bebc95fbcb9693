// TEST INFRASTRUCTURE ONLY -- the exchange protocol (xmp_b200/csrc/exchange_proto.h: the product's source) under random
// timing, without any search behind it.  N threads play the ranks over host atomics; the "engine" of a rank is a deque of
// synthetic jobs: executing job (id, depth) spawns 0-3 children of depth + 1 with fresh ids, after a random pause.  The
// invariants the reference model-checked for its MPI protocol (spin/fastdnamp.mpiprom):
//   every job is executed exactly once (none lost in a mailbox, none duplicated by a steal);
//   an idle rank never gives work away (its replies carry no jobs);
//   every rank terminates, and only after no rank holds work;
//   a bound lowered anywhere is the bound everywhere at the end.
// Usage: proto_check <ranks> <rounds> [seed].  Prints "OK <jobs executed>" or a diagnostic and exits non-zero.
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <random>
#include <thread>
#include <vector>
#include "../xmp_b200/csrc/exchange_proto.h"

using namespace xmp;

struct Job { unsigned id, depth; };

struct Region {
	std::atomic<unsigned> word[XCH_MAX_RANKS][256];
	Job mail[XCH_MAX_RANKS][XCH_MAX_JOBS];
};

static Region *R;
static std::atomic<unsigned> next_id, executed_total, idle_gave;
static std::vector<std::atomic<unsigned char>> *seen;
static unsigned world;

struct Rank {
	unsigned rank;
	std::deque<Job> open;
	std::vector<Job> staged;
	std::mt19937 rng;
	bool idle_now = false;
};

struct Ops {
	Rank &me;
	int snapshot(unsigned *out) { for (unsigned i = 0; i < XW_COUNT; ++i) out[i] = R->word[me.rank][XW_FIRST + i].load(); return 0; }
	int store(unsigned r, unsigned w, unsigned v) { jitter(); R->word[r][w].store(v); return 0; }
	int add(unsigned r, unsigned w, unsigned d, unsigned *old) { jitter(); const unsigned o = R->word[r][w].fetch_add(d); if (old) *old = o; return 0; }
	int gather(unsigned w, unsigned *out) { for (unsigned r = 0; r < world; ++r) out[r] = R->word[r][w].load(); return 0; }
	void pause() { std::this_thread::sleep_for(std::chrono::microseconds(me.rng() % 30)); }
	void jitter() { if (me.rng() % 4 == 0) std::this_thread::sleep_for(std::chrono::microseconds(me.rng() % 20)); }
};

struct Eng {
	Rank &me;
	int export_for_steal(unsigned k, unsigned *n) {
		if (me.idle_now) idle_gave++;                          // must never happen: the protocol asks idle ranks for nothing
		const size_t open = me.open.size();
		unsigned take = 0;
		if (open >= 2) {
			take = (unsigned) ((open * k + k) / (k + 1));
			if (take >= open) take = (unsigned) open - 1;
			if (take > XCH_MAX_JOBS) take = XCH_MAX_JOBS;
		}
		me.staged.clear();
		for (unsigned i = 0; i < take; ++i) { me.staged.push_back(me.open.front()); me.open.pop_front(); }
		*n = take;
		return 0;
	}
	int send_staged(unsigned thief, unsigned i, unsigned k, unsigned n) {
		unsigned out = 0;
		for (unsigned j = i; j < n; j += k) R->mail[thief][out++] = me.staged[j];
		return 0;
	}
	int import_mailbox(unsigned n) {
		for (unsigned i = 0; i < n; ++i) me.open.push_back(R->mail[me.rank][i]);
		return 0;
	}
};

static void atomic_min(std::atomic<unsigned> &w, unsigned v) {
	unsigned cur = w.load();
	while (v < cur && !w.compare_exchange_weak(cur, v)) {}
}

static void run_rank(Rank *me, unsigned max_depth) {
	Ops ops{*me};
	Eng eng{*me};
	for (;;) {
		while (!me->open.empty()) {                             // "search_run": one job per "sweep", requests answered in between
			const Job j = me->open.back();
			me->open.pop_back();
			if ((*seen)[j.id].exchange(1)) { fprintf(stderr, "job %u executed twice\n", j.id); exit(1); }
			executed_total++;
			if (j.depth < max_depth) {
				const unsigned kids = me->rng() % 4;
				for (unsigned c = 0; c < kids; ++c) {
					const unsigned id = next_id++;
					if (id < seen->size()) me->open.push_back(Job{id, j.depth + 1});
				}
			}
			if (me->rng() % 64 == 0) {                           // a better "tree": lower every rank's bound word
				const unsigned v = 1000000u - j.id;
				for (unsigned r = 0; r < world; ++r) atomic_min(R->word[r][XW_BOUND], v);
			}
			if (me->rng() % 3 == 0) std::this_thread::sleep_for(std::chrono::microseconds(me->rng() % 40));
			R->word[me->rank][XW_OPEN].store(me->open.size() < 2 ? (unsigned) me->open.size() : 2u + (unsigned) me->open.size());
			unsigned w[XW_COUNT];
			ops.snapshot(w);
			if (xch_service(ops, eng, me->rank, world, w, false)) exit(2);
			R->word[me->rank][XW_OPEN].store(me->open.size() < 2 ? (unsigned) me->open.size() : 2u + (unsigned) me->open.size());
		}
		R->word[me->rank][XW_OPEN].store(0);
		me->idle_now = true;
		int result = 0;
		if (xch_steal(ops, eng, me->rank, world, &result)) exit(2);
		me->idle_now = false;
		if (result != XCH_GOT_WORK) break;
		if (me->open.empty()) { fprintf(stderr, "rank %u: GOT_WORK without work\n", me->rank); exit(1); }
	}
}

int main(int argc, char **argv) {
	world = argc > 1 ? (unsigned) atoi(argv[1]) : 4;
	const unsigned rounds = argc > 2 ? (unsigned) atoi(argv[2]) : 20;
	unsigned seed = argc > 3 ? (unsigned) atoi(argv[3]) : 1;
	if (world < 2 || world > XCH_MAX_RANKS) return 3;
	R = new Region();
	unsigned long total = 0;
	for (unsigned round = 0; round < rounds; ++round, ++seed) {
		const unsigned cap = 20000;
		std::vector<std::atomic<unsigned char>> s(cap);
		for (auto &x : s) x.store(0);
		seen = &s;
		next_id = 0;
		executed_total = 0;
		idle_gave = 0;
		for (unsigned r = 0; r < world; ++r) {
			for (unsigned w = 0; w < 256; ++w) R->word[r][w].store(0);
			R->word[r][XW_BOUND].store(0x7fffffffu);
			R->word[r][XW_BUSY].store(world);
		}
		std::vector<Rank> ranks(world);
		const bool skew = round % 2 == 1;                       // everything starts on rank 0: the others live on stolen work
		for (unsigned r = 0; r < world; ++r) {
			ranks[r].rank = r;
			ranks[r].rng.seed(seed * 131u + r);
			const unsigned roots = skew ? (r == 0 ? 6u : 0u) : 2u;
			for (unsigned k = 0; k < roots; ++k) ranks[r].open.push_back(Job{next_id++, 0});
		}
		std::vector<std::thread> th;
		for (unsigned r = 0; r < world; ++r) th.emplace_back(run_rank, &ranks[r], 9u + round % 4);
		for (auto &t : th) t.join();
		const unsigned made = std::min<unsigned>(next_id.load(), cap);
		unsigned done = 0;
		for (unsigned i = 0; i < made; ++i) done += s[i].load();
		if (done != made || executed_total.load() != made) { fprintf(stderr, "round %u: %u jobs made, %u executed (%u counted)\n", round, made, done, executed_total.load()); return 1; }
		if (idle_gave.load()) { fprintf(stderr, "round %u: an idle rank was asked to give work away\n", round); return 1; }
		if (R->word[0][XW_BUSY].load() != 0) { fprintf(stderr, "round %u: busy counter ends at %u\n", round, R->word[0][XW_BUSY].load()); return 1; }
		for (unsigned r = 1; r < world; ++r)
			if (R->word[r][XW_BOUND].load() != R->word[0][XW_BOUND].load()) { fprintf(stderr, "round %u: bounds differ at the end\n", round); return 1; }
		for (unsigned r = 0; r < world; ++r)
			if (!ranks[r].open.empty() || !R->word[r][XW_FINISHED].load()) { fprintf(stderr, "round %u: rank %u stopped with work or without the flag\n", round, r); return 1; }
		total += made;
	}
	printf("OK %lu\n", total);
	return 0;
}
