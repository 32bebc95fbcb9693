/* fastdnamp_b200 -- the fastdnamp command line over the B200 search library.
 *
 * Same command line, same PHYLIP input handling, same result files as the reference's fastdnamp
 * (reference src/fastdnamp.c:22-161, option parsing src/common.c:73-250, settings report :252-369,
 * timing report :456-481).  What differs is only what `north_star` says must differ: the branch and
 * bound itself runs on the GPU(s) through include/xmp_cuda.h, and the MPI boss/worker pair
 * (src/mpiboss.c, src/mpiworker.c) is replaced by one host thread per GPU that shares the upper bound
 * and hands subtree jobs to idle GPUs.  Two options are new: --gpus=<n> and --membudget=<MiB>.
 *
 * Not built in (the reference computes them on the CPU before the search; see DESIGN.md "Out of
 * scope"): the -Bi/-Bk/-BK/-Bh lower bounds (load their restBound.txt with -Bf <file>).  --tbr is
 * accepted: in the reference it lowers the bound to the greedy stepwise-addition tree's length and
 * discards the TBR result (src/common.c:2608-2632, src/greedytree.c:92-100); here a greedy beam dive
 * on the device seeds the bound, whatever --tbr says.  --displaynewtrees is accepted too, but trees are not echoed
 * one by one as they are found (src/yanbader.c:260-270): they leave the device in batches and appear in the result.
 */
#include <limits.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>
#include "xmp_cuda.h"
#include "xmp_host.h"

#define PROGNAME "fastDNAmp"
#define PROGLONGNAME "fastDNAmp: Exact Maximum Parsimony Phylogenetic Tree Search v1.1"

typedef struct {
	xmph_settings s;
	unsigned update_secs, seed, n_gpus;
	size_t mem_budget;
	int rand1, rand2, rand3;
	const char *in_name, *out_name;
	FILE *in, *out;
} cli;

static double now_seconds(void) {
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return (double) ts.tv_sec + 1e-9 * (double) ts.tv_nsec;
}

static void die(const char *msg) {
	fprintf(stderr, "%s\n", msg);
	exit(1);
}

static void usage(void) {
	fprintf(stderr,
	        "Syntax: " PROGNAME " [<options>] <phylipfile>\n"
	        "  -o <fname>   File to write results to, instead of standard output\n"
	        "  -b <nn>      Specify an initial upper bound (or 0 for no bound)\n"
	        "  -m <nn>      Maximum number of trees to save (or 0 for no limit)\n"
	        "  -r           Output trees in raw (Newick) format (default: NEXUS format)\n"
	        "  -d           Discard sites containing gaps/ambiguous states\n"
	        "                 (default: like GAPMODE=MISSING in PAUP*)\n"
	        "  -q           Quiet mode: only the final set of trees is output\n"
	        "  -B<letters>  Lower bounds: d = distinct bases (default), m = MST, f <fname> = load restBound[]\n"
	        "  --startcol=<nn>           Starting site number (first site is site 1)\n"
	        "  --endcol=<nn>             Ending site number\n"
	        "  --displaynewbounds={Y|N}  Control printing of updated bounds while running\n"
	        "  --displaynewtrees={Y|N}   Control printing of best trees while running\n"
	        "  --updatesecs=<nn>         Update progress every <nn> seconds (use 0 to\n"
	        "                              turn off periodic progress updates)\n"
	        "  --seed=<nn>               Use <nn> as the random number generation seed --\n"
	        "                              0 (the default) means 'pick a random seed'\n"
	        "  --gpus=<nn>               Number of GPUs to search on (default 1)\n"
	        "  --membudget=<MiB>         Device memory the search frontier may use per GPU\n");
}

/* "--name", "--name=Y|N": returns 1 if arg names this option. */
static int bool_option(const char *arg, const char *name, unsigned *flags, unsigned mask) {
	size_t n = strlen(name);
	if (strncmp(arg, name, n) != 0 || (arg[n] != 0 && arg[n] != '=')) return 0;
	const char *v = arg[n] ? arg + n + 1 : "";
	if (*v == 0 || *v == 'Y' || *v == 'y' || *v == '1' || *v == 'T' || *v == 't') *flags |= mask; else *flags &= ~mask;
	return 1;
}

static int unsigned_option(const char *arg, const char *name, unsigned *out) {
	size_t n = strlen(name);
	if (strncmp(arg, name, n) != 0 || arg[n] != '=' || arg[n + 1] == 0) return 0;
	*out = (unsigned) strtoul(arg + n + 1, NULL, 10);
	return 1;
}

static int parse_args(int argc, char **argv, cli *c) {
	unsigned mb = 0;
	if (argc < 2) { usage(); return 0; }
	for (int i = 1; i < argc; ++i) {
		const char *a = argv[i];
		if (a[0] != '-') {
			c->in_name = a;
			if (!(c->in = fopen(a, "r"))) {
				fprintf(stderr, "File '%s' could not be opened for input, aborting.\n", a);
				exit(1);
			}
			continue;
		}
		switch (a[1]) {
		case 'b':
			if (++i >= argc) die("Option -b needs a value, aborting.");
			c->s.bound = (unsigned) atoi(argv[i]);
			if (c->s.bound == 0) c->s.bound = XMPH_NO_LIMIT;
			break;
		case 'o':
			if (++i >= argc) die("Option -o needs a value, aborting.");
			if (!(c->out = fopen(argv[i], "w"))) {
				fprintf(stderr, "File '%s' could not be opened for output, aborting.\n", argv[i]);
				exit(1);
			}
			c->out_name = argv[i];
			break;
		case 'm':
			if (++i >= argc) die("Option -m needs a value, aborting.");
			c->s.max_trees = (unsigned) atoi(argv[i]);
			if (c->s.max_trees == 0) c->s.max_trees = XMPH_NO_LIMIT;
			break;
		case 'r': c->s.options &= ~XMPH_OPT_NEXUSFMT; break;
		case 'd': c->s.options |= XMPH_OPT_DISCARDMISSAMBIG; break;
		case 'q':
			c->s.options |= XMPH_OPT_QUIET;
			c->s.options &= ~(XMPH_OPT_DISPLAYNEWBOUNDS | XMPH_OPT_DISPLAYNEWTREES);
			c->update_secs = 0;
			break;
		case 'B':
			c->s.options &= ~(XMPH_OPT_USELOADRESTBOUND | XMPH_OPT_USEDISTBASESBOUNDREST | XMPH_OPT_USEINCOMPATBOUNDREST |
			                  XMPH_OPT_USEMSTBOUNDREST | XMPH_OPT_USEKICKASSBOUNDREST | XMPH_OPT_USEKA2BOUNDREST |
			                  XMPH_OPT_USEGREEDYHITTINGSETBOUND | XMPH_OPT_USEPARTBOUND);
			for (const char *p = a + 2; *p; ++p) {
				switch (*p) {
				case 'd': c->s.options |= XMPH_OPT_USEDISTBASESBOUNDREST; break;
				case 'i': c->s.options |= XMPH_OPT_USEINCOMPATBOUNDREST; break;
				case 'm': c->s.options |= XMPH_OPT_USEMSTBOUNDREST; break;
				case 'k': c->s.options |= XMPH_OPT_USEKICKASSBOUNDREST; break;
				case 'K': c->s.options |= XMPH_OPT_USEKA2BOUNDREST; break;
				case 'h': c->s.options |= XMPH_OPT_USEGREEDYHITTINGSETBOUND; break;
				case 'p': c->s.options |= XMPH_OPT_USEPARTBOUND; break;
				case 'C': case 'L':
					if (!(c->s.options & XMPH_OPT_USEPARTBOUND)) {
						fprintf(stderr, "To specify the '%c' option for %s scores only, you must have previously specified -Bp.\n", *p,
						        *p == 'C' ? "clipped" : "lower bound");
						exit(1);
					}
					c->s.part_strategy = *p == 'C' ? 1 : 2;
					break;
				case 'f':
					if (i + 1 >= argc) die("Option -Bf needs a file name, aborting.");
					c->s.options |= XMPH_OPT_USELOADRESTBOUND;
					c->s.rest_bound_file = argv[i + 1];
					break;
				default:
					fprintf(stderr, "Unknown bound type '%c'.\n", *p);
					exit(1);
				}
			}
			if (c->s.options & XMPH_OPT_USELOADRESTBOUND) ++i;
			break;
		case '-': {
			const char *x = a + 2;
			unsigned v;
			if (bool_option(x, "displaynewbounds", &c->s.options, XMPH_OPT_DISPLAYNEWBOUNDS)) break;
			if (bool_option(x, "displaynewtrees", &c->s.options, XMPH_OPT_DISPLAYNEWTREES)) break;
			if (unsigned_option(x, "updatesecs", &c->update_secs)) break;
			if (unsigned_option(x, "startcol", &v)) { c->s.start_col = v - 1; break; }
			if (unsigned_option(x, "endcol", &v)) { c->s.end_col = v - 1; break; }
			if (bool_option(x, "improveinitupperbound", &c->s.options, XMPH_OPT_IMPROVEINITUPPERBOUND)) break;
			if (unsigned_option(x, "seed", &c->seed)) break;
			if (bool_option(x, "onlycomputelowerbound", &c->s.options, XMPH_OPT_ONLYCOMPUTELOWERBOUND)) break;
			if (bool_option(x, "keeptempfiles", &c->s.options, XMPH_OPT_KEEPTEMPFILES)) break;
			if (unsigned_option(x, "pbmaxiters", &c->s.part_max_stale)) break;
			if (bool_option(x, "savereordereddataset", &c->s.options, XMPH_OPT_SAVEREORDEREDDATASET)) break;
			if (bool_option(x, "tbr", &c->s.options, XMPH_OPT_TBR)) break;
			if (unsigned_option(x, "gpus", &c->n_gpus)) break;
			if (unsigned_option(x, "membudget", &mb)) { c->mem_budget = (size_t) mb << 20; break; }
			fprintf(stderr, "Unrecognised option '%s', aborting.\n", a);
			exit(1);
		}
		case '?': case 'h': usage(); return 0;
		case 0: c->in = stdin; break;
		default:
			fprintf(stderr, "Unrecognised option '-%c', aborting.\n", a[1]);
			exit(1);
		}
	}
	if (!c->seed) c->seed = (unsigned) time(NULL) * (unsigned) (getpid() + 1);
	srand(c->seed);
	c->rand1 = rand();
	c->rand2 = rand();
	c->rand3 = rand();
	if (!c->in) die("Must specify PHYLIP-format file to open on the command line (or '-' to read data from standard input).");
	return 1;
}

static void report_settings(const cli *c) {
	const unsigned o = c->s.options;
	if (o & XMPH_OPT_QUIET) return;
	fprintf(stderr, "%s\n\n", PROGLONGNAME);
	fprintf(stderr, "B200 (CUDA) version searching on %u GPU%s.\n", c->n_gpus, c->n_gpus == 1 ? "" : "s");
	if (c->in == stdin) fprintf(stderr, "Reading data from standard input.\n");
	else fprintf(stderr, "Reading data from file '%s'.\n", c->in_name);
	if (c->out == stdout) fprintf(stderr, "Writing result trees to standard output.\n");
	else fprintf(stderr, "Writing result trees to file '%s'.\n", c->out_name);
	if (c->s.bound == XMPH_NO_LIMIT) fprintf(stderr, "No initial bound specified.\n");
	else fprintf(stderr, "Initial bound = %u.\n", c->s.bound);
	if (c->s.max_trees == XMPH_NO_LIMIT) fprintf(stderr, "All optimal trees will be saved.\n");
	else fprintf(stderr, "Up to %u optimal trees will be saved.\n", c->s.max_trees);
	if (c->s.start_col == 0 && c->s.end_col == XMPH_NO_LIMIT) fprintf(stderr, "All sites will be processed.\n");
	else fprintf(stderr, "Sites %u-%u will be processed.\n", c->s.start_col + 1, c->s.end_col + 1);
	fprintf(stderr, "Sites containing gaps or ambiguous characters will be %s.\n", (o & XMPH_OPT_DISCARDMISSAMBIG) ? "discarded" : "retained");
	fprintf(stderr, "Trees will be output in %s format.\n", (o & XMPH_OPT_NEXUSFMT) ? "NEXUS" : "raw (Newick)");
	if (!(o & XMPH_OPT_DISPLAYNEWBOUNDS)) fprintf(stderr, "New bounds will not be reported as they occur.\n");
	if (!(o & XMPH_OPT_DISPLAYNEWTREES)) fprintf(stderr, "New trees will not be reported as they occur.\n");
	if (o & XMPH_OPT_USELOADRESTBOUND) fprintf(stderr, "The restBound[] array will be loaded from the file '%s'.\n", c->s.rest_bound_file);
	if (o & XMPH_OPT_USEDISTBASESBOUNDREST) fprintf(stderr, "The distinct-bases-per-site bound will be used.\n");
	if (o & XMPH_OPT_USEMSTBOUNDREST) fprintf(stderr, "The minimum spanning tree bound will be used.\n");
	if (o & XMPH_OPT_USEPARTBOUND) {
		fprintf(stderr, "The PARTBOUND bound will be used.  %s will be maximised.\n",
		        c->s.part_strategy == 0 ? "Final scores" : c->s.part_strategy == 1 ? "Clipped scores" : "Lower bounds");
		if (c->s.part_max_stale) fprintf(stderr, "PARTBOUND iterations will stop after %u iterations with no improvement.\n", c->s.part_max_stale);
	}
	if (o & XMPH_OPT_ONLYCOMPUTELOWERBOUND) fprintf(stderr, "The program will exit after computing lower bounds.\n");
	if (o & XMPH_OPT_SAVEREORDEREDDATASET) fprintf(stderr, "The dataset produced by reordering taxa and stripping constant and non-PI sites will be saved.\n");
	if (o & XMPH_OPT_KEEPTEMPFILES) fprintf(stderr, "Temporary intermediate files will be kept.\n");
	fprintf(stderr, "Using random number seed %u.  First 3 random numbers: %d, %d, %d.\n", c->seed, c->rand1, c->rand2, c->rand3);
}

/* reordered.phy: the reduced alignment as a legitimate PHYLIP file, weighted columns replicated
 * (reference src/common.c:1190-1202,1233-1286 with options = 0). */
static void save_reordered(const xmph_problem *p) {
	static const char names[16] = { '0', 'A', 'C', 'M', 'G', 'R', 'S', 'V', 'T', 'W', 'Y', 'H', 'K', 'D', 'B', 'N' };
	FILE *f = fopen("reordered.phy", "w");
	unsigned total = 0;
	if (!f) { perror("Could not open 'reordered.phy' for output"); exit(1); }
	for (unsigned i = 0; i < p->num_sites; ++i) total += p->site_weights[i];
	fprintf(f, "%d %d\n", (int) p->num_taxa, (int) total);
	for (unsigned t = 0; t < p->num_taxa; ++t) {
		fprintf(f, "T%-9d", (int) p->taxon_map[t] + 1);
		for (unsigned i = 0; i < p->num_sites; ++i) {
			for (unsigned k = 0; k < p->site_weights[i]; ++k) fputc(names[p->sites[(size_t) t * p->num_sites + i] & 15], f);
		}
		fputc('\n', f);
	}
	fclose(f);
}

/* ---- the search across GPUs -------------------------------------------------------------------- */
/* Replaces Boss()/Worker() (reference src/mpiboss.c:30-429, src/mpiworker.c:214-380).  There is no boss: every GPU
 * searches, one host thread each.  Shards start from a hashed deal of the frontier at a shallow tree size (the
 * library's shard_index/shard_count); everything the GPUs tell each other afterwards goes through their exchange
 * blocks in peer-mapped device memory (include/xmp_cuda.h 2d):
 *   - the upper bound: the kernel that finds a better tree lowers every GPU's bound word itself (the reference relays
 *     MSG_NEWBOUND through the boss, src/mpiboss.c:162-171);
 *   - a GPU that runs dry asks one peer, which detaches open subtrees from its shallowest level (the reference's
 *     "steal the last edge at the shallowest level", src/mpiworker.c:169-181) into the thief's mailbox;
 *   - the search ends when no GPU holds work (xmp_cuda_search_steal reports it).
 * What the threads share on the host is only what the user sees: the best length for --displaynewbounds and the
 * progress report. */
typedef struct {
	pthread_mutex_t mu;
	pthread_barrier_t start;   /* every exchange block is reset before any search begins */
	unsigned bound;
	unsigned n_workers;
	int display_bounds;
	/* progress reports (--updatesecs, reference src/yanbader.c:28-44, src/common.c:525-541) */
	unsigned update_secs;
	double t_start, t_last_update;
	double open_share[64];     /* per worker: the share of the search space its open subtrees still stand for */
} shared_state;

typedef struct {
	shared_state *sh;
	const xmph_problem *p;
	const cli *c;
	unsigned index;
	xmp_ctx *ctx;
	xmp_ctx **all;             /* every worker's context (for the exchange) */
} worker;

#define SLICE_BATCHES 8

/* A failing GPU cannot be waited for by the others (they would wait for its share of the work): fatal, like every
 * error of the reference. */
static void worker_fail(worker *w, const char *what) {
	fprintf(stderr, "GPU %u: %s: %s\n", w->index, what, xmp_cuda_last_error());
	exit(1);
}

/* The reference reports the share of the search space its depth-first walk has left behind.  A frontier search knows
 * the same quantity from what is still open: a partial tree on m taxa stands for 1 / (3 * 5 * ... * (2m-5)) of all trees. */
static void report_progress(worker *w, unsigned bound) {
	shared_state *sh = w->sh;
	unsigned long long pending[XMP_MAX_TAXA + 1];
	if (!sh->update_secs || w->index >= 64 || xmp_cuda_search_pending(w->ctx, pending)) return;
	double share = 0, unit = 1;
	for (unsigned m = 3; m <= XMP_MAX_TAXA; ++m) {
		share += (double) pending[m] * unit;
		unit /= (double) (2 * m - 3);
	}
	pthread_mutex_lock(&sh->mu);
	sh->open_share[w->index] = share;
	const double now = now_seconds();
	if (now - sh->t_last_update >= (double) sh->update_secs) {
		double open = 0;
		for (unsigned g = 0; g < sh->n_workers && g < 64; ++g) open += sh->open_share[g];
		if (open > 1) open = 1;           /* shards expand the shallow levels redundantly before the frontier is dealt out */
		const double done = 1 - open, elapsed = now - sh->t_start;
		fprintf(stderr, "%.2f%% complete.  Best length so far: %u.  Estimated time remaining: ", 100 * done, bound < sh->bound ? bound : sh->bound);
		if (elapsed < 10 || done <= 0) fprintf(stderr, "(too early to tell)\n");
		else fprintf(stderr, "%.0f seconds\n", (1 - done) * elapsed / done);
		sh->t_last_update = now;
	}
	pthread_mutex_unlock(&sh->mu);
}

static void *worker_main(void *arg) {
	worker *w = (worker *) arg;
	shared_state *sh = w->sh;
	const xmph_problem *p = w->p;
	xmp_search_params prm;
	int done = 0;

	if (xmp_cuda_load(w->ctx, p->rows, p->weights, p->num_taxa, p->n_blocks, p->rest_bound, p->constant_weight)) worker_fail(w, "uploading the problem");
	if (sh->n_workers > 1) {
		if (xmp_cuda_exchange_connect_local(w->ctx, w->index, sh->n_workers, w->all)) worker_fail(w, "connecting the GPUs");
		if (xmp_cuda_exchange_reset(w->ctx)) worker_fail(w, "resetting the exchange");
		pthread_barrier_wait(&sh->start);
	}
	memset(&prm, 0, sizeof prm);
	prm.bound = p->bound;
	prm.max_trees = w->c->s.max_trees;
	prm.mem_budget = w->c->mem_budget;
	prm.shard_index = w->index;
	prm.shard_count = sh->n_workers;
	if (xmp_cuda_search_begin(w->ctx, &prm)) worker_fail(w, "starting the search");

	for (;;) {
		unsigned mine = 0;
		/* slices only serve the user (--displaynewbounds, progress reports): the GPUs need no host round to learn a
		 * new bound or to hand over work */
		const unsigned slice = (sh->display_bounds || sh->update_secs) ? SLICE_BATCHES : 0;
		if (!done && xmp_cuda_search_run(w->ctx, slice, &done)) worker_fail(w, "searching");
		if (xmp_cuda_search_get_bound(w->ctx, &mine)) worker_fail(w, "reading the bound");
		report_progress(w, mine);
		pthread_mutex_lock(&sh->mu);
		if (mine < sh->bound) {
			sh->bound = mine;
			if (sh->display_bounds) fprintf(stderr, "New bound = %u\n", mine);
		}
		pthread_mutex_unlock(&sh->mu);
		if (!done) continue;
		/* Out of work: ask a peer, or learn that everybody is done. */
		int state = 0;
		if (xmp_cuda_search_steal(w->ctx, &state)) worker_fail(w, "asking for work");
		if (state != XMP_STEAL_GOT_WORK) break;
		done = 0;
	}
	unsigned mine = 0;
	if (xmp_cuda_search_get_bound(w->ctx, &mine)) worker_fail(w, "reading the bound");
	pthread_mutex_lock(&sh->mu);
	if (mine < sh->bound) {
		sh->bound = mine;
		if (sh->display_bounds) fprintf(stderr, "New bound = %u\n", mine);
	}
	pthread_mutex_unlock(&sh->mu);
	return NULL;
}

int main(int argc, char **argv) {
	cli c;
	xmph_alignment a;
	xmph_problem p;
	double t_start = now_seconds(), t_order, t_upper, t_lower, t_search, t_output, t_end;

	memset(&c, 0, sizeof c);
	xmph_default_settings(&c.s);
	c.update_secs = 2;
	c.n_gpus = 1;
	c.out = stdout;
	if (!parse_args(argc, argv, &c)) return 0;
	if (c.n_gpus == 0) c.n_gpus = 1;
	report_settings(&c);

	if (xmph_read_phylip(c.in, &a)) die(xmph_last_error());
	if (!(c.s.options & XMPH_OPT_QUIET)) fprintf(stderr, "Nucleotide data for %u taxa x %u sites read in.\n", a.num_taxa, a.seq_len);
	if (c.in != stdin) fclose(c.in);

	t_order = now_seconds();   /* recoding and compression are counted as input time, as in the reference */
	const int prepare_failed = xmph_prepare(&a, &c.s, &p);
	if (prepare_failed && p.num_taxa == 0) die(xmph_last_error());
	t_upper = t_lower = now_seconds();
	if (!(c.s.options & XMPH_OPT_QUIET)) {
		fprintf(stderr, "%d constant sites.\n", (int) p.n_constant_sites);
		fprintf(stderr, "%d non-parsimoniously-informative sites, having a total weight of %d.\n", (int) p.n_nonpi_sites, (int) p.constant_weight);
		fprintf(stderr, "%d parsimoniously-informative sites.\n", (int) p.num_sites);
	}
	if ((c.s.options & XMPH_OPT_IMPROVEINITUPPERBOUND) && p.mst_upper_bound < c.s.bound && (c.s.options & XMPH_OPT_DISPLAYNEWBOUNDS)) {
		fprintf(stderr, "Improving initial upper bound to %u, based on minimum spanning tree weight.\n", p.mst_upper_bound);
	}
	if (prepare_failed) die(xmph_last_error());      /* a restBound file that does not parse: the reference gets this far too */
	if (c.s.options & XMPH_OPT_USELOADRESTBOUND) {
		/* the reference's LoadRestBound says so on stderr, whatever -q says (src/common.c:1530-1554): header line + data lines + 1 */
		char line[256];
		unsigned n_lines = 2;
		FILE *f = fopen(c.s.rest_bound_file, "rb");
		if (f) {
			if (fgets(line, sizeof line, f)) while (fgets(line, sizeof line, f)) ++n_lines;
			fclose(f);
		}
		fprintf(stderr, "Read %u lines from '%s'.\n", n_lines, c.s.rest_bound_file);
	}
	if (c.s.options & XMPH_OPT_SAVEREORDEREDDATASET) save_reordered(&p);
	if (xmph_write_rest_bound(&p, "restBound.txt")) die(xmph_last_error());

	if (c.s.options & XMPH_OPT_ONLYCOMPUTELOWERBOUND) {
		if (c.out != stdout) fclose(c.out);
		xmph_free_problem(&p);
		xmph_free_alignment(&a);
		return 0;
	}

	int avail = xmp_cuda_device_count();
	if (avail <= 0) die("No CUDA device found.  This program has no CPU search path, aborting.");
	if ((int) c.n_gpus > avail) {
		fprintf(stderr, "%u GPUs requested but only %d available, aborting.\n", c.n_gpus, avail);
		exit(1);
	}
	if (!(c.s.options & XMPH_OPT_QUIET)) fprintf(stderr, "Beginning branch and bound search...\n");

	shared_state sh;
	memset(&sh, 0, sizeof sh);
	pthread_mutex_init(&sh.mu, NULL);
	pthread_barrier_init(&sh.start, NULL, c.n_gpus);
	sh.bound = p.bound;
	sh.n_workers = c.n_gpus;
	sh.display_bounds = (c.s.options & XMPH_OPT_DISPLAYNEWBOUNDS) != 0;
	sh.update_secs = (c.s.options & XMPH_OPT_QUIET) ? 0 : c.update_secs;
	worker *ws = (worker *) calloc(c.n_gpus, sizeof(worker));
	pthread_t *th = (pthread_t *) calloc(c.n_gpus, sizeof(pthread_t));
	xmp_ctx **all = (xmp_ctx **) calloc(c.n_gpus, sizeof(xmp_ctx *));

	/* Device contexts, the search's device memory and the exchange blocks are set up before the clock starts, like
	 * the reference's arena (InitTreeData, src/common.c:1113-1224) and its MPI_Init; the upload of the problem is
	 * inside the timed search. */
	for (unsigned g = 0; g < c.n_gpus; ++g) {
		unsigned char handle[XMP_EXCHANGE_HANDLE_BYTES];
		ws[g].sh = &sh;
		ws[g].p = &p;
		ws[g].c = &c;
		ws[g].index = g;
		ws[g].all = all;
		if (xmp_cuda_init((int) g, &ws[g].ctx)) die(xmp_cuda_last_error());
		all[g] = ws[g].ctx;
		if (xmp_cuda_search_reserve(ws[g].ctx, p.num_taxa, p.n_blocks, c.mem_budget)) die(xmp_cuda_last_error());
		if (c.n_gpus > 1 && xmp_cuda_exchange_handle(ws[g].ctx, handle)) die(xmp_cuda_last_error());
	}
	t_search = now_seconds();
	sh.t_start = sh.t_last_update = t_search;
	for (unsigned g = 0; g < c.n_gpus; ++g) pthread_create(&th[g], NULL, worker_main, &ws[g]);
	for (unsigned g = 0; g < c.n_gpus; ++g) pthread_join(th[g], NULL);

	/* Gather: every shard reports the trees it holds at the final bound. */
	unsigned char *paths = NULL;
	unsigned long long n_trees = 0, n_paths = 0;   /* optimal trees found / of which the paths are held */
	xmp_search_stats total;
	memset(&total, 0, sizeof total);
	for (unsigned g = 0; g < c.n_gpus; ++g) {
		unsigned score = 0;
		unsigned long long n = 0;
		xmp_search_stats st;
		if (xmp_cuda_search_set_bound(ws[g].ctx, sh.bound)) die(xmp_cuda_last_error());
		unsigned long long stored = 0;     /* < n only when -m made the shard cut its set back to its first max_trees */
		if (xmp_cuda_search_result(ws[g].ctx, &score, &n, NULL, 0)) die(xmp_cuda_last_error());
		if (xmp_cuda_search_stored(ws[g].ctx, &stored)) die(xmp_cuda_last_error());
		paths = (unsigned char *) realloc(paths, (size_t) (n_paths + stored + 1) * p.num_taxa);
		if (stored && xmp_cuda_search_result(ws[g].ctx, &score, &n, paths + (size_t) n_paths * p.num_taxa, (size_t) stored)) die(xmp_cuda_last_error());
		n_paths += stored;
		n_trees += n;
		if (xmp_cuda_search_stats(ws[g].ctx, &st)) die(xmp_cuda_last_error());
		total.nodes += st.nodes;
		total.score_calls += st.score_calls;
		total.children += st.children;
		total.full_trees += st.full_trees;
		total.site_merges += st.site_merges;
		total.launches += st.launches;
		total.iterations += st.iterations;
		total.steals_served += st.steals_served;
		total.steals_received += st.steals_received;
		for (unsigned m = 0; m <= XMP_MAX_TAXA; ++m) {
			total.nodes_by_depth[m] += st.nodes_by_depth[m];
			total.edge_evals_by_depth[m] += st.edge_evals_by_depth[m];
			total.batches_by_depth[m] += st.batches_by_depth[m];
		}
		if (st.seconds > total.seconds) total.seconds = st.seconds;
	}
	t_output = now_seconds();

	xmph_sort_paths_dfs(paths, (size_t) n_paths, p.num_taxa);
	size_t n_out = n_paths > c.s.max_trees ? c.s.max_trees : (size_t) n_paths;
	if (xmph_write_trees(c.out, (c.s.options & XMPH_OPT_NEXUSFMT) != 0, sh.bound, &a, p.taxon_map, paths, n_out)) die(xmph_last_error());
	if (c.out != stdout) fclose(c.out);
	if (!(c.s.options & XMPH_OPT_QUIET) && n_trees == 0) {
		fprintf(stderr, "NO BEST TREE FOUND!  (Have you specified an initial bound that is too low?)\n");
	}
	t_end = now_seconds();

	if (!(c.s.options & XMPH_OPT_QUIET)) {
		double dt = t_output - t_search;
		fprintf(stderr, "Optimal trees found:            %20llu\n", n_trees);
		fprintf(stderr, "BranchAndBound() executed       %20llu times\n", total.nodes);
		fprintf(stderr, "ScoreFitch() executed           %20llu times\n", total.score_calls);
		fprintf(stderr, "Partial trees built             %20llu\n", total.children);
		fprintf(stderr, "Number of full trees considered:%20llu\n", total.full_trees);
		fprintf(stderr, "Site merges                     %20llu (%.3f G/s)\n", total.site_merges, dt > 0 ? (double) total.site_merges / dt * 1e-9 : 0.0);
		fprintf(stderr, "Kernel launches / batches       %20llu / %llu\n", total.launches, total.iterations);
		if (c.n_gpus > 1) fprintf(stderr, "Steals served / received        %20llu / %llu\n", total.steals_served, total.steals_received);
		/* the reference's per-depth counters (src/measure.c:10,79: bAndBScanEdgeCountByDepth[]), by taxa on the tree */
		for (unsigned m = 3; m <= XMP_MAX_TAXA; ++m) {
			if (!total.edge_evals_by_depth[m]) continue;
			fprintf(stderr, "ScoreFitch() on trees of %3u taxa:%19llu times (%llu trees in %llu batches)\n", m, total.edge_evals_by_depth[m],
			        total.nodes_by_depth[m], total.batches_by_depth[m]);
		}
		fprintf(stderr, "%-45s %11.2lfs\n", "Time to read input:", t_order - t_start);
		fprintf(stderr, "%-45s %11.2lfs\n", "Time to order taxa:", t_upper - t_order);
		fprintf(stderr, "%-45s %11.2lfs\n", "Time to determine initial upper bound:", t_lower - t_upper);
		fprintf(stderr, "%-45s %11.2lfs\n", "Time to determine lower bounds:", t_search - t_lower);
		fprintf(stderr, "%-45s %11.2lfs\n", "Time to perform branch and bound search:", t_output - t_search);
		fprintf(stderr, "%-45s %11.2lfs\n", "Time to write output:", t_end - t_output);
		fprintf(stderr, "%-45s %11.2lfs\n", "Total time taken:", t_end - t_start);
	}
	for (unsigned g = 0; g < c.n_gpus; ++g) xmp_cuda_destroy(ws[g].ctx);
	free(paths);
	free(all);
	free(ws);
	free(th);
	xmph_free_problem(&p);
	xmph_free_alignment(&a);
	return 0;
}
