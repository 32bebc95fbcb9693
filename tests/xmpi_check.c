/* Test of the one-node MPI shim (oracle/shim/mpi: test infrastructure for the reference's parallel CPU baseline).
 * Run as `xmpirun -n N xmpi_check`, N >= 3.  Checks what the reference's boss / worker protocol relies on
 * (SURVEY.md 2.3): non-overtaking delivery per (source, destination) pair, matching by tag out of arrival order,
 * wildcards, zero-length messages, MPI_Get_count, eager sends whose requests complete at once, the Wait / Test
 * family on arrays that contain null handles, and messages of the reference's tree-buffer size (2 MiB) through
 * rings that wrap and fill up. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "mpi.h"

#define CHECK(c) do { if (!(c)) { fprintf(stderr, "rank %d: check failed at line %d: %s\n", rank, __LINE__, #c); exit(1); } } while (0)
enum { TAG_A = 1, TAG_B = 2, TAG_BIG = 3, TAG_DONE = 4, TAG_PING = 5, N_SMALL = 2000, BIG = 2 << 20, N_BIG = 5 };

int main(int argc, char **argv) {
	int rank, size;
	MPI_Init(&argc, &argv);
	MPI_Comm_rank(MPI_COMM_WORLD, &rank);
	MPI_Comm_size(MPI_COMM_WORLD, &size);
	CHECK(size >= 3);
	MPI_Status st;
	if (rank != 0) {
		/* 1. a stream of small messages with two tags, through Isend: every request is complete at once */
		for (unsigned i = 0; i < N_SMALL; ++i) {
			unsigned v[2] = {(unsigned) rank, i};
			MPI_Request r = MPI_REQUEST_NULL;
			MPI_Isend(v, 2, MPI_UNSIGNED, 0, (i & 1) ? TAG_B : TAG_A, MPI_COMM_WORLD, &r);
			CHECK(r != MPI_REQUEST_NULL);
			int idx = -1, flag = 0;
			MPI_Testany(1, &r, &idx, &flag, &st);
			CHECK(flag && idx == 0 && r == MPI_REQUEST_NULL);
		}
		/* 2. big messages: more bytes than a ring holds, so the sender has to wait for the receiver */
		unsigned char *big = malloc(BIG);
		for (int k = 0; k < N_BIG; ++k) {
			for (size_t i = 0; i < BIG; ++i) big[i] = (unsigned char) (i * 7 + k + rank);
			MPI_Send(big, BIG - k, MPI_CHAR, 0, TAG_BIG, MPI_COMM_WORLD);
		}
		free(big);
		/* 3. ping-pong with the boss: a zero-length request, a reply received through a wildcard receive */
		unsigned reply[4] = {0, 0, 0, 0};
		MPI_Request rr[3] = {MPI_REQUEST_NULL, MPI_REQUEST_NULL, MPI_REQUEST_NULL};
		MPI_Irecv(reply, 4, MPI_UNSIGNED, 0, MPI_ANY_TAG, MPI_COMM_WORLD, &rr[1]);
		int idx = -1, flag = 1;
		MPI_Testany(3, rr, &idx, &flag, &st);
		CHECK(!flag);                              /* nothing has been sent to us yet */
		MPI_Send(NULL, 0, MPI_UNSIGNED, 0, TAG_PING, MPI_COMM_WORLD);
		MPI_Waitany(3, rr, &idx, &st);
		int count = -1;
		MPI_Get_count(&st, MPI_UNSIGNED, &count);
		CHECK(idx == 1 && st.MPI_SOURCE == 0 && st.MPI_TAG == TAG_PING && count == 3 && reply[2] == 100u + (unsigned) rank);
		MPI_Waitany(3, rr, &idx, &st);             /* only null handles left */
		CHECK(idx == MPI_UNDEFINED);
		/* 4. worker to worker (the reference chains MSG_NTREESLEFT rank -> rank + 1) */
		unsigned token = 0;
		if (rank > 1) MPI_Recv(&token, 1, MPI_UNSIGNED, rank - 1, TAG_DONE, MPI_COMM_WORLD, &st);
		token += (unsigned) rank;
		if (rank + 1 < size) MPI_Send(&token, 1, MPI_UNSIGNED, rank + 1, TAG_DONE, MPI_COMM_WORLD);
		else MPI_Send(&token, 1, MPI_UNSIGNED, 0, TAG_DONE, MPI_COMM_WORLD);
	} else {
		const int W = size - 1;
		/* 1. all of tag B first, then all of tag A: matching by tag skips over queued messages, order kept per tag */
		for (int pass = 0; pass < 2; ++pass)
			for (int w = 1; w <= W; ++w)
				for (unsigned i = pass ? 0 : 1; i < N_SMALL; i += 2) {
					unsigned v[2];
					MPI_Recv(v, 2, MPI_UNSIGNED, w, pass ? TAG_A : TAG_B, MPI_COMM_WORLD, &st);
					CHECK(v[0] == (unsigned) w && v[1] == i && st.MPI_SOURCE == w);
				}
		/* 2. the big ones, one worker after the other (the others' sends wait in their rings meanwhile) */
		unsigned char *big = malloc(BIG);
		for (int w = W; w >= 1; --w)
			for (int k = 0; k < N_BIG; ++k) {
				int count = -1;
				MPI_Recv(big, BIG, MPI_CHAR, w, TAG_BIG, MPI_COMM_WORLD, &st);
				MPI_Get_count(&st, MPI_CHAR, &count);
				CHECK(count == BIG - k);
				for (size_t i = 0; i < (size_t) count; i += 4099) CHECK(big[i] == (unsigned char) (i * 7 + k + w));
			}
		free(big);
		/* 3. one posted receive per worker, Waitsome until everybody has pinged; replies through Isend + Waitall */
		MPI_Request *rq = calloc((size_t) W, sizeof *rq), *sq = calloc((size_t) W, sizeof *sq);   /* zero-filled = null handles */
		MPI_Status *sts = malloc((size_t) W * sizeof *sts);
		int *indices = malloc((size_t) W * sizeof *indices);
		unsigned *bufs = calloc((size_t) W, sizeof *bufs), (*rep)[3] = malloc((size_t) W * sizeof *rep);
		for (int w = 0; w < W; ++w) MPI_Irecv(bufs + w, 1, MPI_UNSIGNED, w + 1, TAG_PING, MPI_COMM_WORLD, rq + w);
		int got = 0;
		while (got < W) {
			int n = 0;
			MPI_Waitsome(W, rq, &n, indices, sts);
			CHECK(n >= 1 && n <= W);
			for (int i = 0; i < n; ++i) {
				int count = -1;
				MPI_Get_count(sts + i, MPI_UNSIGNED, &count);
				CHECK(count == 0 && sts[i].MPI_SOURCE == indices[i] + 1 && rq[indices[i]] == MPI_REQUEST_NULL);
				rep[indices[i]][0] = 1; rep[indices[i]][1] = 2; rep[indices[i]][2] = 100u + (unsigned) indices[i] + 1u;
				MPI_Isend(rep[indices[i]], 3, MPI_UNSIGNED, indices[i] + 1, TAG_PING, MPI_COMM_WORLD, sq + indices[i]);
			}
			got += n;
		}
		int n = 0;
		MPI_Waitsome(W, rq, &n, indices, sts);
		CHECK(n == MPI_UNDEFINED);
		MPI_Waitall(W, sq, sts);
		for (int w = 0; w < W; ++w) CHECK(sq[w] == MPI_REQUEST_NULL);
		/* 4. the token that went round the workers */
		unsigned token = 0;
		MPI_Recv(&token, 1, MPI_UNSIGNED, MPI_ANY_SOURCE, TAG_DONE, MPI_COMM_WORLD, &st);
		CHECK(st.MPI_SOURCE == W && token == (unsigned) (W * (W + 1) / 2));
		printf("xmpi_check ok: %d ranks\n", size);
	}
	MPI_Finalize();
	return 0;
}
