/* TEST INFRASTRUCTURE ONLY.  Launcher of the one-node MPI (mpi.h, xmpi.c):
 *
 *     xmpirun -n N program [arguments]
 *
 * creates the shared segment (an anonymous memory file, so nothing is left behind), starts N copies of the program
 * with XMPI_RANK / XMPI_SIZE / XMPI_FD in their environment and waits for them.  If one of them fails the others are
 * killed; ranks also die with the launcher, so that a caller's timeout takes the whole job down. */
#define _GNU_SOURCE
#include <signal.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/prctl.h>
#include <sys/wait.h>
#include <unistd.h>
#include "xmpi_shm.h"

static pid_t g_pid[XMPI_MAX_RANKS];
static int g_n;

static void kill_all(int sig) {
	(void) sig;
	for (int i = 0; i < g_n; ++i)
		if (g_pid[i] > 0) kill(g_pid[i], SIGKILL);
}

int main(int argc, char **argv) {
	if (argc < 4 || strcmp(argv[1], "-n") != 0 || atoi(argv[2]) < 2 || atoi(argv[2]) > XMPI_MAX_RANKS) {
		fprintf(stderr, "usage: xmpirun -n N program [arguments]      (2 <= N <= %d: one boss and N - 1 workers)\n", XMPI_MAX_RANKS);
		return 2;
	}
	const int n = atoi(argv[2]);
	const int fd = memfd_create("xmpi", 0);
	if (fd < 0 || ftruncate(fd, (off_t) xmpi_shm_bytes(n)) != 0) {
		perror("xmpirun: shared segment");
		return 1;
	}
	struct xmpi_shm *shm = (struct xmpi_shm *) mmap(NULL, XMPI_HEADER_BYTES, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
	if (shm == MAP_FAILED) {
		perror("xmpirun: mmap");
		return 1;
	}
	shm->n_ranks = (uint32_t) n;               /* the rest of the segment starts out as zeros: empty rings */
	shm->ring_bytes = (uint32_t) XMPI_RING_BYTES;
	signal(SIGTERM, kill_all);
	signal(SIGINT, kill_all);
	char buf[32];
	snprintf(buf, sizeof buf, "%d", fd);
	setenv("XMPI_FD", buf, 1);
	snprintf(buf, sizeof buf, "%d", n);
	setenv("XMPI_SIZE", buf, 1);
	g_n = n;
	for (int r = 0; r < n; ++r) {
		const pid_t p = fork();
		if (p < 0) {
			perror("xmpirun: fork");
			kill_all(0);
			return 1;
		}
		if (p == 0) {
			prctl(PR_SET_PDEATHSIG, SIGKILL);
			snprintf(buf, sizeof buf, "%d", r);
			setenv("XMPI_RANK", buf, 1);
			execvp(argv[3], argv + 3);
			perror("xmpirun: exec");
			_exit(127);
		}
		g_pid[r] = p;
	}
	int rc = 0, left = n;
	while (left > 0) {
		int st = 0;
		const pid_t p = wait(&st);
		if (p < 0) break;
		for (int r = 0; r < n; ++r)
			if (g_pid[r] == p) g_pid[r] = 0;
		--left;
		if (!WIFEXITED(st) || WEXITSTATUS(st) != 0) {
			if (!rc) rc = WIFEXITED(st) ? WEXITSTATUS(st) : 128 + WTERMSIG(st);
			kill_all(0);
		}
	}
	return rc;
}
