/* TEST INFRASTRUCTURE ONLY -- never linked into the product.
 *
 * A minimal MPI-1 for ONE node, enough to compile and run the reference's own parallel program (TARGETMULTI:
 * src/mpiboss.c, src/mpiworker.c, the hooks in src/yanbader.c:78-94, src/common.c:397-413, src/measure.c:51-118)
 * in a container that has no MPI installation: the reference's boss / worker build is then the multi-core CPU
 * baseline that north_star asks for ("its MPI build where it compiles, timed on the box's own host cores").
 * Ranks are processes started by oracle/_ref/xmpirun; messages travel through single-producer / single-consumer
 * rings in shared memory (xmpi.c).  Only what the reference calls is provided (SURVEY.md 2.3 lists every call
 * site): point-to-point sends and receives on MPI_COMM_WORLD with MPI_UNSIGNED / MPI_CHAR payloads, the Wait /
 * Test family, MPI_Get_count.  Sends are eager (buffered on arrival, so a send request is complete when the call
 * returns) and non-overtaking per (source, destination) pair, which is all the reference's protocol relies on. */
#ifndef XMPI_MPI_H
#define XMPI_MPI_H

#ifdef __cplusplus
extern "C" {
#endif

typedef int MPI_Comm;
typedef int MPI_Datatype;              /* the value is the size of one element in bytes */
typedef int MPI_Request;               /* 0 = MPI_REQUEST_NULL, so that zero-filled request arrays hold null handles */
typedef struct {
	int MPI_SOURCE, MPI_TAG, MPI_ERROR;
	int xmpi_bytes;                    /* length of the message received (MPI_Get_count) */
} MPI_Status;

#define MPI_COMM_WORLD   0
#define MPI_CHAR         1
#define MPI_UNSIGNED     4
#define MPI_ANY_SOURCE   (-1)
#define MPI_ANY_TAG      (-1)
#define MPI_REQUEST_NULL 0
#define MPI_UNDEFINED    (-32766)
#define MPI_SUCCESS      0

int MPI_Init(int *argc, char ***argv);
int MPI_Finalize(void);
int MPI_Comm_rank(MPI_Comm comm, int *rank);
int MPI_Comm_size(MPI_Comm comm, int *size);
int MPI_Send(const void *buf, int count, MPI_Datatype dt, int dest, int tag, MPI_Comm comm);
int MPI_Ssend(const void *buf, int count, MPI_Datatype dt, int dest, int tag, MPI_Comm comm);
int MPI_Recv(void *buf, int count, MPI_Datatype dt, int source, int tag, MPI_Comm comm, MPI_Status *status);
int MPI_Isend(const void *buf, int count, MPI_Datatype dt, int dest, int tag, MPI_Comm comm, MPI_Request *request);
int MPI_Issend(const void *buf, int count, MPI_Datatype dt, int dest, int tag, MPI_Comm comm, MPI_Request *request);
int MPI_Irecv(void *buf, int count, MPI_Datatype dt, int source, int tag, MPI_Comm comm, MPI_Request *request);
int MPI_Wait(MPI_Request *request, MPI_Status *status);
int MPI_Waitany(int count, MPI_Request *requests, int *index, MPI_Status *status);
int MPI_Waitall(int count, MPI_Request *requests, MPI_Status *statuses);
int MPI_Waitsome(int incount, MPI_Request *requests, int *outcount, int *indices, MPI_Status *statuses);
int MPI_Testany(int count, MPI_Request *requests, int *index, int *flag, MPI_Status *status);
int MPI_Get_count(const MPI_Status *status, MPI_Datatype dt, int *count);

#ifdef __cplusplus
}
#endif
#endif
