/* Minimal MPI shim (header) so that the reference hot-path sources compile and link
 * without an MPI installation.  TEST INFRASTRUCTURE ONLY (oracle/_ref).
 *
 * Ranks are THREADS of one process (see mpi_shim.c): each rank thread calls
 * pincShimSetRank(rank) once; the "world" size is set with pincShimInit(size).
 * Only the 14 entry points the reference's hot-path objects reference exist.
 */
#ifndef PINC_SHIM_MPI_H
#define PINC_SHIM_MPI_H
#include <stddef.h>

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Info;
typedef struct { int MPI_SOURCE; int MPI_TAG; int MPI_ERROR; int count_; } MPI_Status;
typedef struct pincShimReq_ *MPI_Request;

#define MPI_COMM_WORLD   0
#define MPI_INFO_NULL    0
#define MPI_DOUBLE       1
#define MPI_LONG         2
#define MPI_INT          3
#define MPI_SUM          1
#define MPI_MAX          2
#define MPI_ANY_SOURCE   (-1)
#define MPI_ANY_TAG      (-1)
#define MPI_SUCCESS      0
#define MPI_REQUEST_NULL ((MPI_Request)0)
#define MPI_STATUS_IGNORE   ((MPI_Status*)0)
#define MPI_STATUSES_IGNORE ((MPI_Status*)0)
#define MPI_IN_PLACE     ((void*)-1)

int MPI_Init(int *argc, char ***argv);
int MPI_Finalize(void);
int MPI_Comm_rank(MPI_Comm comm, int *rank);
int MPI_Comm_size(MPI_Comm comm, int *size);
int MPI_Barrier(MPI_Comm comm);
int MPI_Send(const void *buf, int count, MPI_Datatype t, int dest, int tag, MPI_Comm comm);
int MPI_Recv(void *buf, int count, MPI_Datatype t, int src, int tag, MPI_Comm comm, MPI_Status *st);
int MPI_Isend(const void *buf, int count, MPI_Datatype t, int dest, int tag, MPI_Comm comm, MPI_Request *req);
int MPI_Irecv(void *buf, int count, MPI_Datatype t, int src, int tag, MPI_Comm comm, MPI_Request *req);
int MPI_Waitall(int n, MPI_Request *reqs, MPI_Status *sts);
int MPI_Sendrecv(const void *sbuf, int scount, MPI_Datatype st, int dest, int stag,
                 void *rbuf, int rcount, MPI_Datatype rt, int src, int rtag,
                 MPI_Comm comm, MPI_Status *status);
int MPI_Allreduce(const void *sbuf, void *rbuf, int count, MPI_Datatype t, MPI_Op op, MPI_Comm comm);
int MPI_Reduce(const void *sbuf, void *rbuf, int count, MPI_Datatype t, MPI_Op op, int root, MPI_Comm comm);
int MPI_Allgather(const void *sbuf, int scount, MPI_Datatype st, void *rbuf, int rcount, MPI_Datatype rt, MPI_Comm comm);

/* shim control (not MPI) */
void pincShimInit(int worldSize);
void pincShimSetRank(int rank);
#endif
