/*
 * Minimal MPI interface used ONLY to compile the unmodified reference sources
 * into oracle/_ref (test infrastructure; never linked into the product).
 *
 * Datatype handles encode the element size in bytes (low 8 bits) so the
 * implementation needs no lookup tables.  Two back ends implement this header:
 *   mpi_single.c : one rank, collectives are memcpy
 *   mpi_procs.c  : N forked ranks over socketpairs (launched by tacs_mpiexec)
 */
#ifndef TACSB200_ORACLE_MPI_H
#define TACSB200_ORACLE_MPI_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Request;
typedef int MPI_Info;
typedef long long MPI_Offset;
typedef struct tacsb200_mpi_file *MPI_File;

typedef struct {
  int MPI_SOURCE;
  int MPI_TAG;
  int MPI_ERROR;
  int nbytes_;
} MPI_Status;

typedef void(MPI_User_function)(void *, void *, int *, MPI_Datatype *);

#define MPI_COMM_NULL 0
#define MPI_COMM_WORLD 1
#define MPI_COMM_SELF 2

/* handle = (kind << 8) | sizeof(element) */
#define MPI_CHAR ((1 << 8) | 1)
#define MPI_INT ((2 << 8) | 4)
#define MPI_FLOAT ((3 << 8) | 4)
#define MPI_DOUBLE ((4 << 8) | 8)
#define MPI_DOUBLE_COMPLEX ((5 << 8) | 16)

#define MPI_SUM 1
#define MPI_MAX 2
#define MPI_MIN 3

#define MPI_IN_PLACE ((void *)-1)
#define MPI_STATUS_IGNORE ((MPI_Status *)0)
#define MPI_STATUSES_IGNORE ((MPI_Status *)0)
#define MPI_INFO_NULL 0
#define MPI_SUCCESS 0
#define MPI_ANY_TAG (-1)
#define MPI_ANY_SOURCE (-1)
#define MPI_UNDEFINED (-32766)
#define MPI_MAX_ERROR_STRING 256
#define MPI_IDENT 0
#define MPI_CONGRUENT 1
#define MPI_UNEQUAL 3
#define MPI_MODE_RDONLY 2
#define MPI_MODE_WRONLY 4
#define MPI_MODE_CREATE 1

int MPI_Init(int *argc, char ***argv);
int MPI_Finalize(void);
int MPI_Initialized(int *flag);
int MPI_Abort(MPI_Comm comm, int code);
int MPI_Comm_rank(MPI_Comm comm, int *rank);
int MPI_Comm_size(MPI_Comm comm, int *size);
int MPI_Comm_compare(MPI_Comm a, MPI_Comm b, int *result);
double MPI_Wtime(void);
int MPI_Barrier(MPI_Comm comm);

int MPI_Send(const void *buf, int n, MPI_Datatype t, int dest, int tag, MPI_Comm comm);
int MPI_Recv(void *buf, int n, MPI_Datatype t, int src, int tag, MPI_Comm comm, MPI_Status *st);
int MPI_Isend(const void *buf, int n, MPI_Datatype t, int dest, int tag, MPI_Comm comm, MPI_Request *req);
int MPI_Irecv(void *buf, int n, MPI_Datatype t, int src, int tag, MPI_Comm comm, MPI_Request *req);
int MPI_Wait(MPI_Request *req, MPI_Status *st);
int MPI_Waitall(int n, MPI_Request *reqs, MPI_Status *sts);
int MPI_Waitany(int n, MPI_Request *reqs, int *index, MPI_Status *st);
int MPI_Probe(int src, int tag, MPI_Comm comm, MPI_Status *st);
int MPI_Get_count(const MPI_Status *st, MPI_Datatype t, int *count);

int MPI_Bcast(void *buf, int n, MPI_Datatype t, int root, MPI_Comm comm);
int MPI_Reduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op op, int root, MPI_Comm comm);
int MPI_Allreduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op op, MPI_Comm comm);
int MPI_Gather(const void *s, int sn, MPI_Datatype st, void *r, int rn, MPI_Datatype rt, int root, MPI_Comm comm);
int MPI_Gatherv(const void *s, int sn, MPI_Datatype st, void *r, const int *rc, const int *displs, MPI_Datatype rt, int root, MPI_Comm comm);
int MPI_Allgather(const void *s, int sn, MPI_Datatype st, void *r, int rn, MPI_Datatype rt, MPI_Comm comm);
int MPI_Scatter(const void *s, int sn, MPI_Datatype st, void *r, int rn, MPI_Datatype rt, int root, MPI_Comm comm);
int MPI_Scatterv(const void *s, const int *sc, const int *displs, MPI_Datatype st, void *r, int rn, MPI_Datatype rt, int root, MPI_Comm comm);
int MPI_Alltoall(const void *s, int sn, MPI_Datatype st, void *r, int rn, MPI_Datatype rt, MPI_Comm comm);
int MPI_Alltoallv(const void *s, const int *sc, const int *sd, MPI_Datatype st, void *r, const int *rc, const int *rd, MPI_Datatype rt, MPI_Comm comm);

int MPI_Error_string(int code, char *str, int *len);
int MPI_Op_create(MPI_User_function *fn, int commute, MPI_Op *op);
int MPI_Op_free(MPI_Op *op);

int MPI_File_open(MPI_Comm comm, const char *name, int mode, MPI_Info info, MPI_File *fp);
int MPI_File_close(MPI_File *fp);
int MPI_File_set_view(MPI_File fp, MPI_Offset disp, MPI_Datatype et, MPI_Datatype ft, const char *rep, MPI_Info info);
int MPI_File_set_size(MPI_File fp, MPI_Offset size);
int MPI_File_write(MPI_File fp, const void *buf, int n, MPI_Datatype t, MPI_Status *st);
int MPI_File_read(MPI_File fp, void *buf, int n, MPI_Datatype t, MPI_Status *st);
int MPI_File_write_at_all(MPI_File fp, MPI_Offset off, const void *buf, int n, MPI_Datatype t, MPI_Status *st);
int MPI_File_read_at_all(MPI_File fp, MPI_Offset off, void *buf, int n, MPI_Datatype t, MPI_Status *st);

#ifdef __cplusplus
}
#endif
#endif
