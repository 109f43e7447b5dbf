/*
 * One-rank implementation of oracle/mpistub/mpi.h (test infrastructure).
 * Collectives degenerate to copies; point-to-point to another rank cannot
 * happen and aborts loudly so a mis-use is never silent.
 */
#include "mpi.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

struct tacsb200_mpi_file {
  FILE *fp;
  MPI_Offset disp;
};

static size_t esize(MPI_Datatype t) { return (size_t)(t & 0xff); }

static void fatal(const char *what) {
  fprintf(stderr, "[oracle mpi_single] %s is not available with one rank\n", what);
  abort();
}

static void self_copy(const void *s, void *r, size_t nbytes) {
  if (s != MPI_IN_PLACE && s != r && nbytes) memmove(r, s, nbytes);
}

int MPI_Init(int *argc, char ***argv) { (void)argc; (void)argv; return MPI_SUCCESS; }
int MPI_Finalize(void) { return MPI_SUCCESS; }
int MPI_Initialized(int *flag) { *flag = 1; return MPI_SUCCESS; }
int MPI_Abort(MPI_Comm comm, int code) {
  (void)comm;
  fprintf(stderr, "[oracle mpi_single] MPI_Abort(%d)\n", code);
  abort();
  return code;
}
int MPI_Comm_rank(MPI_Comm comm, int *rank) { (void)comm; *rank = 0; return MPI_SUCCESS; }
int MPI_Comm_size(MPI_Comm comm, int *size) { (void)comm; *size = 1; return MPI_SUCCESS; }
int MPI_Comm_compare(MPI_Comm a, MPI_Comm b, int *result) {
  *result = (a == b) ? MPI_IDENT : MPI_CONGRUENT;
  return MPI_SUCCESS;
}
double MPI_Wtime(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}
int MPI_Barrier(MPI_Comm comm) { (void)comm; return MPI_SUCCESS; }

int MPI_Send(const void *b, int n, MPI_Datatype t, int d, int tag, MPI_Comm c) {
  (void)b; (void)n; (void)t; (void)d; (void)tag; (void)c; fatal("MPI_Send"); return 1;
}
int MPI_Recv(void *b, int n, MPI_Datatype t, int s, int tag, MPI_Comm c, MPI_Status *st) {
  (void)b; (void)n; (void)t; (void)s; (void)tag; (void)c; (void)st; fatal("MPI_Recv"); return 1;
}
int MPI_Isend(const void *b, int n, MPI_Datatype t, int d, int tag, MPI_Comm c, MPI_Request *r) {
  (void)b; (void)n; (void)t; (void)d; (void)tag; (void)c; (void)r; fatal("MPI_Isend"); return 1;
}
int MPI_Irecv(void *b, int n, MPI_Datatype t, int s, int tag, MPI_Comm c, MPI_Request *r) {
  (void)b; (void)n; (void)t; (void)s; (void)tag; (void)c; (void)r; fatal("MPI_Irecv"); return 1;
}
int MPI_Wait(MPI_Request *r, MPI_Status *s) { (void)r; (void)s; return MPI_SUCCESS; }
int MPI_Waitall(int n, MPI_Request *r, MPI_Status *s) { (void)n; (void)r; (void)s; return MPI_SUCCESS; }
int MPI_Waitany(int n, MPI_Request *r, int *index, MPI_Status *s) {
  (void)n; (void)r; (void)s; *index = MPI_UNDEFINED; return MPI_SUCCESS;
}
int MPI_Probe(int s, int tag, MPI_Comm c, MPI_Status *st) {
  (void)s; (void)tag; (void)c; (void)st; fatal("MPI_Probe"); return 1;
}
int MPI_Get_count(const MPI_Status *st, MPI_Datatype t, int *count) {
  *count = st ? (int)(st->nbytes_ / (int)esize(t)) : 0;
  return MPI_SUCCESS;
}

int MPI_Bcast(void *b, int n, MPI_Datatype t, int root, MPI_Comm c) {
  (void)b; (void)n; (void)t; (void)root; (void)c; return MPI_SUCCESS;
}
int MPI_Reduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op op, int root, MPI_Comm c) {
  (void)op; (void)root; (void)c; self_copy(s, r, n * esize(t)); return MPI_SUCCESS;
}
int MPI_Allreduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op op, MPI_Comm c) {
  (void)op; (void)c; self_copy(s, r, n * esize(t)); return MPI_SUCCESS;
}
int MPI_Gather(const void *s, int sn, MPI_Datatype st, void *r, int rn, MPI_Datatype rt, int root, MPI_Comm c) {
  (void)rn; (void)rt; (void)root; (void)c; self_copy(s, r, sn * esize(st)); return MPI_SUCCESS;
}
int MPI_Gatherv(const void *s, int sn, MPI_Datatype st, void *r, const int *rc, const int *displs,
                MPI_Datatype rt, int root, MPI_Comm c) {
  (void)rc; (void)root; (void)c;
  self_copy(s, (char *)r + displs[0] * esize(rt), sn * esize(st));
  return MPI_SUCCESS;
}
int MPI_Allgather(const void *s, int sn, MPI_Datatype st, void *r, int rn, MPI_Datatype rt, MPI_Comm c) {
  (void)rn; (void)rt; (void)c; self_copy(s, r, sn * esize(st)); return MPI_SUCCESS;
}
int MPI_Scatter(const void *s, int sn, MPI_Datatype st, void *r, int rn, MPI_Datatype rt, int root, MPI_Comm c) {
  (void)sn; (void)st; (void)root; (void)c; self_copy(s, r, rn * esize(rt)); return MPI_SUCCESS;
}
int MPI_Scatterv(const void *s, const int *sc, const int *displs, MPI_Datatype st, void *r, int rn,
                 MPI_Datatype rt, int root, MPI_Comm c) {
  (void)sc; (void)root; (void)c;
  self_copy((const char *)s + displs[0] * esize(st), r, rn * esize(rt));
  return MPI_SUCCESS;
}
int MPI_Alltoall(const void *s, int sn, MPI_Datatype st, void *r, int rn, MPI_Datatype rt, MPI_Comm c) {
  (void)rn; (void)rt; (void)c; self_copy(s, r, sn * esize(st)); return MPI_SUCCESS;
}
int MPI_Alltoallv(const void *s, const int *sc, const int *sd, MPI_Datatype st, void *r, const int *rc,
                  const int *rd, MPI_Datatype rt, MPI_Comm c) {
  (void)rc; (void)c;
  self_copy((const char *)s + sd[0] * esize(st), (char *)r + rd[0] * esize(rt), sc[0] * esize(st));
  return MPI_SUCCESS;
}

int MPI_Error_string(int code, char *str, int *len) {
  *len = snprintf(str, MPI_MAX_ERROR_STRING, "oracle mpi error %d", code);
  return MPI_SUCCESS;
}
int MPI_Op_create(MPI_User_function *fn, int commute, MPI_Op *op) {
  (void)fn; (void)commute; *op = 100; return MPI_SUCCESS;
}
int MPI_Op_free(MPI_Op *op) { *op = 0; return MPI_SUCCESS; }

int MPI_File_open(MPI_Comm c, const char *name, int mode, MPI_Info info, MPI_File *fp) {
  (void)c; (void)info;
  FILE *f = fopen(name, (mode & MPI_MODE_WRONLY) ? "wb" : "rb");
  if (!f) { *fp = NULL; return 1; }
  *fp = (MPI_File)calloc(1, sizeof(**fp));
  (*fp)->fp = f;
  return MPI_SUCCESS;
}
int MPI_File_close(MPI_File *fp) {
  if (*fp) { fclose((*fp)->fp); free(*fp); *fp = NULL; }
  return MPI_SUCCESS;
}
int MPI_File_set_view(MPI_File fp, MPI_Offset disp, MPI_Datatype et, MPI_Datatype ft, const char *rep, MPI_Info info) {
  (void)et; (void)ft; (void)rep; (void)info; fp->disp = disp; fseek(fp->fp, (long)disp, SEEK_SET);
  return MPI_SUCCESS;
}
int MPI_File_set_size(MPI_File fp, MPI_Offset size) { (void)fp; (void)size; return MPI_SUCCESS; }
int MPI_File_write(MPI_File fp, const void *b, int n, MPI_Datatype t, MPI_Status *st) {
  (void)st; return fwrite(b, esize(t), n, fp->fp) == (size_t)n ? MPI_SUCCESS : 1;
}
int MPI_File_read(MPI_File fp, void *b, int n, MPI_Datatype t, MPI_Status *st) {
  (void)st; return fread(b, esize(t), n, fp->fp) == (size_t)n ? MPI_SUCCESS : 1;
}
int MPI_File_write_at_all(MPI_File fp, MPI_Offset off, const void *b, int n, MPI_Datatype t, MPI_Status *st) {
  (void)st; fseek(fp->fp, (long)(fp->disp + off * (MPI_Offset)esize(t)), SEEK_SET);
  return fwrite(b, esize(t), n, fp->fp) == (size_t)n ? MPI_SUCCESS : 1;
}
int MPI_File_read_at_all(MPI_File fp, MPI_Offset off, void *b, int n, MPI_Datatype t, MPI_Status *st) {
  (void)st; fseek(fp->fp, (long)(fp->disp + off * (MPI_Offset)esize(t)), SEEK_SET);
  return fread(b, esize(t), n, fp->fp) == (size_t)n ? MPI_SUCCESS : 1;
}
