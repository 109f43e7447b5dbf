/*
 * int32 -> int64 adapter between the reference's METIS calls (32-bit idx_t,
 * /root/reference/src/TACSCreator.cpp:1105-1124, TACSAssembler.cpp:1620-1624)
 * and the CUDA toolkit's libmetis_static.a (64-bit idx_t).  The toolkit archive
 * is partially linked into one object and its four entry points renamed to
 * metis64_* (objcopy --redefine-sym in the build), so these 32-bit wrappers can
 * carry the original names.  The product library links this adapter for
 * TACSCreator-style element partitioning; oracle/Makefile compiles the same file
 * into oracle/_ref so both sides consume the partition of one METIS binary.
 */
#include <stdint.h>
#include <stdlib.h>
#include <sys/stat.h>

#include "metis.h"

typedef int64_t idx64;
int metis64_SetDefaultOptions(idx64 *options);
int metis64_PartGraphRecursive(idx64 *, idx64 *, idx64 *, idx64 *, idx64 *, idx64 *, idx64 *, idx64 *,
                               float *, float *, idx64 *, idx64 *, idx64 *);
int metis64_PartGraphKway(idx64 *, idx64 *, idx64 *, idx64 *, idx64 *, idx64 *, idx64 *, idx64 *,
                          float *, float *, idx64 *, idx64 *, idx64 *);
int metis64_NodeND(idx64 *, idx64 *, idx64 *, idx64 *, idx64 *, idx64 *, idx64 *);

/* GKlib inside the archive references the pre-2.33 glibc stat entry point. */
int __xstat64(int ver, const char *path, void *buf) {
  (void)ver;
  return stat(path, (struct stat *)buf);
}

static idx64 *widen(const int *a, size_t n) {
  if (!a) return NULL;
  idx64 *w = (idx64 *)malloc((n ? n : 1) * sizeof(idx64));
  for (size_t i = 0; i < n; i++) w[i] = a[i];
  return w;
}

int METIS_SetDefaultOptions(int *options) {
  idx64 o[METIS_NOPTIONS];
  int rc = metis64_SetDefaultOptions(o);
  for (int i = 0; i < METIS_NOPTIONS; i++) options[i] = (int)o[i];
  return rc;
}

typedef int (*part_fn)(idx64 *, idx64 *, idx64 *, idx64 *, idx64 *, idx64 *, idx64 *, idx64 *, float *,
                       float *, idx64 *, idx64 *, idx64 *);

static int part_graph(part_fn fn, int *nvtxs, int *ncon, int *xadj, int *adjncy, int *vwgt, int *vsize,
                      int *adjwgt, int *nparts, float *tpwgts, float *ubvec, int *options, int *edgecut,
                      int *part) {
  idx64 n = *nvtxs, nc = *ncon, np = *nparts, cut = 0;
  size_t nedge = (size_t)xadj[n];
  idx64 *x = widen(xadj, (size_t)n + 1), *a = widen(adjncy, nedge);
  idx64 *vw = widen(vwgt, (size_t)(n * nc)), *vs = widen(vsize, (size_t)n), *aw = widen(adjwgt, nedge);
  idx64 *op = widen(options, METIS_NOPTIONS);
  idx64 *p = (idx64 *)malloc((size_t)(n ? n : 1) * sizeof(idx64));
  int rc = fn(&n, &nc, x, a, vw, vs, aw, &np, tpwgts, ubvec, op, &cut, p);
  for (idx64 i = 0; i < n; i++) part[i] = (int)p[i];
  *edgecut = (int)cut;
  free(x); free(a); free(vw); free(vs); free(aw); free(op); free(p);
  return rc;
}

int METIS_PartGraphRecursive(int *nvtxs, int *ncon, int *xadj, int *adjncy, int *vwgt, int *vsize,
                             int *adjwgt, int *nparts, float *tpwgts, float *ubvec, int *options,
                             int *edgecut, int *part) {
  return part_graph(metis64_PartGraphRecursive, nvtxs, ncon, xadj, adjncy, vwgt, vsize, adjwgt, nparts,
                    tpwgts, ubvec, options, edgecut, part);
}

int METIS_PartGraphKway(int *nvtxs, int *ncon, int *xadj, int *adjncy, int *vwgt, int *vsize,
                        int *adjwgt, int *nparts, float *tpwgts, float *ubvec, int *options,
                        int *edgecut, int *part) {
  return part_graph(metis64_PartGraphKway, nvtxs, ncon, xadj, adjncy, vwgt, vsize, adjwgt, nparts, tpwgts,
                    ubvec, options, edgecut, part);
}

int METIS_NodeND(int *nvtxs, int *xadj, int *adjncy, int *vwgt, int *options, int *perm, int *iperm) {
  idx64 n = *nvtxs;
  size_t nedge = (size_t)xadj[n];
  idx64 *x = widen(xadj, (size_t)n + 1), *a = widen(adjncy, nedge), *vw = widen(vwgt, (size_t)n);
  idx64 *op = widen(options, METIS_NOPTIONS);
  idx64 *p = (idx64 *)malloc((size_t)(n ? n : 1) * sizeof(idx64));
  idx64 *ip = (idx64 *)malloc((size_t)(n ? n : 1) * sizeof(idx64));
  int rc = metis64_NodeND(&n, x, a, vw, op, p, ip);
  for (idx64 i = 0; i < n; i++) { perm[i] = (int)p[i]; iperm[i] = (int)ip[i]; }
  free(x); free(a); free(vw); free(op); free(p); free(ip);
  return rc;
}
