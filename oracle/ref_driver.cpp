/*
 * oracle/ref_driver.cpp -- TEST INFRASTRUCTURE. Standalone driver that runs the UNMODIFIED reference
 * on N ranks (oracle/mpistub/mpi_procs.c: TACSB200_MPI_NP forked processes) for one mesh written by
 * the Python tests, and dumps what each rank holds so the product's integer pipeline can be compared
 * bit-for-bit on N ranks: node renumbering and partition (root), owner range, local->global node map,
 * element connectivity, Aloc/Bext patterns, np, external column nodes; plus A, residual and A*x.
 *
 *   TACSB200_MPI_NP=4 ref_driver <indir> <outdir> [timing_reps]
 *
 * <indir>: meta.txt ("vars_per_node num_nodes num_elements nodes_per_elem num_bcs elem_kind con_kind")
 *          ptr.bin conn.bin ids.bin bc.bin (int32), X.bin (float64)       -- read on rank 0 only
 * elem_kind 1 Quad4 2 Quad9 3 hex8 4 hex27; con_kind 0 iso shell (t=0.01) 1 composite [0/45/30]s 2 solid
 * State and input vectors are deterministic functions of the global dof index (tests/common hash),
 * so no rank needs data from another.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "TACSAssembler.h"
#include "TACSCompositeShellConstitutive.h"
#include "TACSCreator.h"
#include "TACSElement3D.h"
#include "TACSHexaBasis.h"
#include "TACSIsoShellConstitutive.h"
#include "TACSLinearElasticity.h"
#include "TACSShellElementDefs.h"
#include "TACSSolidConstitutive.h"

template <class T>
static std::vector<T> read_bin(const std::string &path, size_t n) {
  std::vector<T> v(n);
  FILE *f = fopen(path.c_str(), "rb");
  if (!f || fread(v.data(), sizeof(T), n, f) != n) {
    fprintf(stderr, "ref_driver: cannot read %s\n", path.c_str());
    abort();
  }
  fclose(f);
  return v;
}
template <class T>
static void write_bin(const std::string &dir, int rank, const char *name, const T *data, size_t n) {
  char path[1024];
  snprintf(path, sizeof(path), "%s/r%d_%s.bin", dir.c_str(), rank, name);
  FILE *f = fopen(path, "wb");
  if (!f) abort();
  if (n) fwrite(data, sizeof(T), n, f);
  fclose(f);
}
static double hashval(long i) { return 1e-3 * (double)(((unsigned long)i * 2654435761ul % 4294967296ul) % 1000ul) / 1000.0; }

int main(int argc, char **argv) {
  MPI_Init(&argc, &argv);
  int rank, size;
  MPI_Comm_rank(MPI_COMM_WORLD, &rank);
  MPI_Comm_size(MPI_COMM_WORLD, &size);
  if (argc < 3) return 1;
  std::string in = argv[1], out = argv[2];
  int reps = argc > 3 ? atoi(argv[3]) : 0;
  int meta[7];
  if (rank == 0) {
    FILE *f = fopen((in + "/meta.txt").c_str(), "r");
    if (!f || fscanf(f, "%d %d %d %d %d %d %d", &meta[0], &meta[1], &meta[2], &meta[3], &meta[4], &meta[5], &meta[6]) != 7) abort();
    fclose(f);
  }
  MPI_Bcast(meta, 7, MPI_INT, 0, MPI_COMM_WORLD);
  const int vpn = meta[0], nnodes = meta[1], nelems = meta[2], npe = meta[3], nbcs = meta[4], kind = meta[5], con = meta[6];

  TACSCreator *creator = new TACSCreator(MPI_COMM_WORLD, vpn);
  creator->incref();
  if (rank == 0) {
    std::vector<int> ptr = read_bin<int>(in + "/ptr.bin", nelems + 1), conn = read_bin<int>(in + "/conn.bin", (size_t)nelems * npe);
    std::vector<int> ids = read_bin<int>(in + "/ids.bin", nelems), bc = read_bin<int>(in + "/bc.bin", nbcs);
    std::vector<double> X = read_bin<double>(in + "/X.bin", 3 * (size_t)nnodes);
    creator->setGlobalConnectivity(nnodes, nelems, ptr.data(), conn.data(), ids.data());
    creator->setBoundaryConditions(nbcs, bc.data());
    creator->setNodes(X.data());
  }
  TACSElement *elem = NULL;
  if (kind <= 2) {
    TACSShellConstitutive *c = NULL;
    if (con == 0) {
      TACSMaterialProperties *p = new TACSMaterialProperties(2700.0, 921.096, 70e3, 0.3, 270.0, 24e-6, 230.0);
      c = new TACSIsoShellConstitutive(p, 0.01);
    } else {
      TACSMaterialProperties *p = new TACSMaterialProperties(1550.0, 0.0, 54e3, 18e3, 18e3, 0.25, 0.25, 0.25, 9e3, 9e3, 9e3);
      TACSOrthotropicPly *ply = new TACSOrthotropicPly(1.25e-4, p);
      TACSOrthotropicPly *plies[6] = {ply, ply, ply, ply, ply, ply};
      TacsScalar th[6], ang[6] = {0.0, 45.0, 30.0, 30.0, 45.0, 0.0};
      for (int i = 0; i < 6; i++) { th[i] = 1.25e-4; ang[i] *= M_PI / 180.0; }
      c = new TACSCompositeShellConstitutive(6, plies, th, ang, 5.0 / 6.0, 0.0);
    }
    TacsScalar axis[3] = {1.0, 0.0, 0.0};
    TACSShellTransform *t = new TACSShellRefAxisTransform(axis);
    elem = kind == 1 ? (TACSElement *)new TACSQuad4Shell(t, c) : (TACSElement *)new TACSQuad9Shell(t, c);
  } else {
    TACSMaterialProperties *p = new TACSMaterialProperties(2700.0, 921.096, 70e3, 0.3, 270.0, 24e-6, 230.0);
    TACSSolidConstitutive *s = new TACSSolidConstitutive(p, 1.0, -1);
    TACSElementModel *m = new TACSLinearElasticity3D(s, TACS_LINEAR_STRAIN);
    TACSElementBasis *b = kind == 3 ? (TACSElementBasis *)new TACSLinearHexaBasis() : (TACSElementBasis *)new TACSQuadraticHexaBasis();
    elem = new TACSElement3D(m, b);
  }
  creator->setElements(1, &elem);
  TACSAssembler *a = creator->createTACS();
  a->incref();
  if (rank == 0) {
    const int *nn = NULL, *part = NULL;
    creator->getNodeNums(&nn);
    creator->getElementPartition(&part);
    write_bin(out, 0, "new_nodes", nn, nnodes);
    write_bin(out, 0, "partition", part, nelems);
  }
  const int bs = vpn;
  const int *range;
  a->getNodeMap()->getOwnerRange(&range);
  const int lo = range[rank], hi = range[rank + 1];
  write_bin(out, rank, "owner_range", range, size + 1);
  {
    int n = a->getNumNodes();
    std::vector<int> l2g(n);
    for (int i = 0; i < n; i++) l2g[i] = a->getGlobalNodeNum(i);
    write_bin(out, rank, "local_to_global", l2g.data(), n);
    const int *ptr, *conn;
    a->getElementConnectivity(&ptr, &conn);
    write_bin(out, rank, "elem_conn", conn, ptr[a->getNumElements()]);
  }
  TACSParallelMat *A = a->createMat();
  A->incref();
  TACSBVec *res = a->createVec(), *u = a->createVec(), *x = a->createVec(), *y = a->createVec();
  res->incref(); u->incref(); x->incref(); y->incref();
  TacsScalar *ua, *xa;
  int n = u->getArray(&ua);
  x->getArray(&xa);
  const long ntot = (long)bs * nnodes;
  for (int i = 0; i < n; i++) {
    long g = (long)bs * lo + i;
    ua[i] = hashval(g);
    xa[i] = hashval(ntot - 1 - g);
  }
  a->applyBCs(u);
  a->applyBCs(x);
  a->setVariables(u);
  a->assembleJacobian(1.0, 0.0, 0.0, res, A);
  A->mult(x, y);
  BCSRMat *Al, *Bx;
  A->getBCSRMat(&Al, &Bx);
  int b, nr, nc;
  const int *rowp, *cols;
  TacsScalar *vals;
  Al->getArrays(&b, &nr, &nc, &rowp, &cols, &vals);
  write_bin(out, rank, "Aloc_rowp", rowp, nr + 1);
  write_bin(out, rank, "Aloc_cols", cols, rowp[nr]);
  write_bin(out, rank, "Aloc_vals", vals, (size_t)b * b * rowp[nr]);
  Bx->getArrays(&b, &nr, &nc, &rowp, &cols, &vals);
  write_bin(out, rank, "Bext_rowp", rowp, nr + 1);
  write_bin(out, rank, "Bext_cols", cols, rowp[nr]);
  write_bin(out, rank, "Bext_vals", vals, (size_t)b * b * rowp[nr]);
  {
    TACSBVecDistribute *dist;
    A->getExtColMap(&dist);
    const int *idx;
    int ne = dist->getIndices()->getIndices(&idx);
    write_bin(out, rank, "ext_col_nodes", idx, ne);
  }
  TacsScalar *ra, *ya;
  res->getArray(&ra);
  y->getArray(&ya);
  write_bin(out, rank, "res", ra, n);
  write_bin(out, rank, "y", ya, n);
  write_bin(out, rank, "u", ua, n);
  write_bin(out, rank, "x", xa, n);
  double ynorm = y->norm();
  if (reps > 0) {
    /* all-ranks CPU baseline: assembleJacobian wall time, max over ranks */
    MPI_Barrier(MPI_COMM_WORLD);
    double t0 = MPI_Wtime();
    for (int k = 0; k < reps; k++) a->assembleJacobian(1.0, 0.0, 0.0, res, A);
    MPI_Barrier(MPI_COMM_WORLD);
    double tj = (MPI_Wtime() - t0) / reps;
    t0 = MPI_Wtime();
    for (int k = 0; k < 10; k++) A->mult(x, y);
    MPI_Barrier(MPI_COMM_WORLD);
    double tm = (MPI_Wtime() - t0) / 10;
    if (rank == 0) printf("{\"ranks\": %d, \"elements\": %d, \"jac_s\": %.6e, \"elements_per_s\": %.6e, \"spmv_s\": %.6e, \"ynorm\": %.15e}\n", size, nelems, tj, nelems / tj, tm, ynorm);
  } else if (rank == 0) {
    printf("{\"ranks\": %d, \"elements\": %d, \"ynorm\": %.15e}\n", size, nelems, ynorm);
  }
  fflush(stdout);
  MPI_Finalize();
  return 0;
}
