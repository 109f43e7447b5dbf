// Reference-side binding of tacs_b200: thin C++ classes that a TACS application links next to its own libtacs.
//
// This file is compiled against the UNMODIFIED reference headers (TACSAssembler.h, KSM.h, ...). It adapts the three
// seams the hot path sits behind to the flat C ABI of include/tacs_b200.h:
//
//   TACSVec  (src/bpmat/KSM.h:91-115)   ->  TACSB200Vec   device vector, every Krylov operation forwarded
//   TACSMat  (src/bpmat/KSM.h:147-189)  ->  TACSB200Mat   device block-CSR matrix, mult on the B200
//   TACSPc   (src/bpmat/KSM.h:200-221)  ->  TACSB200ChebyshevPc   device polynomial smoother
//   TACSAssembler::assembleRes / assembleJacobian / assembleMatType (src/TACSAssembler.h:221-229)
//                                       ->  TACSB200Assembler, built from a live TACSAssembler
//
// so that the reference's own drivers (examples/plate/plate.cpp, examples/tutorial/tutorial.cpp) and its own Krylov
// solvers (GMRES, PCG, GCROT: they only see TACSVec / TACSMat / TACSPc) run on the device objects unchanged.
// TACSB200Assembler::create reads the mesh, the element objects, the node locations and the boundary conditions out of
// the application's TACSAssembler; every element must be one of the families of the path (TACSQuad4Shell,
// TACSQuad9Shell, TACSElement3D with TACSLinearElasticity3D on a linear / quadratic hexahedral basis), otherwise create
// returns NULL and the caller keeps the reference path (the error convention of SURVEY 8b).
#ifndef TACS_B200_SHIM_H
#define TACS_B200_SHIM_H

#include "KSM.h"
#include "TACSAssembler.h"
#include "TACSSchurMat.h"
#include "tacs_b200.h"

class TACSB200Assembler;

class TACSB200Vec : public TACSVec {
 public:
  TACSB200Vec(tacsb200_handle handle);  // adopts the handle
  ~TACSB200Vec();

  // TACSVec
  TacsScalar norm();
  void scale(TacsScalar alpha);
  TacsScalar dot(TACSVec *x);
  void mdot(TACSVec **x, TacsScalar *ans, int m);
  void axpy(TacsScalar alpha, TACSVec *x);
  void copyValues(TACSVec *x);  // x: TACSB200Vec, or a host TACSBVec of the same size (uploaded)
  void axpby(TacsScalar alpha, TacsScalar beta, TACSVec *x);
  void zeroEntries();
  void applyBCs(TACSBcMap *map, TACSVec *vec = NULL, const TacsScalar lambda = 1.0);
  void setBCs(TACSBcMap *map, const TacsScalar lambda = 1.0);

  // host <-> device
  int getSize();
  void getValues(TacsScalar *host);        // download
  void setValues(const TacsScalar *host);  // upload
  void copyTo(TACSBVec *host);             // download into a host vector of the reference
  tacsb200_handle getHandle() { return handle; }
  void setAssembler(TACSB200Assembler *a) { owner = a; }
  const char *getObjectName() { return "TACSB200Vec"; }

 private:
  tacsb200_handle handle;
  TACSB200Assembler *owner;  // not ref-counted (the assembler outlives what it creates in the drivers)
};

class TACSB200Mat : public TACSMat {
 public:
  TACSB200Mat(TACSB200Assembler *assembler, tacsb200_handle handle);
  ~TACSB200Mat();

  // TACSMat. addValues is not part of the device path: the assembler fills the matrix (assembleJacobian).
  void zeroEntries();
  void applyBCs(TACSBcMap *bcmap);
  void getSize(int *nr, int *nc);
  TACSVec *createVec();
  TACSMat *createDuplicate();
  void mult(TACSVec *x, TACSVec *y);
  void copyValues(TACSMat *mat);
  void scale(TacsScalar alpha);
  void axpy(TacsScalar alpha, TACSMat *mat);

  // BCSRMatData view of the owned rows (host copies): bsize, nrows, rowp, cols, values of Aloc
  void getArrays(int *bsize, int *nrows, int *nnzb, int **rowp, int **cols, TacsScalar **vals);
  tacsb200_handle getHandle() { return handle; }
  const char *getObjectName() { return "TACSB200Mat"; }

 private:
  TACSB200Assembler *assembler;
  tacsb200_handle handle;
};

// TACSSchurMat (src/bpmat/TACSSchurMat.h:58) on the device: the blocks [B E; F C] of a live reference TACSSchurMat --
// its non-zero patterns and its local ordering (AMD / nested dissection, interior before interface unknowns) are read
// from the object the application created with assembler->createSchurMat() -- filled from a device-assembled matrix.
// update() after TACSB200Assembler::assembleJacobian; mult on the device; copyValuesTo writes the device-assembled
// values into a reference TACSSchurMat of the same pattern so that TACSSchurPc::factor / applyFactor run unchanged
// (the assembly the reference spends its time in is replaced, its direct solver is kept).
class TACSB200SchurMat : public TACSMat {
 public:
  TACSB200SchurMat(TACSB200Mat *source, TACSSchurMat *pattern);
  ~TACSB200SchurMat();
  bool valid() { return handle != NULL; }
  int update();
  void copyValuesTo(TACSSchurMat *host);
  TACSVec *createVec();
  void mult(TACSVec *x, TACSVec *y);
  void getSize(int *nr, int *nc);
  tacsb200_handle getHandle() { return handle; }
  const char *getObjectName() { return "TACSB200SchurMat"; }

 private:
  TACSB200Mat *source;
  tacsb200_handle handle;
};

class TACSB200ChebyshevPc : public TACSPc {
 public:
  // TACSChebyshevSmoother(mat, degree, lower_factor, upper_factor, iters) (src/bpmat/TACSParallelMat.h:180-217)
  TACSB200ChebyshevPc(TACSB200Mat *mat, int degree, double lower_factor, double upper_factor, int iters);
  ~TACSB200ChebyshevPc();
  void applyFactor(TACSVec *x, TACSVec *y);
  void factor();
  void getMat(TACSMat **mat);
  tacsb200_handle getHandle() { return handle; }
  const char *getObjectName() { return "TACSB200ChebyshevPc"; }

 private:
  TACSB200Mat *mat;
  tacsb200_handle handle;
};

// GMRES on the device (tacs_b200's own solver) behind the reference's TACSKsm interface
class TACSB200GMRES : public TACSKsm {
 public:
  TACSB200GMRES(TACSB200Mat *mat, TACSB200ChebyshevPc *pc, int m, int nrestart, int is_flexible);
  ~TACSB200GMRES();
  TACSVec *createVec();
  void setOperators(TACSMat *mat, TACSPc *pc);
  void getOperators(TACSMat **mat, TACSPc **pc);
  int solve(TACSVec *b, TACSVec *x, int zero_guess = 1);
  void setTolerances(double rtol, double atol);
  void setMonitor(KSMPrint *monitor);
  const char *getObjectName() { return "TACSB200GMRES"; }

 private:
  TACSB200Mat *mat;
  TACSB200ChebyshevPc *pc;
  tacsb200_handle handle;
  KSMPrint *monitor;
};

class TACSB200Assembler : public TACSObject {
 public:
  // NULL (with a message on stderr) when the assembler holds something outside the device path
  static TACSB200Assembler *create(TACSAssembler *assembler);
  ~TACSB200Assembler();

  TACSB200Vec *createVec();
  TACSB200Mat *createMat();

  // TACSAssembler::setVariables / assembleRes / assembleJacobian / assembleMatType / applyBCs / setBCs
  // (src/TACSAssembler.h:197-229). Vector arguments: TACSB200Vec, or the reference's host TACSBVec (copied).
  void setVariables(TACSVec *q, TACSVec *qdot = NULL, TACSVec *qddot = NULL);
  void setNodes(TACSBVec *X);
  void assembleRes(TACSB200Vec *res);
  void assembleJacobian(TacsScalar alpha, TacsScalar beta, TacsScalar gamma, TACSB200Vec *res, TACSB200Mat *A);
  int assembleMatType(ElementMatrixType matType, TACSB200Mat *A);
  // setVariables(vars) + assembleJacobian for drivers whose state and residual are the reference's host vectors
  // (integrators, Newton loops): one call of tacsb200_assembler_assemble_jacobian_host, transfers pipelined against the
  // kernels. Returns when `res` is complete; the matrix may still be in flight (device-side consumers are ordered
  // behind it). Not tested on hardware in round 2 (added after the last GPU run; a forwarding call).
  int assembleJacobianHost(TacsScalar alpha, TacsScalar beta, TacsScalar gamma, TACSBVec *vars, TACSBVec *res,
                           TACSB200Mat *A);
  // The drop-in for drivers that assemble into a TACSSchurMat (examples/plate/plate.cpp:133-146, pyTACS
  // StaticProblem): element loop, scatter and boundary conditions on the device, then the values land in the blocks
  // of `mat`; `res` (host vector, may be NULL) receives the residual.
  int assembleJacobian(TacsScalar alpha, TacsScalar beta, TacsScalar gamma, TACSBVec *res, TACSSchurMat *mat);
  void applyBCs(TACSVec *vec);
  void setBCs(TACSVec *vec);

  TACSAssembler *getAssembler() { return assembler; }
  tacsb200_handle getHandle() { return handle; }
  int getNumOwnedNodes() { return num_nodes; }
  int getVarsPerNode() { return vars_per_node; }
  const char *getObjectName() { return "TACSB200Assembler"; }

 private:
  TACSB200Assembler()
      : assembler(NULL), handle(NULL), creator(NULL), num_nodes(0), vars_per_node(0), scratch_q(NULL), scratch_qd(NULL),
        scratch_qdd(NULL), schur_source(NULL), schur_view(NULL), schur_res(NULL) {}
  TACSB200Vec *stage(TACSVec *v, TACSB200Vec **scratch);
  TACSAssembler *assembler;
  tacsb200_handle handle, creator;
  int num_nodes, vars_per_node;
  TACSB200Vec *scratch_q, *scratch_qd, *scratch_qdd;
  // device objects behind assembleJacobian(..., TACSSchurMat *): created on first use, keyed by the Schur ordering
  TACSB200Mat *schur_source;
  TACSB200SchurMat *schur_view;
  TACSB200Vec *schur_res;
};

#endif  // TACS_B200_SHIM_H
