/*
 * oracle/ref_capi.cpp  --  TEST INFRASTRUCTURE, not product code.
 *
 * A flat C interface over the UNMODIFIED reference library (compiled from
 * /root/reference/src by oracle/Makefile into oracle/_ref/libtacs_ref.so) so
 * that tests/ and bench.py's cpu_baseline / --impl reference legs can drive
 * the reference's own TACSCreator / TACSAssembler / TACSParallelMat / TACSBVec
 * / GMRES objects through ctypes.  Every entry point has the same name and
 * argument list as the product entry point declared in include/tacs_b200.h,
 * with the prefix `ref_` in place of `tacsb200_`, so one Python binding drives
 * both sides of a parity test.
 *
 * Nothing in tacs_b200/ may link or load this file.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "KSM.h"
#include "TACSAssembler.h"
#include "TACSCompositeShellConstitutive.h"
#include "TACSCreator.h"
#include "TACSElement3D.h"
#include "TACSHexaBasis.h"
#include "TACSIsoShellConstitutive.h"
#include "TACSLinearElasticity.h"
#include "TACSMaterialProperties.h"
#include "TACSShellElementDefs.h"
#include "TACSAuxElements.h"
#include "TACSShellPressure.h"
#include "TACSShellTraction.h"
#include "TACSSolidConstitutive.h"

typedef void *ref_handle;

static bool mpi_ready = false;
static void ensure_mpi() {
  if (!mpi_ready) {
    MPI_Init(NULL, NULL);
    mpi_ready = true;
  }
}

template <class T>
static T *as(ref_handle h) {
  return dynamic_cast<T *>(static_cast<TACSObject *>(h));
}

static ref_handle keep(TACSObject *o) {
  if (o) o->incref();
  return static_cast<ref_handle>(o);
}

extern "C" {

int ref_abi_version(void) { return 1; }

void ref_release(ref_handle h) {
  if (h) static_cast<TACSObject *>(h)->decref();
}

/* ---- constitutive ---------------------------------------------------- */
ref_handle ref_material_properties_create(double rho, double specific_heat, double E, double nu,
                                          double ys, double alpha, double kappa) {
  ensure_mpi();
  return keep(new TACSMaterialProperties(rho, specific_heat, E, nu, ys, alpha, kappa));
}

ref_handle ref_material_properties_create_ortho(double rho, double specific_heat, double E1, double E2,
                                                double E3, double nu12, double nu13, double nu23,
                                                double G12, double G13, double G23) {
  ensure_mpi();
  return keep(new TACSMaterialProperties(rho, specific_heat, E1, E2, E3, nu12, nu13, nu23, G12, G13, G23));
}

ref_handle ref_orthotropic_ply_create(double ply_thickness, ref_handle props) {
  return keep(new TACSOrthotropicPly(ply_thickness, as<TACSMaterialProperties>(props)));
}

ref_handle ref_iso_shell_constitutive_create(ref_handle props, double t, double tOffset, double kcorr) {
  return keep(new TACSIsoShellConstitutive(as<TACSMaterialProperties>(props), t, -1, 0.0, 10.0, tOffset,
                                           kcorr));
}

ref_handle ref_composite_shell_constitutive_create(int num_plies, ref_handle *plies,
                                                   const double *ply_thickness, const double *ply_angles,
                                                   double kcorr, double tOffset) {
  std::vector<TACSOrthotropicPly *> p(num_plies);
  for (int i = 0; i < num_plies; i++) p[i] = as<TACSOrthotropicPly>(plies[i]);
  return keep(new TACSCompositeShellConstitutive(num_plies, p.data(), ply_thickness, ply_angles, kcorr,
                                                 tOffset));
}

ref_handle ref_solid_constitutive_create(ref_handle props, double t) {
  return keep(new TACSSolidConstitutive(as<TACSMaterialProperties>(props), t, -1));
}

void ref_shell_set_drilling_regularization(double k) { TACSShellConstitutive::setDrillingRegularization(k); }

/* C has 22 entries for a shell constitutive object, 21 for a solid one. */
int ref_constitutive_eval_tangent_stiffness(ref_handle con, double *C) {
  double pt[3] = {0.0, 0.0, 0.0};
  TacsScalar X[3] = {0.0, 0.0, 0.0};
  TACSConstitutive *c = as<TACSConstitutive>(con);
  if (!c) return 1;
  c->evalTangentStiffness(0, pt, X, C);
  return 0;
}

int ref_shell_constitutive_eval_mass_moments(ref_handle con, double *moments) {
  double pt[3] = {0.0, 0.0, 0.0};
  TacsScalar X[3] = {0.0, 0.0, 0.0};
  TACSShellConstitutive *c = as<TACSShellConstitutive>(con);
  if (!c) return 1;
  c->evalMassMoments(0, pt, X, moments);
  return 0;
}

/* ---- transforms / elements -------------------------------------------- */
ref_handle ref_shell_natural_transform_create(void) {
  ensure_mpi();
  return keep(new TACSShellNaturalTransform());
}

ref_handle ref_shell_ref_axis_transform_create(const double axis[3]) {
  ensure_mpi();
  return keep(new TACSShellRefAxisTransform(axis));
}

ref_handle ref_quad4_shell_create(ref_handle transform, ref_handle con) {
  return keep(new TACSQuad4Shell(as<TACSShellTransform>(transform), as<TACSShellConstitutive>(con)));
}

ref_handle ref_quad9_shell_create(ref_handle transform, ref_handle con) {
  return keep(new TACSQuad9Shell(as<TACSShellTransform>(transform), as<TACSShellConstitutive>(con)));
}

ref_handle ref_linear_hexa_basis_create(void) {
  ensure_mpi();
  return keep(new TACSLinearHexaBasis());
}

ref_handle ref_quadratic_hexa_basis_create(void) {
  ensure_mpi();
  return keep(new TACSQuadraticHexaBasis());
}

ref_handle ref_linear_elasticity3d_create(ref_handle con) {
  return keep(new TACSLinearElasticity3D(as<TACSSolidConstitutive>(con), TACS_LINEAR_STRAIN));
}

ref_handle ref_element3d_create(ref_handle model, ref_handle basis) {
  return keep(new TACSElement3D(as<TACSElementModel>(model), as<TACSElementBasis>(basis)));
}

int ref_element_num_nodes(ref_handle e) { return as<TACSElement>(e)->getNumNodes(); }
int ref_element_vars_per_node(ref_handle e) { return as<TACSElement>(e)->getVarsPerNode(); }

/* Element-level evaluation of a batch of `count` elements that all use the
   descriptor `e`.  Arrays are element-major: Xpts[count][3*nn], vars[count][nv],
   res[count][nv], mat[count][nv*nv]; res/mat are overwritten. dvars/ddvars may be NULL. */
int ref_element_add_jacobian(ref_handle e, int count, double alpha, double beta, double gamma,
                             const double *Xpts, const double *vars, const double *dvars,
                             const double *ddvars, double *res, double *mat) {
  TACSElement *el = as<TACSElement>(e);
  if (!el) return 1;
  const int nn = el->getNumNodes(), nv = el->getNumVariables();
  std::vector<TacsScalar> zero(nv, 0.0);
  for (int k = 0; k < count; k++) {
    TacsScalar *r = res + (size_t)k * nv, *m = mat + (size_t)k * nv * nv;
    memset(r, 0, nv * sizeof(TacsScalar));
    memset(m, 0, (size_t)nv * nv * sizeof(TacsScalar));
    el->addJacobian(k, 0.0, alpha, beta, gamma, Xpts + (size_t)k * 3 * nn, vars + (size_t)k * nv,
                    dvars ? dvars + (size_t)k * nv : zero.data(),
                    ddvars ? ddvars + (size_t)k * nv : zero.data(), r, m);
  }
  return 0;
}

int ref_element_add_residual(ref_handle e, int count, const double *Xpts, const double *vars,
                             const double *dvars, const double *ddvars, double *res) {
  TACSElement *el = as<TACSElement>(e);
  if (!el) return 1;
  const int nn = el->getNumNodes(), nv = el->getNumVariables();
  std::vector<TacsScalar> zero(nv, 0.0);
  for (int k = 0; k < count; k++) {
    TacsScalar *r = res + (size_t)k * nv;
    memset(r, 0, nv * sizeof(TacsScalar));
    el->addResidual(k, 0.0, Xpts + (size_t)k * 3 * nn, vars + (size_t)k * nv,
                    dvars ? dvars + (size_t)k * nv : zero.data(),
                    ddvars ? ddvars + (size_t)k * nv : zero.data(), r);
  }
  return 0;
}

/* ---- creator ----------------------------------------------------------- */
ref_handle ref_creator_create(int vars_per_node) {
  ensure_mpi();
  return keep(new TACSCreator(MPI_COMM_WORLD, vars_per_node));
}

int ref_comm_rank(void) {
  ensure_mpi();
  int r;
  MPI_Comm_rank(MPI_COMM_WORLD, &r);
  return r;
}

int ref_comm_size(void) {
  ensure_mpi();
  int s;
  MPI_Comm_size(MPI_COMM_WORLD, &s);
  return s;
}

int ref_creator_set_global_connectivity(ref_handle c, int num_nodes, int num_elements, const int *ptr,
                                        const int *conn, const int *elem_id_nums) {
  as<TACSCreator>(c)->setGlobalConnectivity(num_nodes, num_elements, ptr, conn, elem_id_nums);
  return 0;
}

int ref_creator_set_boundary_conditions(ref_handle c, int num_bcs, const int *bc_nodes, const int *bc_ptr,
                                        const int *bc_vars, const double *bc_vals) {
  as<TACSCreator>(c)->setBoundaryConditions(num_bcs, bc_nodes, bc_ptr, bc_vars, bc_vals);
  return 0;
}

int ref_creator_set_nodes(ref_handle c, const double *Xpts) {
  as<TACSCreator>(c)->setNodes(Xpts);
  return 0;
}

int ref_creator_set_elements(ref_handle c, int num_elems, ref_handle *elems) {
  std::vector<TACSElement *> e(num_elems);
  for (int i = 0; i < num_elems; i++) e[i] = as<TACSElement>(elems[i]);
  as<TACSCreator>(c)->setElements(num_elems, e.data());
  return 0;
}

int ref_creator_partition_mesh(ref_handle c, int split_size, const int *part) {
  as<TACSCreator>(c)->partitionMesh(split_size, part);
  return 0;
}

/* Copies out new_nodes[num_nodes] (root rank only; returns the length). */
int ref_creator_get_node_nums(ref_handle c, int *new_nodes) {
  const int *nn = NULL;
  int n = as<TACSCreator>(c)->getNodeNums(&nn);
  if (new_nodes && nn) memcpy(new_nodes, nn, n * sizeof(int));
  return nn ? n : 0;
}

int ref_creator_get_element_partition(ref_handle c, int *partition) {
  const int *p = NULL;
  int n = as<TACSCreator>(c)->getElementPartition(&p);
  if (partition && p) memcpy(partition, p, n * sizeof(int));
  return p ? n : 0;
}

ref_handle ref_creator_create_tacs(ref_handle c) { return keep(as<TACSCreator>(c)->createTACS()); }

/* ---- assembler ---------------------------------------------------------- */
int ref_assembler_get_vars_per_node(ref_handle a) { return as<TACSAssembler>(a)->getVarsPerNode(); }
int ref_assembler_get_num_nodes(ref_handle a) { return as<TACSAssembler>(a)->getNumNodes(); }
int ref_assembler_get_num_owned_nodes(ref_handle a) { return as<TACSAssembler>(a)->getNumOwnedNodes(); }
int ref_assembler_get_num_elements(ref_handle a) { return as<TACSAssembler>(a)->getNumElements(); }

/* owner range of this rank: [lo, hi) in global node numbers */
int ref_assembler_get_owner_range(ref_handle a, int *lo, int *hi) {
  TACSAssembler *t = as<TACSAssembler>(a);
  const int *range;
  t->getNodeMap()->getOwnerRange(&range);
  int rank;
  MPI_Comm_rank(t->getMPIComm(), &rank);
  *lo = range[rank];
  *hi = range[rank + 1];
  return 0;
}

/* local element connectivity in GLOBAL node numbers */
int ref_assembler_get_element_connectivity(ref_handle a, int *ptr, int *conn) {
  TACSAssembler *t = as<TACSAssembler>(a);
  const int *p, *c;
  t->getElementConnectivity(&p, &c);
  int ne = t->getNumElements();
  if (ptr) memcpy(ptr, p, (ne + 1) * sizeof(int));
  if (conn) memcpy(conn, c, p[ne] * sizeof(int));
  return p[ne];
}

/* local -> global node number map, length getNumNodes() */
int ref_assembler_get_local_to_global(ref_handle a, int *global) {
  TACSAssembler *t = as<TACSAssembler>(a);
  int n = t->getNumNodes();
  for (int i = 0; i < n; i++) global[i] = t->getGlobalNodeNum(i);
  return n;
}

ref_handle ref_assembler_create_vec(ref_handle a) { return keep(as<TACSAssembler>(a)->createVec()); }
ref_handle ref_assembler_create_node_vec(ref_handle a) { return keep(as<TACSAssembler>(a)->createNodeVec()); }
ref_handle ref_assembler_create_mat(ref_handle a) { return keep(as<TACSAssembler>(a)->createMat()); }

int ref_assembler_get_nodes(ref_handle a, ref_handle X) {
  as<TACSAssembler>(a)->getNodes(as<TACSBVec>(X));
  return 0;
}

int ref_assembler_set_nodes(ref_handle a, ref_handle X) {
  as<TACSAssembler>(a)->setNodes(as<TACSBVec>(X));
  return 0;
}

int ref_assembler_set_variables(ref_handle a, ref_handle q, ref_handle qdot, ref_handle qddot) {
  as<TACSAssembler>(a)->setVariables(as<TACSBVec>(q), qdot ? as<TACSBVec>(qdot) : NULL,
                                     qddot ? as<TACSBVec>(qddot) : NULL);
  return 0;
}

int ref_assembler_zero_variables(ref_handle a) {
  TACSAssembler *t = as<TACSAssembler>(a);
  t->zeroVariables();
  t->zeroDotVariables();
  t->zeroDDotVariables();
  return 0;
}

int ref_assembler_apply_bcs_vec(ref_handle a, ref_handle v) {
  as<TACSAssembler>(a)->applyBCs(as<TACSVec>(v));
  return 0;
}

int ref_assembler_apply_bcs_mat(ref_handle a, ref_handle m) {
  as<TACSAssembler>(a)->applyBCs(as<TACSMat>(m));
  return 0;
}

int ref_assembler_set_bcs(ref_handle a, ref_handle v) {
  as<TACSAssembler>(a)->setBCs(as<TACSVec>(v));
  return 0;
}

int ref_assembler_set_num_threads(ref_handle a, int t) {
  as<TACSAssembler>(a)->setNumThreads(t);
  return 0;
}

int ref_assembler_assemble_res(ref_handle a, ref_handle res) {
  as<TACSAssembler>(a)->assembleRes(as<TACSBVec>(res));
  return 0;
}

int ref_assembler_assemble_jacobian(ref_handle a, double alpha, double beta, double gamma, ref_handle res,
                                    ref_handle mat) {
  as<TACSAssembler>(a)->assembleJacobian(alpha, beta, gamma, res ? as<TACSBVec>(res) : NULL,
                                         as<TACSMat>(mat));
  return 0;
}

int ref_assembler_assemble_jacobian_async(ref_handle a, double alpha, double beta, double gamma, ref_handle res,
                                          ref_handle mat) {
  return ref_assembler_assemble_jacobian(a, alpha, beta, gamma, res, mat);  /* the CPU path has no asynchrony */
}

int ref_assembler_assemble_mat_type(ref_handle a, int mat_type, ref_handle mat, int apply_bcs) {
  as<TACSAssembler>(a)->assembleMatType((ElementMatrixType)mat_type, as<TACSMat>(mat), TACS_MAT_NORMAL, 1.0,
                                        apply_bcs != 0);
  return 0;
}

int ref_assembler_add_jacobian_vec_product(ref_handle a, double scale, double alpha, double beta, double gamma,
                                           ref_handle x, ref_handle y, int apply_bcs) {
  as<TACSAssembler>(a)->addJacobianVecProduct(scale, alpha, beta, gamma, as<TACSBVec>(x), as<TACSBVec>(y),
                                              TACS_MAT_NORMAL, 1.0, apply_bcs != 0);
  return 0;
}

/* ---- vectors ------------------------------------------------------------- */
int ref_vec_get_size(ref_handle v) {
  TacsScalar *x;
  return as<TACSBVec>(v)->getArray(&x);
}

int ref_vec_get_array(ref_handle v, double *out) {
  TacsScalar *x;
  int n = as<TACSBVec>(v)->getArray(&x);
  memcpy(out, x, n * sizeof(double));
  return 0;
}

int ref_vec_set_array(ref_handle v, const double *in) {
  TacsScalar *x;
  int n = as<TACSBVec>(v)->getArray(&x);
  memcpy(x, in, n * sizeof(double));
  return 0;
}

double ref_vec_norm(ref_handle v) { return as<TACSBVec>(v)->norm(); }
double ref_vec_dot(ref_handle x, ref_handle y) { return as<TACSBVec>(x)->dot(as<TACSBVec>(y)); }
int ref_vec_mdot(ref_handle x, int n, ref_handle *ys, double *out) {
  std::vector<TACSVec *> v(n);
  for (int i = 0; i < n; i++) v[i] = as<TACSBVec>(ys[i]);
  as<TACSBVec>(x)->mdot(v.data(), out, n);
  return 0;
}
int ref_vec_axpy(ref_handle y, double alpha, ref_handle x) {
  as<TACSBVec>(y)->axpy(alpha, as<TACSBVec>(x));
  return 0;
}
int ref_vec_axpby(ref_handle y, double alpha, double beta, ref_handle x) {
  as<TACSBVec>(y)->axpby(alpha, beta, as<TACSBVec>(x));
  return 0;
}
int ref_vec_scale(ref_handle y, double alpha) {
  as<TACSBVec>(y)->scale(alpha);
  return 0;
}
int ref_vec_copy_values(ref_handle y, ref_handle x) {
  as<TACSBVec>(y)->copyValues(as<TACSBVec>(x));
  return 0;
}
int ref_vec_zero_entries(ref_handle y) {
  as<TACSBVec>(y)->zeroEntries();
  return 0;
}

/* ---- matrix --------------------------------------------------------------- */
/* which: 0 = Aloc (owned rows x owned cols), 1 = Bext (rows >= np x external cols) */
static BCSRMat *pick(ref_handle m, int which) {
  BCSRMat *A, *B;
  as<TACSParallelMat>(m)->getBCSRMat(&A, &B);
  return which ? B : A;
}

int ref_mat_get_sizes(ref_handle m, int which, int *bsize, int *nrows, int *ncols, int *nnzb) {
  const int *rowp, *cols;
  TacsScalar *vals;
  pick(m, which)->getArrays(bsize, nrows, ncols, &rowp, &cols, &vals);
  *nnzb = rowp[*nrows];
  return 0;
}

int ref_mat_get_pattern(ref_handle m, int which, int *rowp_out, int *cols_out) {
  int bs, nr, nc;
  const int *rowp, *cols;
  TacsScalar *vals;
  pick(m, which)->getArrays(&bs, &nr, &nc, &rowp, &cols, &vals);
  memcpy(rowp_out, rowp, (nr + 1) * sizeof(int));
  memcpy(cols_out, cols, rowp[nr] * sizeof(int));
  return 0;
}

int ref_mat_get_values(ref_handle m, int which, double *out) {
  int bs, nr, nc;
  const int *rowp, *cols;
  TacsScalar *vals;
  pick(m, which)->getArrays(&bs, &nr, &nc, &rowp, &cols, &vals);
  memcpy(out, vals, (size_t)bs * bs * rowp[nr] * sizeof(double));
  return 0;
}

/* the ascending global node ids of the external columns of Bext */
int ref_mat_get_ext_col_nodes(ref_handle m, int *nodes) {
  TACSBVecDistribute *dist;
  as<TACSParallelMat>(m)->getExtColMap(&dist);
  const int *idx;
  int n = dist->getIndices()->getIndices(&idx);
  if (nodes) memcpy(nodes, idx, n * sizeof(int));
  return n;
}

int ref_mat_zero_entries(ref_handle m) {
  as<TACSMat>(m)->zeroEntries();
  return 0;
}

int ref_mat_mult(ref_handle m, ref_handle x, ref_handle y) {
  as<TACSMat>(m)->mult(as<TACSBVec>(x), as<TACSBVec>(y));
  return 0;
}

int ref_mat_mult_transpose(ref_handle m, ref_handle x, ref_handle y) {
  as<TACSMat>(m)->multTranspose(as<TACSBVec>(x), as<TACSBVec>(y));
  return 0;
}

/* ---- auxiliary load elements -------------------------------------------------- */
ref_handle ref_aux_elements_create(void) { return keep(new TACSAuxElements()); }
int ref_aux_elements_add_shell_traction(ref_handle aux, int elem_num, int order, const double *t, int use_const) {
  TACSElement *e = NULL;
  if (order == 2) e = new TACSShellTraction<6, TACSQuadLinearQuadrature, TACSShellQuadBasis<2> >(t, use_const);
  else e = new TACSShellTraction<6, TACSQuadQuadraticQuadrature, TACSShellQuadBasis<3> >(t, use_const);
  as<TACSAuxElements>(aux)->addElement(elem_num, e);
  return 0;
}
int ref_aux_elements_add_shell_pressure(ref_handle aux, int elem_num, int order, const double *p, int use_const) {
  TACSElement *e = NULL;
  if (order == 2) {
    if (use_const) e = new TACSShellPressure<6, TACSQuadLinearQuadrature, TACSShellQuadBasis<2> >(p[0]);
    else e = new TACSShellPressure<6, TACSQuadLinearQuadrature, TACSShellQuadBasis<2> >(p);
  } else {
    if (use_const) e = new TACSShellPressure<6, TACSQuadQuadraticQuadrature, TACSShellQuadBasis<3> >(p[0]);
    else e = new TACSShellPressure<6, TACSQuadQuadraticQuadrature, TACSShellQuadBasis<3> >(p);
  }
  as<TACSAuxElements>(aux)->addElement(elem_num, e);
  return 0;
}
int ref_assembler_set_aux_elements(ref_handle a, ref_handle aux) {
  as<TACSAssembler>(a)->setAuxElements(aux ? as<TACSAuxElements>(aux) : NULL);
  return 0;
}

/* ---- Chebyshev smoother ---------------------------------------------------- */
ref_handle ref_chebyshev_create(ref_handle mat, int degree, double lower_factor, double upper_factor, int iters) {
  return keep(new TACSChebyshevSmoother(as<TACSMat>(mat), degree, lower_factor, upper_factor, iters));
}
int ref_chebyshev_factor(ref_handle pc) {
  as<TACSPc>(pc)->factor();
  return 0;
}
int ref_chebyshev_apply_factor(ref_handle pc, ref_handle x, ref_handle y) {
  as<TACSPc>(pc)->applyFactor(as<TACSBVec>(x), as<TACSBVec>(y));
  return 0;
}
double ref_chebyshev_get_spectral_radius(ref_handle) { return -1.0; /* private in the reference */ }

/* ---- GMRES (unpreconditioned or with the reference additive Schwarz PC) ----- */
ref_handle ref_gmres_create(ref_handle mat, int m, int nrestart) {
  return keep(new GMRES(as<TACSMat>(mat), m, nrestart));
}

ref_handle ref_gmres_create_pc(ref_handle mat, ref_handle pc, int m, int nrestart, int is_flexible) {
  return keep(new GMRES(as<TACSMat>(mat), as<TACSPc>(pc), m, nrestart, is_flexible));
}

int ref_gmres_set_tolerances(ref_handle k, double rtol, double atol) {
  as<TACSKsm>(k)->setTolerances(rtol, atol);
  return 0;
}

int ref_gmres_solve(ref_handle k, ref_handle b, ref_handle x, int zero_guess) {
  return as<TACSKsm>(k)->solve(as<TACSBVec>(b), as<TACSBVec>(x), zero_guess);
}

int ref_gmres_get_iter_count(ref_handle k) { return as<TACSKsm>(k)->getIterCount(); }
double ref_gmres_get_residual_norm(ref_handle k) { return as<TACSKsm>(k)->getResidualNorm(); }

/* ---- timing helpers for the CPU baseline (seconds, wall clock) ---------------- */
double ref_wtime(void) { return MPI_Wtime(); }

}  // extern "C"
