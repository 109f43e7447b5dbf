// extern "C" boundary of libtacs_b200.so (declared in include/tacs_b200.h).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/tacs_b200.h"
#include "tb2_host.h"

using namespace tb2;

namespace tb2 {
int comm_unique_id(unsigned char id[128]);
int comm_init(int rank, int size, const unsigned char id[128]);
}  // namespace tb2

template <class T>
static T *as(tacsb200_handle h) {
  T *p = h ? dynamic_cast<T *>(static_cast<Object *>(h)) : nullptr;
  return p;
}
static tacsb200_handle keep(Object *o) {
  if (o) o->incref();
  return static_cast<tacsb200_handle>(o);
}
#define REQUIRE(ptr, what)                                              \
  if (!(ptr)) {                                                         \
    fprintf(stderr, "tacs_b200: %s: invalid %s handle\n", __func__, what); \
    return 1;                                                           \
  }
#define REQUIRE_H(ptr, what)                                            \
  if (!(ptr)) {                                                         \
    fprintf(stderr, "tacs_b200: %s: invalid %s handle\n", __func__, what); \
    return nullptr;                                                     \
  }

template <class F>
static double timed(int reps, F fn) {
  if (ctx().device < 0) return -1.0;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaStreamSynchronize(ctx().stream.s);
  cudaEventRecord(e0, ctx().stream);
  int fail = 0;
  for (int i = 0; i < reps && !fail; i++) fail = fn();
  cudaEventRecord(e1, ctx().stream);
  cudaEventSynchronize(e1);
  float ms = 0.0f;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (fail || cudaGetLastError() != cudaSuccess) return -1.0;
  return (double)ms;
}

extern "C" {

int tacsb200_abi_version(void) { return 1; }

int tacsb200_init(int device) {
  if (device < 0) {
    const char *lr = getenv("LOCAL_RANK");
    device = lr ? atoi(lr) : 0;
  }
  return ctx_init(device);
}
int tacsb200_comm_unique_id(unsigned char id[128]) { return comm_unique_id(id); }
int tacsb200_comm_init(int rank, int size, const unsigned char id[128]) { return comm_init(rank, size, id); }
int tacsb200_comm_rank(void) { return ctx().rank; }
int tacsb200_comm_size(void) { return ctx().size; }
int tacsb200_synchronize(void) {
  if (ctx().device < 0) return 1;
  return cuda_ok(cudaStreamSynchronize(ctx().stream.s), "synchronize") ? 0 : 1;
}
long tacsb200_kernel_launches(int reset) {
  long n = ctx().kernel_launches;
  if (reset) ctx().kernel_launches = 0;
  return n;
}
void tacsb200_release(tacsb200_handle h) {
  if (h) static_cast<Object *>(h)->decref();
}

/* ---- constitutive ---------------------------------------------------------------------- */
tacsb200_handle tacsb200_material_properties_create(double rho, double cp, double E, double nu, double ys,
                                                    double alpha, double kappa) {
  return keep(new TACSMaterialProperties(rho, cp, E, nu, ys, alpha, kappa));
}
tacsb200_handle tacsb200_material_properties_create_ortho(double rho, double cp, double E1, double E2,
                                                          double E3, double nu12, double nu13, double nu23,
                                                          double G12, double G13, double G23) {
  return keep(new TACSMaterialProperties(rho, cp, E1, E2, E3, nu12, nu13, nu23, G12, G13, G23));
}
tacsb200_handle tacsb200_orthotropic_ply_create(double t, tacsb200_handle props) {
  TACSMaterialProperties *p = as<TACSMaterialProperties>(props);
  REQUIRE_H(p, "material properties");
  return keep(new TACSOrthotropicPly(t, p));
}
tacsb200_handle tacsb200_iso_shell_constitutive_create(tacsb200_handle props, double t, double tOffset,
                                                       double kcorr) {
  TACSMaterialProperties *p = as<TACSMaterialProperties>(props);
  REQUIRE_H(p, "material properties");
  return keep(new TACSIsoShellConstitutive(p, t, tOffset, kcorr));
}
tacsb200_handle tacsb200_composite_shell_constitutive_create(int n, tacsb200_handle *plies, const double *thick,
                                                             const double *angles, double kcorr,
                                                             double tOffset) {
  std::vector<TACSOrthotropicPly *> p(n);
  for (int i = 0; i < n; i++) {
    p[i] = as<TACSOrthotropicPly>(plies[i]);
    REQUIRE_H(p[i], "orthotropic ply");
  }
  return keep(new TACSCompositeShellConstitutive(n, p.data(), thick, angles, kcorr, tOffset));
}
tacsb200_handle tacsb200_solid_constitutive_create(tacsb200_handle props, double t) {
  TACSMaterialProperties *p = as<TACSMaterialProperties>(props);
  REQUIRE_H(p, "material properties");
  return keep(new TACSSolidConstitutive(p, t));
}
tacsb200_handle tacsb200_shell_constitutive_create_raw(const double *C22, const double *moments3) {
  if (!C22 || !moments3) return nullptr;
  return keep(new TACSRawShellConstitutive(C22, moments3));
}
tacsb200_handle tacsb200_solid_constitutive_create_raw(const double *C21, double density) {
  if (!C21) return nullptr;
  return keep(new TACSSolidConstitutive(C21, density));
}
void tacsb200_shell_set_drilling_regularization(double k) { TACSShellConstitutive::setDrillingRegularization(k); }
int tacsb200_constitutive_eval_tangent_stiffness(tacsb200_handle con, double *C) {
  TACSConstitutive *c = as<TACSConstitutive>(con);
  REQUIRE(c, "constitutive");
  c->evalTangentStiffness(C);
  return 0;
}
int tacsb200_shell_constitutive_eval_mass_moments(tacsb200_handle con, double *m) {
  TACSShellConstitutive *c = as<TACSShellConstitutive>(con);
  REQUIRE(c, "shell constitutive");
  c->evalMassMoments(m);
  return 0;
}

/* ---- transforms / elements -------------------------------------------------------------- */
tacsb200_handle tacsb200_shell_natural_transform_create(void) { return keep(new TACSShellNaturalTransform()); }
tacsb200_handle tacsb200_shell_ref_axis_transform_create(const double axis[3]) {
  return keep(new TACSShellRefAxisTransform(axis));
}
static tacsb200_handle make_shell(int order, tacsb200_handle transform, tacsb200_handle con) {
  TACSShellTransform *t = as<TACSShellTransform>(transform);
  TACSShellConstitutive *c = as<TACSShellConstitutive>(con);
  if (!t || !c) {
    fprintf(stderr, "tacs_b200: shell element needs a recognised transform and shell constitutive object\n");
    return nullptr;
  }
  return keep(new TACSShellElement(order, t, c));
}
tacsb200_handle tacsb200_quad4_shell_create(tacsb200_handle t, tacsb200_handle c) { return make_shell(2, t, c); }
tacsb200_handle tacsb200_quad9_shell_create(tacsb200_handle t, tacsb200_handle c) { return make_shell(3, t, c); }
tacsb200_handle tacsb200_linear_hexa_basis_create(void) { return keep(new TACSLinearHexaBasis()); }
tacsb200_handle tacsb200_quadratic_hexa_basis_create(void) { return keep(new TACSQuadraticHexaBasis()); }
tacsb200_handle tacsb200_linear_elasticity3d_create(tacsb200_handle con) {
  TACSSolidConstitutive *c = as<TACSSolidConstitutive>(con);
  REQUIRE_H(c, "solid constitutive");
  return keep(new TACSLinearElasticity3D(c));
}
tacsb200_handle tacsb200_element3d_create(tacsb200_handle model, tacsb200_handle basis) {
  TACSElementModel *m = as<TACSElementModel>(model);
  TACSElementBasis *b = as<TACSElementBasis>(basis);
  if (!m || !b || !dynamic_cast<TACSLinearElasticity3D *>(m)) {
    fprintf(stderr, "tacs_b200: TACSElement3D needs TACSLinearElasticity3D and a hexahedral basis\n");
    return nullptr;
  }
  return keep(new TACSElement3D(m, b));
}
int tacsb200_element_num_nodes(tacsb200_handle e) {
  TACSElement *el = as<TACSElement>(e);
  return el ? el->getNumNodes() : -1;
}
int tacsb200_element_vars_per_node(tacsb200_handle e) {
  TACSElement *el = as<TACSElement>(e);
  return el ? el->getVarsPerNode() : -1;
}
int tacsb200_element_add_jacobian(tacsb200_handle e, int count, double alpha, double beta, double gamma,
                                  const double *Xpts, const double *vars, const double *dvars,
                                  const double *ddvars, double *res, double *mat) {
  TACSElement *el = as<TACSElement>(e);
  REQUIRE(el, "element");
  return el->addJacobianBatch(count, alpha, beta, gamma, Xpts, vars, dvars, ddvars, res, mat);
}
int tacsb200_element_add_residual(tacsb200_handle e, int count, const double *Xpts, const double *vars,
                                  const double *dvars, const double *ddvars, double *res) {
  TACSElement *el = as<TACSElement>(e);
  REQUIRE(el, "element");
  return el->addJacobianBatch(count, 1.0, 0.0, 0.0, Xpts, vars, dvars, ddvars, res, nullptr);
}

/* ---- creator ------------------------------------------------------------------------------ */
tacsb200_handle tacsb200_creator_create(int vars_per_node) { return keep(new TACSCreator(vars_per_node)); }
int tacsb200_creator_set_global_connectivity(tacsb200_handle c, int nn, int ne, const int *ptr, const int *conn,
                                             const int *ids) {
  TACSCreator *cr = as<TACSCreator>(c);
  REQUIRE(cr, "creator");
  cr->setGlobalConnectivity(nn, ne, ptr, conn, ids);
  return 0;
}
int tacsb200_creator_set_boundary_conditions(tacsb200_handle c, int nb, const int *nodes, const int *ptr,
                                             const int *vars, const double *vals) {
  TACSCreator *cr = as<TACSCreator>(c);
  REQUIRE(cr, "creator");
  cr->setBoundaryConditions(nb, nodes, ptr, vars, vals);
  return 0;
}
int tacsb200_creator_set_nodes(tacsb200_handle c, const double *X) {
  TACSCreator *cr = as<TACSCreator>(c);
  REQUIRE(cr, "creator");
  cr->setNodes(X);
  return 0;
}
int tacsb200_creator_set_elements(tacsb200_handle c, int n, tacsb200_handle *elems) {
  TACSCreator *cr = as<TACSCreator>(c);
  REQUIRE(cr, "creator");
  std::vector<TACSElement *> e(n);
  for (int i = 0; i < n; i++) {
    e[i] = as<TACSElement>(elems[i]);
    REQUIRE(e[i], "element");
  }
  cr->setElements(n, e.data());
  return 0;
}
int tacsb200_creator_set_keep_numbering(tacsb200_handle c, int keep_numbering) {
  TACSCreator *cr = as<TACSCreator>(c);
  REQUIRE(cr, "creator");
  if (keep_numbering && cr->comm_size() > 1) {
    fprintf(stderr, "tacs_b200: keep_numbering adopts the numbering of a single-rank assembler only\n");
    return 1;
  }
  cr->keep_numbering = keep_numbering != 0;
  return 0;
}
int tacsb200_creator_partition_mesh(tacsb200_handle c, int split, const int *part) {
  TACSCreator *cr = as<TACSCreator>(c);
  REQUIRE(cr, "creator");
  return cr->partitionMesh(split, part);
}
int tacsb200_creator_get_node_nums(tacsb200_handle c, int *out) {
  TACSCreator *cr = as<TACSCreator>(c);
  if (!cr) return 0;
  const int *nn = nullptr;
  int n = cr->getNodeNums(&nn);
  if (out && nn) memcpy(out, nn, n * sizeof(int));
  return n;
}
int tacsb200_creator_get_element_partition(tacsb200_handle c, int *out) {
  TACSCreator *cr = as<TACSCreator>(c);
  if (!cr) return 0;
  const int *p = nullptr;
  int n = cr->getElementPartition(&p);
  if (out && p) memcpy(out, p, n * sizeof(int));
  return n;
}
tacsb200_handle tacsb200_creator_create_tacs(tacsb200_handle c) {
  TACSCreator *cr = as<TACSCreator>(c);
  REQUIRE_H(cr, "creator");
  return keep(cr->createTACS());
}

/* ---- host-only plan -------------------------------------------------------------------------- */
tacsb200_handle tacsb200_creator_create_plan(tacsb200_handle c, int rank, int size) {
  TACSCreator *cr = as<TACSCreator>(c);
  REQUIRE_H(cr, "creator");
  return keep(cr->createPlan(rank, size));
}
int tacsb200_plan_get_array(tacsb200_handle plan, const char *name, int *out) {
  PlanObject *po = as<PlanObject>(plan);
  if (!po) return -1;
  HostPlan &P = po->plan;
  std::vector<int> tmp;
  const std::vector<int> *v = nullptr;
  std::string n(name);
  auto ex = [&](const char *prefix, ExchangePlan &x) -> const std::vector<int> * {
    std::string p(prefix);
    if (n == p + "_send_peers") return &x.send_peers;
    if (n == p + "_send_ptr") return &x.send_ptr;
    if (n == p + "_send_idx") return &x.send_idx;
    if (n == p + "_recv_peers") return &x.recv_peers;
    if (n == p + "_recv_ptr") return &x.recv_ptr;
    return nullptr;
  };
  if (n == "elem_global") v = &P.elem_global;
  else if (n == "elem_ptr") v = &P.elem_ptr;
  else if (n == "elem_conn_global") v = &P.elem_conn_global;
  else if (n == "elem_conn_local") v = &P.elem_conn_local;
  else if (n == "ext_nodes") v = &P.ext_nodes;
  else if (n == "owner_range") v = &P.owner_range;
  else if (n == "Aloc_rowp") v = &P.Aloc.rowp;
  else if (n == "Aloc_cols") v = &P.Aloc.cols;
  else if (n == "Bext_rowp") v = &P.Bext.rowp;
  else if (n == "Bext_cols") v = &P.Bext.cols;
  else if (n == "ext_col_nodes") v = &P.ext_col_nodes;
  else if (n == "dmap") v = &P.dmap;
  else if (n == "gb_blk") v = &P.gb_blk;
  else if (n == "gb_ptr") v = &P.gb_ptr;
  else if (n == "gb_src") v = &P.gb_src;
  else if (n == "elem_block_base") { tmp.assign(P.elem_block_base.begin(), P.elem_block_base.end()); v = &tmp; }
  else if (n == "elem_pair_base") { tmp.assign(P.elem_pair_base.begin(), P.elem_pair_base.end()); v = &tmp; }
  else if (n == "r_ptr") v = &P.r_ptr;
  else if (n == "r_src") v = &P.r_src;
  else if (n == "scalars") {
    tmp = {P.nelems, P.nowned, P.nlocal, P.ext_before, P.ext_after, P.np, (int)P.local_blocks, (int)P.recv_blocks,
           (int)P.local_node_slots, (int)P.recv_node_slots, (int)P.direct_blocks,
           (int)P.gatherEnd(P.local_blocks)};  // gathered blocks below this index read no received staging slot
    v = &tmp;
  } else if (!(v = ex("state", P.state)) && !(v = ex("cols", P.cols)) && !(v = ex("rows", P.rows)) &&
             !(v = ex("blocks", P.blocks))) {
    fprintf(stderr, "tacs_b200: unknown plan array '%s'\n", name);
    return -1;
  }
  if (out && !v->empty()) memcpy(out, v->data(), v->size() * sizeof(int));
  return (int)v->size();
}

/* ---- assembler ------------------------------------------------------------------------------ */
#define ASM(a)                           \
  TACSAssembler *t = as<TACSAssembler>(a); \
  REQUIRE(t, "assembler")
int tacsb200_assembler_get_vars_per_node(tacsb200_handle a) { ASM(a); return t->getVarsPerNode(); }
int tacsb200_assembler_get_num_nodes(tacsb200_handle a) { ASM(a); return t->getNumNodes(); }
int tacsb200_assembler_get_num_owned_nodes(tacsb200_handle a) { ASM(a); return t->getNumOwnedNodes(); }
int tacsb200_assembler_get_num_elements(tacsb200_handle a) { ASM(a); return t->getNumElements(); }
/* plan statistics behind the roofline byte counts: {staging slots (local), staging slots (received), blocks written
   directly by the element kernels, upper node-pair blocks staged by the element kernels, gathered blocks, gather
   sources, total blocks [Aloc | Bext], local node pairs}; valid after the first createMat */
int tacsb200_assembler_get_plan_stats(tacsb200_handle a, long *out) {
  ASM(a);
  HostPlan &P = *t->plan;
  out[0] = P.local_blocks;
  out[1] = P.recv_blocks;
  out[2] = P.direct_blocks;
  out[3] = P.staged_blocks;
  out[4] = (long)P.gb_blk.size();
  out[5] = P.gb_ptr.empty() ? 0 : (long)P.gb_ptr.back();
  out[6] = P.Aloc.nnzb() + P.Bext.nnzb();
  out[7] = P.local_pairs;
  return P.has_matrix ? 0 : 1;
}
int tacsb200_assembler_get_owner_range(tacsb200_handle a, int *lo, int *hi) {
  ASM(a);
  *lo = t->owner_range[t->rank];
  *hi = t->owner_range[t->rank + 1];
  return 0;
}
int tacsb200_assembler_get_element_connectivity(tacsb200_handle a, int *ptr, int *conn) {
  TACSAssembler *t = as<TACSAssembler>(a);
  if (!t) return -1;
  if (ptr) memcpy(ptr, t->plan->elem_ptr.data(), t->plan->elem_ptr.size() * sizeof(int));
  if (conn) memcpy(conn, t->plan->elem_conn_global.data(), t->plan->elem_conn_global.size() * sizeof(int));
  return (int)t->plan->elem_conn_global.size();
}
int tacsb200_assembler_get_local_to_global(tacsb200_handle a, int *global) {
  TACSAssembler *t = as<TACSAssembler>(a);
  if (!t) return -1;
  for (int l = 0; l < t->nlocal; l++) global[l] = t->plan->globalNode(l);
  return t->nlocal;
}
// a vector whose device allocation failed is not handed out (the message was printed by cudaMalloc's check)
static tacsb200_handle keep_vec(TACSBVec *v) {
  if (v && v->localSize() > 0 && !v->data.ptr) {
    delete v;
    return nullptr;
  }
  return keep(v);
}
tacsb200_handle tacsb200_assembler_create_vec(tacsb200_handle a) {
  TACSAssembler *t = as<TACSAssembler>(a);
  REQUIRE_H(t, "assembler");
  return keep_vec(t->createVec());
}
tacsb200_handle tacsb200_assembler_create_node_vec(tacsb200_handle a) {
  TACSAssembler *t = as<TACSAssembler>(a);
  REQUIRE_H(t, "assembler");
  return keep_vec(t->createNodeVec());
}
tacsb200_handle tacsb200_assembler_create_mat(tacsb200_handle a) {
  TACSAssembler *t = as<TACSAssembler>(a);
  REQUIRE_H(t, "assembler");
  return keep(t->createMat());
}
int tacsb200_assembler_get_nodes(tacsb200_handle a, tacsb200_handle X) {
  ASM(a);
  TACSBVec *v = as<TACSBVec>(X);
  REQUIRE(v, "vector");
  t->getNodes(v);
  return 0;
}
int tacsb200_assembler_set_nodes(tacsb200_handle a, tacsb200_handle X) {
  ASM(a);
  TACSBVec *v = as<TACSBVec>(X);
  REQUIRE(v, "vector");
  return t->setNodes(v);
}
int tacsb200_assembler_set_variables(tacsb200_handle a, tacsb200_handle q, tacsb200_handle qd, tacsb200_handle qdd) {
  ASM(a);
  return t->setVariables(as<TACSBVec>(q), as<TACSBVec>(qd), as<TACSBVec>(qdd));
}
int tacsb200_assembler_zero_variables(tacsb200_handle a) { ASM(a); t->zeroVariables(); return 0; }
int tacsb200_assembler_apply_bcs_vec(tacsb200_handle a, tacsb200_handle v) {
  ASM(a);
  TACSBVec *x = as<TACSBVec>(v);
  REQUIRE(x, "vector");
  t->applyBCs(x);
  return 0;
}
int tacsb200_assembler_apply_bcs_mat(tacsb200_handle a, tacsb200_handle m) {
  ASM(a);
  TACSParallelMat *A = as<TACSParallelMat>(m);
  REQUIRE(A, "matrix");
  t->applyBCs(A);
  return 0;
}
int tacsb200_assembler_set_bcs(tacsb200_handle a, tacsb200_handle v) {
  ASM(a);
  TACSBVec *x = as<TACSBVec>(v);
  REQUIRE(x, "vector");
  t->setBCs(x);
  return 0;
}
int tacsb200_assembler_set_num_threads(tacsb200_handle a, int n) { ASM(a); (void)n; return 0; }
int tacsb200_assembler_assemble_res(tacsb200_handle a, tacsb200_handle res) {
  ASM(a);
  TACSBVec *r = as<TACSBVec>(res);
  REQUIRE(r, "vector");
  if (t->assembleRes(r, 1.0)) return 1;
  return tacsb200_synchronize();
}
int tacsb200_assembler_assemble_jacobian(tacsb200_handle a, double alpha, double beta, double gamma,
                                         tacsb200_handle res, tacsb200_handle mat) {
  ASM(a);
  TACSParallelMat *A = as<TACSParallelMat>(mat);
  REQUIRE(A, "matrix");
  if (t->assembleJacobian(alpha, beta, gamma, as<TACSBVec>(res), A, 1.0)) return 1;
  return tacsb200_synchronize();
}
int tacsb200_assembler_assemble_jacobian_host(tacsb200_handle a, double alpha, double beta, double gamma,
                                              const double *q_host, double *res_host, tacsb200_handle mat) {
  ASM(a);
  TACSParallelMat *A = as<TACSParallelMat>(mat);
  REQUIRE(A, "matrix");
  if (!q_host || !res_host) return 1;
  return t->assembleJacobianHost(alpha, beta, gamma, q_host, res_host, A, 1.0) ? 1 : 0;
}
int tacsb200_assembler_assemble_jacobian_async(tacsb200_handle a, double alpha, double beta, double gamma,
                                               tacsb200_handle res, tacsb200_handle mat) {
  ASM(a);
  TACSParallelMat *A = as<TACSParallelMat>(mat);
  REQUIRE(A, "matrix");
  return t->assembleJacobian(alpha, beta, gamma, as<TACSBVec>(res), A, 1.0) ? 1 : 0;
}
int tacsb200_assembler_assemble_mat_type(tacsb200_handle a, int mat_type, tacsb200_handle mat, int apply_bcs) {
  ASM(a);
  TACSParallelMat *A = as<TACSParallelMat>(mat);
  REQUIRE(A, "matrix");
  if (t->assembleMatType(mat_type, A, apply_bcs != 0)) return 1;
  return tacsb200_synchronize();
}
int tacsb200_assembler_add_jacobian_vec_product(tacsb200_handle a, double scale, double alpha, double beta,
                                                double gamma, tacsb200_handle x, tacsb200_handle y, int apply_bcs) {
  ASM(a);
  TACSBVec *xv = as<TACSBVec>(x), *yv = as<TACSBVec>(y);
  REQUIRE(xv && yv, "vector");
  if (t->addJacobianVecProduct(scale, alpha, beta, gamma, xv, yv, apply_bcs != 0)) return 1;
  return tacsb200_synchronize();
}

/* ---- vectors ---------------------------------------------------------------------------------- */
#define VEC(v, name)              \
  TACSBVec *name = as<TACSBVec>(v); \
  REQUIRE(name, "vector")
int tacsb200_vec_get_size(tacsb200_handle v) { VEC(v, x); return (int)x->ownedSize(); }
int tacsb200_vec_get_array(tacsb200_handle v, double *out) { VEC(v, x); return x->getArray(out); }
int tacsb200_vec_set_array(tacsb200_handle v, const double *in) { VEC(v, x); return x->setArray(in); }
double *tacsb200_vec_device_ptr(tacsb200_handle v) {
  TACSBVec *x = as<TACSBVec>(v);
  return x ? x->owned() : nullptr;
}
double tacsb200_vec_norm(tacsb200_handle v) {
  TACSBVec *x = as<TACSBVec>(v);
  return x ? x->norm() : -1.0;
}
double tacsb200_vec_dot(tacsb200_handle a, tacsb200_handle b) {
  TACSBVec *x = as<TACSBVec>(a), *y = as<TACSBVec>(b);
  return (x && y) ? x->dot(y) : 0.0;
}
int tacsb200_vec_mdot(tacsb200_handle v, int n, tacsb200_handle *ys, double *out) {
  VEC(v, x);
  std::vector<TACSBVec *> y(n);
  for (int i = 0; i < n; i++) {
    y[i] = as<TACSBVec>(ys[i]);
    REQUIRE(y[i], "vector");
  }
  return x->mdot(y.data(), out, n);
}
int tacsb200_vec_axpy(tacsb200_handle yv, double alpha, tacsb200_handle xv) {
  VEC(yv, y); VEC(xv, x);
  y->axpy(alpha, x);
  return 0;
}
int tacsb200_vec_axpby(tacsb200_handle yv, double alpha, double beta, tacsb200_handle xv) {
  VEC(yv, y); VEC(xv, x);
  y->axpby(alpha, beta, x);
  return 0;
}
int tacsb200_vec_scale(tacsb200_handle yv, double alpha) { VEC(yv, y); y->scale(alpha); return 0; }
int tacsb200_vec_copy_values(tacsb200_handle yv, tacsb200_handle xv) {
  VEC(yv, y); VEC(xv, x);
  y->copyValues(x);
  return 0;
}
int tacsb200_vec_zero_entries(tacsb200_handle yv) { VEC(yv, y); y->zeroEntries(); return 0; }

/* ---- matrix ------------------------------------------------------------------------------------ */
#define MAT(m)                                  \
  TACSParallelMat *A = as<TACSParallelMat>(m); \
  REQUIRE(A, "matrix")
int tacsb200_mat_get_sizes(tacsb200_handle m, int which, int *bsize, int *nrows, int *ncols, int *nnzb) {
  MAT(m);
  BCSRPattern &P = which ? A->Bext : A->Aloc;
  *bsize = P.bsize; *nrows = P.nrows; *ncols = P.ncols; *nnzb = (int)P.nnzb();
  return 0;
}
int tacsb200_mat_get_pattern(tacsb200_handle m, int which, int *rowp, int *cols) {
  MAT(m);
  BCSRPattern &P = which ? A->Bext : A->Aloc;
  memcpy(rowp, P.rowp.data(), P.rowp.size() * sizeof(int));
  memcpy(cols, P.cols.data(), P.cols.size() * sizeof(int));
  return 0;
}
int tacsb200_mat_get_values(tacsb200_handle m, int which, double *out) {
  MAT(m);
  BCSRPattern &P = which ? A->Bext : A->Aloc;
  return P.d_vals.download(out, P.d_vals.count) ? 0 : 1;
}
double *tacsb200_mat_device_values(tacsb200_handle m, int which) {
  TACSParallelMat *A = as<TACSParallelMat>(m);
  if (!A) return nullptr;
  return which ? A->Bext.d_vals.ptr : A->Aloc.d_vals.ptr;
}
int tacsb200_mat_get_ext_col_nodes(tacsb200_handle m, int *nodes) {
  TACSParallelMat *A = as<TACSParallelMat>(m);
  if (!A) return -1;
  if (nodes) memcpy(nodes, A->ext_col_nodes.data(), A->ext_col_nodes.size() * sizeof(int));
  return (int)A->ext_col_nodes.size();
}
int tacsb200_mat_zero_entries(tacsb200_handle m) { MAT(m); A->zeroEntries(); return 0; }
int tacsb200_mat_copy_values(tacsb200_handle m, tacsb200_handle other) {
  MAT(m);
  return A->copyValues(as<TACSParallelMat>(other));
}
int tacsb200_mat_scale(tacsb200_handle m, double alpha) { MAT(m); return A->scale(alpha); }
int tacsb200_mat_axpy(tacsb200_handle m, double alpha, tacsb200_handle other) {
  MAT(m);
  return A->axpy(alpha, as<TACSParallelMat>(other));
}
int tacsb200_mat_mult_async(tacsb200_handle m, tacsb200_handle xv, tacsb200_handle yv) {
  MAT(m);
  VEC(xv, x); VEC(yv, y);
  return A->mult(x, y);
}
int tacsb200_mat_mult_transpose(tacsb200_handle m, tacsb200_handle xv, tacsb200_handle yv) {
  MAT(m);
  VEC(xv, x); VEC(yv, y);
  if (A->multTranspose(x, y)) return 1;
  return tacsb200_synchronize();
}
int tacsb200_mat_mult(tacsb200_handle m, tacsb200_handle xv, tacsb200_handle yv) {
  if (tacsb200_mat_mult_async(m, xv, yv)) return 1;
  return tacsb200_synchronize();
}
tacsb200_handle tacsb200_mat_create_vec(tacsb200_handle m) {
  TACSParallelMat *A = as<TACSParallelMat>(m);
  REQUIRE_H(A, "matrix");
  return keep_vec(A->createVec());
}

/* ---- auxiliary load elements ------------------------------------------------------------------------ */
tacsb200_handle tacsb200_aux_elements_create(void) { return keep(new TACSAuxElements()); }
int tacsb200_aux_elements_add_shell_traction(tacsb200_handle aux, int elem_num, int order, const double *t,
                                             int use_const_trac) {
  TACSAuxElements *a = as<TACSAuxElements>(aux);
  REQUIRE(a, "auxiliary elements");
  if ((order != 2 && order != 3) || !t) return 1;
  a->addShellTraction(elem_num, order, t, use_const_trac != 0);
  return 0;
}
int tacsb200_aux_elements_add_shell_pressure(tacsb200_handle aux, int elem_num, int order, const double *p,
                                             int use_const_pressure) {
  TACSAuxElements *a = as<TACSAuxElements>(aux);
  REQUIRE(a, "auxiliary elements");
  if ((order != 2 && order != 3) || !p) return 1;
  a->addShellPressure(elem_num, order, p, use_const_pressure != 0);
  return 0;
}
int tacsb200_assembler_set_aux_elements(tacsb200_handle asmb, tacsb200_handle aux) {
  ASM(asmb);
  return t->setAuxElements(as<TACSAuxElements>(aux));
}

/* ---- TACSSchurMat view ------------------------------------------------------------------------------- */
tacsb200_handle tacsb200_schur_mat_create(tacsb200_handle mat, int nb, const int *b_nodes, int nc, const int *c_nodes,
                                          const int *Browp, const int *Bcols, const int *Erowp, const int *Ecols,
                                          const int *Frowp, const int *Fcols, const int *Crowp, const int *Ccols) {
  TACSParallelMat *A = as<TACSParallelMat>(mat);
  REQUIRE_H(A, "matrix");
  const int *rowp[4] = {Browp, Erowp, Frowp, Crowp}, *cols[4] = {Bcols, Ecols, Fcols, Ccols};
  for (int k = 0; k < 4; k++)
    if (!rowp[k]) {
      fprintf(stderr, "tacs_b200: schur_mat_create: row pointers of all four blocks are required\n");
      return nullptr;
    }
  TACSSchurMat *S = new TACSSchurMat(A, nb, b_nodes, nc, c_nodes, rowp, cols);
  if (!S->ok) {
    S->incref();
    S->decref();
    return nullptr;
  }
  return keep(S);
}
int tacsb200_schur_mat_update(tacsb200_handle s) {
  TACSSchurMat *S = as<TACSSchurMat>(s);
  REQUIRE(S, "Schur matrix");
  return S->update();
}
int tacsb200_schur_mat_get_values(tacsb200_handle s, int which, double *vals) {
  TACSSchurMat *S = as<TACSSchurMat>(s);
  REQUIRE(S, "Schur matrix");
  return S->getValues(which, vals);
}
int tacsb200_schur_mat_mult(tacsb200_handle s, tacsb200_handle xv, tacsb200_handle yv) {
  TACSSchurMat *S = as<TACSSchurMat>(s);
  REQUIRE(S, "Schur matrix");
  VEC(xv, x); VEC(yv, y);
  if (S->mult(x, y)) return 1;
  return tacsb200_synchronize();
}

/* ---- GMRES --------------------------------------------------------------------------------------- */
tacsb200_handle tacsb200_gmres_create(tacsb200_handle mat, int m, int nrestart) {
  TACSParallelMat *A = as<TACSParallelMat>(mat);
  REQUIRE_H(A, "matrix");
  return keep(new GMRES(A, m, nrestart));
}
tacsb200_handle tacsb200_gmres_create_pc(tacsb200_handle mat, tacsb200_handle pc, int m, int nrestart,
                                         int is_flexible) {
  TACSParallelMat *A = as<TACSParallelMat>(mat);
  TACSChebyshevSmoother *P = as<TACSChebyshevSmoother>(pc);
  REQUIRE_H(A && P, "matrix / preconditioner");
  return keep(new GMRES(A, m, nrestart, P, is_flexible != 0));
}
tacsb200_handle tacsb200_chebyshev_create(tacsb200_handle mat, int degree, double lower_factor,
                                          double upper_factor, int iters) {
  TACSParallelMat *A = as<TACSParallelMat>(mat);
  REQUIRE_H(A, "matrix");
  return keep(new TACSChebyshevSmoother(A, degree, lower_factor, upper_factor, iters));
}
int tacsb200_chebyshev_factor(tacsb200_handle pc) {
  TACSChebyshevSmoother *P = as<TACSChebyshevSmoother>(pc);
  REQUIRE(P, "Chebyshev smoother");
  return P->factor();
}
int tacsb200_chebyshev_apply_factor(tacsb200_handle pc, tacsb200_handle x, tacsb200_handle y) {
  TACSChebyshevSmoother *P = as<TACSChebyshevSmoother>(pc);
  TACSBVec *xv = as<TACSBVec>(x), *yv = as<TACSBVec>(y);
  REQUIRE(P && xv && yv, "Chebyshev smoother / vector");
  if (P->applyFactor(xv, yv)) return 1;
  return tacsb200_synchronize();
}
double tacsb200_chebyshev_get_spectral_radius(tacsb200_handle pc) {
  TACSChebyshevSmoother *P = as<TACSChebyshevSmoother>(pc);
  return P ? P->rho : -1.0;
}
int tacsb200_gmres_set_tolerances(tacsb200_handle k, double rtol, double atol) {
  GMRES *g = as<GMRES>(k);
  REQUIRE(g, "GMRES");
  g->setTolerances(rtol, atol);
  return 0;
}
int tacsb200_gmres_solve(tacsb200_handle k, tacsb200_handle b, tacsb200_handle x, int zero_guess) {
  GMRES *g = as<GMRES>(k);
  TACSBVec *bv = as<TACSBVec>(b), *xv = as<TACSBVec>(x);
  if (!g || !bv || !xv) return -1;
  return g->solve(bv, xv, zero_guess);
}
int tacsb200_gmres_set_ortho_type(tacsb200_handle k, int classical) {
  GMRES *g = as<GMRES>(k);
  REQUIRE(g, "GMRES");
  g->setOrthoType(classical ? GMRES::CLASSICAL_GRAM_SCHMIDT : GMRES::MODIFIED_GRAM_SCHMIDT);
  return 0;
}
int tacsb200_gmres_set_monitor(tacsb200_handle k, const char *descript, int freq) {
  GMRES *g = as<GMRES>(k);
  REQUIRE(g, "GMRES");
  g->setMonitor(descript, freq);
  return 0;
}
int tacsb200_gmres_set_time_monitor(tacsb200_handle k) {
  GMRES *g = as<GMRES>(k);
  REQUIRE(g, "GMRES");
  g->setTimeMonitor();
  return 0;
}
int tacsb200_gmres_get_iter_count(tacsb200_handle k) {
  GMRES *g = as<GMRES>(k);
  return g ? g->getIterCount() : -1;
}
double tacsb200_gmres_get_residual_norm(tacsb200_handle k) {
  GMRES *g = as<GMRES>(k);
  return g ? g->getResidualNorm() : -1.0;
}

/* ---- device-timed helpers --------------------------------------------------------------------------- */
double tacsb200_time_assemble_jacobian(tacsb200_handle a, double alpha, double beta, double gamma,
                                       tacsb200_handle res, tacsb200_handle mat, int reps) {
  TACSAssembler *t = as<TACSAssembler>(a);
  TACSParallelMat *A = as<TACSParallelMat>(mat);
  TACSBVec *r = as<TACSBVec>(res);
  if (!t || !A) return -1.0;
  return timed(reps, [&]() { return t->assembleJacobian(alpha, beta, gamma, r, A, 1.0); });
}
double tacsb200_time_assemble_res(tacsb200_handle a, tacsb200_handle res, int reps) {
  TACSAssembler *t = as<TACSAssembler>(a);
  TACSBVec *r = as<TACSBVec>(res);
  if (!t || !r) return -1.0;
  return timed(reps, [&]() { return t->assembleRes(r, 1.0); });
}
double tacsb200_time_mat_mult(tacsb200_handle m, tacsb200_handle xv, tacsb200_handle yv, int reps) {
  TACSParallelMat *A = as<TACSParallelMat>(m);
  TACSBVec *x = as<TACSBVec>(xv), *y = as<TACSBVec>(yv);
  if (!A || !x || !y) return -1.0;
  return timed(reps, [&]() { return A->mult(x, y); });
}

int tacsb200_profile_enable(int on) {
  profile_enable(on);
  return 0;
}
int tacsb200_profile_collect(double *ms, long *count) { return profile_collect(ms, count); }
const char *tacsb200_profile_named(void) { return profile_named(); }

double tacsb200_measure_fp64_tflops(void) {
  if (ctx_init(-1)) return -1.0;
  DeviceArray<double> out;
  if (!out.alloc(8)) return -1.0;
  const int iters = 20000, blocks = ctx().num_sms * 8;
  launch_dfma_peak(out.ptr, 200, blocks, ctx().stream);
  double best = 0.0;
  for (int rep = 0; rep < 3; rep++) {
    double ms = timed(1, [&]() { return launch_dfma_peak(out.ptr, iters, blocks, ctx().stream) != cudaSuccess; });
    if (ms <= 0.0) return -1.0;
    double tf = 2.0 * 16.0 * iters * 256.0 * blocks / (ms * 1e-3) * 1e-12;
    if (tf > best) best = tf;
  }
  return best;
}

double tacsb200_measure_copy_gbs(void) {
  if (ctx_init(-1)) return -1.0;
  const long n = 1L << 28;  // 2 GiB per buffer, far larger than L2
  DeviceArray<double> a, b;
  if (!a.alloc(n) || !b.alloc(n)) return -1.0;
  cudaMemsetAsync(a.ptr, 0, n * sizeof(double), ctx().stream);
  launch_copy(n, a.ptr, b.ptr, ctx().num_sms, ctx().stream);
  double best = 0.0;
  for (int rep = 0; rep < 5; rep++) {
    double ms = timed(1, [&]() { return launch_copy(n, a.ptr, b.ptr, ctx().num_sms, ctx().stream) != cudaSuccess; });
    if (ms <= 0.0) return -1.0;
    double gbs = 2.0 * n * sizeof(double) / (ms * 1e-3) * 1e-9;
    if (gbs > best) best = gbs;
  }
  return best;
}

}  // extern "C"
