// Reference-side binding of tacs_b200 (see tacs_b200_shim.h). Compiled against the unmodified reference headers;
// everything that computes is behind the C ABI of include/tacs_b200.h.
#include "tacs_b200_shim.h"

#include <stdio.h>
#include <string.h>

#include <map>
#include <vector>

#include "TACSElement3D.h"
#include "TACSHexaBasis.h"
#include "TACSLinearElasticity.h"
#include "TACSAuxElements.h"
#include "TACSShellElementDefs.h"
#include "TACSShellPressure.h"
#include "TACSShellTraction.h"
#include "TACSSolidConstitutive.h"

// ---------------------------------------------------------------------------------------------------------------
// The shell element keeps its transform and constitutive objects in private members without accessors
// (src/elements/shell/TACSShellElement.h:181-188). An explicit template instantiation may name a private member
// (access checking does not apply to its arguments), which gives the shim a read-only path to the two pointers
// without touching the reference headers.
// ---------------------------------------------------------------------------------------------------------------
namespace {
template <class Tag, typename Tag::type Member>
struct PrivateMember {
  friend typename Tag::type get(Tag) { return Member; }
};
#define TB2_PRIVATE_MEMBER(TAG, CLASS, TYPE, NAME) \
  struct TAG {                                     \
    typedef TYPE CLASS::*type;                     \
    friend type get(TAG);                          \
  };                                               \
  template struct PrivateMember<TAG, &CLASS::NAME>;
TB2_PRIVATE_MEMBER(Quad4Con, TACSQuad4Shell, TACSShellConstitutive *, con)
TB2_PRIVATE_MEMBER(Quad4Transform, TACSQuad4Shell, TACSShellTransform *, transform)
TB2_PRIVATE_MEMBER(Quad9Con, TACSQuad9Shell, TACSShellConstitutive *, con)
TB2_PRIVATE_MEMBER(Quad9Transform, TACSQuad9Shell, TACSShellTransform *, transform)
TB2_PRIVATE_MEMBER(ElasticityStrainType, TACSLinearElasticity3D, ElementStrainType, strain_type)

// the nodal traction / pressure arrays of the shell load elements are private as well
typedef TACSShellTraction<6, TACSQuadLinearQuadrature, TACSShellQuadBasis<2> > Quad4TractionElem;
typedef TACSShellTraction<6, TACSQuadQuadraticQuadrature, TACSShellQuadBasis<3> > Quad9TractionElem;
typedef TACSShellPressure<6, TACSQuadLinearQuadrature, TACSShellQuadBasis<2> > Quad4PressureElem;
typedef TACSShellPressure<6, TACSQuadQuadraticQuadrature, TACSShellQuadBasis<3> > Quad9PressureElem;
typedef TacsScalar TractionArray4[12];
typedef TacsScalar TractionArray9[27];
typedef TacsScalar PressureArray4[4];
typedef TacsScalar PressureArray9[9];
TB2_PRIVATE_MEMBER(Quad4TractionT, Quad4TractionElem, TractionArray4, t)
TB2_PRIVATE_MEMBER(Quad9TractionT, Quad9TractionElem, TractionArray9, t)
TB2_PRIVATE_MEMBER(Quad4PressureP, Quad4PressureElem, PressureArray4, p)
TB2_PRIVATE_MEMBER(Quad9PressureP, Quad9PressureElem, PressureArray9, p)

// device-side copy of the application's auxiliary elements, or NULL when one of them is not a shell load of the path
tacsb200_handle convert_aux_elements(TACSAuxElements *aux) {
  TACSAuxElem *list = NULL;
  const int naux = aux->getAuxElements(&list);
  tacsb200_handle out = tacsb200_aux_elements_create();
  for (int k = 0; k < naux && out; k++) {
    TACSElement *e = list[k].elem;
    int rc = 1;
    if (Quad4TractionElem *q = dynamic_cast<Quad4TractionElem *>(e))
      rc = tacsb200_aux_elements_add_shell_traction(out, list[k].num, 2, q->*get(Quad4TractionT()), 0);
    else if (Quad9TractionElem *q = dynamic_cast<Quad9TractionElem *>(e))
      rc = tacsb200_aux_elements_add_shell_traction(out, list[k].num, 3, q->*get(Quad9TractionT()), 0);
    else if (Quad4PressureElem *q = dynamic_cast<Quad4PressureElem *>(e))
      rc = tacsb200_aux_elements_add_shell_pressure(out, list[k].num, 2, q->*get(Quad4PressureP()), 0);
    else if (Quad9PressureElem *q = dynamic_cast<Quad9PressureElem *>(e))
      rc = tacsb200_aux_elements_add_shell_pressure(out, list[k].num, 3, q->*get(Quad9PressureP()), 0);
    if (rc) {
      tacsb200_release(out);
      out = NULL;
    }
  }
  return out;
}

void fail(const char *what) { fprintf(stderr, "TACSB200Assembler: %s; the reference path stays in charge\n", what); }

// device-side element object for one element object of the application, or NULL
tacsb200_handle convert_shell(int order, TACSShellTransform *transform, TACSShellConstitutive *con) {
  // the tangent stiffness and the mass moments of the classes on the path do not depend on the point
  const double pt[3] = {0.0, 0.0, 0.0};
  const TacsScalar X[3] = {0.0, 0.0, 0.0};
  TacsScalar C[TACSShellConstitutive::NUM_TANGENT_STIFFNESS_ENTRIES], moments[3];
  con->evalTangentStiffness(0, pt, X, C);
  con->evalMassMoments(0, pt, X, moments);
  tacsb200_handle dcon = tacsb200_shell_constitutive_create_raw(C, moments);
  tacsb200_handle dtr = NULL;
  if (TACSShellRefAxisTransform *ra = dynamic_cast<TACSShellRefAxisTransform *>(transform)) {
    TacsScalar axis[3];
    ra->getRefAxis(axis);
    dtr = tacsb200_shell_ref_axis_transform_create(axis);
  } else if (dynamic_cast<TACSShellNaturalTransform *>(transform)) {
    dtr = tacsb200_shell_natural_transform_create();
  }
  tacsb200_handle elem = NULL;
  if (dcon && dtr) elem = order == 2 ? tacsb200_quad4_shell_create(dtr, dcon) : tacsb200_quad9_shell_create(dtr, dcon);
  if (dcon) tacsb200_release(dcon);
  if (dtr) tacsb200_release(dtr);
  return elem;
}

tacsb200_handle convert_element(TACSElement *e) {
  if (TACSQuad4Shell *s = dynamic_cast<TACSQuad4Shell *>(e))
    return convert_shell(2, s->*get(Quad4Transform()), s->*get(Quad4Con()));
  if (TACSQuad9Shell *s = dynamic_cast<TACSQuad9Shell *>(e))
    return convert_shell(3, s->*get(Quad9Transform()), s->*get(Quad9Con()));
  if (TACSElement3D *s = dynamic_cast<TACSElement3D *>(e)) {
    TACSLinearElasticity3D *model = dynamic_cast<TACSLinearElasticity3D *>(s->getElementModel());
    if (!model || model->*get(ElasticityStrainType()) != TACS_LINEAR_STRAIN) return NULL;
    TACSSolidConstitutive *con = dynamic_cast<TACSSolidConstitutive *>(model->getConstitutive());
    if (!con) return NULL;
    tacsb200_handle basis = NULL;
    if (dynamic_cast<TACSLinearHexaBasis *>(s->getElementBasis())) basis = tacsb200_linear_hexa_basis_create();
    else if (dynamic_cast<TACSQuadraticHexaBasis *>(s->getElementBasis())) basis = tacsb200_quadratic_hexa_basis_create();
    if (!basis) return NULL;
    const double pt[3] = {0.0, 0.0, 0.0};
    const TacsScalar X[3] = {0.0, 0.0, 0.0};
    TacsScalar C[21];
    con->evalTangentStiffness(0, pt, X, C);
    tacsb200_handle dcon = tacsb200_solid_constitutive_create_raw(C, con->evalDensity(0, pt, X));
    tacsb200_handle dmodel = dcon ? tacsb200_linear_elasticity3d_create(dcon) : NULL;
    tacsb200_handle elem = dmodel ? tacsb200_element3d_create(dmodel, basis) : NULL;
    if (dcon) tacsb200_release(dcon);
    if (dmodel) tacsb200_release(dmodel);
    tacsb200_release(basis);
    return elem;
  }
  return NULL;
}

TACSB200Vec *device_vec(TACSVec *v) { return dynamic_cast<TACSB200Vec *>(v); }
}  // namespace

// ---------------------------------------------------------------------------------------------------------------
// TACSB200Vec
// ---------------------------------------------------------------------------------------------------------------
TACSB200Vec::TACSB200Vec(tacsb200_handle h) : handle(h), owner(NULL) {}
TACSB200Vec::~TACSB200Vec() {
  if (handle) tacsb200_release(handle);
}
TacsScalar TACSB200Vec::norm() { return tacsb200_vec_norm(handle); }
void TACSB200Vec::scale(TacsScalar alpha) { tacsb200_vec_scale(handle, alpha); }
TacsScalar TACSB200Vec::dot(TACSVec *x) {
  TACSB200Vec *xv = device_vec(x);
  if (!xv) {
    fprintf(stderr, "TACSB200Vec::dot: the argument is not a device vector\n");
    return 0.0;
  }
  return tacsb200_vec_dot(handle, xv->handle);
}
void TACSB200Vec::mdot(TACSVec **x, TacsScalar *ans, int m) {
  std::vector<tacsb200_handle> hs(m);
  for (int i = 0; i < m; i++) {
    TACSB200Vec *xv = device_vec(x[i]);
    if (!xv) {
      fprintf(stderr, "TACSB200Vec::mdot: argument %d is not a device vector\n", i);
      for (int k = 0; k < m; k++) ans[k] = 0.0;
      return;
    }
    hs[i] = xv->handle;
  }
  tacsb200_vec_mdot(handle, m, hs.data(), ans);
}
void TACSB200Vec::axpy(TacsScalar alpha, TACSVec *x) {
  TACSB200Vec *xv = device_vec(x);
  if (xv) tacsb200_vec_axpy(handle, alpha, xv->handle);
  else fprintf(stderr, "TACSB200Vec::axpy: the argument is not a device vector\n");
}
void TACSB200Vec::copyValues(TACSVec *x) {
  if (TACSB200Vec *xv = device_vec(x)) {
    tacsb200_vec_copy_values(handle, xv->handle);
  } else if (TACSBVec *hv = dynamic_cast<TACSBVec *>(x)) {
    TacsScalar *vals = NULL;
    const int n = hv->getArray(&vals);
    if (n == getSize()) setValues(vals);
    else fprintf(stderr, "TACSB200Vec::copyValues: size mismatch (%d vs %d)\n", n, getSize());
  } else {
    fprintf(stderr, "TACSB200Vec::copyValues: unsupported vector type\n");
  }
}
void TACSB200Vec::axpby(TacsScalar alpha, TacsScalar beta, TACSVec *x) {
  TACSB200Vec *xv = device_vec(x);
  if (xv) tacsb200_vec_axpby(handle, alpha, beta, xv->handle);
  else fprintf(stderr, "TACSB200Vec::axpby: the argument is not a device vector\n");
}
void TACSB200Vec::zeroEntries() { tacsb200_vec_zero_entries(handle); }
int TACSB200Vec::getSize() { return tacsb200_vec_get_size(handle); }
void TACSB200Vec::getValues(TacsScalar *host) { tacsb200_vec_get_array(handle, host); }
void TACSB200Vec::setValues(const TacsScalar *host) { tacsb200_vec_set_array(handle, host); }
void TACSB200Vec::copyTo(TACSBVec *host) {
  TacsScalar *vals = NULL;
  const int n = host->getArray(&vals);
  if (n == getSize()) getValues(vals);
  else fprintf(stderr, "TACSB200Vec::copyTo: size mismatch (%d vs %d)\n", n, getSize());
}
// The boundary conditions live with the device assembler (they were read from the application's TACSBcMap when it
// was created); the map argument identifies nothing else.
void TACSB200Vec::applyBCs(TACSBcMap *map, TACSVec *vec, const TacsScalar lambda) {
  (void)map;
  if (!owner) return;
  if (vec || lambda != 1.0) {
    fprintf(stderr, "TACSB200Vec::applyBCs: only the homogeneous form (vec = NULL) is on the device path\n");
    return;
  }
  tacsb200_assembler_apply_bcs_vec(owner->getHandle(), handle);
}
void TACSB200Vec::setBCs(TACSBcMap *map, const TacsScalar lambda) {
  (void)map;
  if (!owner || lambda != 1.0) return;
  tacsb200_assembler_set_bcs(owner->getHandle(), handle);
}

// ---------------------------------------------------------------------------------------------------------------
// TACSB200Mat
// ---------------------------------------------------------------------------------------------------------------
TACSB200Mat::TACSB200Mat(TACSB200Assembler *a, tacsb200_handle h) : assembler(a), handle(h) { assembler->incref(); }
TACSB200Mat::~TACSB200Mat() {
  if (handle) tacsb200_release(handle);
  assembler->decref();
}
void TACSB200Mat::zeroEntries() { tacsb200_mat_zero_entries(handle); }
void TACSB200Mat::applyBCs(TACSBcMap *bcmap) {
  (void)bcmap;
  tacsb200_assembler_apply_bcs_mat(assembler->getHandle(), handle);
}
void TACSB200Mat::getSize(int *nr, int *nc) {
  int bs = 0, rows = 0, cols = 0, nnzb = 0;
  tacsb200_mat_get_sizes(handle, 0, &bs, &rows, &cols, &nnzb);
  if (nr) *nr = bs * rows;
  if (nc) *nc = bs * rows;
}
TACSVec *TACSB200Mat::createVec() { return assembler->createVec(); }
TACSMat *TACSB200Mat::createDuplicate() { return assembler->createMat(); }
void TACSB200Mat::mult(TACSVec *x, TACSVec *y) {
  TACSB200Vec *xv = device_vec(x), *yv = device_vec(y);
  if (xv && yv) {
    tacsb200_mat_mult(handle, xv->getHandle(), yv->getHandle());
    return;
  }
  // host vectors of the reference: through device scratch copies
  TACSBVec *xh = dynamic_cast<TACSBVec *>(x), *yh = dynamic_cast<TACSBVec *>(y);
  if (!xh || !yh) {
    fprintf(stderr, "TACSB200Mat::mult: unsupported vector types\n");
    return;
  }
  TACSB200Vec *xd = assembler->createVec(), *yd = assembler->createVec();
  xd->incref();
  yd->incref();
  xd->copyValues(xh);
  tacsb200_mat_mult(handle, xd->getHandle(), yd->getHandle());
  yd->copyTo(yh);
  xd->decref();
  yd->decref();
}
void TACSB200Mat::copyValues(TACSMat *mat) {
  TACSB200Mat *m = dynamic_cast<TACSB200Mat *>(mat);
  if (m) tacsb200_mat_copy_values(handle, m->handle);
  else fprintf(stderr, "TACSB200Mat::copyValues: the argument is not a device matrix\n");
}
void TACSB200Mat::scale(TacsScalar alpha) { tacsb200_mat_scale(handle, alpha); }
void TACSB200Mat::axpy(TacsScalar alpha, TACSMat *mat) {
  TACSB200Mat *m = dynamic_cast<TACSB200Mat *>(mat);
  if (m) tacsb200_mat_axpy(handle, alpha, m->handle);
  else fprintf(stderr, "TACSB200Mat::axpy: the argument is not a device matrix\n");
}
void TACSB200Mat::getArrays(int *bsize, int *nrows, int *nnzb, int **rowp, int **cols, TacsScalar **vals) {
  int bs = 0, rows = 0, ncols = 0, nz = 0;
  tacsb200_mat_get_sizes(handle, 0, &bs, &rows, &ncols, &nz);
  if (bsize) *bsize = bs;
  if (nrows) *nrows = rows;
  if (nnzb) *nnzb = nz;
  int *rp = new int[rows + 1], *cl = new int[nz > 0 ? nz : 1];
  tacsb200_mat_get_pattern(handle, 0, rp, cl);
  if (rowp) *rowp = rp; else delete[] rp;
  if (cols) *cols = cl; else delete[] cl;
  if (vals) {
    *vals = new TacsScalar[(size_t)(nz > 0 ? nz : 1) * bs * bs];
    tacsb200_mat_get_values(handle, 0, *vals);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// TACSB200SchurMat
// ---------------------------------------------------------------------------------------------------------------
TACSB200SchurMat::TACSB200SchurMat(TACSB200Mat *src, TACSSchurMat *pattern) : source(src), handle(NULL) {
  source->incref();
  BCSRMat *blk[4] = {NULL, NULL, NULL, NULL};
  pattern->getBCSRMat(&blk[0], &blk[1], &blk[2], &blk[3]);
  const int *rowp[4], *cols[4];
  for (int k = 0; k < 4; k++) {
    int bs, nr, ncol;
    TacsScalar *vals;
    blk[k]->getArrays(&bs, &nr, &ncol, &rowp[k], &cols[k], &vals);
  }
  // local index -> global node of the interior (b) and interface (c) unknowns
  const int *b_nodes = NULL, *c_nodes = NULL;
  const int nb = pattern->getLocalMap()->getIndices()->getIndices(&b_nodes);
  const int nc = pattern->getSchurMap()->getIndices()->getIndices(&c_nodes);
  handle = tacsb200_schur_mat_create(source->getHandle(), nb, b_nodes, nc, c_nodes, rowp[0], cols[0], rowp[1], cols[1],
                                     rowp[2], cols[2], rowp[3], cols[3]);
  if (!handle) fprintf(stderr, "TACSB200SchurMat: the device view could not be created\n");
}
TACSB200SchurMat::~TACSB200SchurMat() {
  if (handle) tacsb200_release(handle);
  source->decref();
}
int TACSB200SchurMat::update() { return handle ? tacsb200_schur_mat_update(handle) : 1; }
void TACSB200SchurMat::copyValuesTo(TACSSchurMat *host) {
  if (!handle) return;
  BCSRMat *blk[4] = {NULL, NULL, NULL, NULL};
  host->getBCSRMat(&blk[0], &blk[1], &blk[2], &blk[3]);
  for (int k = 0; k < 4; k++) {
    int bs, nr, ncol;
    const int *rowp, *cols;
    TacsScalar *vals = NULL;
    blk[k]->getArrays(&bs, &nr, &ncol, &rowp, &cols, &vals);
    if (rowp[nr] > 0) tacsb200_schur_mat_get_values(handle, k, vals);
  }
}
TACSVec *TACSB200SchurMat::createVec() { return source->createVec(); }
void TACSB200SchurMat::getSize(int *nr, int *nc) { source->getSize(nr, nc); }
void TACSB200SchurMat::mult(TACSVec *x, TACSVec *y) {
  TACSB200Vec *xv = device_vec(x), *yv = device_vec(y);
  if (xv && yv && handle) tacsb200_schur_mat_mult(handle, xv->getHandle(), yv->getHandle());
  else fprintf(stderr, "TACSB200SchurMat::mult: device vectors required\n");
}

// ---------------------------------------------------------------------------------------------------------------
// preconditioner and Krylov solver
// ---------------------------------------------------------------------------------------------------------------
TACSB200ChebyshevPc::TACSB200ChebyshevPc(TACSB200Mat *m, int degree, double lower_factor, double upper_factor,
                                         int iters)
    : mat(m) {
  mat->incref();
  handle = tacsb200_chebyshev_create(mat->getHandle(), degree, lower_factor, upper_factor, iters);
}
TACSB200ChebyshevPc::~TACSB200ChebyshevPc() {
  if (handle) tacsb200_release(handle);
  mat->decref();
}
void TACSB200ChebyshevPc::factor() { tacsb200_chebyshev_factor(handle); }
void TACSB200ChebyshevPc::applyFactor(TACSVec *x, TACSVec *y) {
  TACSB200Vec *xv = device_vec(x), *yv = device_vec(y);
  if (xv && yv) tacsb200_chebyshev_apply_factor(handle, xv->getHandle(), yv->getHandle());
  else fprintf(stderr, "TACSB200ChebyshevPc::applyFactor: device vectors required\n");
}
void TACSB200ChebyshevPc::getMat(TACSMat **m) { *m = mat; }

TACSB200GMRES::TACSB200GMRES(TACSB200Mat *m, TACSB200ChebyshevPc *p, int msub, int nrestart, int is_flexible)
    : mat(m), pc(p), monitor(NULL) {
  mat->incref();
  if (pc) pc->incref();
  handle = pc ? tacsb200_gmres_create_pc(mat->getHandle(), pc->getHandle(), msub, nrestart, is_flexible)
              : tacsb200_gmres_create(mat->getHandle(), msub, nrestart);
}
TACSB200GMRES::~TACSB200GMRES() {
  if (handle) tacsb200_release(handle);
  if (pc) pc->decref();
  if (monitor) monitor->decref();
  mat->decref();
}
TACSVec *TACSB200GMRES::createVec() { return mat->createVec(); }
void TACSB200GMRES::setOperators(TACSMat *m, TACSPc *p) {
  (void)m;
  (void)p;
  fprintf(stderr, "TACSB200GMRES::setOperators: create a new solver for new operators\n");
}
void TACSB200GMRES::getOperators(TACSMat **m, TACSPc **p) {
  if (m) *m = mat;
  if (p) *p = pc;
}
int TACSB200GMRES::solve(TACSVec *b, TACSVec *x, int zero_guess) {
  TACSB200Vec *bv = device_vec(b), *xv = device_vec(x);
  if (!bv || !xv) {
    fprintf(stderr, "TACSB200GMRES::solve: device vectors required\n");
    return 0;
  }
  const int flag = tacsb200_gmres_solve(handle, bv->getHandle(), xv->getHandle(), zero_guess);
  iterCount = tacsb200_gmres_get_iter_count(handle);
  resNorm = tacsb200_gmres_get_residual_norm(handle);
  if (monitor) monitor->printResidual(iterCount, resNorm);
  return flag;
}
void TACSB200GMRES::setTolerances(double rtol, double atol) { tacsb200_gmres_set_tolerances(handle, rtol, atol); }
void TACSB200GMRES::setMonitor(KSMPrint *m) {
  if (m) m->incref();
  if (monitor) monitor->decref();
  monitor = m;
}

// ---------------------------------------------------------------------------------------------------------------
// TACSB200Assembler
// ---------------------------------------------------------------------------------------------------------------
TACSB200Assembler *TACSB200Assembler::create(TACSAssembler *assembler) {
  int mpi_size = 1;
  MPI_Comm_size(assembler->getMPIComm(), &mpi_size);
  if (mpi_size != 1) {
    fail("adopting an existing assembler is implemented for one rank (multi-GPU models go through tacsb200_creator_*)");
    return NULL;
  }
  if (assembler->getNumDependentNodes() > 0) {
    fail("dependent nodes are not on the device path");
    return NULL;
  }
  if (tacsb200_init(0) != 0) {
    fail("no usable B200");
    return NULL;
  }
  // auxiliary elements: the shell traction / pressure loads are carried over, anything else keeps the reference path
  tacsb200_handle dev_aux = NULL;
  if (assembler->getAuxElements()) {
    dev_aux = convert_aux_elements(assembler->getAuxElements());
    if (!dev_aux) {
      fail("an auxiliary element is not a TACSShellTraction / TACSShellPressure");
      return NULL;
    }
  }
  const int vpn = assembler->getVarsPerNode(), nnodes = assembler->getNumNodes(), nelems = assembler->getNumElements();
  const int *ptr = NULL, *conn = NULL;
  assembler->getElementConnectivity(&ptr, &conn);
  TACSElement **elements = assembler->getElements();

  // one device element object per distinct element object of the application
  std::map<TACSElement *, int> index;
  std::vector<tacsb200_handle> dev_elems;
  std::vector<int> elem_ids(nelems);
  bool ok = true;
  for (int e = 0; e < nelems && ok; e++) {
    std::map<TACSElement *, int>::iterator it = index.find(elements[e]);
    if (it == index.end()) {
      tacsb200_handle h = convert_element(elements[e]);
      if (!h) {
        char msg[256];
        snprintf(msg, sizeof(msg), "element %d (%s) is not one of the families of the device path", e,
                 elements[e]->getObjectName());
        fail(msg);
        ok = false;
        break;
      }
      index[elements[e]] = (int)dev_elems.size();
      elem_ids[e] = (int)dev_elems.size();
      dev_elems.push_back(h);
    } else {
      elem_ids[e] = it->second;
    }
  }
  TACSB200Assembler *self = NULL;
  if (ok) {
    self = new TACSB200Assembler();
    self->assembler = assembler;
    assembler->incref();
    self->num_nodes = nnodes;
    self->vars_per_node = vpn;
    tacsb200_handle cr = tacsb200_creator_create(vpn);
    self->creator = cr;
    ok = cr && tacsb200_creator_set_keep_numbering(cr, 1) == 0 &&
         tacsb200_creator_set_global_connectivity(cr, nnodes, nelems, ptr, conn, elem_ids.data()) == 0;
    // boundary conditions: TACSBcMap keeps (node, bit mask of constrained dofs, values[vpn]) (KSM.cpp:38-210)
    if (ok) {
      const int *bc_nodes = NULL, *bc_vars = NULL;
      TacsScalar *bc_vals = NULL;
      const int nbcs = assembler->getBcMap()->getBCs(&bc_nodes, &bc_vars, &bc_vals);
      std::vector<int> bptr(nbcs + 1, 0), bvars;
      std::vector<double> bvals;
      for (int i = 0; i < nbcs; i++) {
        for (int k = 0; k < vpn; k++)
          if (bc_vars[i] & (1 << k)) {
            bvars.push_back(k);
            bvals.push_back(bc_vals[vpn * i + k]);
          }
        bptr[i + 1] = (int)bvars.size();
      }
      ok = tacsb200_creator_set_boundary_conditions(cr, nbcs, bc_nodes, bptr.data(), bvars.data(), bvals.data()) == 0;
    }
    if (ok) {
      TACSBVec *X = NULL;
      assembler->getNodes(&X);
      TacsScalar *xp = NULL;
      X->getArray(&xp);
      ok = tacsb200_creator_set_nodes(cr, xp) == 0 &&
           tacsb200_creator_set_elements(cr, (int)dev_elems.size(), dev_elems.data()) == 0;
    }
    if (ok) {
      self->handle = tacsb200_creator_create_tacs(cr);
      ok = self->handle != NULL;
    }
    if (ok && dev_aux) ok = tacsb200_assembler_set_aux_elements(self->handle, dev_aux) == 0;
    if (!ok) {
      fail("the device assembler could not be created");
      self->incref();
      self->decref();
      self = NULL;
    }
  }
  for (size_t k = 0; k < dev_elems.size(); k++) tacsb200_release(dev_elems[k]);
  if (dev_aux) tacsb200_release(dev_aux);
  return self;
}

int TACSB200Assembler::assembleJacobian(TacsScalar alpha, TacsScalar beta, TacsScalar gamma, TACSBVec *res,
                                        TACSSchurMat *mat) {
  if (!schur_view) {
    schur_source = createMat();
    schur_source->incref();
    schur_view = new TACSB200SchurMat(schur_source, mat);
    schur_view->incref();
    schur_res = createVec();
    schur_res->incref();
  }
  if (!schur_view->valid()) return 1;
  assembleJacobian(alpha, beta, gamma, schur_res, schur_source);
  if (schur_view->update()) return 1;
  schur_view->copyValuesTo(mat);
  if (res) schur_res->copyTo(res);
  return 0;
}

TACSB200Assembler::~TACSB200Assembler() {
  if (schur_res) schur_res->decref();
  if (schur_view) schur_view->decref();
  if (schur_source) schur_source->decref();
  if (scratch_q) scratch_q->decref();
  if (scratch_qd) scratch_qd->decref();
  if (scratch_qdd) scratch_qdd->decref();
  if (handle) tacsb200_release(handle);
  if (creator) tacsb200_release(creator);
  if (assembler) assembler->decref();
}

TACSB200Vec *TACSB200Assembler::createVec() {
  TACSB200Vec *v = new TACSB200Vec(tacsb200_assembler_create_vec(handle));
  v->setAssembler(this);
  return v;
}
TACSB200Mat *TACSB200Assembler::createMat() { return new TACSB200Mat(this, tacsb200_assembler_create_mat(handle)); }

// a device vector for `v`: itself, or a scratch copy of a host vector of the reference
TACSB200Vec *TACSB200Assembler::stage(TACSVec *v, TACSB200Vec **scratch) {
  if (!v) return NULL;
  if (TACSB200Vec *d = device_vec(v)) return d;
  if (!*scratch) {
    *scratch = createVec();
    (*scratch)->incref();
  }
  (*scratch)->copyValues(v);
  return *scratch;
}

void TACSB200Assembler::setVariables(TACSVec *q, TACSVec *qdot, TACSVec *qddot) {
  TACSB200Vec *dq = stage(q, &scratch_q), *dqd = stage(qdot, &scratch_qd), *dqdd = stage(qddot, &scratch_qdd);
  tacsb200_assembler_set_variables(handle, dq ? dq->getHandle() : NULL, dqd ? dqd->getHandle() : NULL,
                                   dqdd ? dqdd->getHandle() : NULL);
}
void TACSB200Assembler::setNodes(TACSBVec *X) {
  TacsScalar *xp = NULL;
  X->getArray(&xp);
  tacsb200_handle nv = tacsb200_assembler_create_node_vec(handle);
  tacsb200_vec_set_array(nv, xp);
  tacsb200_assembler_set_nodes(handle, nv);
  tacsb200_release(nv);
}
void TACSB200Assembler::assembleRes(TACSB200Vec *res) { tacsb200_assembler_assemble_res(handle, res->getHandle()); }
void TACSB200Assembler::assembleJacobian(TacsScalar alpha, TacsScalar beta, TacsScalar gamma, TACSB200Vec *res,
                                         TACSB200Mat *A) {
  tacsb200_assembler_assemble_jacobian(handle, alpha, beta, gamma, res ? res->getHandle() : NULL, A->getHandle());
}
int TACSB200Assembler::assembleJacobianHost(TacsScalar alpha, TacsScalar beta, TacsScalar gamma, TACSBVec *vars,
                                            TACSBVec *res, TACSB200Mat *A) {
  TacsScalar *q = NULL, *r = NULL;
  vars->getArray(&q);
  res->getArray(&r);
  return tacsb200_assembler_assemble_jacobian_host(handle, alpha, beta, gamma, q, r, A->getHandle());
}
int TACSB200Assembler::assembleMatType(ElementMatrixType matType, TACSB200Mat *A) {
  return tacsb200_assembler_assemble_mat_type(handle, (int)matType, A->getHandle(), 1);
}
void TACSB200Assembler::applyBCs(TACSVec *vec) {
  if (TACSB200Vec *d = device_vec(vec)) tacsb200_assembler_apply_bcs_vec(handle, d->getHandle());
  else assembler->applyBCs(vec);
}
void TACSB200Assembler::setBCs(TACSVec *vec) {
  if (TACSB200Vec *d = device_vec(vec)) tacsb200_assembler_set_bcs(handle, d->getHandle());
  else assembler->setBCs(vec);
}
