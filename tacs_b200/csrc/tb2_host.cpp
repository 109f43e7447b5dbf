// Host objects of tacs_b200 (see tb2_host.h): constitutive constants, mesh partition / numbering,
// sparsity and gather plans, and the drivers that launch the sm_100a kernels.
#include "tb2_host.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <functional>
#include <map>
#include <thread>

#include "metis_shim/metis.h"
#include "plan.h"

namespace tb2 {

// ---------------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------------
static Context g_ctx;
Context &ctx() { return g_ctx; }

bool cuda_ok(cudaError_t err, const char *what) {
  if (err == cudaSuccess) return true;
  fprintf(stderr, "tacs_b200: CUDA error in %s: %s\n", what, cudaGetErrorString(err));
  return false;
}

int ctx_init(int device) {
  Context &c = g_ctx;
  if (c.device >= 0) return 0;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    fprintf(stderr, "tacs_b200: no CUDA device is visible; this library has no CPU path\n");
    return 1;
  }
  if (device < 0) device = 0;
  if (device >= ndev) {
    // a wrong LOCAL_RANK must not silently put two ranks on one GPU (ncclCommInitRank would fail much later)
    fprintf(stderr, "tacs_b200: device index %d is out of range (%d visible device%s)\n", device, ndev,
            ndev == 1 ? "" : "s");
    return 1;
  }
  if (!cuda_ok(cudaSetDevice(device), "cudaSetDevice")) return 1;
  cudaDeviceProp prop;
  if (!cuda_ok(cudaGetDeviceProperties(&prop, device), "cudaGetDeviceProperties")) return 1;
  if (prop.major < 10) {
    fprintf(stderr, "tacs_b200: device %d (%s, sm_%d%d) is not a Blackwell sm_100 part\n", device, prop.name,
            prop.major, prop.minor);
    return 1;
  }
  c.num_sms = prop.multiProcessorCount;
  // The gather stream carries bandwidth-bound work that is meant to fill in behind the critical path (element chunks,
  // the pack kernels and NCCL kernels of the exchanges): it gets the lowest CTA dispatch priority, the others the highest.
  // Without this the gather, once started, keeps every SM slot until it has drained and the exchange it should overlap
  // runs after it (C4 on 8 GPUs: 0.5 ms of 5.0).
  int prio_least = 0, prio_greatest = 0;
  if (!cuda_ok(cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest), "cudaDeviceGetStreamPriorityRange")) return 1;
  if (getenv("TACSB200_FLAT_PRIORITY")) prio_greatest = prio_least;   // (for A/B measurements)
  if (!cuda_ok(cudaStreamCreateWithPriority(&c.stream.s, cudaStreamNonBlocking, prio_greatest), "cudaStreamCreate")) return 1;
  if (!cuda_ok(cudaStreamCreateWithPriority(&c.comm_stream, cudaStreamNonBlocking, prio_greatest), "cudaStreamCreate")) return 1;
  if (!cuda_ok(cudaStreamCreateWithFlags(&c.copy_stream, cudaStreamNonBlocking), "cudaStreamCreate")) return 1;
  if (!cuda_ok(cudaStreamCreateWithPriority(&c.gather_stream, cudaStreamNonBlocking, prio_least), "cudaStreamCreate")) return 1;
  if (!cuda_ok(cudaEventCreateWithFlags(&c.gather_evt, cudaEventDisableTiming), "cudaEventCreate")) return 1;
  if (!cuda_ok(cudaEventCreateWithFlags(&c.tail_evt, cudaEventDisableTiming), "cudaEventCreate")) return 1;
  c.device = device;
  return 0;
}

// ---- event profiler ------------------------------------------------------------------------
static bool g_prof_on = false;
struct ProfRec { KernelId id; const char *name; cudaEvent_t e0, e1; };
static std::vector<ProfRec> g_prof_log;
static std::vector<cudaEvent_t> g_prof_pool;
static cudaEvent_t prof_event() {
  if (!g_prof_pool.empty()) {
    cudaEvent_t e = g_prof_pool.back();
    g_prof_pool.pop_back();
    return e;
  }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}
void profile_enable(int on) { g_prof_on = on != 0; }
KernelTimer::KernelTimer(KernelId id, const char *name, cudaStream_t st) : slot(-1), stream(st ? st : g_ctx.stream.s) {
  g_ctx.kernel_launches++;
  if (!g_prof_on) return;
  ProfRec r;
  r.id = id;
  r.name = name;
  r.e0 = prof_event();
  r.e1 = prof_event();
  cudaEventRecord(r.e0, stream);
  slot = (int)g_prof_log.size();
  g_prof_log.push_back(r);
}
KernelTimer::~KernelTimer() {
  if (slot >= 0) cudaEventRecord(g_prof_log[slot].e1, stream);
}
// per-name totals of the last collect: "name|launches|ms" lines (kernel names as launched, not literals of the caller)
static std::string g_prof_named;
const char *profile_named() { return g_prof_named.c_str(); }
int profile_collect(double *ms, long *count) {
  for (int k = 0; k < K_COUNT; k++) { ms[k] = 0.0; count[k] = 0; }
  if (g_ctx.device < 0) return 1;
  cudaStreamSynchronize(g_ctx.stream.s);
  cudaStreamSynchronize(g_ctx.gather_stream);
  std::map<std::string, std::pair<long, double>> named;
  for (auto &r : g_prof_log) {
    float t = 0.0f;
    cudaEventElapsedTime(&t, r.e0, r.e1);
    ms[r.id] += t;
    count[r.id]++;
    auto &slot = named[r.name ? r.name : "(unnamed)"];
    slot.first++;
    slot.second += t;
    g_prof_pool.push_back(r.e0);
    g_prof_pool.push_back(r.e1);
  }
  g_prof_log.clear();
  g_prof_named.clear();
  for (auto &kv : named) {
    char line[256];
    snprintf(line, sizeof(line), "%s|%ld|%.6f\n", kv.first.c_str(), kv.second.first, kv.second.second);
    g_prof_named += line;
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------
// constitutive objects
// ---------------------------------------------------------------------------------------------
TACSMaterialProperties::TACSMaterialProperties(double _rho, double _cp, double _E, double _nu, double _ys,
                                               double _alpha, double _kappa) {
  isotropic = true;
  rho = _rho; specific_heat = _cp; E = _E; nu = _nu; ys = _ys; alpha = _alpha; kappa = _kappa;
  G = 0.5 * E / (1.0 + nu);
  E1 = E2 = E3 = E;
  nu12 = nu13 = nu23 = nu;
  G12 = G13 = G23 = G;
}

TACSMaterialProperties::TACSMaterialProperties(double _rho, double _cp, double _E1, double _E2, double _E3,
                                               double _nu12, double _nu13, double _nu23, double _G12,
                                               double _G13, double _G23) {
  isotropic = false;
  rho = _rho; specific_heat = _cp;
  E = nu = G = 0.0; ys = alpha = kappa = 0.0;
  E1 = _E1; E2 = _E2; E3 = _E3;
  nu12 = _nu12; nu13 = _nu13; nu23 = _nu23;
  G12 = _G12; G13 = _G13; G23 = _G23;
}

// TACSMaterialProperties.cpp:323-341
void TACSMaterialProperties::evalTangentStiffness2D(double C[6]) const {
  if (isotropic) {
    double D = E / (1.0 - nu * nu);
    C[0] = D; C[1] = nu * D; C[2] = 0.0; C[3] = D; C[4] = 0.0; C[5] = G;
  } else {
    double nu21 = nu12 * E2 / E1;
    C[0] = E1 / (1.0 - nu12 * nu21);
    C[1] = nu12 * E2 / (1.0 - nu12 * nu21);
    C[2] = 0.0;
    C[3] = E2 / (1.0 - nu12 * nu21);
    C[4] = 0.0;
    C[5] = G12;
  }
}

// TACSMaterialProperties.cpp:270-321
void TACSMaterialProperties::evalTangentStiffness3D(double C[21]) const {
  for (int i = 0; i < 21; i++) C[i] = 0.0;
  if (isotropic) {
    double D = E / ((1.0 + nu) * (1.0 - 2.0 * nu));
    C[0] = (1.0 - nu) * D; C[1] = nu * D; C[2] = nu * D;
    C[6] = (1.0 - nu) * D; C[7] = nu * D;
    C[11] = (1.0 - nu) * D;
    C[15] = G; C[18] = G; C[20] = G;
  } else {
    double nu21 = (E2 / E1) * nu12, nu31 = (E3 / E1) * nu13, nu32 = (E3 / E2) * nu23;
    double D = 1.0 / (1.0 - nu12 * nu21 - nu13 * nu31 - nu23 * nu32 - 2.0 * nu21 * nu32 * nu13);
    C[0] = (1.0 - nu23 * nu32) * E1 * D;
    C[1] = (nu21 + nu31 * nu23) * E1 * D;
    C[2] = (nu31 + nu21 * nu32) * E1 * D;
    C[6] = (1.0 - nu13 * nu31) * E2 * D;
    C[7] = (nu32 + nu12 * nu31) * E2 * D;
    C[11] = (1.0 - nu12 * nu21) * E3 * D;
    C[15] = G23; C[18] = G13; C[20] = G12;
  }
}

// TACSMaterialProperties.cpp:523-545
TACSOrthotropicPly::TACSOrthotropicPly(double t, TACSMaterialProperties *p) {
  plyThickness = t;
  properties = p;
  properties->incref();
  rho = p->getDensity();
  double E1 = p->E1, E2 = p->E2, nu12 = p->nu12;
  double nu21 = nu12 * E2 / E1;
  Q11 = E1 / (1.0 - nu12 * nu21);
  Q22 = E2 / (1.0 - nu12 * nu21);
  Q12 = nu12 * E2 / (1.0 - nu12 * nu21);
  Q44 = p->G23; Q55 = p->G13; Q66 = p->G12;
  C12 = (Q11 + Q22 - 4.0 * Q66);
  C16 = (Q11 - Q12 - 2.0 * Q66);
  C26 = (Q12 - Q22 + 2.0 * Q66);
  C66 = (Q11 + Q22 - 2.0 * Q12 - 2.0 * Q66);
}
TACSOrthotropicPly::~TACSOrthotropicPly() { properties->decref(); }

// TACSMaterialProperties.cpp:752-773 (Jones, Mechanics of Composite Materials p.51)
void TACSOrthotropicPly::calculateQbar(double angle, double Qbar[6]) const {
  double cos1 = cos(angle), sin1 = sin(angle);
  double cos2 = cos1 * cos1, sin2 = sin1 * sin1, cos4 = cos2 * cos2, sin4 = sin2 * sin2;
  Qbar[0] = Q11 * cos4 + 2.0 * (Q12 + 2.0 * Q66) * sin2 * cos2 + Q22 * sin4;
  Qbar[1] = C12 * sin2 * cos2 + Q12 * (sin4 + cos4);
  Qbar[2] = C16 * sin1 * cos2 * cos1 + C26 * sin2 * sin1 * cos1;
  Qbar[3] = Q11 * sin4 + 2.0 * (Q12 + 2.0 * Q66) * sin2 * cos2 + Q22 * cos4;
  Qbar[4] = C16 * sin2 * sin1 * cos1 + C26 * sin1 * cos2 * cos1;
  Qbar[5] = C66 * sin2 * cos2 + Q66 * (sin4 + cos4);
}

// TACSMaterialProperties.cpp:724-734
void TACSOrthotropicPly::calculateAbar(double angle, double Abar[3]) const {
  double cos1 = cos(angle), sin1 = sin(angle);
  double cos2 = cos1 * cos1, sin2 = sin1 * sin1;
  Abar[0] = cos2 * Q44 + sin2 * Q55;
  Abar[1] = cos1 * sin1 * (Q55 - Q44);
  Abar[2] = sin2 * Q44 + cos2 * Q55;
}

double TACSShellConstitutive::DRILLING_REGULARIZATION = 0.1;  // TACSShellConstitutive.cpp:61
long TACSShellConstitutive::drill_version = 0;

void TACSShellConstitutive::fillDescriptor(double d[]) {
  evalTangentStiffness(d);
  evalMassMoments(d + 22);
}

TACSIsoShellConstitutive::TACSIsoShellConstitutive(TACSMaterialProperties *props, double _t, double _tOffset,
                                                   double _kcorr) {
  properties = props;
  if (properties) properties->incref();
  t = _t; tOffset = _tOffset; kcorr = _kcorr;
}
TACSIsoShellConstitutive::~TACSIsoShellConstitutive() {
  if (properties) properties->decref();
}

// TACSIsoShellConstitutive.cpp:192-226
void TACSIsoShellConstitutive::evalTangentStiffness(double C[]) {
  if (!properties) {
    memset(C, 0, 22 * sizeof(double));
    return;
  }
  double *A = &C[0], *B = &C[6], *D = &C[12], *As = &C[18];
  properties->evalTangentStiffness2D(A);
  for (int i = 0; i < 6; i++) B[i] = 0.0;
  double I = t * t * t / 12.0;
  for (int i = 0; i < 6; i++) {
    D[i] = I * A[i];
    A[i] *= t;
    B[i] += -tOffset * t * A[i];
    D[i] += tOffset * tOffset * t * t * A[i];
  }
  As[0] = As[2] = kcorr * A[5];
  As[1] = 0.0;
  C[21] = 0.5 * DRILLING_REGULARIZATION * (As[0] + As[2]);
}

// TACSIsoShellConstitutive.cpp:120-129
void TACSIsoShellConstitutive::evalMassMoments(double m[3]) {
  m[0] = m[1] = m[2] = 0.0;
  if (properties) {
    double rho = properties->getDensity();
    m[0] = rho * t;
    m[1] = -rho * t * t * tOffset;
    m[2] = rho * t * t * t * (tOffset * tOffset + 1.0 / 12.0);
  }
}

TACSCompositeShellConstitutive::TACSCompositeShellConstitutive(int n, TACSOrthotropicPly **plies,
                                                               const double *thick, const double *angles,
                                                               double _kcorr, double _tOffset) {
  for (int i = 0; i < n; i++) {
    plies[i]->incref();
    ply_props.push_back(plies[i]);
    ply_thickness.push_back(thick[i]);
    ply_angles.push_back(angles[i]);
  }
  kcorr = _kcorr;
  tOffset = _tOffset;
}
TACSCompositeShellConstitutive::~TACSCompositeShellConstitutive() {
  for (auto p : ply_props) p->decref();
}

// TACSCompositeShellConstitutive.cpp:249-301
void TACSCompositeShellConstitutive::evalTangentStiffness(double C[]) {
  double *A = &C[0], *B = &C[6], *D = &C[12], *As = &C[18];
  for (int k = 0; k < 6; k++) A[k] = B[k] = D[k] = 0.0;
  for (int k = 0; k < 3; k++) As[k] = 0.0;
  double t = 0.0;
  const int np = (int)ply_props.size();
  for (int i = 0; i < np; i++) t += ply_thickness[i];
  double t0 = -(0.5 + tOffset) * t;
  for (int k = 0; k < np; k++) {
    double Qbar[6], Abar[3];
    ply_props[k]->calculateQbar(ply_angles[k], Qbar);
    ply_props[k]->calculateAbar(ply_angles[k], Abar);
    double t1 = t0 + ply_thickness[k];
    double a = (t1 - t0), b = 0.5 * (t1 * t1 - t0 * t0), d = 1.0 / 3.0 * (t1 * t1 * t1 - t0 * t0 * t0);
    for (int i = 0; i < 6; i++) {
      A[i] += a * Qbar[i];
      B[i] += b * Qbar[i];
      D[i] += d * Qbar[i];
    }
    for (int i = 0; i < 3; i++) As[i] += kcorr * a * Abar[i];
    t0 = t1;
  }
  C[21] = 0.5 * DRILLING_REGULARIZATION * (As[0] + As[2]);
}

// TACSCompositeShellConstitutive.cpp:70-100
void TACSCompositeShellConstitutive::evalMassMoments(double m[3]) {
  m[0] = m[1] = m[2] = 0.0;
  double t = 0.0;
  const int np = (int)ply_props.size();
  for (int i = 0; i < np; i++) t += ply_thickness[i];
  double t0 = -(0.5 + tOffset) * t;
  for (int i = 0; i < np; i++) {
    double rho_ply = ply_props[i]->getDensity();
    double t1 = t0 + ply_thickness[i];
    double a = (t1 - t0), b = 0.5 * (t1 * t1 - t0 * t0), d = 1.0 / 3.0 * (t1 * t1 * t1 - t0 * t0 * t0);
    m[0] += a * rho_ply;
    m[1] += b * rho_ply;
    m[2] += d * rho_ply;
    t0 = t1;
  }
}

TACSSolidConstitutive::TACSSolidConstitutive(TACSMaterialProperties *props, double _t) {
  properties = props;
  if (properties) properties->incref();
  t = _t;
}
TACSSolidConstitutive::~TACSSolidConstitutive() {
  if (properties) properties->decref();
}
TACSSolidConstitutive::TACSSolidConstitutive(const double C[21], double density) {
  properties = nullptr;
  t = 1.0;
  raw = true;
  memcpy(Craw, C, sizeof(Craw));
  rho_raw = density;
}
TACSRawShellConstitutive::TACSRawShellConstitutive(const double C[22], const double moments[3]) {
  memcpy(Cs, C, sizeof(Cs));
  memcpy(mom, moments, sizeof(mom));
}
void TACSRawShellConstitutive::evalTangentStiffness(double C[]) { memcpy(C, Cs, sizeof(Cs)); }
void TACSRawShellConstitutive::evalMassMoments(double moments[3]) { memcpy(moments, mom, sizeof(mom)); }
void TACSRawShellConstitutive::fillDescriptor(double d[]) {
  memcpy(d, Cs, sizeof(Cs));
  memcpy(d + 22, mom, sizeof(mom));
}

// TACSSolidConstitutive.cpp:166-178
void TACSSolidConstitutive::evalTangentStiffness(double C[]) {
  if (raw) {
    memcpy(C, Craw, sizeof(Craw));
    return;
  }
  if (!properties) {
    memset(C, 0, 21 * sizeof(double));
    return;
  }
  properties->evalTangentStiffness3D(C);
  for (int i = 0; i < 21; i++) C[i] *= t;
}
void TACSSolidConstitutive::fillDescriptor(double d[]) {
  evalTangentStiffness(d);
  d[21] = raw ? rho_raw : (properties ? evalDensity() : 0.0);
}

// ---------------------------------------------------------------------------------------------
// transforms and elements
// ---------------------------------------------------------------------------------------------
TACSShellNaturalTransform::TACSShellNaturalTransform() {
  kind = 0;
  axis[0] = axis[1] = axis[2] = 0.0;
}
// TACSShellElementTransform.h:97-108: the axis is normalised once in the constructor
TACSShellRefAxisTransform::TACSShellRefAxisTransform(const double a[3]) {
  kind = 1;
  axis[0] = a[0]; axis[1] = a[1]; axis[2] = a[2];
  double norm = sqrt(axis[0] * axis[0] + axis[1] * axis[1] + axis[2] * axis[2]);
  double inv = 0.0;
  if (norm != 0.0) inv = 1.0 / norm;
  axis[0] *= inv; axis[1] *= inv; axis[2] *= inv;
}

TACSLinearElasticity3D::TACSLinearElasticity3D(TACSSolidConstitutive *con) {
  stiff = con;
  stiff->incref();
}
TACSLinearElasticity3D::~TACSLinearElasticity3D() { stiff->decref(); }

TACSShellElement::TACSShellElement(int o, TACSShellTransform *t, TACSShellConstitutive *c) {
  order = o; transform = t; con = c;
  transform->incref();
  con->incref();
}
TACSShellElement::~TACSShellElement() {
  transform->decref();
  con->decref();
}
void TACSShellElement::fillDescriptor(double d[]) {
  for (int i = 0; i < 32; i++) d[i] = 0.0;
  con->fillDescriptor(d);
  d[25] = (double)transform->kind;
  d[26] = transform->axis[0]; d[27] = transform->axis[1]; d[28] = transform->axis[2];
}

TACSElement3D::TACSElement3D(TACSElementModel *m, TACSElementBasis *b) {
  model = dynamic_cast<TACSLinearElasticity3D *>(m);
  basis = b;
  if (model) model->incref();
  basis->incref();
}
TACSElement3D::~TACSElement3D() {
  if (model) model->decref();
  basis->decref();
}
void TACSElement3D::fillDescriptor(double d[]) {
  for (int i = 0; i < 32; i++) d[i] = 0.0;
  model->stiff->fillDescriptor(d);
}

// Staging layout is node-pair-major with a row-major bs x bs block inside; convert to the dense
// row-major element matrix mat[nv*row + col] of TACSElement::addJacobian.
int TACSElement::addJacobianBatch(int count, double alpha, double beta, double gamma, const double *Xpts,
                                  const double *vars, const double *dvars, const double *ddvars, double *res,
                                  double *mat) {
  (void)beta; (void)dvars;
  if (ctx_init(-1)) return 1;
  const int nn = getNumNodes(), bs = getVarsPerNode(), nv = nn * bs, kind = kernelKind();
  std::vector<int> conn((size_t)count * nn), desc(count, 0);
  for (size_t k = 0; k < conn.size(); k++) conn[k] = (int)k;
  double drow[32];
  fillDescriptor(drow);
  std::vector<unsigned char> tab(elem_tables_bytes(kind));
  elem_tables_build(kind, tab.data());
  DeviceArray<int> d_conn, d_desc;
  DeviceArray<double> d_table, d_X, d_u, d_a, d_Ke, d_Re;
  DeviceArray<unsigned char> d_tab;
  if (!d_conn.upload(conn) || !d_desc.upload(desc) || !d_table.upload(drow, 32) ||
      !d_tab.upload(tab.data(), tab.size()) || !d_X.upload(Xpts, (size_t)count * 3 * nn) ||
      !d_u.upload(vars, (size_t)count * nv))
    return 1;
  if (ddvars && !d_a.upload(ddvars, (size_t)count * nv)) return 1;
  if (mat && !d_Ke.alloc((size_t)count * nv * nv)) return 1;
  if (!d_Re.alloc((size_t)count * nv)) return 1;
  ElemGroupArgs g;
  g.kind = kind; g.nelem = count; g.conn = d_conn.ptr; g.desc_index = d_desc.ptr; g.desc_table = d_table.ptr;
  g.tables = d_tab.ptr; g.Xpts = d_X.ptr; g.vars = d_u.ptr; g.ddvars = ddvars ? d_a.ptr : nullptr;
  g.alpha = alpha; g.gamma = gamma; g.Ke = mat ? d_Ke.ptr : nullptr; g.Re = d_Re.ptr;
  g.upper = 0; g.dmap = nullptr; g.direct = nullptr;  // element-level interface: every node pair is returned
  g.geometric = 0;
  g.uncoupled = (kind == ELEM_QUAD4_SHELL || kind == ELEM_QUAD9_SHELL) && shell_desc_uncoupled(drow) ? 1 : 0;
  if (!cuda_ok(launch_element_group(g, ctx().num_sms, ctx().stream), "element kernel")) return 1;
  ctx().kernel_launches++;
  if (!cuda_ok(cudaStreamSynchronize(ctx().stream), "element kernel sync")) return 1;
  if (res && !d_Re.download(res, (size_t)count * nv)) return 1;
  if (mat) {
    std::vector<double> stage((size_t)count * nv * nv);
    if (!d_Ke.download(stage.data(), stage.size())) return 1;
    const int b2 = bs * bs;
    for (int e = 0; e < count; e++)
      for (int i = 0; i < nn; i++)
        for (int j = 0; j < nn; j++) {
          const double *blk = &stage[((size_t)(e * nn + i) * nn + j) * b2];
          for (int a = 0; a < bs; a++)
            for (int b = 0; b < bs; b++)
              mat[(size_t)e * nv * nv + (size_t)nv * (bs * i + a) + bs * j + b] = blk[bs * a + b];
        }
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------
// creator: partition, first-touch numbering, local meshes
// ---------------------------------------------------------------------------------------------
TACSCreator::TACSCreator(int vpn) { vars_per_node = vpn; }
TACSCreator::~TACSCreator() {
  for (auto e : elements)
    if (e) e->decref();
}

void TACSCreator::setGlobalConnectivity(int nn, int ne, const int *ptr, const int *conn, const int *ids) {
  num_nodes = nn;
  num_elements = ne;
  elem_node_ptr.assign(ptr, ptr + ne + 1);
  elem_node_conn.assign(conn, conn + ptr[ne]);
  elem_id_nums.assign(ids, ids + ne);
}

// TACSCreator.cpp:226-271
void TACSCreator::setBoundaryConditions(int nb, const int *nodes, const int *ptr, const int *vars,
                                        const double *vals) {
  bc_nodes.assign(nodes, nodes + nb);
  bc_ptr.resize(nb + 1);
  if (ptr && vars) {
    bc_ptr.assign(ptr, ptr + nb + 1);
    bc_vars.assign(vars, vars + ptr[nb]);
    if (vals) bc_vals.assign(vals, vals + ptr[nb]);
    else bc_vals.assign(ptr[nb], 0.0);
  } else {
    bc_vars.resize((size_t)nb * vars_per_node);
    bc_ptr[0] = 0;
    for (int i = 0; i < nb; i++) {
      bc_ptr[i + 1] = bc_ptr[i] + vars_per_node;
      for (int j = 0; j < vars_per_node; j++) bc_vars[bc_ptr[i] + j] = j;
    }
    bc_vals.assign(bc_ptr[nb], 0.0);
  }
}

void TACSCreator::setNodes(const double *X) { Xpts.assign(X, X + 3 * (size_t)num_nodes); }

void TACSCreator::setElements(int n, TACSElement **elems) {
  for (auto e : elements)
    if (e) e->decref();
  elements.assign(elems, elems + n);
  for (auto e : elements)
    if (e) e->incref();
}

// TacsUniqueSort (TacsUtilities.cpp:76-99)
static int unique_sort(int len, int *a) {
  std::sort(a, a + len);
  int i = 0;
  while (i < len && a[i] < 0) i++;
  int n = 0;
  for (; i < len; i++)
    if (n == 0 || a[n - 1] != a[i]) a[n++] = a[i];
  return n;
}

// TACSCreator::partitionMesh (TACSCreator.cpp:923-1209): element dual graph (elements adjacent when
// they share a node, rows sorted ascending, diagonal removed) -> METIS recursive (< 8 parts) or
// k-way; then first-touch node renumbering in global element order.
int TACSCreator::comm_size() const { return plan_size > 0 ? plan_size : ctx().size; }

int TACSCreator::partitionMesh(int split_size, const int *part) {
  const int mpi_size = comm_size();
  if (split_size <= 0 || split_size > mpi_size) split_size = mpi_size;
  partition.clear();
  if (part) {
    bool legit = true;
    for (int i = 0; i < num_elements; i++)
      if (part[i] < 0 || part[i] >= split_size) { legit = false; break; }
    if (legit) partition.assign(part, part + num_elements);
  }
  if (partition.empty()) {
    partition.assign(num_elements, 0);
    if (split_size > 1) {
      std::vector<int> node_elem_ptr(num_nodes + 1, 0);
      for (int i = 0; i < num_elements; i++)
        for (int j = elem_node_ptr[i]; j < elem_node_ptr[i + 1]; j++)
          if (elem_node_conn[j] >= 0) node_elem_ptr[elem_node_conn[j] + 1]++;
      for (int i = 0; i < num_nodes; i++) node_elem_ptr[i + 1] += node_elem_ptr[i];
      std::vector<int> node_elem_conn(node_elem_ptr[num_nodes]);
      {
        std::vector<int> cursor(node_elem_ptr.begin(), node_elem_ptr.end() - 1);
        for (int i = 0; i < num_elements; i++)
          for (int j = elem_node_ptr[i]; j < elem_node_ptr[i + 1]; j++)
            if (elem_node_conn[j] >= 0) node_elem_conn[cursor[elem_node_conn[j]]++] = i;
      }
      std::vector<int> elem_ptr(num_elements + 1, 0), elem_conn, row;
      elem_conn.reserve((size_t)27 * num_elements);
      for (int i = 0; i < num_elements; i++) {
        row.clear();
        for (int jp = elem_node_ptr[i]; jp < elem_node_ptr[i + 1]; jp++) {
          int node = elem_node_conn[jp];
          if (node >= 0)
            for (int kp = node_elem_ptr[node]; kp < node_elem_ptr[node + 1]; kp++) row.push_back(node_elem_conn[kp]);
        }
        int len = unique_sort((int)row.size(), row.data());
        for (int j = 0; j < len; j++)
          if (row[j] != i) elem_conn.push_back(row[j]);
        elem_ptr[i + 1] = (int)elem_conn.size();
      }
      int ncon = 1, options[METIS_NOPTIONS], objval = 0, nparts = split_size, ne = num_elements;
      METIS_SetDefaultOptions(options);
      options[METIS_OPTION_NUMBERING] = 0;
      if (split_size < 8)
        METIS_PartGraphRecursive(&ne, &ncon, elem_ptr.data(), elem_conn.data(), NULL, NULL, NULL, &nparts, NULL,
                                 NULL, options, &objval, partition.data());
      else
        METIS_PartGraphKway(&ne, &ncon, elem_ptr.data(), elem_conn.data(), NULL, NULL, NULL, &nparts, NULL, NULL,
                            options, &objval, partition.data());
    }
  }
  if (keep_numbering) {
    // the caller's numbering is final (single rank): every node and element belongs to rank 0 as numbered
    new_nodes.resize(num_nodes);
    for (int i = 0; i < num_nodes; i++) new_nodes[i] = i;
    owned_elements.assign(mpi_size, 0);
    owned_nodes.assign(mpi_size, 0);
    owned_elements[0] = num_elements;
    owned_nodes[0] = num_nodes;
    std::fill(partition.begin(), partition.end(), 0);
    return 0;
  }
  // first-touch numbering (TACSCreator.cpp:1141-1205)
  new_nodes.assign(num_nodes, 0);
  owned_elements.assign(mpi_size, 0);
  owned_nodes.assign(mpi_size, 0);
  for (int j = 0; j < num_elements; j++) {
    int owner = partition[j];
    owned_elements[owner]++;
    for (int i = elem_node_ptr[j]; i < elem_node_ptr[j + 1]; i++) {
      int node = elem_node_conn[i];
      if (node >= 0 && !new_nodes[node]) {
        new_nodes[node] = 1;
        owned_nodes[owner]++;
      }
    }
  }
  std::fill(new_nodes.begin(), new_nodes.end(), -1);
  std::vector<int> split_offset(split_size, 0);
  for (int k = 1; k < split_size; k++) split_offset[k] = split_offset[k - 1] + owned_nodes[k - 1];
  for (int j = 0; j < num_elements; j++) {
    int owner = partition[j];
    for (int i = elem_node_ptr[j]; i < elem_node_ptr[j + 1]; i++) {
      int node = elem_node_conn[i];
      if (node >= 0 && new_nodes[node] < 0) new_nodes[node] = split_offset[owner]++;
    }
  }
  return 0;
}

int TACSCreator::getNodeNums(const int **nn) {
  if (nn) *nn = new_nodes.empty() ? nullptr : new_nodes.data();
  return new_nodes.empty() ? 0 : num_nodes;
}
int TACSCreator::getElementPartition(const int **p) {
  if (p) *p = partition.empty() ? nullptr : partition.data();
  return partition.empty() ? 0 : num_elements;
}

// TACSCreator::createTACS (TACSCreator.cpp:436-909). Every rank holds the global mesh, so the
// root->rank scatter of the reference becomes a local selection (plan.cpp): the elements of this rank
// in ascending global order, connectivity in the new global numbering, nodes of the owned range.
int TACSCreator::prepareMesh(int rank, int size, std::shared_ptr<GlobalMesh> &gm, std::vector<int> &kinds) {
  if (elements.empty()) {
    fprintf(stderr, "[%d] TACSCreator: Elements and callback not defined\n", rank);
    return 1;
  }
  if (partition.empty() || new_nodes.empty() || (int)owned_nodes.size() != size) partitionMesh(size, nullptr);
  gm = std::make_shared<GlobalMesh>();
  gm->num_nodes = num_nodes;
  gm->num_elements = num_elements;
  gm->size = size;
  gm->ptr = elem_node_ptr;
  gm->conn.resize(elem_node_conn.size());
  for (size_t k = 0; k < elem_node_conn.size(); k++) {
    if (elem_node_conn[k] < 0) {
      fprintf(stderr, "[%d] tacs_b200: dependent nodes are not supported on the device path\n", rank);
      return 1;
    }
    gm->conn[k] = new_nodes[elem_node_conn[k]];
  }
  gm->part = partition;
  gm->owner_range.assign(size + 1, 0);
  for (int k = 0; k < size; k++) gm->owner_range[k + 1] = gm->owner_range[k] + owned_nodes[k];
  kinds.assign(num_elements, 0);
  for (int e = 0; e < num_elements; e++) {
    int id = elem_id_nums[e];
    TACSElement *el = (id >= 0 && id < (int)elements.size()) ? elements[id] : nullptr;
    if (!el) {
      fprintf(stderr, "[%d] TACSCreator: Element undefined for element ID %d\n", rank, id);
      return 1;
    }
    kinds[e] = el->kernelKind();
    if (el->getVarsPerNode() != vars_per_node || el->getNumNodes() != elem_node_ptr[e + 1] - elem_node_ptr[e]) {
      fprintf(stderr, "[%d] TACSAssembler: Element %s does not match variables per node / connectivity\n", rank,
              el->getObjectName());
      return 1;
    }
    TACSElement3D *solid = dynamic_cast<TACSElement3D *>(el);
    if (solid && !solid->model) {
      fprintf(stderr, "[%d] tacs_b200: unsupported element model on the device path\n", rank);
      return 1;
    }
  }
  return 0;
}

PlanObject *TACSCreator::createPlan(int rank, int size) {
  plan_size = size;
  std::shared_ptr<GlobalMesh> gm;
  std::vector<int> kinds;
  if (prepareMesh(rank, size, gm, kinds)) return nullptr;
  PlanObject *po = new PlanObject();
  if (po->plan.build(gm, vars_per_node, rank, kinds) || po->plan.buildMatrix()) {
    delete po;
    return nullptr;
  }
  return po;
}

TACSAssembler *TACSCreator::createTACS() {
  if (ctx_init(-1)) return nullptr;
  const int rank = ctx().rank, size = ctx().size;
  plan_size = 0;
  std::shared_ptr<GlobalMesh> gm;
  std::vector<int> kinds;
  if (prepareMesh(rank, size, gm, kinds)) return nullptr;
  TACSAssembler *a = new TACSAssembler();
  a->plan.reset(new HostPlan());
  if (a->plan->build(gm, vars_per_node, rank, kinds)) {
    delete a;
    return nullptr;
  }
  HostPlan &P = *a->plan;
  a->bs = vars_per_node;
  a->rank = rank;
  a->size = size;
  a->owner_range = P.owner_range;
  a->nowned = P.nowned;
  a->nlocal = P.nlocal;
  a->nelems = P.nelems;
  a->ext_before = P.ext_before;
  a->ext_after = P.ext_after;
  for (int e = 0; e < P.nelems; e++) a->elems.push_back(elements[elem_id_nums[P.elem_global[e]]]);
  for (auto e : a->elems) e->incref();
  // boundary conditions in the new numbering (TACSCreator.cpp:478-481, 836-851)
  for (size_t k = 0; k < bc_nodes.size(); k++) {
    int node = new_nodes[bc_nodes[k]];
    if (node < 0) continue;
    int mask = 0;
    std::vector<double> vals(vars_per_node, 0.0);
    int n = 0;
    for (int j = bc_ptr[k]; j < bc_ptr[k + 1]; j++)
      if (bc_vars[j] >= 0 && bc_vars[j] < vars_per_node) {
        mask |= (1 << bc_vars[j]);
        vals[bc_vars[j]] = bc_vals[j];
        n++;
      }
    if (n > 0) {
      a->bc_nodes.push_back(node);
      a->bc_vars.push_back(mask);
      a->bc_vals.insert(a->bc_vals.end(), vals.begin(), vals.end());
    }
  }
  // node locations of every local node (the reference fills the ghosts with a halo exchange in
  // setNodes, TACSAssembler.cpp:920-926; here every rank already holds the global coordinates)
  std::vector<double> Xl((size_t)3 * a->nlocal, 0.0);
  if (!Xpts.empty()) {
    std::vector<int> inv(num_nodes, -1);
    for (int i = 0; i < num_nodes; i++)
      if (new_nodes[i] >= 0) inv[new_nodes[i]] = i;
    for (int l = 0; l < a->nlocal; l++) {
      int old = inv[P.globalNode(l)];
      for (int c = 0; c < 3; c++) Xl[3 * (size_t)l + c] = Xpts[3 * (size_t)old + c];
    }
  }
  if (a->finalize()) {
    delete a;
    return nullptr;
  }
  if (!a->xpts->data.upload(Xl)) {
    delete a;
    return nullptr;
  }
  return a;
}

// ---------------------------------------------------------------------------------------------
// vectors
// ---------------------------------------------------------------------------------------------
TACSBVec::TACSBVec(int bs, int no, int eb, int ea) {
  bsize = bs; nowned = no; ext_before = eb; ext_after = ea;
  data.alloc((size_t)localSize());
  zeroEntries();
}

double *g_dot_partial = nullptr, *g_dot_out = nullptr, *g_dot_host = nullptr;
bool dot_buffers() {
  if (g_dot_partial) return true;
  const size_t np = (size_t)dot_num_partials(ctx().num_sms) * 8;
  return cuda_ok(cudaMalloc(&g_dot_partial, np * sizeof(double)), "cudaMalloc") &&
         cuda_ok(cudaMalloc(&g_dot_out, 8 * sizeof(double)), "cudaMalloc") &&
         cuda_ok(cudaMallocHost(&g_dot_host, 8 * sizeof(double)), "cudaMallocHost");
}

int comm_allreduce_sum(double *dev_buf, int n);  // comm.cpp

int TACSBVec::mdot(TACSBVec **ys, double *out, int n) {
  if (!dot_buffers()) {
    for (int v = 0; v < n; v++) out[v] = NAN;
    return 1;
  }
  int rc = 0;
  for (int done = 0; done < n; done += 8) {
    const int nv = std::min(8, n - done);
    const double *ptrs[8];
    for (int v = 0; v < nv; v++) ptrs[v] = ys[done + v]->owned();
    {
      KernelTimer kt(K_DOT);
      if (!cuda_ok(launch_mdot(ownedSize(), owned(), nv, ptrs, g_dot_partial, g_dot_out, ctx().num_sms, ctx().stream),
                   "mdot")) rc = 1;
      ctx().kernel_launches++;
    }
    if (ctx().size > 1 && comm_allreduce_sum(g_dot_out, nv)) rc = 1;
    if (!cuda_ok(cudaMemcpyAsync(g_dot_host, g_dot_out, nv * sizeof(double), cudaMemcpyDeviceToHost, ctx().stream),
                 "mdot D2H") ||
        !cuda_ok(cudaStreamSynchronize(ctx().stream), "mdot sync"))
      rc = 1;
    for (int v = 0; v < nv; v++) out[done + v] = rc ? NAN : g_dot_host[v];
  }
  return rc;
}
double TACSBVec::dot(TACSBVec *y) {
  double r = 0.0;
  mdot(&y, &r, 1);
  return r;
}
double TACSBVec::norm() {
  TACSBVec *self = this;
  double r = 0.0;
  mdot(&self, &r, 1);
  return sqrt(r);
}
void TACSBVec::axpy(double alpha, TACSBVec *x) {
  KernelTimer kt(K_VEC);
  cuda_ok(launch_axpy(ownedSize(), alpha, x->owned(), owned(), ctx().num_sms, ctx().stream), "axpy");
}
void TACSBVec::axpby(double alpha, double beta, TACSBVec *x) {
  KernelTimer kt(K_VEC);
  cuda_ok(launch_axpby(ownedSize(), alpha, beta, x->owned(), owned(), ctx().num_sms, ctx().stream), "axpby");
}
void TACSBVec::scale(double alpha) {
  KernelTimer kt(K_VEC);
  cuda_ok(launch_scale(ownedSize(), alpha, owned(), ctx().num_sms, ctx().stream), "scale");
}
void TACSBVec::copyValues(TACSBVec *x) {
  cuda_ok(cudaMemcpyAsync(owned(), x->owned(), ownedSize() * sizeof(double), cudaMemcpyDeviceToDevice,
                          ctx().stream), "copyValues");
}
void TACSBVec::zeroEntries() {
  if (data.count) cuda_ok(cudaMemsetAsync(data.ptr, 0, data.count * sizeof(double), ctx().stream), "zeroEntries");
}
// copies between a host array and the owned slice; behind a matrix-only tail of the compute stream they overlap it
static bool host_copy(void *dst, const void *src, size_t bytes, cudaMemcpyKind kind, const char *what) {
  Context &c = ctx();
  if (c.tail_is_matrix_only()) {
    return cuda_ok(cudaStreamWaitEvent(c.copy_stream, c.tail_evt, 0), what) &&
           cuda_ok(cudaMemcpyAsync(dst, src, bytes, kind, c.copy_stream), what) &&
           cuda_ok(cudaStreamSynchronize(c.copy_stream), what);
  }
  return cuda_ok(cudaMemcpyAsync(dst, src, bytes, kind, c.stream), what) &&
         cuda_ok(cudaStreamSynchronize(c.stream.s), what);
}
int TACSBVec::getArray(double *out) {
  if (ownedSize() == 0) return 0;
  return host_copy(out, owned(), ownedSize() * sizeof(double), cudaMemcpyDeviceToHost, "getArray") ? 0 : 1;
}
int TACSBVec::setArray(const double *in) {
  if (ownedSize() == 0) return 0;
  return host_copy(owned(), in, ownedSize() * sizeof(double), cudaMemcpyHostToDevice, "setArray") ? 0 : 1;
}

// ---------------------------------------------------------------------------------------------
// assembler
// ---------------------------------------------------------------------------------------------
TACSAssembler::TACSAssembler() {}
TACSAssembler::~TACSAssembler() {
  for (auto e : elems) e->decref();
  if (xpts) xpts->decref();
  if (vars) vars->decref();
  if (dvars) dvars->decref();
  if (ddvars) ddvars->decref();
  if (jvp_x) jvp_x->decref();
  if (jvp_a) jvp_a->decref();
  if (jvp_t) jvp_t->decref();
  if (aux_elements) aux_elements->decref();
}

int TACSAssembler::localNode(int g) const { return plan->localNode(g); }

int comm_setup_exchange(DeviceExchange &dx, const ExchangePlan &x);  // comm.cpp

int TACSAssembler::finalize() {
  HostPlan &P = *plan;
  // state vectors in local order
  xpts = new TACSBVec(3, nowned, ext_before, ext_after);
  vars = new TACSBVec(bs, nowned, ext_before, ext_after);
  dvars = new TACSBVec(bs, nowned, ext_before, ext_after);
  ddvars = new TACSBVec(bs, nowned, ext_before, ext_after);
  xpts->incref(); vars->incref(); dvars->incref(); ddvars->incref();

  // distinct descriptors -> device table
  std::map<TACSElement *, int> index;
  elem_desc.resize(nelems);
  for (int e = 0; e < nelems; e++) {
    TACSElement *el = elems[e];
    auto it = index.find(el);
    if (it == index.end()) {
      int row = (int)distinct.size();
      index[el] = row;
      distinct.push_back(el);
      elem_desc[e] = row;
    } else {
      elem_desc[e] = it->second;
    }
  }
  if (buildDescriptorTable()) return 1;
  // element groups by kernel family (local order preserved inside a group)
  groups.clear();
  for (size_t gi = 0; gi < P.group_kinds.size(); gi++) {
    groups.emplace_back();
    ElemGroup &g = groups.back();
    g.kind = P.group_kinds[gi];
    g.nn = elem_kind_nodes(g.kind);
    g.local_elems = P.group_elems[gi];
    g.nelem = (long)g.local_elems.size();
    g.block_base = P.group_block_base[gi];
    g.node_base = P.group_node_base[gi];
    std::vector<int> conn((size_t)g.nelem * g.nn), desc(g.nelem);
    for (long k = 0; k < g.nelem; k++) {
      int e = g.local_elems[k];
      for (int i = 0; i < g.nn; i++) conn[(size_t)k * g.nn + i] = P.elem_conn_local[P.elem_ptr[e] + i];
      desc[k] = elem_desc[e];
    }
    std::vector<unsigned char> tab(elem_tables_bytes(g.kind));
    elem_tables_build(g.kind, tab.data());
    if (!g.d_conn.upload(conn) || !g.d_desc.upload(desc) || !g.d_tables.upload(tab.data(), tab.size())) return 1;
  }
  total_blocks = P.local_blocks + P.recv_blocks;
  total_node_slots = P.local_node_slots + P.recv_node_slots;

  // boundary conditions: merge duplicates in application order (sequential semantics of TACSBcMap)
  {
    std::map<int, int> pos;
    std::vector<int> nodes, masks;
    std::vector<double> vals;
    for (size_t k = 0; k < bc_nodes.size(); k++) {
      int g = bc_nodes[k];
      auto it = pos.find(g);
      int p;
      if (it == pos.end()) {
        p = (int)nodes.size();
        pos[g] = p;
        nodes.push_back(g);
        masks.push_back(0);
        vals.resize((size_t)bs * (p + 1), 0.0);
      } else {
        p = it->second;
      }
      masks[p] |= bc_vars[k];
      for (int c = 0; c < bs; c++)
        if (bc_vars[k] & (1 << c)) vals[(size_t)bs * p + c] = bc_vals[(size_t)bs * k + c];
    }
    const int lo = owner_range[rank], hi = owner_range[rank + 1];
    std::vector<int> rows(nodes.size()), local(nodes.size());
    for (size_t k = 0; k < nodes.size(); k++) {
      rows[k] = (nodes[k] >= lo && nodes[k] < hi) ? nodes[k] - lo : -1;
      local[k] = localNode(nodes[k]);
    }
    nbc_dev = (int)nodes.size();
    h_bc_rows = rows;
    if (!d_bc_rows.upload(rows) || !d_bc_vars.upload(masks) || !d_bc_vals.upload(vals) || !d_bc_local.upload(local))
      return 1;
  }
  if (!r_ptr.upload(P.r_ptr) || !r_src.upload(P.r_src)) return 1;
  if (!Re.alloc((size_t)total_node_slots * bs)) return 1;
  if (size > 1) {
    if (comm_setup_exchange(x_state, P.state) || comm_setup_exchange(x_rows, P.rows) ||
        comm_setup_exchange(x_blocks, P.blocks))
      return 1;
  }
  return 0;
}

int TACSAssembler::buildDescriptorTable() {
  std::vector<double> table((size_t)32 * distinct.size(), 0.0);
  for (size_t row = 0; row < distinct.size(); row++) distinct[row]->fillDescriptor(&table[32 * row]);
  if (!d_desc_table.upload(table)) return 1;
  // shells whose constitutive B block (entries 6..11) vanishes take the cheaper uncoupled kernel path
  shells_uncoupled = true;
  for (size_t row = 0; row < distinct.size(); row++) {
    const int k = distinct[row]->kernelKind();
    if ((k == ELEM_QUAD4_SHELL || k == ELEM_QUAD9_SHELL) && !shell_desc_uncoupled(&table[32 * row]))
      shells_uncoupled = false;
  }
  desc_version = TACSShellConstitutive::drill_version;
  return 0;
}

int TACSAssembler::refreshDescriptors() {
  if (desc_version == TACSShellConstitutive::drill_version) return 0;
  // the table may still be read by kernels in flight
  if (!cuda_ok(cudaStreamSynchronize(ctx().stream.s), "descriptor refresh")) return 1;
  return buildDescriptorTable();
}

TACSBVec *TACSAssembler::createVec() { return new TACSBVec(bs, nowned, ext_before, ext_after); }
TACSBVec *TACSAssembler::createNodeVec() { return new TACSBVec(3, nowned, ext_before, ext_after); }

int halo_forward(TACSAssembler *a, TACSBVec *v);  // comm.cpp: fill the external blocks of v; non-zero on failure

int TACSAssembler::setVariables(TACSBVec *q, TACSBVec *qdot, TACSBVec *qddot) {
  if (q) {
    vars->copyValues(q);
    vars_zero = false;
    if (size > 1 && halo_forward(this, vars)) return 1;
  }
  if (qdot) dvars->copyValues(qdot);
  if (qddot) {
    ddvars->copyValues(qddot);
    ddvars_zero = false;
    if (size > 1 && halo_forward(this, ddvars)) return 1;
  }
  return 0;
}
void TACSAssembler::zeroVariables() {
  vars->zeroEntries();
  dvars->zeroEntries();
  ddvars->zeroEntries();
  vars_zero = ddvars_zero = true;
}
void TACSAssembler::getNodes(TACSBVec *X) { X->copyValues(xpts); }
int TACSAssembler::setNodes(TACSBVec *X) {
  xpts->copyValues(X);
  if (size > 1 && halo_forward(this, xpts)) return 1;
  return evaluateAuxLoads();
}

// ---------------------------------------------------------------------------------------------
// auxiliary load elements
// ---------------------------------------------------------------------------------------------
void TACSAuxElements::addShellTraction(int elem_num, int order, const double *t, bool constant) {
  Load l;
  l.elem_num = elem_num;
  l.type = 0;
  l.order = order;
  const int nn = order * order;
  l.data.resize(3 * nn);
  for (int i = 0; i < nn; i++)
    for (int c = 0; c < 3; c++) l.data[3 * i + c] = constant ? t[c] : t[3 * i + c];
  loads.push_back(l);
}
void TACSAuxElements::addShellPressure(int elem_num, int order, const double *p, bool constant) {
  Load l;
  l.elem_num = elem_num;
  l.type = 1;
  l.order = order;
  const int nn = order * order;
  l.data.assign(3 * nn, 0.0);
  for (int i = 0; i < nn; i++) l.data[i] = constant ? p[0] : p[i];
  loads.push_back(l);
}

int TACSAssembler::setAuxElements(TACSAuxElements *aux) {
  if (aux) aux->incref();
  if (aux_elements) aux_elements->decref();
  aux_elements = aux;
  aux_groups.clear();
  if (!aux) return 0;
  HostPlan &P = *plan;
  // global element number -> local element, then (group, index inside the group)
  std::map<int, int> local_of;
  for (int e = 0; e < P.nelems; e++) local_of[P.elem_global[e]] = e;
  std::vector<std::pair<int, long>> where(P.nelems);  // local element -> (group, index in group)
  for (size_t gi = 0; gi < groups.size(); gi++)
    for (long k = 0; k < groups[gi].nelem; k++) where[groups[gi].local_elems[k]] = {(int)gi, k};
  for (size_t gi = 0; gi < groups.size(); gi++) {
    const ElemGroup &g = groups[gi];
    if (g.kind != ELEM_QUAD4_SHELL && g.kind != ELEM_QUAD9_SHELL) continue;
    const int order = g.kind == ELEM_QUAD4_SHELL ? 2 : 3;
    // loads of this group sorted by element (stable: insertion order inside an element, TACSAuxElements::sort)
    std::vector<std::pair<long, const TACSAuxElements::Load *>> mine;
    for (const auto &l : aux->loads) {
      auto it = local_of.find(l.elem_num);
      if (it == local_of.end() || where[it->second].first != (int)gi) continue;
      if (l.order != order) {
        fprintf(stderr, "tacs_b200: auxiliary load on element %d does not match the element's order\n", l.elem_num);
        return 1;
      }
      mine.emplace_back(where[it->second].second, &l);
    }
    if (mine.empty()) continue;
    std::stable_sort(mine.begin(), mine.end(), [](const auto &a, const auto &b) { return a.first < b.first; });
    std::unique_ptr<AuxGroup> ag(new AuxGroup());
    ag->group = (int)gi;
    ag->nloads = (int)mine.size();
    std::vector<int> elem, type, run_ptr;
    std::vector<long> slot;
    std::vector<double> data;
    for (size_t k = 0; k < mine.size(); k++) {
      if (k == 0 || mine[k].first != mine[k - 1].first) {
        run_ptr.push_back((int)k);
        slot.push_back(((long)g.node_base + mine[k].first * g.nn) * bs);
      }
      elem.push_back((int)mine[k].first);
      type.push_back(mine[k].second->type);
      data.insert(data.end(), mine[k].second->data.begin(), mine[k].second->data.end());
    }
    run_ptr.push_back((int)mine.size());
    ag->nruns = (int)slot.size();
    if (!ag->d_elem.upload(elem) || !ag->d_type.upload(type) || !ag->d_run_ptr.upload(run_ptr) ||
        !ag->d_slot.upload(slot) || !ag->d_data.upload(data) || !ag->d_loads.alloc((size_t)ag->nloads * 3 * g.nn))
      return 1;
    aux_groups.push_back(std::move(ag));
  }
  // every load must sit on a shell element of some rank; a load on a solid / unknown element is an error here
  for (const auto &l : aux->loads) {
    if (l.elem_num < 0 || l.elem_num >= (int)plan->gm->num_elements) {
      fprintf(stderr, "tacs_b200: auxiliary load on element %d: no such element\n", l.elem_num);
      return 1;
    }
  }
  return evaluateAuxLoads();
}

// the loads depend on the node locations only: evaluated when they are set and when the nodes change
int TACSAssembler::evaluateAuxLoads() {
  for (auto &ag : aux_groups) {
    const ElemGroup &g = groups[ag->group];
    KernelTimer kt(K_ELEMENT, g.nn == 4 ? "shell_aux_loads_kernel<2>" : "shell_aux_loads_kernel<3>");
    if (!cuda_ok(launch_shell_aux_loads(g.nn == 4 ? 2 : 3, ag->nloads, g.d_conn.ptr, ag->d_elem.ptr, ag->d_type.ptr,
                                        ag->d_data.ptr, g.d_tables.ptr, xpts->local(), ag->d_loads.ptr, ctx().stream),
                 "auxiliary loads")) return 1;
  }
  return 0;
}

// after the element kernels: lambda * load added to the residual staging slots of the loaded elements
int TACSAssembler::addAuxLoads(double lambda) {
  for (auto &ag : aux_groups) {
    KernelTimer kt(K_ELEMENT, "aux_add_kernel");
    if (!cuda_ok(launch_aux_add(ag->nruns, groups[ag->group].nn, ag->d_run_ptr.ptr, ag->d_slot.ptr, ag->d_loads.ptr,
                                lambda, Re.ptr, ctx().stream), "auxiliary loads")) return 1;
  }
  return 0;
}

void TACSAssembler::applyBCs(TACSBVec *v) {
  KernelTimer kt(K_BCS);
  cuda_ok(launch_vec_apply_bcs(bs, nbc_dev, d_bc_rows.ptr, d_bc_vars.ptr, d_bc_vals.ptr, nullptr, 1.0, v->owned(),
                               ctx().stream), "vec applyBCs");
}
void TACSAssembler::setBCs(TACSBVec *v) {
  KernelTimer kt(K_BCS);
  cuda_ok(launch_vec_set_bcs(bs, nbc_dev, d_bc_rows.ptr, d_bc_vars.ptr, d_bc_vals.ptr, 1.0, v->owned(),
                             ctx().stream), "vec setBCs");
}
void TACSAssembler::applyBCs(TACSParallelMat *m) { m->applyBCs(); }

int TACSAssembler::launchGroupRange(const ElemGroup &g, long e0, long e1, double alpha, double gamma,
                                    TACSParallelMat *mat, const double *vars_p, const double *ddvars_p) {
  const bool want_mat = mat != nullptr;
  const long b2 = (long)bs * bs, nu = upper_pairs(g.nn);
  ElemGroupArgs a;
  a.kind = g.kind;
  a.nelem = e1 - e0;
  a.conn = g.d_conn.ptr + e0 * g.nn;
  a.desc_index = g.d_desc.ptr + e0;
  a.desc_table = d_desc_table.ptr;
  a.tables = g.d_tables.ptr;
  a.Xpts = xpts->local();
  a.vars = vars_p;
  a.ddvars = ddvars_p;
  a.alpha = alpha;
  a.gamma = gamma;
  a.uncoupled = shells_uncoupled ? 1 : 0;
  a.Ke = want_mat ? Ke.ptr + ((size_t)g.block_base + (size_t)e0 * nu) * b2 : nullptr;
  a.Re = Re.ptr + ((size_t)g.node_base + (size_t)e0 * g.nn) * bs;
  a.upper = 1;
  a.geometric = geometric_pass ? 1 : 0;
  a.dmap = want_mat ? g.d_dmap.ptr + e0 * g.nn * g.nn : nullptr;
  a.direct = want_mat ? mat->vals_all.ptr : nullptr;
  KernelTimer kt(K_ELEMENT, element_kernel_name(a));
  return cuda_ok(launch_element_group(a, ctx().num_sms, ctx().stream), "element kernel") ? 0 : 1;
}

int TACSAssembler::launchElements(double alpha, double gamma, TACSParallelMat *mat, const double *vars_override,
                                  const double *ddvars_override, bool use_override) {
  if (mat && Ke.count < (size_t)total_blocks * bs * bs) {
    if (!Ke.alloc((size_t)total_blocks * bs * bs)) return 1;
  }
  const double *vp = use_override ? vars_override : (vars_zero ? nullptr : vars->local());
  const double *ap = use_override ? ddvars_override : (ddvars_zero ? nullptr : ddvars->local());
  for (auto &g : groups)
    if (launchGroupRange(g, 0, g.nelem, alpha, gamma, mat, vp, ap)) return 1;
  return 0;
}

int staging_exchange(TACSAssembler *a, bool with_blocks);  // comm.cpp: off-rank rows of Re (and Ke)

// TACSAssembler::assembleRes (TACSAssembler.cpp:4133-4242)
int TACSAssembler::assembleRes(TACSBVec *res, double lambda) {
  NvtxRange nvtx_range("tacs_b200::assembleRes");
  if (refreshDescriptors()) return 1;
  if (launchElements(1.0, 0.0, nullptr)) return 1;
  if (addAuxLoads(lambda)) return 1;
  if (size > 1 && staging_exchange(this, false)) return 1;
  {
    KernelTimer kt(K_GATHER_RES, gather_residual_kernel_name(bs));
    if (!cuda_ok(launch_gather_residual(bs, nowned, r_ptr.ptr, r_src.ptr, Re.ptr, res->owned(), ctx().num_sms,
                                        ctx().stream), "gather residual")) return 1;
  }
  {
    KernelTimer kt(K_BCS, "vec_apply_bcs_kernel");
    if (!cuda_ok(launch_vec_apply_bcs(bs, nbc_dev, d_bc_rows.ptr, d_bc_vars.ptr, d_bc_vals.ptr, vars->owned(),
                                      lambda, res->owned(), ctx().stream), "residual BCs")) return 1;
  }
  return 0;
}


// TACSAssembler::assembleJacobian (TACSAssembler.cpp:4291-4406)
//
// The elements are evaluated in chunks on the compute stream. The block gather is ordered by the last staging slot
// a block reads (HostPlan::buildMatrix), so the blocks completed by chunk c can be summed on the gather stream while
// chunk c+1 is being computed: the element kernels are FP64 / shared-memory bound, the gather is HBM bound. Blocks that
// need rows of other ranks are gathered after the exchange.
int TACSAssembler::assembleJacobian(double alpha, double beta, double gamma, TACSBVec *res, TACSParallelMat *A,
                                    double lambda, bool apply_bcs) {
  return assembleJacobianImpl(alpha, beta, gamma, res, A, lambda, apply_bcs, nullptr);
}

// setVariables(q) + assembleJacobian + the residual back on the host, with the state vector taken from (pinned) host
// memory: the upload is pipelined against the element kernels (see host_chunks). Returns when the residual has
// arrived; the block gather of the matrix may still be running (tacsb200_synchronize before the matrix is read).
int TACSAssembler::assembleJacobianHost(double alpha, double beta, double gamma, const double *q_host, double *res_host,
                                        TACSParallelMat *A, double lambda) {
  Context &c = ctx();
  if (!jvp_t) {
    // residual target of this entry point (shares the scratch vectors of addJacobianVecProduct)
    jvp_x = createVec(); jvp_a = createVec(); jvp_t = createVec();
    jvp_x->incref(); jvp_a->incref(); jvp_t->incref();
    if (!jvp_x->data.ptr || !jvp_a->data.ptr || !jvp_t->data.ptr) return 1;
  }
  TACSBVec *res = jvp_t;
  if (size > 1) {
    // several ranks: the halo needs the whole owned state first -- plain upload, then the usual path
    if (!cuda_ok(cudaMemcpyAsync(vars->owned(), q_host, vars->ownedSize() * sizeof(double), cudaMemcpyHostToDevice,
                                 c.stream), "state H2D")) return 1;
    vars_zero = false;
    if (halo_forward(this, vars)) return 1;
    if (assembleJacobianImpl(alpha, beta, gamma, res, A, lambda, true, nullptr)) return 1;
  } else {
    if (assembleJacobianImpl(alpha, beta, gamma, res, A, lambda, true, q_host)) return 1;
  }
  // residual to the host behind the residual kernels only (tail_evt), while the matrix is still being gathered
  return cuda_ok(cudaStreamWaitEvent(c.copy_stream, c.tail_evt, 0), "residual D2H") &&
                 cuda_ok(cudaMemcpyAsync(res_host, res->owned(), res->ownedSize() * sizeof(double),
                                         cudaMemcpyDeviceToHost, c.copy_stream), "residual D2H") &&
                 cuda_ok(cudaStreamSynchronize(c.copy_stream), "residual D2H")
             ? 0 : 1;
}

int TACSAssembler::assembleJacobianImpl(double alpha, double beta, double gamma, TACSBVec *res, TACSParallelMat *A,
                                        double lambda, bool apply_bcs, const double *q_host) {
  (void)beta;
  Context &c = ctx();
  NvtxRange nvtx_range("tacs_b200::assembleJacobian");
  if (refreshDescriptors()) return 1;
  if (Ke.count < (size_t)total_blocks * bs * bs && !Ke.alloc((size_t)total_blocks * bs * bs)) return 1;
  if (!elem_done_evt && !cuda_ok(cudaEventCreateWithFlags(&elem_done_evt, cudaEventDisableTiming), "event")) return 1;
  const std::vector<ElemChunk> &chunks = q_host ? host_chunks : this->chunks;
  if (q_host) {
    // state upload in pieces on the copy stream, behind the element kernels of the previous assembly (they read vars)
    cudaStreamWaitEvent(c.copy_stream, elem_done_evt, 0);
    const long nb = nowned;
    for (int k = 0; k < kStateChunks; k++) {
      if (!state_evt[k] && !cuda_ok(cudaEventCreateWithFlags(&state_evt[k], cudaEventDisableTiming), "event")) return 1;
      const long n0 = nb * k / kStateChunks * bs, n1 = nb * (k + 1) / kStateChunks * bs;
      if (n1 > n0 && !cuda_ok(cudaMemcpyAsync(vars->owned() + n0, q_host + n0, (size_t)(n1 - n0) * sizeof(double),
                                               cudaMemcpyHostToDevice, c.copy_stream), "state H2D")) return 1;
      cudaEventRecord(state_evt[k], c.copy_stream);
    }
    vars_zero = false;
  }
  const double *vp = vars_zero ? nullptr : vars->local(), *ap = ddvars_zero ? nullptr : ddvars->local();
  auto gather = [&](long g0, long g1, cudaStream_t st) -> int {
    if (g1 <= g0) return 0;
    KernelTimer kt(K_GATHER_MAT, gather_blocks_kernel_name(bs), st);
    return cuda_ok(launch_gather_blocks(bs, g1 - g0, gb_blk.ptr + g0, gb_ptr.ptr + g0, gb_src.ptr, Ke.ptr,
                                        A->vals_all.ptr, c.num_sms, st), "gather blocks") ? 0 : 1;
  };
  long gdone = 0;
  bool forked = false;
  int state_waited = -1;
  for (size_t k = 0; k < chunks.size(); k++) {
    const ElemChunk &ch = chunks[k];
    if (q_host && ch.need_state > state_waited) {
      cudaStreamWaitEvent(c.stream.s, state_evt[ch.need_state], 0);
      state_waited = ch.need_state;
    }
    if (launchGroupRange(groups[ch.group], ch.e0, ch.e1, alpha, gamma, A, vp, ap)) return 1;
    if (overlap_gather && k + 1 < chunks.size() && ch.gather_end > gdone) {
      if (c.chunk_evt.size() <= k) {
        c.chunk_evt.resize(k + 1, nullptr);
      }
      if (!c.chunk_evt[k] && !cuda_ok(cudaEventCreateWithFlags(&c.chunk_evt[k], cudaEventDisableTiming), "event"))
        return 1;
      cudaEventRecord(c.chunk_evt[k], c.stream.s);
      cudaStreamWaitEvent(c.gather_stream, c.chunk_evt[k], 0);
      if (gather(gdone, ch.gather_end, c.gather_stream)) return 1;
      gdone = ch.gather_end;
      forked = true;
    }
  }
  if (q_host && state_waited < kStateChunks - 1) cudaStreamWaitEvent(c.stream.s, state_evt[kStateChunks - 1], 0);
  cudaEventRecord(elem_done_evt, c.stream.s);
  if (res && addAuxLoads(lambda)) return 1;
  if (size > 1) {
    // the blocks that read local staging slots only come first in the plan: they are gathered on the second stream
    // while the staged rows owned elsewhere are packed, sent and received on this one
    // (the exchange is enqueued first: its pack kernels and the NCCL kernel must get their CTAs before the gather
    // fills the GPU)
    if (staging_exchange(this, true)) return 1;
    if (local_gather_end > gdone) {
      cudaStreamWaitEvent(c.gather_stream, elem_done_evt, 0);
      if (gather(gdone, local_gather_end, c.gather_stream)) return 1;
      gdone = local_gather_end;
      forked = true;
    }
  }
  if (res) {
    {
      KernelTimer kt(K_GATHER_RES, gather_residual_kernel_name(bs));
      if (!cuda_ok(launch_gather_residual(bs, nowned, r_ptr.ptr, r_src.ptr, Re.ptr, res->owned(), c.num_sms,
                                          c.stream), "gather residual")) return 1;
    }
    KernelTimer kt(K_BCS, "vec_apply_bcs_kernel");
    if (!cuda_ok(launch_vec_apply_bcs(bs, nbc_dev, d_bc_rows.ptr, d_bc_vars.ptr, d_bc_vals.ptr, vars->owned(),
                                      lambda, res->owned(), c.stream), "residual BCs")) return 1;
  }
  // from here on only the staging area and the matrix are touched (Context::tail_evt)
  cudaEventRecord(c.tail_evt, c.stream.s);
  if (gather(gdone, num_gather_blocks, c.stream)) return 1;
  if (forked) {
    cudaEventRecord(c.gather_evt, c.gather_stream);
    cudaStreamWaitEvent(c.stream.s, c.gather_evt, 0);
  }
  if (apply_bcs) A->applyBCs();
  c.tail_seq = c.stream.seq;
  return 0;
}

// TACSAssembler::assembleMatType (TACSAssembler.cpp:4418-4504). The element matrices come from getMatType:
// shells TACSShellElement.h:644-660 (stiffness: alpha = 1, mass: gamma = 1), solids TACSElement3D.cpp:316-380 --
// for the linear models of this path these are the alpha / gamma parts of addJacobian. The geometric stiffness
// needs the nonlinear strain terms and is not on the device path: non-zero return, the caller keeps the reference.
int TACSAssembler::assembleMatType(int matType, TACSParallelMat *A, bool apply_bcs) {
  double alpha = 0.0, gamma = 0.0;
  if (matType == 1) {
    alpha = 1.0;
  } else if (matType == 2) {
    gamma = 1.0;
  } else if (matType == 3) {
    // TACS_GEOMETRIC_STIFFNESS_MATRIX: solids evaluate it from the stress of the current state (elem_phases.cuh
    // solid_geo_*). The shell's (a directional derivative of the tangent of its nonlinear model,
    // TACSShellElement.h:643-760) is not on the device path.
    for (auto &g : groups)
      if (g.kind != ELEM_HEX8 && g.kind != ELEM_HEX27) {
        fprintf(stderr, "tacs_b200: assembleMatType(TACS_GEOMETRIC_STIFFNESS_MATRIX) is evaluated on the device for "
                        "solid elements only\n");
        return 1;
      }
    alpha = 1.0;
    geometric_pass = true;
    const int rc = assembleJacobian(alpha, 0.0, 0.0, nullptr, A, 1.0, apply_bcs);
    geometric_pass = false;
    return rc;
  } else {
    fprintf(stderr, "tacs_b200: assembleMatType(%d): stiffness (1), mass (2) and geometric stiffness (3) matrices are "
                    "evaluated on the device\n", matType);
    return 1;
  }
  return assembleJacobian(alpha, 0.0, gamma, nullptr, A, 1.0, apply_bcs);
}

// TACSAssembler::addJacobianVecProduct (TACSAssembler.cpp:5416-5496): y <- y + scale * J x without forming J.
// Both element models are linear, J = alpha K + gamma M does not depend on the state, and the residual-only
// element kernels evaluate K v + M a for any pair of nodal vectors: J x = K (alpha x) + M (gamma x).
int TACSAssembler::addJacobianVecProduct(double scale, double alpha, double beta, double gamma, TACSBVec *x,
                                         TACSBVec *y, bool apply_bcs) {
  (void)beta;
  NvtxRange nvtx_range("tacs_b200::addJacobianVecProduct");
  if (refreshDescriptors()) return 1;
  if (!jvp_x) {
    jvp_x = createVec();
    jvp_a = createVec();
    jvp_t = createVec();
    jvp_x->incref();
    jvp_a->incref();
    jvp_t->incref();
    if (!jvp_x->data.ptr || !jvp_a->data.ptr || !jvp_t->data.ptr) return 1;
  }
  jvp_x->copyValues(x);
  jvp_x->scale(alpha);
  if (size > 1 && halo_forward(this, jvp_x)) return 1;
  if (gamma != 0.0) {
    jvp_a->copyValues(x);
    jvp_a->scale(gamma);
    if (size > 1 && halo_forward(this, jvp_a)) return 1;
  }
  if (launchElements(1.0, 0.0, nullptr, jvp_x->local(), gamma != 0.0 ? jvp_a->local() : nullptr, true)) return 1;
  if (size > 1 && staging_exchange(this, false)) return 1;
  {
    KernelTimer kt(K_GATHER_RES, gather_residual_kernel_name(bs));
    if (!cuda_ok(launch_gather_residual(bs, nowned, r_ptr.ptr, r_src.ptr, Re.ptr, jvp_t->owned(), ctx().num_sms,
                                        ctx().stream), "gather residual")) return 1;
  }
  y->axpy(scale, jvp_t);
  if (apply_bcs) applyBCs(y);
  return 0;
}

// ---------------------------------------------------------------------------------------------
// matrix: sparsity (TACSAssembler::createMat :3384, computeLocalNodeToNodeCSR :1899-2053,
// TACSMatDistribute::computeLocalCSR :453-692) and the block gather plan
// ---------------------------------------------------------------------------------------------
TACSParallelMat *TACSAssembler::createMat() {
  TACSParallelMat *m = new TACSParallelMat(this);
  if (m->Aloc.bsize == 0) {
    delete m;
    return nullptr;
  }
  return m;
}

TACSParallelMat::TACSParallelMat(TACSAssembler *a) {
  assembler = a;
  a->incref();
  HostPlan &P = *a->plan;
  if (P.buildMatrix()) return;
  np = P.np;
  ext_col_nodes = P.ext_col_nodes;
  auto copy = [](BCSRPattern &d, const HostBCSR &h) {
    d.bsize = h.bsize; d.nrows = h.nrows; d.ncols = h.ncols;
    d.rowp = h.rowp; d.cols = h.cols;
  };
  copy(Aloc, P.Aloc);
  copy(Bext, P.Bext);
  const long nnzA = Aloc.nnzb(), nnzB = Bext.nnzb();
  const size_t b2 = (size_t)Aloc.bsize * Aloc.bsize;
  // one value array [Aloc | Bext]: the direct map of the element kernels and the gather plan index into it
  bool ok = Aloc.d_rowp.upload(Aloc.rowp) && Aloc.d_cols.upload(Aloc.cols) && vals_all.alloc(b2 * (nnzA + nnzB));
  if (ok) {
    Aloc.d_vals.ptr = vals_all.ptr;
    Aloc.d_vals.count = b2 * nnzA;
    Bext.d_vals.ptr = vals_all.ptr + b2 * nnzA;
    Bext.d_vals.count = b2 * nnzB;
  }
  if (ok && nnzB > 0)
    ok = Bext.d_rowp.upload(Bext.rowp) && Bext.d_cols.upload(Bext.cols) && x_ext.alloc((size_t)Aloc.bsize * Bext.ncols);
  if (ok) ok = a->uploadMatPlan() == 0;
  if (ok && !getenv("TACSB200_SPMV_NATURAL_ORDER")) ok = Aloc.buildRowOrder();
  if (ok && nnzB > 0) ok = Bext.buildNonEmptyRows();
  if (ok && a->size > 1) ok = comm_setup_exchange(x_cols, P.cols) == 0;
  if (!ok) {
    Aloc.bsize = 0;
    return;
  }
  // Bext rows of constrained nodes: row index relative to np
  {
    std::vector<int> rows;
    for (int r : a->h_bc_rows) rows.push_back((r >= np) ? r - np : -1);
    if (!d_bc_rows_ext.upload(rows)) {
      Aloc.bsize = 0;
      return;
    }
  }
  zeroEntries();
}

TACSParallelMat::~TACSParallelMat() { assembler->decref(); }

void TACSParallelMat::zeroEntries() {
  if (vals_all.count)
    cuda_ok(cudaMemsetAsync(vals_all.ptr, 0, vals_all.count * sizeof(double), ctx().stream), "zeroEntries");
}

// Rows of a mesh of quadratic elements come in a few length classes that alternate from one node to the next (hex27:
// 27 / 45 / 75 / 125 blocks); with one thread per scalar row a warp then waits for its longest row. When the mean row
// length is well below the maximum the SpMV visits the rows sorted by length instead.
// Measured on the full-size configurations: Quad9 cylinder (rows of 9 / 15 / 25 blocks) 7.12 -> 6.77 ms, 0.81 -> 0.86 of
// the HBM peak; 100^3 hex27 9.4 -> 11.2 ms -- with 3x3 blocks the rows of one class are two nodes apart and the lanes
// of a warp lose the x entries their natural neighbours share -- so the order is used for 6x6 blocks only.
bool BCSRPattern::buildRowOrder() {
  if (nrows < 1024 || bsize != 6) return true;
  long maxlen = 0;
  for (int i = 0; i < nrows; i++) maxlen = std::max<long>(maxlen, rowp[i + 1] - rowp[i]);
  if (maxlen == 0 || (double)nnzb() / nrows > 0.85 * (double)maxlen) return true;  // uniform enough
  std::vector<int> order(nrows);
  for (int i = 0; i < nrows; i++) order[i] = i;
  std::stable_sort(order.begin(), order.end(),
                   [&](int a, int b) { return rowp[a + 1] - rowp[a] > rowp[b + 1] - rowp[b]; });
  return d_order.upload(order);
}

bool BCSRPattern::buildNonEmptyRows() {
  std::vector<int> rows;
  for (int i = 0; i < nrows; i++)
    if (rowp[i + 1] > rowp[i]) rows.push_back(i);
  order_rows = (int)rows.size();
  if (rows.empty()) return true;
  return d_order.upload(rows);
}

int TACSParallelMat::copyValues(TACSParallelMat *o) {
  if (!o || o->vals_all.count != vals_all.count) {
    fprintf(stderr, "tacs_b200: copyValues: matrices do not share a non-zero pattern\n");
    return 1;
  }
  return cuda_ok(cudaMemcpyAsync(vals_all.ptr, o->vals_all.ptr, vals_all.count * sizeof(double),
                                 cudaMemcpyDeviceToDevice, ctx().stream), "copyValues") ? 0 : 1;
}
int TACSParallelMat::scale(double alpha) {
  KernelTimer kt(K_VEC, "scale_kernel");
  return cuda_ok(launch_scale((long)vals_all.count, alpha, vals_all.ptr, ctx().num_sms, ctx().stream), "scale") ? 0 : 1;
}
int TACSParallelMat::axpy(double alpha, TACSParallelMat *o) {
  if (!o || o->vals_all.count != vals_all.count) {
    fprintf(stderr, "tacs_b200: axpy: matrices do not share a non-zero pattern\n");
    return 1;
  }
  KernelTimer kt(K_VEC, "axpy_kernel");
  return cuda_ok(launch_axpy((long)vals_all.count, alpha, o->vals_all.ptr, vals_all.ptr, ctx().num_sms, ctx().stream),
                 "axpy") ? 0 : 1;
}

// ---------------------------------------------------------------------------------------------
// TACSSchurMat: [B E; F C] view of an assembled matrix in the reference's local ordering
// ---------------------------------------------------------------------------------------------
TACSSchurMat::TACSSchurMat(TACSParallelMat *src, int _nb, const int *b_nodes, int _nc, const int *c_nodes,
                           const int *const rowp[4], const int *const cols[4]) {
  source = src;
  source->incref();
  nb = _nb;
  nc = _nc;
  bsize = src->Aloc.bsize;
  TACSAssembler *a = src->assembler;
  if (a->size != 1) {
    fprintf(stderr, "tacs_b200: the TACSSchurMat view is implemented for one rank\n");
    return;
  }
  const int nrows[4] = {nb, nb, nc, nc}, ncols[4] = {nb, nc, nb, nc};
  const int *rnodes[4] = {b_nodes, b_nodes, c_nodes, c_nodes}, *cnodes[4] = {b_nodes, c_nodes, b_nodes, c_nodes};
  const BCSRPattern &A = src->Aloc;
  std::vector<char> used(A.nnzb(), 0);
  for (int k = 0; k < 4; k++) {
    BCSRPattern &P = blk[k];
    P.bsize = bsize;
    P.nrows = nrows[k];
    P.ncols = ncols[k];
    P.rowp.assign(rowp[k], rowp[k] + nrows[k] + 1);
    P.cols.assign(cols[k], cols[k] + P.rowp[nrows[k]]);
    std::vector<int> srcv(P.nnzb(), -1);
    for (int i = 0; i < nrows[k]; i++) {
      const int gr = rnodes[k][i];
      if (gr < 0 || gr >= A.nrows) continue;
      const int *cb = A.cols.data() + A.rowp[gr], *ce = A.cols.data() + A.rowp[gr + 1];
      for (int p = P.rowp[i]; p < P.rowp[i + 1]; p++) {
        const int gc = cnodes[k][P.cols[p]];
        const int *it = std::lower_bound(cb, ce, gc);
        if (it != ce && *it == gc) {
          srcv[p] = (int)(it - A.cols.data());
          used[srcv[p]] = 1;
        }
      }
    }
    const size_t b2 = (size_t)bsize * bsize;
    if (!P.d_rowp.upload(P.rowp) || !P.d_cols.upload(P.cols) || !src_blk[k].upload(srcv) ||
        !vals[k].alloc(b2 * P.nnzb()))
      return;
    P.d_vals.ptr = vals[k].ptr;
    P.d_vals.count = vals[k].count;
  }
  for (char u : used)
    if (!u) missing++;
  if (missing) {
    fprintf(stderr, "tacs_b200: TACSSchurMat view: %ld blocks of the assembled matrix have no place in [B E; F C]\n",
            missing);
    return;
  }
  std::vector<int> bn(b_nodes, b_nodes + nb), cn(c_nodes, c_nodes + nc);
  if (!d_bnodes.upload(bn) || !d_cnodes.upload(cn)) return;
  xb = new TACSBVec(bsize, nb, 0, 0);
  yb = new TACSBVec(bsize, nb, 0, 0);
  xc = new TACSBVec(bsize, nc, 0, 0);
  yc = new TACSBVec(bsize, nc, 0, 0);
  xb->incref(); yb->incref(); xc->incref(); yc->incref();
  ok = true;
}

TACSSchurMat::~TACSSchurMat() {
  if (xb) xb->decref();
  if (yb) yb->decref();
  if (xc) xc->decref();
  if (yc) yc->decref();
  source->decref();
}

int TACSSchurMat::update() {
  if (!ok) return 1;
  for (int k = 0; k < 4; k++) {
    KernelTimer kt(K_GATHER_MAT, "permute_blocks_kernel");
    if (!cuda_ok(launch_permute_blocks(bsize * bsize, blk[k].nnzb(), src_blk[k].ptr, source->vals_all.ptr, vals[k].ptr,
                                       ctx().num_sms, ctx().stream), "schur view update")) return 1;
  }
  return 0;
}

int TACSSchurMat::getValues(int which, double *host) {
  if (!ok || which < 0 || which > 3) return 1;
  return blk[which].d_vals.download(host, blk[which].d_vals.count) ? 0 : 1;
}

int TACSSchurMat::mult(TACSBVec *x, TACSBVec *y) {
  if (!ok) return 1;
  Context &c = ctx();
  bool good = true;
  auto spmv = [&](int k, TACSBVec *in, TACSBVec *out, int add) {
    if (blk[k].nrows == 0) return;
    if (blk[k].nnzb() == 0 && add) return;
    KernelTimer kt(K_SPMV, spmv_kernel_name(bsize, blk[k].nrows, blk[k].nnzb(), blk[k].d_cols.ptr, blk[k].d_vals.ptr, add, nullptr));
    good = good && cuda_ok(launch_spmv(bsize, blk[k].nrows, blk[k].nnzb(), blk[k].d_rowp.ptr, blk[k].d_cols.ptr, blk[k].d_vals.ptr,
                                       in->owned(), out->owned(), add, c.num_sms, c.stream), "schur spmv");
  };
  {
    KernelTimer kt(K_HALO, "pack_blocks_kernel");
    good = good && cuda_ok(launch_pack_blocks(bsize, nb, d_bnodes.ptr, x->owned(), xb->owned(), c.num_sms, c.stream), "x_b");
  }
  if (nc > 0) {
    KernelTimer kt(K_HALO, "pack_blocks_kernel");
    good = good && cuda_ok(launch_pack_blocks(bsize, nc, d_cnodes.ptr, x->owned(), xc->owned(), c.num_sms, c.stream), "x_c");
  }
  spmv(0, xb, yb, 0);  // y_b = B x_b
  spmv(2, xb, yc, 0);  // y_c = F x_b
  spmv(3, xc, yc, 1);  // y_c += C x_c
  spmv(1, xc, yb, 1);  // y_b += E x_c
  {
    KernelTimer kt(K_HALO, "unpack_blocks_kernel");
    good = good && cuda_ok(launch_unpack_blocks(bsize, nb, d_bnodes.ptr, yb->owned(), y->owned(), 0, c.num_sms, c.stream),
                           "y_b");
  }
  if (nc > 0) {
    KernelTimer kt(K_HALO, "unpack_blocks_kernel");
    good = good && cuda_ok(launch_unpack_blocks(bsize, nc, d_cnodes.ptr, yc->owned(), y->owned(), 0, c.num_sms, c.stream),
                           "y_c");
  }
  return good ? 0 : 1;
}

bool BCSRPattern::ValuesView::download(double *host, size_t n) const {
  if (n == 0) return true;
  return cuda_ok(cudaMemcpyAsync(host, ptr, n * sizeof(double), cudaMemcpyDeviceToHost, ctx().stream), "D2H") &&
         cuda_ok(cudaStreamSynchronize(ctx().stream), "D2H sync");
}

// device copies of the direct map and the gather plan (HostPlan::buildMatrix), shared by all matrices
int TACSAssembler::uploadMatPlan() {
  if (mat_plan_ready) return 0;
  HostPlan &P = *plan;
  for (size_t gi = 0; gi < groups.size(); gi++) {
    ElemGroup &g = groups[gi];
    const size_t count = (size_t)g.nelem * g.nn * g.nn;
    if (!g.d_dmap.upload(P.dmap.data() + P.group_pair_base[gi], count)) return 1;
  }
  if (!gb_blk.upload(P.gb_blk) || !gb_ptr.upload(P.gb_ptr) || !gb_src.upload(P.gb_src)) return 1;
  num_gather_blocks = (long)P.gb_blk.size();
  local_gather_end = size > 1 ? P.gatherEnd(P.local_blocks) : 0;
  // Element chunks. TACSB200_CHUNKS: chunks per group (default 8; chunks of fewer than 2^19 elements are merged).
  // TACSB200_OVERLAP_KINDS: element families (bit kind-1) whose gather overlaps the element kernel. Default hex8
  // only: its kernel leaves room for a gather CTA on every SM (2 CTAs x 128 threads x 222 registers); the kernels of
  // the other families fill the register file, and a co-resident gather would cost them a CTA per SM.
  int nchunk = 8;
  unsigned overlap_kinds = 1u << (ELEM_HEX8 - 1);
  long chunk_min = 1L << 19;   // TACSB200_CHUNK_MIN: the tests force small chunks on small meshes
  if (const char *env = getenv("TACSB200_CHUNK_MIN")) chunk_min = std::max(1L, atol(env));
  if (const char *env = getenv("TACSB200_CHUNKS")) nchunk = std::max(1, atoi(env));
  if (const char *env = getenv("TACSB200_OVERLAP_KINDS")) overlap_kinds = (unsigned)strtoul(env, nullptr, 0);
  chunks.clear();
  overlap_gather = false;
  for (size_t gi = 0; gi < groups.size(); gi++) {
    const ElemGroup &g = groups[gi];
    const bool ov = (overlap_kinds >> (g.kind - 1)) & 1u;
    // at least 2^19 elements per chunk: on 1M hex8 elements the overlap gains nothing (4.39 against 4.37 ms), on 8M
    // (C4 on one GPU) 3 % (33.4 against 34.5 ms)
    long n = ov ? std::min<long>(nchunk, std::max<long>(1, g.nelem / chunk_min)) : 1;
    const long nu = upper_pairs(g.nn);
    for (long k = 0; k < n; k++) {
      ElemChunk ch;
      ch.group = (int)gi;
      ch.e0 = g.nelem * k / n;
      ch.e1 = g.nelem * (k + 1) / n;
      ch.gather_end = P.gatherEnd(g.block_base + ch.e1 * nu);
      if (ch.e1 > ch.e0) chunks.push_back(ch);
    }
    if (n > 1) overlap_gather = true;
  }
  // chunks of the host-state entry point: kStateChunks per group, each knowing the last state piece it reads
  host_chunks.clear();
  for (size_t gi = 0; gi < groups.size(); gi++) {
    const ElemGroup &g = groups[gi];
    const long nu = upper_pairs(g.nn);
    const long n = std::min<long>(kStateChunks, std::max<long>(1, g.nelem / 8192));
    for (long k = 0; k < n; k++) {
      ElemChunk ch;
      ch.group = (int)gi;
      ch.e0 = g.nelem * k / n;
      ch.e1 = g.nelem * (k + 1) / n;
      if (ch.e1 <= ch.e0) continue;
      ch.gather_end = ((overlap_kinds >> (g.kind - 1)) & 1u) ? P.gatherEnd(g.block_base + ch.e1 * nu) : 0;
      int max_node = 0;
      for (long e = ch.e0; e < ch.e1; e++) {
        const int le = g.local_elems[e];
        for (int q = P.elem_ptr[le]; q < P.elem_ptr[le + 1]; q++) max_node = std::max(max_node, P.elem_conn_local[q]);
      }
      // piece k holds the owned nodes [nowned k / K, nowned (k+1) / K)
      int piece = 0;
      while (piece < kStateChunks - 1 && (long)nowned * (piece + 1) / kStateChunks <= max_node) piece++;
      ch.need_state = piece;
      host_chunks.push_back(ch);
    }
  }
  mat_plan_ready = true;
  return 0;
}

TACSBVec *TACSParallelMat::createVec() { return new TACSBVec(Aloc.bsize, Aloc.nrows, 0, 0); }

// TACSParallelMat::applyBCs (TACSParallelMat.cpp:343-374)
void TACSParallelMat::applyBCs() {
  TACSAssembler *a = assembler;
  KernelTimer kt(K_BCS, "mat_apply_bcs_kernel");
  cuda_ok(launch_mat_apply_bcs(Aloc.bsize, a->nbc_dev, a->d_bc_rows.ptr, a->d_bc_vars.ptr, Aloc.d_rowp.ptr,
                               Aloc.d_cols.ptr, Aloc.d_vals.ptr, 0, ctx().stream), "mat applyBCs");
  if (Bext.nnzb() > 0) {
    KernelTimer kt2(K_BCS, "mat_apply_bcs_kernel");
    cuda_ok(launch_mat_apply_bcs(Bext.bsize, a->nbc_dev, d_bc_rows_ext.ptr, a->d_bc_vars.ptr, Bext.d_rowp.ptr,
                                 Bext.d_cols.ptr, Bext.d_vals.ptr, -1, ctx().stream), "mat applyBCs ext");
  }
}

int spmv_halo_begin(TACSParallelMat *A, TACSBVec *x);  // comm.cpp
void spmv_halo_end(TACSParallelMat *A);

int TACSParallelMat::multTranspose(TACSBVec *x, TACSBVec *y) {
  if (assembler->size > 1) {
    fprintf(stderr, "tacs_b200: multTranspose is implemented for one rank\n");
    return 1;
  }
  if (!d_tidx.ptr) {
    std::vector<int> tidx(Aloc.nnzb(), -1);
    for (int i = 0; i < Aloc.nrows; i++)
      for (int k = Aloc.rowp[i]; k < Aloc.rowp[i + 1]; k++) {
        const int j = Aloc.cols[k];
        const int *b = Aloc.cols.data() + Aloc.rowp[j], *e = Aloc.cols.data() + Aloc.rowp[j + 1];
        const int *it = std::lower_bound(b, e, i);
        if (it == e || *it != i) {
          fprintf(stderr, "tacs_b200: multTranspose: the pattern is not structurally symmetric at (%d, %d)\n", i, j);
          return 1;
        }
        tidx[k] = (int)(it - Aloc.cols.data());
      }
    if (!d_tidx.upload(tidx)) return 1;
  }
  KernelTimer kt(K_SPMV, Aloc.bsize == 6 ? "spmv_transpose_kernel<6>" : "spmv_transpose_kernel<3>");
  return cuda_ok(launch_spmv_transpose(Aloc.bsize, Aloc.nrows, Aloc.d_rowp.ptr, Aloc.d_cols.ptr, d_tidx.ptr,
                                       Aloc.d_vals.ptr, x->owned(), y->owned(), ctx().num_sms, ctx().stream),
                 "spmv transpose") ? 0 : 1;
}

// y = zs z + sign (A x): the product with the vector update in its epilogue (smoother steps, Krylov residuals)
int TACSParallelMat::multFused(TACSBVec *x, TACSBVec *y, double sign, double zs, TACSBVec *z) {
  const bool dist = assembler->size > 1;
  int rc = 0;
  if (dist) rc = spmv_halo_begin(this, x);
  {
    KernelTimer kt(K_SPMV, spmv_kernel_name(Aloc.bsize, Aloc.nrows, Aloc.nnzb(), Aloc.d_cols.ptr, Aloc.d_vals.ptr, 2, Aloc.d_order.ptr));
    if (!cuda_ok(launch_spmv_fused(Aloc.bsize, Aloc.nrows, Aloc.nnzb(), Aloc.d_rowp.ptr, Aloc.d_cols.ptr, Aloc.d_vals.ptr,
                                   x->owned(), y->owned(), 2, sign, zs, z->owned(), Aloc.d_order.ptr, ctx().num_sms,
                                   ctx().stream), "spmv fused")) rc = 1;
  }
  if (dist) {
    spmv_halo_end(this);
    if (Bext.nnzb() > 0) {
      KernelTimer kt(K_SPMV, spmv_kernel_name(Bext.bsize, Bext.order_rows, Bext.nnzb(), Bext.d_cols.ptr, Bext.d_vals.ptr, 3, Bext.d_order.ptr));
      if (!cuda_ok(launch_spmv_fused(Bext.bsize, Bext.order_rows, Bext.nnzb(), Bext.d_rowp.ptr, Bext.d_cols.ptr, Bext.d_vals.ptr,
                                     x_ext.ptr, y->owned() + (size_t)Bext.bsize * np, 3, sign, 0.0, nullptr,
                                     Bext.d_order.ptr, ctx().num_sms, ctx().stream), "spmv ext fused")) rc = 1;
    }
  }
  return rc;
}

// TACSParallelMat::mult (TACSParallelMat.cpp:248-265): y = Aloc x + Bext x_ext, halo overlapped
int TACSParallelMat::mult(TACSBVec *x, TACSBVec *y) {
  NvtxRange nvtx_range("tacs_b200::TACSParallelMat::mult");
  const bool dist = assembler->size > 1;
  int rc = 0;
  // every rank takes part in the column halo (a rank without external columns may still have to send)
  if (dist) rc = spmv_halo_begin(this, x);
  {
    KernelTimer kt(K_SPMV, spmv_kernel_name(Aloc.bsize, Aloc.nrows, Aloc.nnzb(), Aloc.d_cols.ptr, Aloc.d_vals.ptr, 0, Aloc.d_order.ptr));
    if (!cuda_ok(launch_spmv_fused(Aloc.bsize, Aloc.nrows, Aloc.nnzb(), Aloc.d_rowp.ptr, Aloc.d_cols.ptr, Aloc.d_vals.ptr,
                                   x->owned(), y->owned(), 0, 1.0, 0.0, nullptr, Aloc.d_order.ptr, ctx().num_sms,
                                   ctx().stream), "spmv")) rc = 1;
  }
  if (dist) {
    spmv_halo_end(this);
    if (Bext.nnzb() > 0) {
      KernelTimer kt(K_SPMV, spmv_kernel_name(Bext.bsize, Bext.order_rows, Bext.nnzb(), Bext.d_cols.ptr, Bext.d_vals.ptr, 1, Bext.d_order.ptr));
      // rows with an off-rank column only (Bext.d_order): the others would be read and rewritten for nothing
      if (!cuda_ok(launch_spmv_fused(Bext.bsize, Bext.order_rows, Bext.nnzb(), Bext.d_rowp.ptr, Bext.d_cols.ptr, Bext.d_vals.ptr,
                                     x_ext.ptr, y->owned() + (size_t)Bext.bsize * np, 1, 1.0, 0.0, nullptr,
                                     Bext.d_order.ptr, ctx().num_sms, ctx().stream), "spmv ext"))
        rc = 1;
    }
  }
  return rc;
}

}  // namespace tb2
