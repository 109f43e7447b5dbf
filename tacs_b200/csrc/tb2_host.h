// Host-side object model of tacs_b200: a C++ mirror of the reference's plug-in interface for the
// assembly + Krylov-operator path (same class names, argument meaning and error behaviour), with the
// data living in B200 HBM and the work done by the kernels in kernels.cu.
//
//   TACSMaterialProperties / TACSOrthotropicPly      src/constitutive/TACSMaterialProperties.h:42-200
//   TACSIsoShellConstitutive                         src/constitutive/TACSIsoShellConstitutive.h:31-36
//   TACSCompositeShellConstitutive                   src/constitutive/TACSCompositeShellConstitutive.h:27-32
//   TACSSolidConstitutive                            src/constitutive/TACSSolidConstitutive.h:34-36
//   TACSShellNaturalTransform / RefAxisTransform     src/elements/shell/TACSShellElementTransform.h:21-215
//   TACSQuad4Shell / TACSQuad9Shell                  src/elements/shell/TACSShellElementDefs.h:19-25
//   TACSElement3D + TACSLinearElasticity3D + HexaBasis  src/elements/TACSElement3D.h:26
//   TACSCreator                                      src/TACSCreator.h:46-120
//   TACSAssembler                                    src/TACSAssembler.h:61-523
//   TACSBVec / TACSParallelMat / BCSRMat / GMRES     src/bpmat/TACSBVec.h:67, TACSParallelMat.h:52, BCSRMat.h:34, KSM.h:392
//
// Elements are descriptors, not virtual device code: each recognised family has a hand-written
// kernel, and per-object constants (tangent stiffness, mass moments, transform) are evaluated once
// on the host and stored in a device table row. Unsupported combinations are rejected when the
// assembler is created (non-zero return / stderr message, like TACSAssembler.cpp:513-581).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <string>
#include <vector>

#include <memory>

#include <algorithm>
#include <cmath>

#include "kernels.h"
#include "plan.h"

#include <nvtx3/nvToolsExt.h>  // header-only NVTX 3: a no-op unless a profiler is attached

namespace tb2 {

// NVTX range around a host-side phase (assembleJacobian, assembleRes, mult, GMRES::solve ...): what Nsight Systems /
// ncu --nvtx show instead of the reference's MPI_Wtime monitors (KSM.cpp:765, TACSAssembler timing prints)
struct NvtxRange {
  explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};

// ---------------------------------------------------------------------------------------------
// runtime context: one process drives one GPU
// ---------------------------------------------------------------------------------------------
// The compute stream counts its uses: every enqueue site reads it through the implicit conversion, so
// `stream.seq` identifies the last operation put on it (see Context::tail_seq).
struct CountedStream {
  cudaStream_t s = nullptr;
  mutable long seq = 0;
  operator cudaStream_t() const {
    ++seq;
    return s;
  }
};

struct Context {
  int device = -1;
  int num_sms = 0;
  CountedStream stream;                // compute stream
  cudaStream_t comm_stream = nullptr;  // halo / collective stream
  // Host <-> device vector copies normally ride the compute stream. assembleJacobian ends with the block gather and
  // the matrix boundary conditions, which touch no vector: it records `tail_evt` in front of them and remembers the
  // stream position after them, so that a getArray/setArray issued while that tail is still the last thing on the
  // compute stream runs on `copy_stream` behind `tail_evt` only -- the residual returns to the host while the matrix
  // is still being gathered.
  cudaStream_t copy_stream = nullptr;
  // The block gather of a chunk of elements runs on `gather_stream` while the compute stream evaluates the next chunk
  // (TACSAssembler::assembleJacobian); chunk_evt orders the two, gather_evt joins them again.
  cudaStream_t gather_stream = nullptr;
  std::vector<cudaEvent_t> chunk_evt;
  cudaEvent_t gather_evt = nullptr;
  cudaEvent_t tail_evt = nullptr;
  long tail_seq = -1;
  bool tail_is_matrix_only() const { return tail_seq >= 0 && stream.seq == tail_seq; }
  int rank = 0, size = 1;
  void *nccl_comm = nullptr;  // ncclComm_t when size > 1
  // scalar all-reduce over NVLink peer memory (kernels.h PeerExchange): device copy of the descriptor, null when the
  // peers' buffers could not be mapped (the NCCL all-reduce is used then)
  PeerExchange *d_peer = nullptr;
  long kernel_launches = 0;   // launches of tacs_b200 kernels since the last reset
};
Context &ctx();

// Optional per-kernel timing with CUDA events on the launching stream (bench.py's roofline numbers).
enum KernelId { K_ELEMENT = 0, K_GATHER_RES, K_GATHER_MAT, K_BCS, K_SPMV, K_VEC, K_DOT, K_HALO, K_COUNT };
struct KernelTimer {
  // name: static string of the kernel launched; stream: where it is launched (default: the compute stream)
  explicit KernelTimer(KernelId id, const char *name = nullptr, cudaStream_t stream = nullptr);
  cudaStream_t stream;
  ~KernelTimer();
  int slot;
};
void profile_enable(int on);
// sums the recorded launches: ms[K_COUNT], count[K_COUNT]; clears the log
int profile_collect(double *ms, long *count);
const char *profile_named();  // "kernel name|launches|ms" lines of the last profile_collect
int ctx_init(int device);  // 0 ok; prints to stderr and returns non-zero when no usable GPU is present
bool cuda_ok(cudaError_t err, const char *what);

// A shell descriptor row is membrane-bending uncoupled when its B block (entries 6..11) vanishes against the scale
// sqrt(max|A| max|D|) of a coupling term. Symmetric laminates evaluated with the reference's ply sum leave
// |B| ~ 1e-20 of rounding noise there; dropping it changes no tangent entry at double precision.
inline bool shell_desc_uncoupled(const double *row) {
  double amax = 0.0, dmax = 0.0, bmax = 0.0;
  for (int i = 0; i < 6; i++) {
    amax = std::max(amax, std::fabs(row[i]));
    bmax = std::max(bmax, std::fabs(row[6 + i]));
    dmax = std::max(dmax, std::fabs(row[12 + i]));
  }
  return bmax <= 1e-14 * std::sqrt(amax * dmax);
}

template <class T>
struct DeviceArray {
  T *ptr = nullptr;
  size_t count = 0;
  DeviceArray() {}
  DeviceArray(const DeviceArray &) = delete;
  DeviceArray &operator=(const DeviceArray &) = delete;
  DeviceArray(DeviceArray &&o) noexcept : ptr(o.ptr), count(o.count) {
    o.ptr = nullptr;
    o.count = 0;
  }
  ~DeviceArray() { release(); }
  void release() {
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    count = 0;
  }
  bool alloc(size_t n) {
    release();
    count = n;
    if (n == 0) return true;
    // 16 bytes of slack: bulk copies of spans that start or end off a 16-byte boundary are rounded outwards
    return cuda_ok(cudaMalloc(&ptr, n * sizeof(T) + 16), "cudaMalloc");
  }
  bool upload(const T *host, size_t n) {
    if (n != count && !alloc(n)) return false;
    if (n == 0) return true;
    return cuda_ok(cudaMemcpyAsync(ptr, host, n * sizeof(T), cudaMemcpyHostToDevice, ctx().stream), "H2D") &&
           cuda_ok(cudaStreamSynchronize(ctx().stream), "H2D sync");
  }
  bool upload(const std::vector<T> &v) { return upload(v.data(), v.size()); }
  bool download(T *host, size_t n) const {
    if (n == 0) return true;
    return cuda_ok(cudaMemcpyAsync(host, ptr, n * sizeof(T), cudaMemcpyDeviceToHost, ctx().stream), "D2H") &&
           cuda_ok(cudaStreamSynchronize(ctx().stream), "D2H sync");
  }
};

// device side of an ExchangePlan (comm.cpp): packed send indices and buffers
struct DeviceExchange {
  std::vector<int> send_peers, send_ptr, recv_peers, recv_ptr;
  DeviceArray<int> d_send_idx;
  DeviceArray<double> sendbuf;  // sized on first use for the largest chunk
  int sendTotal() const { return send_ptr.empty() ? 0 : send_ptr.back(); }
  int recvTotal() const { return recv_ptr.empty() ? 0 : recv_ptr.back(); }
};

// ---------------------------------------------------------------------------------------------
// reference-counted base (TACSObject, src/TACSObject.h:106-135)
// ---------------------------------------------------------------------------------------------
class Object {
 public:
  Object() : refs(0) {}
  virtual ~Object() {}
  void incref() { refs++; }
  void decref() {
    if (--refs <= 0) delete this;
  }
  virtual const char *getObjectName() { return "TACSObject"; }

 private:
  std::atomic<int> refs;
};

// ---------------------------------------------------------------------------------------------
// constitutive
// ---------------------------------------------------------------------------------------------
class TACSMaterialProperties : public Object {
 public:
  TACSMaterialProperties(double rho, double specific_heat, double E, double nu, double ys, double alpha,
                         double kappa);
  TACSMaterialProperties(double rho, double specific_heat, double E1, double E2, double E3, double nu12,
                         double nu13, double nu23, double G12, double G13, double G23);
  double getDensity() const { return rho; }
  void evalTangentStiffness2D(double C[6]) const;
  void evalTangentStiffness3D(double C[21]) const;
  bool isotropic;
  double rho, specific_heat, E, nu, G, ys, alpha, kappa;
  double E1, E2, E3, nu12, nu13, nu23, G12, G13, G23;
};

class TACSOrthotropicPly : public Object {
 public:
  TACSOrthotropicPly(double ply_thickness, TACSMaterialProperties *props);
  ~TACSOrthotropicPly();
  void calculateQbar(double angle, double Qbar[6]) const;
  void calculateAbar(double angle, double Abar[3]) const;
  double getDensity() const { return rho; }
  double plyThickness, rho;
  double Q11, Q12, Q22, Q44, Q55, Q66, C12, C16, C26, C66;
  TACSMaterialProperties *properties;
};

class TACSConstitutive : public Object {
 public:
  virtual int getNumStresses() = 0;
  // 22 entries for shells (A,B,D,As,drill), 21 for solids; constant per object on this path
  virtual void evalTangentStiffness(double C[]) = 0;
  // fill one 32-double device descriptor row (layout in elem_phases.cuh)
  virtual void fillDescriptor(double d[]) = 0;
};

class TACSShellConstitutive : public TACSConstitutive {
 public:
  int getNumStresses() { return 9; }
  virtual void evalMassMoments(double moments[3]) = 0;
  void fillDescriptor(double d[]);
  // The reference evaluates the tangent stiffness (drill entry included) on every assembly, so a change takes effect
  // at once (TACSShellConstitutive.cpp:57-73). Here the constants live in a device table: the version counter lets
  // every assembler rebuild its table at the start of the next assembly.
  static void setDrillingRegularization(double k) {
    DRILLING_REGULARIZATION = k;
    drill_version++;
  }
  static long drill_version;
  static double getDrillingRegularization() { return DRILLING_REGULARIZATION; }
  static double DRILLING_REGULARIZATION;
};

class TACSIsoShellConstitutive : public TACSShellConstitutive {
 public:
  TACSIsoShellConstitutive(TACSMaterialProperties *props, double t, double tOffset, double kcorr);
  ~TACSIsoShellConstitutive();
  void evalTangentStiffness(double C[]);
  void evalMassMoments(double moments[3]);
  TACSMaterialProperties *properties;
  double t, tOffset, kcorr;
};

class TACSCompositeShellConstitutive : public TACSShellConstitutive {
 public:
  TACSCompositeShellConstitutive(int num_plies, TACSOrthotropicPly **plies, const double *thickness,
                                 const double *angles, double kcorr, double tOffset);
  ~TACSCompositeShellConstitutive();
  void evalTangentStiffness(double C[]);
  void evalMassMoments(double moments[3]);
  std::vector<TACSOrthotropicPly *> ply_props;
  std::vector<double> ply_thickness, ply_angles;
  double kcorr, tOffset;
};

// A shell constitutive object given by its constant tangent stiffness (22 entries A, B, D, As, drill) and mass moments:
// what any TACSShellConstitutive of the reference evaluates to on this path (evalTangentStiffness / evalMassMoments
// do not depend on the point for the classes of the path). Used by the reference-side shim, which reads these numbers
// off the caller's own constitutive objects.
class TACSRawShellConstitutive : public TACSShellConstitutive {
 public:
  TACSRawShellConstitutive(const double C[22], const double moments[3]);
  void evalTangentStiffness(double C[]);
  void evalMassMoments(double moments[3]);
  void fillDescriptor(double d[]);
  double Cs[22], mom[3];
};

class TACSSolidConstitutive : public TACSConstitutive {
 public:
  TACSSolidConstitutive(TACSMaterialProperties *props, double t);
  ~TACSSolidConstitutive();
  int getNumStresses() { return 6; }
  void evalTangentStiffness(double C[]);
  double evalDensity() { return raw ? rho_raw : t * properties->getDensity(); }
  void fillDescriptor(double d[]);
  TACSMaterialProperties *properties;  // null: raw form, stiffness and density given directly
  double t;
  // raw form (shim): constant 21-entry tangent stiffness and density
  TACSSolidConstitutive(const double C[21], double density);
  bool raw = false;
  double Craw[21], rho_raw = 0.0;
};

// ---------------------------------------------------------------------------------------------
// transforms, models, bases, elements
// ---------------------------------------------------------------------------------------------
class TACSShellTransform : public Object {
 public:
  int kind;  // 0 natural, 1 reference axis
  double axis[3];
};
class TACSShellNaturalTransform : public TACSShellTransform {
 public:
  TACSShellNaturalTransform();
};
class TACSShellRefAxisTransform : public TACSShellTransform {
 public:
  explicit TACSShellRefAxisTransform(const double axis[3]);
  void getRefAxis(double a[3]) { a[0] = axis[0]; a[1] = axis[1]; a[2] = axis[2]; }
};

class TACSElementBasis : public Object {
 public:
  int order;  // 2: TACSLinearHexaBasis, 3: TACSQuadraticHexaBasis
};
class TACSLinearHexaBasis : public TACSElementBasis {
 public:
  TACSLinearHexaBasis() { order = 2; }
};
class TACSQuadraticHexaBasis : public TACSElementBasis {
 public:
  TACSQuadraticHexaBasis() { order = 3; }
};

class TACSElementModel : public Object {};
class TACSLinearElasticity3D : public TACSElementModel {
 public:
  explicit TACSLinearElasticity3D(TACSSolidConstitutive *con);
  ~TACSLinearElasticity3D();
  TACSSolidConstitutive *stiff;
};

class TACSElement : public Object {
 public:
  virtual int getVarsPerNode() = 0;
  virtual int getNumNodes() = 0;
  int getNumVariables() { return getVarsPerNode() * getNumNodes(); }
  virtual int kernelKind() = 0;  // ElemKind
  virtual void fillDescriptor(double d[]) = 0;
  // Evaluate `count` elements with this descriptor on the device (host arrays in/out), element-major:
  // Xpts[count][3nn], vars/ddvars[count][nv] (ddvars may be null), res[count][nv], mat[count][nv*nv]
  // (res or mat may be null). Mirrors TACSElement::addJacobian (TACSElement.h:450) for a batch.
  int addJacobianBatch(int count, double alpha, double beta, double gamma, const double *Xpts,
                       const double *vars, const double *dvars, const double *ddvars, double *res, double *mat);
};

class TACSShellElement : public TACSElement {
 public:
  TACSShellElement(int order, TACSShellTransform *t, TACSShellConstitutive *c);
  ~TACSShellElement();
  int getVarsPerNode() { return 6; }
  int getNumNodes() { return order * order; }
  int kernelKind() { return order == 2 ? ELEM_QUAD4_SHELL : ELEM_QUAD9_SHELL; }
  void fillDescriptor(double d[]);
  int order;
  TACSShellTransform *transform;
  TACSShellConstitutive *con;
};

class TACSElement3D : public TACSElement {
 public:
  TACSElement3D(TACSElementModel *model, TACSElementBasis *basis);
  ~TACSElement3D();
  int getVarsPerNode() { return 3; }
  int getNumNodes() { return basis->order * basis->order * basis->order; }
  int kernelKind() { return basis->order == 2 ? ELEM_HEX8 : ELEM_HEX27; }
  void fillDescriptor(double d[]);
  TACSLinearElasticity3D *model;
  TACSElementBasis *basis;
};

// ---------------------------------------------------------------------------------------------
// block-CSR storage (BCSRMatData, src/bpmat/BCSRMatImpl.h:27-46)
// ---------------------------------------------------------------------------------------------
struct BCSRPattern {
  int bsize = 0, nrows = 0, ncols = 0;
  std::vector<int> rowp, cols;  // host copies (bit-exact parity target)
  DeviceArray<int> d_rowp, d_cols;
  // rows listed by descending length (stable), built when the row lengths differ much: the SpMV then hands the lanes
  // of a warp rows of (nearly) equal length. Null pointer: natural order.
  DeviceArray<int> d_order;
  int order_rows = -1;  // number of rows listed in d_order (-1: none; < nrows: the empty rows are left out)
  bool buildRowOrder();
  bool buildNonEmptyRows();  // Bext: most rows of a partition have no off-rank column; list the ones that do
  // bsize^2 * nnzb values, 64-bit offsets: a view into the owning matrix's single value array [Aloc | Bext]
  struct ValuesView {
    double *ptr = nullptr;
    size_t count = 0;
    bool download(double *host, size_t n) const;
  } d_vals;
  long nnzb() const { return rowp.empty() ? 0 : rowp[nrows]; }
};

class TACSAssembler;

// TACSBVec: owned block vector with room for the external (ghost) blocks in local order
// [ext < range | owned | ext >= range] so that element kernels index it directly.
class TACSBVec : public Object {
 public:
  TACSBVec(int bs, int nowned, int next_before, int next_after);
  int bsize, nowned, ext_before, ext_after;
  DeviceArray<double> data;  // (ext_before + nowned + ext_after) * bsize
  double *owned() { return data.ptr + (size_t)ext_before * bsize; }
  double *local() { return data.ptr; }
  long ownedSize() const { return (long)nowned * bsize; }
  long localSize() const { return (long)(ext_before + nowned + ext_after) * bsize; }
  // TACSVec interface (KSM.h:91-115); reductions are over owned entries and all ranks
  double norm();
  double dot(TACSBVec *y);
  int mdot(TACSBVec **ys, double *out, int n);  // non-zero (and NaN results) when a launch / reduction failed
  void axpy(double alpha, TACSBVec *x);
  void axpby(double alpha, double beta, TACSBVec *x);
  void scale(double alpha);
  void copyValues(TACSBVec *x);
  void zeroEntries();
  int getArray(double *host_out);        // device -> host copy of the owned entries
  int setArray(const double *host_in);   // host -> device
};

class TACSParallelMat : public Object {
 public:
  explicit TACSParallelMat(TACSAssembler *a);
  ~TACSParallelMat();
  TACSAssembler *assembler;
  BCSRPattern Aloc, Bext;
  int np = 0;                      // first owned row with an off-rank column (rows of Bext start here)
  std::vector<int> ext_col_nodes;  // ascending global node ids of the external columns
  DeviceArray<double> vals_all;  // [Aloc blocks | Bext blocks]: the block index space of the direct map / gather plan
  DeviceArray<double> x_ext;  // external column values for the SpMV halo
  DeviceArray<int> d_bc_rows_ext;  // Bext row (owned row - np) of each merged BC, or -1
  DeviceExchange x_cols;
  void zeroEntries();
  // TACSParallelMat::copyValues / scale / axpy / addDiag (TACSParallelMat.cpp:198-246): same-pattern matrices only
  int copyValues(TACSParallelMat *other);
  int scale(double alpha);
  int axpy(double alpha, TACSParallelMat *other);
  int mult(TACSBVec *x, TACSBVec *y);
  int multFused(TACSBVec *x, TACSBVec *y, double sign, double zs, TACSBVec *z);  // y = zs z + sign (A x)
  // TACSParallelMat::multTranspose (TACSParallelMat.cpp:267-290), one rank: the mirror-block index is built on first use
  int multTranspose(TACSBVec *x, TACSBVec *y);
  DeviceArray<int> d_tidx;
  void applyBCs();
  TACSBVec *createVec();
};

// ---------------------------------------------------------------------------------------------
// creator and assembler
// ---------------------------------------------------------------------------------------------
class TACSCreator : public Object {
 public:
  explicit TACSCreator(int vars_per_node);
  ~TACSCreator();
  void setGlobalConnectivity(int num_nodes, int num_elements, const int *ptr, const int *conn,
                             const int *elem_id_nums);
  void setBoundaryConditions(int num_bcs, const int *bc_nodes, const int *bc_ptr, const int *bc_vars,
                             const double *bc_vals);
  void setNodes(const double *Xpts);
  void setElements(int num_elems, TACSElement **elems);
  int partitionMesh(int split_size, const int *part);
  // The mesh is already in its final numbering (an existing single-rank TACSAssembler): no first-touch renumbering
  bool keep_numbering = false;
  int getNodeNums(const int **new_nodes);
  int getElementPartition(const int **part);
  TACSAssembler *createTACS();
  // host-only plan of `rank` of `size` (no GPU): shares the mesh preparation with createTACS
  class PlanObject *createPlan(int rank, int size);
  int prepareMesh(int rank, int size, std::shared_ptr<GlobalMesh> &gm, std::vector<int> &kinds);
  int comm_size() const;
  int plan_size = 0;  // > 0: partition for this many ranks without a communicator (host-only planning)

  int vars_per_node;
  int num_nodes = 0, num_elements = 0;
  std::vector<int> elem_node_ptr, elem_node_conn, elem_id_nums;
  std::vector<int> bc_nodes, bc_ptr, bc_vars;
  std::vector<double> bc_vals;
  std::vector<double> Xpts;
  std::vector<TACSElement *> elements;
  std::vector<int> partition, new_nodes, owned_nodes, owned_elements;
};

class PlanObject : public Object {
 public:
  HostPlan plan;
};

struct ElemGroup {
  int kind = 0, nn = 0;
  long nelem = 0;
  std::vector<int> local_elems;  // local element indices in this group (ascending)
  DeviceArray<int> d_conn, d_desc;
  DeviceArray<int> d_dmap;  // [nelem][nn*nn] direct map (HostPlan::dmap), uploaded with the first matrix
  DeviceArray<unsigned char> d_tables;
  long block_base = 0;  // first staging slot (units of one bs x bs block) of this group
  long node_base = 0;   // first residual staging slot (units of one node block)
};

// TACSAuxElements (src/TACSAuxElements.h:52-100) restricted to the state-independent shell loads of the reference:
// TACSShellTraction and TACSShellPressure (src/elements/shell/TACSShellTraction.h, TACSShellPressure.h), each bound to
// an element number of the creator's global numbering.
class TACSAuxElements : public Object {
 public:
  struct Load {
    int elem_num;   // global element number (TACSAuxElements::addElement(num, elem))
    int type;       // 0 traction, 1 pressure
    int order;      // 2 Quad4, 3 Quad9
    std::vector<double> data;  // 3 nn tractions (node by node) or nn pressures
  };
  std::vector<Load> loads;
  void addShellTraction(int elem_num, int order, const double *t, bool constant);
  void addShellPressure(int elem_num, int order, const double *p, bool constant);
};

class TACSAssembler : public Object {
 public:
  TACSAssembler();
  ~TACSAssembler();
  int getVarsPerNode() { return bs; }
  int getNumNodes() { return nlocal; }
  int getNumOwnedNodes() { return nowned; }
  int getNumElements() { return nelems; }
  TACSBVec *createVec();
  TACSBVec *createNodeVec();
  TACSParallelMat *createMat();
  int setVariables(TACSBVec *q, TACSBVec *qdot, TACSBVec *qddot);
  void zeroVariables();
  void getNodes(TACSBVec *X);
  int setNodes(TACSBVec *X);
  // TACSAssembler::setAuxElements (src/TACSAssembler.h:158): loads added to the residual of assembleRes /
  // assembleJacobian, scaled by the load factor lambda
  int setAuxElements(TACSAuxElements *aux);
  void applyBCs(TACSBVec *v);
  void applyBCs(TACSParallelMat *m);
  void setBCs(TACSBVec *v);
  int assembleRes(TACSBVec *res, double lambda);
  int assembleJacobian(double alpha, double beta, double gamma, TACSBVec *res, TACSParallelMat *A,
                       double lambda, bool apply_bcs = true);
  // matType as ElementMatrixType (elements/TACSElementTypes.h:113): 1 stiffness, 2 mass
  int assembleMatType(int matType, TACSParallelMat *A, bool apply_bcs);
  // y <- y + scale (alpha K + gamma M) x, matrix free (TACSAssembler.cpp:5416-5496)
  int addJacobianVecProduct(double scale, double alpha, double beta, double gamma, TACSBVec *x, TACSBVec *y,
                            bool apply_bcs);

  // --- data -------------------------------------------------------------------------------
  int bs = 0, rank = 0, size = 1;
  int nowned = 0, nlocal = 0, nelems = 0;
  int ext_before = 0, ext_after = 0;
  std::vector<int> owner_range;  // size+1
  std::vector<TACSElement *> elems;          // per local element
  std::vector<TACSElement *> distinct;       // distinct descriptors (table rows)
  std::vector<int> elem_desc;                // per local element: row of the descriptor table
  // boundary conditions in the creator's order (global node, bit mask, values[bs]) + merged device form
  std::vector<int> bc_nodes, bc_vars;
  std::vector<double> bc_vals;
  int nbc_dev = 0;
  std::vector<int> h_bc_rows;  // owned-row index of each merged BC (or -1)
  DeviceArray<int> d_bc_rows, d_bc_vars;  // owned-row index of each merged BC (or -1), mask
  DeviceArray<double> d_bc_vals;
  DeviceArray<int> d_bc_local;            // local node index (for state vectors in local order)
  // state
  TACSBVec *xpts = nullptr, *vars = nullptr, *dvars = nullptr, *ddvars = nullptr;
  bool vars_zero = true, ddvars_zero = true;
  TACSBVec *jvp_x = nullptr, *jvp_a = nullptr, *jvp_t = nullptr;  // scratch of addJacobianVecProduct
  bool shells_uncoupled = false;  // all shell descriptors have a zero membrane-bending block
  // element groups, descriptor table, staging and residual gather plan
  std::vector<ElemGroup> groups;
  DeviceArray<double> d_desc_table;
  DeviceArray<double> Ke, Re;
  long total_blocks = 0, total_node_slots = 0;
  DeviceArray<int> r_ptr, r_src;
  // matrix gather plan (HostPlan::gb_*), shared by every matrix of this assembler; uploaded with the first matrix
  DeviceArray<int> gb_blk, gb_ptr, gb_src;
  long num_gather_blocks = 0;
  bool mat_plan_ready = false;
  // element chunks of assembleJacobian: [e0, e1) of group `group`; every gathered block below gather_end is complete
  // once the chunk (and all chunks before it) has been evaluated
  struct ElemChunk { int group; long e0, e1, gather_end; int need_state = -1; };
  long local_gather_end = 0;  // gathered blocks below it read no received staging slot (multi-rank)
  std::vector<ElemChunk> chunks;
  // Finer chunks of the host-state entry point (assembleJacobianHost): the state vector arrives from pinned host memory
  // in kStateChunks pieces on the copy stream, and element chunk k only waits for the piece that holds the last node
  // it references (need_state), so the upload hides behind the element kernels of the chunks before it.
  static constexpr int kStateChunks = 8;
  std::vector<ElemChunk> host_chunks;
  cudaEvent_t state_evt[kStateChunks] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t elem_done_evt = nullptr;
  int assembleJacobianHost(double alpha, double beta, double gamma, const double *q_host, double *res_host,
                           TACSParallelMat *A, double lambda);
  int assembleJacobianImpl(double alpha, double beta, double gamma, TACSBVec *res, TACSParallelMat *A, double lambda,
                           bool apply_bcs, const double *q_host);
  bool overlap_gather = false;
  bool geometric_pass = false;
  // auxiliary loads of the local elements, per shell group: sorted by element, one run of loads per loaded element
  TACSAuxElements *aux_elements = nullptr;
  struct AuxGroup {
    int group = 0, nloads = 0, nruns = 0;
    DeviceArray<int> d_elem, d_type, d_run_ptr;
    DeviceArray<long> d_slot;
    DeviceArray<double> d_data, d_loads;
  };
  std::vector<std::unique_ptr<AuxGroup>> aux_groups;
  int evaluateAuxLoads();
  int addAuxLoads(double lambda);  // set by assembleMatType(TACS_GEOMETRIC_STIFFNESS_MATRIX) around its element launches
  int launchGroupRange(const ElemGroup &g, long e0, long e1, double alpha, double gamma, TACSParallelMat *mat,
                       const double *vars_p, const double *ddvars_p);
  int uploadMatPlan();
  std::unique_ptr<HostPlan> plan;
  DeviceExchange x_state, x_rows, x_blocks;
  int localNode(int global) const;
  int finalize();  // build device data after the creator filled the host arrays
  int buildDescriptorTable();   // constitutive / transform constants of the distinct element objects -> device
  int refreshDescriptors();     // rebuilds the table when setDrillingRegularization was called since
  long desc_version = -1;
  int launchElements(double alpha, double gamma, TACSParallelMat *mat, const double *vars_override = nullptr,
                     const double *ddvars_override = nullptr, bool use_override = false);
};

// Device mirror of TACSSchurMat (src/bpmat/TACSSchurMat.h:58-125): the four blocks [B E; F C] in the reference's own
// local ordering -- interior unknowns "b" first, interface unknowns "c" -- with the non-zero patterns and the
// node -> local index maps read from a live reference TACSSchurMat (the shim does that). The values are not
// assembled a second time: update() gathers them from the value array of an assembled TACSParallelMat through a
// block permutation computed once, so the matrix carries the boundary conditions of the source. mult follows
// TACSSchurMat::mult (TACSSchurMat.cpp:883-938): y_b = B x_b + E x_c, y_c = F x_b + C x_c on permuted copies of x.
class TACSSchurMat : public Object {
 public:
  TACSSchurMat(TACSParallelMat *src, int nb, const int *b_nodes, int nc, const int *c_nodes,
               const int *const rowp[4], const int *const cols[4]);
  ~TACSSchurMat();
  bool ok = false;
  TACSParallelMat *source;
  int nb = 0, nc = 0, bsize = 0;
  BCSRPattern blk[4];           // B (nb x nb), E (nb x nc), F (nc x nb), C (nc x nc)
  DeviceArray<double> vals[4];
  DeviceArray<int> src_blk[4];  // per block: index into the source's [Aloc | Bext] value array, or -1 (stays zero)
  long missing = 0;             // blocks of the source pattern that have no place in the four blocks (must be 0)
  DeviceArray<int> d_bnodes, d_cnodes;
  TACSBVec *xb = nullptr, *xc = nullptr, *yb = nullptr, *yc = nullptr;
  int update();
  int mult(TACSBVec *x, TACSBVec *y);
  int getValues(int which, double *host);
};

// reduction scratch shared by the vector kernels (tb2_host.cpp)
extern double *g_dot_partial, *g_dot_out, *g_dot_host;
bool dot_buffers();

// TACSChebyshevSmoother (bpmat/TACSParallelMat.h:180-217, TACSParallelMat.cpp:871-1113): polynomial smoother /
// preconditioner built from products with the matrix only; the spectral radius comes from Gershgorin's discs.
// applyFactor runs as a chain of SpMV launches whose epilogues carry the vector updates (krylov.cpp).
class TACSChebyshevSmoother : public Object {
 public:
  TACSChebyshevSmoother(TACSParallelMat *mat, int degree, double lower_factor, double upper_factor, int iters);
  ~TACSChebyshevSmoother();
  int factor();
  int applyFactor(TACSBVec *x, TACSBVec *y);
  double gershgorin();
  TACSParallelMat *mat;
  int degree, iters;
  double lower_factor, upper_factor, alpha = 0.0, beta = 0.0, rho = 0.0;
  std::vector<double> roots, coef;  // roots of the shifted Chebyshev polynomial, monomial coefficients of q
  TACSBVec *res = nullptr, *h0 = nullptr, *h1 = nullptr;
};

// GMRES (reference interface: src/bpmat/KSM.h:392-440): restarted, right-preconditioned (optionally flexible)
// GMRES. Device-resident design (krylov.cpp): the Hessenberg column, the plane rotations and the least-squares
// right-hand side live in device memory, every Gram-Schmidt coefficient is consumed by the next kernel through a
// device pointer, the body of an iteration is a CUDA graph, and the host reads the residual history once every
// `check_every` iterations.
class GMRES : public Object {
 public:
  enum OrthoType { CLASSICAL_GRAM_SCHMIDT = 0, MODIFIED_GRAM_SCHMIDT = 1 };
  GMRES(TACSParallelMat *mat, int m, int nrestart, TACSChebyshevSmoother *pc = nullptr, bool flexible = false);
  ~GMRES();
  void setTolerances(double rtol, double atol) { this->rtol = rtol; this->atol = atol; }
  void setOrthoType(int t) { ortho = t; dropGraphs(); }
  // KSMPrintStdout(descript, rank, freq) (KSM.cpp:239-283): "%s[%3d]: %15.8e" lines on rank 0 every freq iterations
  void setMonitor(const char *descript, int freq);
  // GMRES::setTimeMonitor (KSM.cpp:765): per-solve device times of the preconditioner / orthogonalisation / total
  void setTimeMonitor() { monitor_time = true; }
  int solve(TACSBVec *b, TACSBVec *x, int zero_guess);
  int getIterCount() { return iters; }
  double getResidualNorm() { return resnorm; }
  TACSParallelMat *mat;
  int m, nrestart, iters = 0;
  double rtol = 1e-8, atol = 1e-30, resnorm = 0.0;
  int ortho = MODIFIED_GRAM_SCHMIDT;
  int check_every = 4;            // iterations between reads of the residual history (TACSB200_GMRES_CHECK)
  bool use_graphs = true;         // TACSB200_GMRES_GRAPHS=0 launches the kernels one by one
  std::vector<TACSBVec *> W, Z;   // Krylov basis; Z: preconditioned directions of the flexible variant
  TACSBVec *work = nullptr;       // M^{-1} W[i] of the regular variant
  TACSChebyshevSmoother *pc = nullptr;
  bool flexible = false;
  // device state of the least-squares problem: hcol[m+2], R[(m+1) x m] column major, cs/sn[m], g[m+1], resnorm[m],
  // y[m], sumsq[1]; pinned host mirror of the residual history
  DeviceArray<double> d_state;
  DeviceArray<unsigned> d_ticket;
  double *d_hcol = nullptr, *d_R = nullptr, *d_cs = nullptr, *d_sn = nullptr, *d_g = nullptr, *d_res = nullptr,
         *d_y = nullptr, *d_sumsq = nullptr;
  double *h_res = nullptr;  // pinned
  std::vector<cudaGraphExec_t> graphs;  // one per iteration index, captured on first use
  std::vector<int> graph_launches;      // kernels inside each graph (for the launch counter)
  std::string monitor_name;
  int monitor_freq = 0;
  bool monitor_time = false;
  double t_pc = 0.0, t_ortho = 0.0, t_total = 0.0;  // ms, last solve (setTimeMonitor)

 private:
  int iterationBody(int i);   // enqueue iteration i on the compute stream
  int runIteration(int i);    // through its graph when enabled
  void dropGraphs();
};

}  // namespace tb2
