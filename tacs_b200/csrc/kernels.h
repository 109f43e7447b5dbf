// Launchers of the sm_100a kernels (kernels.cu). Plain C++ interface used by the host objects.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace tb2 {

enum ElemKind { ELEM_QUAD4_SHELL = 1, ELEM_QUAD9_SHELL = 2, ELEM_HEX8 = 3, ELEM_HEX27 = 4 };

inline int elem_kind_nodes(int kind) {
  switch (kind) {
    case ELEM_QUAD4_SHELL: return 4;
    case ELEM_QUAD9_SHELL: return 9;
    case ELEM_HEX8: return 8;
    case ELEM_HEX27: return 27;
  }
  return 0;
}
inline int elem_kind_bs(int kind) { return (kind == ELEM_QUAD4_SHELL || kind == ELEM_QUAD9_SHELL) ? 6 : 3; }

// One homogeneous group of elements (same family); all pointers are device pointers.
struct ElemGroupArgs {
  int kind;
  long nelem;
  const int *conn;           // [nelem][nn] local node numbers
  const int *desc_index;     // [nelem] row of desc_table
  const double *desc_table;  // [ndesc][32] constitutive / transform constants
  const void *tables;        // family shape-function tables
  const double *Xpts;        // [nlocal][3]
  const double *vars;        // [nlocal][bs] or null (zero state)
  const double *ddvars;      // [nlocal][bs] or null
  double alpha, gamma;
  int uncoupled;             // 1: every shell descriptor of the group has a zero membrane-bending block
  double *Ke;                // matrix staging or null (residual only); layout selected by `upper`
  double *Re;                // staging [nelem][nn*bs] or null
  // upper = 0: Ke is [nelem][nn][nn][bs*bs], every directed node pair (element-level interface).
  // upper = 1: Ke is [nelem][nn(nn+1)/2][bs*bs], node pairs i <= j only (plan.h upper_index); a directed pair (i, j)
  //            whose dmap entry is >= 0 is written to block dmap[..] of `direct` instead (the BCSR value array) and
  //            pairs i > j without a direct target are not written at all: the gather reads the mirror transposed.
  int upper;
  // 1: the tangent is the geometric stiffness of the current state (solids only; TACS_GEOMETRIC_STIFFNESS_MATRIX)
  int geometric;
  const int *dmap;           // [nelem][nn*nn] or null (no direct targets)
  double *direct;            // value array of the matrix being assembled
};

// destination of the bs x bs block of the directed node pair (i, j) of element e; null: not stored
template <int NN, int B2>
__host__ __device__ inline double *pair_block_dst(const ElemGroupArgs &g, long e, int i, int j, int dm) {
  if (!g.upper) return g.Ke + ((e * NN + i) * NN + j) * B2;
  if (dm >= 0) return g.direct + (long)dm * B2;
  if (i <= j) return g.Ke + (e * (NN * (NN + 1) / 2) + (i * NN - i * (i - 1) / 2 + (j - i))) * B2;
  return nullptr;
}

size_t elem_tables_bytes(int kind);
void elem_tables_build(int kind, void *host_dst);
cudaError_t launch_element_group(const ElemGroupArgs &g, int num_sms, cudaStream_t s);
const char *element_kernel_name(const ElemGroupArgs &g);
inline const char *gather_blocks_kernel_name(int bs) { return bs == 6 ? "gather_blocks36_kernel" : "gather_blocks9_kernel"; }
inline const char *gather_residual_kernel_name(int bs) { return bs == 6 ? "gather_residual_kernel<6>" : "gather_residual_kernel<3>"; }

// blocks gb_blk[0..nblocks) of A sum their staging sources src[ptr[g]..ptr[g+1]) (2*slot + transposed flag)
cudaError_t launch_gather_blocks(int bs, long nblocks, const int *blk, const int *ptr, const int *src,
                                 const double *Ke, double *A, int num_sms, cudaStream_t s);
// auxiliary shell loads (traction type 0: data = t[3 nn]; pressure type 1: data[0..nn) = p): loads[k][3 nn]
cudaError_t launch_shell_aux_loads(int order, int nloads, const int *conn, const int *elem, const int *type,
                                   const double *data, const void *tables, const double *Xpts, double *loads,
                                   cudaStream_t s);
// Re[slot[r] + 6 a + c] += lambda * sum of the loads run_ptr[r] .. run_ptr[r+1]) of one element
cudaError_t launch_aux_add(int nruns, int nn, const int *run_ptr, const long *slot, const double *loads, double lambda,
                           double *Re, cudaStream_t s);
cudaError_t launch_gather_residual(int bs, long nnodes, const int *ptr, const int *src, const double *Re,
                                   double *res, int num_sms, cudaStream_t s);

cudaError_t launch_mat_apply_bcs(int bs, int nbcs, const int *bc_rows, const int *bc_vars, const int *rowp,
                                 const int *cols, double *A, int diag_offset, cudaStream_t s);
cudaError_t launch_vec_apply_bcs(int bs, int nbcs, const int *bc_rows, const int *bc_vars,
                                 const double *bc_vals, const double *u, double lambda, double *x,
                                 cudaStream_t s);
cudaError_t launch_vec_set_bcs(int bs, int nbcs, const int *bc_rows, const int *bc_vars, const double *bc_vals,
                               double lambda, double *x, cudaStream_t s);

cudaError_t launch_spmv(int bs, int nrows, long nnzb, const int *rowp, const int *cols, const double *A,
                        const double *x, double *y, int add, int num_sms, cudaStream_t s);

// mode 0: y = A x; 1: y += A x (block by block, the multAdd order); 2: y = zs z + sign (A x); 3: y = y + sign (A x)
// order (optional): the rows listed by length class (BCSRPattern::d_order), so that the lanes of a warp run the same
// number of steps on meshes whose rows differ in length (quadratic elements)
// nnzb: the number of blocks (picks the 3x3 kernel: rows of 40 blocks and more on average are streamed through shared
// memory, spmv3_stream_kernel)
cudaError_t launch_spmv_fused(int bs, int nrows, long nnzb, const int *rowp, const int *cols, const double *A,
                              const double *x, double *y, int mode, double sign, double zs, const double *z,
                              const int *order, int num_sms, cudaStream_t s);
// the name of the kernel launch_spmv_fused picks (profile records)
const char *spmv_kernel_name(int bs, int nrows, long nnzb, const int *cols, const double *A, int mode, const int *order);

// y = A^T x; tidx[k] = position of the mirror block of block k (structurally symmetric pattern)
cudaError_t launch_spmv_transpose(int bs, int nrows, const int *rowp, const int *cols, const int *tidx, const double *A,
                                  const double *x, double *y, int num_sms, cudaStream_t s);

// Scalar all-reduce over NVLink peer memory, fused into the kernels that produce and consume the scalar (kernels.cu).
// Every rank owns an exchange buffer that all peers map (CUDA IPC): vals[slot][rank], flags[slot][rank]. The block that
// finishes a reduction stores its rank's partial into every peer's buffer and then raises the flag there; the kernel
// that needs the sum spins on its own (local) flags and adds the partials in rank order -- deterministic, no NCCL call,
// no host round trip. Sequence numbers live in device memory (produced / consumed counters advanced by the kernels), so
// a captured CUDA graph can be replayed.
struct PeerExchange {
  static constexpr int kMaxRanks = 8, kSlots = 4;
  double *vals[kMaxRanks];                 // vals[p]: rank p's buffer, kSlots x size doubles
  unsigned long long *flags[kMaxRanks];    // flags[p]: rank p's flags, kSlots x size
  unsigned long long *produced, *consumed; // local counters
  int rank, size;
};

// Krylov building blocks with device-resident scalars (kernels.cu): see the kernels for the contracts
cudaError_t launch_orth_step(long n, double *w, const double *vprev, const double *coef, const double *vnext,
                             double *partial, unsigned *ticket, double *out, int num_sms, cudaStream_t s);
// The same sweep on several GPUs: the coefficient is the all-rank sum of the previous sweep's partials (taken from the
// peer exchange; its value is also written to coef_out for later readers), and this sweep's partial is pushed to the
// peers instead of being written to `out`.
cudaError_t launch_orth_step_peer(long n, double *w, const double *vprev, double *coef_out, const double *vnext,
                                  double *partial, unsigned *ticket, const PeerExchange *px, int num_sms,
                                  cudaStream_t s);
// completes the last pending peer reduction: out[0] = sum over ranks (one thread)
cudaError_t launch_peer_finish(const PeerExchange *px, double *out, cudaStream_t s);
cudaError_t launch_scale_rsqrt(long n, double *v, const double *sumsq, double sign, int num_sms, cudaStream_t s);
cudaError_t launch_multi_axpy(long n, double *x, int nv, const double *const *vs, const double *coef, double scale,
                              int num_sms, cudaStream_t s);
cudaError_t launch_gmres_rotate(int i, int ldr, const double *hcol, double *R, double *cs, double *sn, double *g,
                                double *resnorm, cudaStream_t s);
cudaError_t launch_gmres_backsolve(int k, int ldr, const double *R, const double *g, double *y, cudaStream_t s);
cudaError_t launch_gmres_start(const double *sumsq, double *g, int m, cudaStream_t s);
cudaError_t launch_axpbz(long n, double zs, const double *z, double ys, double *y, int num_sms, cudaStream_t s);

// out[0] = max over scalar rows of (signed diagonal + sum of magnitudes); browp/B: Bext rows for owned rows >= np
cudaError_t launch_gershgorin(int bs, int nrows, const int *rowp, const int *cols, const double *A, int np,
                              const int *browp, const double *B, double *out, int num_sms, cudaStream_t s);

cudaError_t launch_axpy(long n, double alpha, const double *x, double *y, int num_sms, cudaStream_t s);
cudaError_t launch_axpby(long n, double alpha, double beta, const double *x, double *y, int num_sms,
                         cudaStream_t s);
cudaError_t launch_scale(long n, double alpha, double *y, int num_sms, cudaStream_t s);
int dot_num_partials(int num_sms);
// out[v] = x . ys[v] for v < nv <= 8; partial has nv*dot_num_partials doubles
cudaError_t launch_mdot(long n, const double *x, int nv, const double *const *ys, double *partial, double *out,
                        int num_sms, cudaStream_t s);

cudaError_t launch_pack_blocks(int bs, long count, const int *idx, const double *x, double *buf, int num_sms,
                               cudaStream_t s);
cudaError_t launch_unpack_blocks(int bs, long count, const int *idx, const double *buf, double *x, int add,
                                 int num_sms, cudaStream_t s);

// out[k] = src[k] >= 0 ? in[src[k]] : 0 for blocks of b2 doubles (values of a matrix view in another block order)
cudaError_t launch_permute_blocks(int b2, long nblocks, const int *src, const double *in, double *out, int num_sms,
                                  cudaStream_t s);

cudaError_t launch_dfma_peak(double *out, int iters, int blocks, cudaStream_t s);
cudaError_t launch_copy(long n, const double *src, double *dst, int num_sms, cudaStream_t s);

}  // namespace tb2
