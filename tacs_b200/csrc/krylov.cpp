// Krylov side of the hot path: the polynomial (Chebyshev) smoother and restarted GMRES, designed around the device.
//
// Reference interfaces (what a caller sees): TACSChebyshevSmoother  src/bpmat/TACSParallelMat.h:180-217,
// GMRES  src/bpmat/KSM.h:392-440 (solve / setTolerances / setMonitor / setOrthoType / setTimeMonitor).
//
// What is different from a host-driven solver:
//  * no scalar ever travels to the host inside an iteration. A Gram-Schmidt coefficient is produced by the reduction
//    of one kernel and consumed by the next through a device pointer (orth_step_kernel does "subtract the previous
//    projection, then reduce against the next basis vector" in one sweep); the Hessenberg column, the plane rotations
//    and the least-squares right-hand side are updated by a one-thread kernel.
//  * the body of iteration i (preconditioner, SpMV with halo, i+2 orthogonalisation sweeps, rotation, normalisation)
//    is captured once into a CUDA graph and replayed.
//  * the host looks at the residual history once every `check_every` iterations. Iterations run past the point of
//    convergence are discarded: the update uses the first k columns only, k being the first iterate that meets the
//    tolerance, so the result is the one a check after every iteration would have produced.
//  * the smoother's vector updates ride in the epilogue of the SpMV (y = zs z + sign A x), two buffers alternate.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>

#include "tb2_host.h"

namespace tb2 {

int comm_allreduce_sum(double *dev_buf, int n);  // comm.cpp
int comm_allreduce_max(double *dev_buf, int n);

// ---------------------------------------------------------------------------------------------
// TACSChebyshevSmoother
// ---------------------------------------------------------------------------------------------
TACSChebyshevSmoother::TACSChebyshevSmoother(TACSParallelMat *_mat, int _degree, double _lower, double _upper,
                                             int _iters) {
  mat = _mat;
  mat->incref();
  degree = _degree > 0 ? _degree : 1;
  iters = _iters;
  lower_factor = _lower;
  upper_factor = _upper;
  roots.assign(degree, 0.0);
  coef.assign(degree + 1, 1.0);
  res = mat->createVec();
  h0 = mat->createVec();
  h1 = mat->createVec();
  res->incref();
  h0->incref();
  h1->incref();
}

TACSChebyshevSmoother::~TACSChebyshevSmoother() {
  res->decref();
  h0->decref();
  h1->decref();
  mat->decref();
}

// Gershgorin bound of the spectral radius: one kernel over the rows (summation order per row as in
// TACSParallelMat.cpp:1024-1113, order-independent maximum), allreduce(max) over the ranks.
double TACSChebyshevSmoother::gershgorin() {
  if (!dot_buffers()) return 0.0;
  const BCSRPattern &A = mat->Aloc, &B = mat->Bext;
  {
    KernelTimer kt(K_VEC, "gershgorin_kernel");
    cuda_ok(launch_gershgorin(A.bsize, A.nrows, A.d_rowp.ptr, A.d_cols.ptr, A.d_vals.ptr, mat->np,
                              B.nnzb() > 0 ? B.d_rowp.ptr : nullptr, B.d_vals.ptr, g_dot_out, ctx().num_sms,
                              ctx().stream), "gershgorin");
  }
  if (ctx().size > 1) comm_allreduce_max(g_dot_out, 1);
  cuda_ok(cudaMemcpyAsync(g_dot_host, g_dot_out, sizeof(double), cudaMemcpyDeviceToHost, ctx().stream), "gershgorin");
  cuda_ok(cudaStreamSynchronize(ctx().stream.s), "gershgorin sync");
  return g_dot_host[0];
}

// The smoother applies y <- y + s(A) (x - A y) with s the polynomial for which q(t) = 1 - t s(t) is the Chebyshev
// polynomial of the interval [alpha, beta] = [lower, upper] * rho scaled to q(0) = 1. coef[] are the monomial
// coefficients of q from its roots (leading coefficient first), normalised by the constant term; the arithmetic
// follows TACSParallelMat.cpp:930-976 so that the coefficients agree with the reference's to the last bit.
int TACSChebyshevSmoother::factor() {
  rho = gershgorin();
  alpha = lower_factor * rho;
  beta = upper_factor * rho;
  const int d = degree;
  for (int k = 0; k < d; k++) {
    roots[k] = cos(M_PI * (0.5 + k) / d);
    roots[k] = 0.5 * (beta - alpha) * (roots[k] + 1.0) + alpha;
  }
  // expand prod_j (t - roots[j]) one factor at a time
  std::fill(coef.begin(), coef.end(), 0.0);
  coef[0] = 1.0;
  for (int j = 0; j < d; j++)
    for (int k = j; k >= 0; k--) coef[k + 1] = coef[k + 1] - roots[j] * coef[k];
  const double constant = coef[d];
  for (int k = 0; k < d; k++) coef[k] = coef[k] / constant;
  coef[d] = 1.0;
  return 0;
}

// y enters as the initial guess (TACSParallelMat.cpp:981-1014). Horner's rule on s(A) r with the updates fused into
// the products: r = x - A y;  h = -c0 r;  h <- A h - c_k r (k = 1 .. d-1);  y += h.
int TACSChebyshevSmoother::applyFactor(TACSBVec *x, TACSBVec *y) {
  NvtxRange nvtx_range("tacs_b200::TACSChebyshevSmoother::applyFactor");
  int rc = 0;
  for (int it = 0; it < iters; it++) {
    rc |= mat->multFused(y, res, -1.0, 1.0, x);  // res = x - A y
    TACSBVec *cur = h0, *nxt = h1;
    {
      KernelTimer kt(K_VEC, "axpbz_kernel");
      if (!cuda_ok(launch_axpbz(cur->ownedSize(), -coef[0], res->owned(), 0.0, cur->owned(), ctx().num_sms,
                                ctx().stream), "smoother start")) rc = 1;
    }
    for (int k = 1; k < degree; k++) {
      rc |= mat->multFused(cur, nxt, 1.0, -coef[k], res);  // nxt = A cur - c_k res
      std::swap(cur, nxt);
    }
    y->axpy(1.0, cur);
  }
  return rc;
}

// ---------------------------------------------------------------------------------------------
// GMRES
// ---------------------------------------------------------------------------------------------
GMRES::GMRES(TACSParallelMat *_mat, int _m, int _nrestart, TACSChebyshevSmoother *_pc, bool _flexible) {
  mat = _mat;
  mat->incref();
  pc = _pc;
  if (pc) pc->incref();
  flexible = _flexible && pc;
  m = _m > 0 ? _m : 1;
  nrestart = _nrestart >= 0 ? _nrestart : 0;
  for (int i = 0; i < m + 1; i++) {
    W.push_back(mat->createVec());
    W.back()->incref();
  }
  if (flexible) {
    for (int i = 0; i < m; i++) {
      Z.push_back(mat->createVec());
      Z.back()->incref();
    }
  } else if (pc) {
    work = mat->createVec();
    work->incref();
  }
  if (const char *env = getenv("TACSB200_GMRES_CHECK")) check_every = std::max(1, atoi(env));
  if (const char *env = getenv("TACSB200_GMRES_GRAPHS")) use_graphs = atoi(env) != 0;
  // device state, one allocation: hcol[m+2] R[(m+1) m] cs[m] sn[m] g[m+1] resnorm[m] y[m] sumsq[1]
  const size_t n = (size_t)(m + 2) + (size_t)(m + 1) * m + 2 * m + (m + 1) + m + m + 1;
  if (d_state.alloc(n) && d_ticket.alloc(1)) {
    cudaMemsetAsync(d_state.ptr, 0, n * sizeof(double), ctx().stream);
    cudaMemsetAsync(d_ticket.ptr, 0, sizeof(unsigned), ctx().stream);
    double *p = d_state.ptr;
    d_hcol = p; p += m + 2;
    d_R = p; p += (size_t)(m + 1) * m;
    d_cs = p; p += m;
    d_sn = p; p += m;
    d_g = p; p += m + 1;
    d_res = p; p += m;
    d_y = p; p += m;
    d_sumsq = p;
  }
  cuda_ok(cudaMallocHost(&h_res, (size_t)(m + 1) * sizeof(double)), "cudaMallocHost");
  graphs.assign(m, nullptr);
  graph_launches.assign(m, 0);
}

void GMRES::dropGraphs() {
  for (auto &g : graphs) {
    if (g) cudaGraphExecDestroy(g);
    g = nullptr;
  }
  std::fill(graph_launches.begin(), graph_launches.end(), 0);
}

GMRES::~GMRES() {
  dropGraphs();
  for (auto w : W) w->decref();
  for (auto z : Z) z->decref();
  if (work) work->decref();
  if (pc) pc->decref();
  if (h_res) cudaFreeHost(h_res);
  mat->decref();
}

void GMRES::setMonitor(const char *descript, int freq) {
  monitor_name = descript ? descript : "";
  monitor_freq = freq < 1 ? 1 : freq;
}

// w . v (or w . w) of the whole distributed vector into *out, optionally after w -= (*coef) vprev.
// On several GPUs with peer-mapped exchange buffers the reduction over the ranks is fused into the sweeps themselves
// (orth_step_peer_kernel): the partial of this sweep goes to the peers over NVLink from its last block, the next sweep
// picks the all-rank sum up in its prologue (and leaves it in *coef for the rotation kernel), and `out` is written by
// whoever completes the chain -- orth_finish after the last sweep. Otherwise: one ncclAllReduce per sweep.
static int orth_step(TACSBVec *w, TACSBVec *vprev, double *coef, TACSBVec *vnext, unsigned *ticket, double *out) {
  if (ctx().size > 1 && ctx().d_peer) {
    KernelTimer kt(K_DOT, "orth_step_peer_kernel");
    return cuda_ok(launch_orth_step_peer(w->ownedSize(), w->owned(), vprev ? vprev->owned() : nullptr, coef,
                                         vnext ? vnext->owned() : nullptr, g_dot_partial, ticket, ctx().d_peer,
                                         ctx().num_sms, ctx().stream), "orth step") ? 0 : 1;
  }
  {
    KernelTimer kt(K_DOT, "orth_step_kernel");
    if (!cuda_ok(launch_orth_step(w->ownedSize(), w->owned(), vprev ? vprev->owned() : nullptr, coef,
                                  vnext ? vnext->owned() : nullptr, g_dot_partial, ticket, out, ctx().num_sms,
                                  ctx().stream), "orth step")) return 1;
  }
  return ctx().size > 1 ? comm_allreduce_sum(out, 1) : 0;
}

// L2 residency of the vector being orthogonalised. Every Gram-Schmidt sweep reads and rewrites w and streams one basis
// vector past it; with w pinned in the 126 MB L2 (access-policy window on the compute stream: hits persist, everything
// else is streamed through) a sweep costs HBM one basis vector instead of three vector passes. The window is a stream
// attribute, so the kernels of a captured iteration carry it as a node attribute.
// The persisting carve-out takes L2 away from everything else (the SpMV alone ran 8 % slower with it left in place), so
// it exists only for the duration of a solve: l2_carveout(true) at the start, (false) at the end.
static size_t g_l2_persist_bytes = 0;
static void l2_carveout(bool on) {
  Context &c = ctx();
  if (getenv("TACSB200_NO_L2_WINDOW")) return;
  if (on) {
    int max_persist = 0, max_window = 0;
    cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, c.device);
    cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, c.device);
    size_t want = (size_t)max_persist;
    if (max_window > 0 && want > (size_t)max_window) want = (size_t)max_window;
    g_l2_persist_bytes = (want > 0 && cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) == cudaSuccess) ? want : 0;
  } else if (g_l2_persist_bytes) {
    cudaCtxResetPersistingL2Cache();
    cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, 0);
    g_l2_persist_bytes = 0;
  }
  cudaGetLastError();
}
static void l2_window(const void *base, size_t bytes) {
  Context &c = ctx();
  if (g_l2_persist_bytes == 0) return;
  cudaStreamAttrValue attr;
  memset(&attr, 0, sizeof(attr));
  if (base) {
    attr.accessPolicyWindow.base_ptr = const_cast<void *>(base);
    attr.accessPolicyWindow.num_bytes = bytes < g_l2_persist_bytes ? bytes : g_l2_persist_bytes;
    attr.accessPolicyWindow.hitRatio = 1.0f;
    attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
  } else {
    attr.accessPolicyWindow.num_bytes = 0;
    attr.accessPolicyWindow.hitProp = cudaAccessPropertyNormal;
    attr.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
  }
  cudaStreamSetAttribute(c.stream.s, cudaStreamAttributeAccessPolicyWindow, &attr);
  cudaGetLastError();
}

// completes the reduction chain of orth_step on the peer path: *out = all-rank sum of the last sweep's partials
static int orth_finish(double *out) {
  if (!(ctx().size > 1 && ctx().d_peer)) return 0;
  KernelTimer kt(K_DOT, "peer_finish_kernel");
  return cuda_ok(launch_peer_finish(ctx().d_peer, out, ctx().stream), "peer finish") ? 0 : 1;
}

// Iteration i: w = A M^{-1} v_i, orthogonalised against v_0..v_i, new rotation, v_{i+1} = w / |w|
int GMRES::iterationBody(int i) {
  int rc = 0;
  TACSBVec *w = W[i + 1];
  if (flexible) {
    rc |= pc->applyFactor(W[i], Z[i]);
    rc |= mat->mult(Z[i], w);
  } else if (pc) {
    rc |= pc->applyFactor(W[i], work);  // `work` keeps its previous contents as the smoother's initial guess
    rc |= mat->mult(work, w);
  } else {
    rc |= mat->mult(W[i], w);
  }
  l2_window(w->owned(), (size_t)w->ownedSize() * sizeof(double));
  if (ortho == MODIFIED_GRAM_SCHMIDT) {
    // h_j = v_j . w, w -= h_j v_j for j = 0..i: sweep j subtracts projection j-1 and reduces against v_j
    rc |= orth_step(w, nullptr, nullptr, W[0], d_ticket.ptr, d_hcol);
    for (int j = 1; j <= i; j++) rc |= orth_step(w, W[j - 1], d_hcol + (j - 1), W[j], d_ticket.ptr, d_hcol + j);
    rc |= orth_step(w, W[i], d_hcol + i, nullptr, d_ticket.ptr, d_hcol + i + 1);
    rc |= orth_finish(d_hcol + i + 1);
  } else {
    // classical Gram-Schmidt: all projections from one sweep (batches of 8), one sweep to subtract them
    for (int j0 = 0; j0 <= i; j0 += 8) {
      const int nv = std::min(8, i + 1 - j0);
      const double *ptrs[8];
      for (int v = 0; v < nv; v++) ptrs[v] = W[j0 + v]->owned();
      KernelTimer kt(K_DOT, "dot_partial_kernel");
      if (!cuda_ok(launch_mdot(w->ownedSize(), w->owned(), nv, ptrs, g_dot_partial, d_hcol + j0, ctx().num_sms,
                               ctx().stream), "mdot")) rc = 1;
    }
    if (ctx().size > 1) rc |= comm_allreduce_sum(d_hcol, i + 1);
    for (int j0 = 0; j0 <= i; j0 += 8) {
      const int nv = std::min(8, i + 1 - j0);
      const double *ptrs[8];
      for (int v = 0; v < nv; v++) ptrs[v] = W[j0 + v]->owned();
      KernelTimer kt(K_VEC, "multi_axpy_kernel");
      if (!cuda_ok(launch_multi_axpy(w->ownedSize(), w->owned(), nv, ptrs, d_hcol + j0, -1.0, ctx().num_sms,
                                     ctx().stream), "multi axpy")) rc = 1;
    }
    rc |= orth_step(w, nullptr, nullptr, nullptr, d_ticket.ptr, d_hcol + i + 1);
    rc |= orth_finish(d_hcol + i + 1);
  }
  {
    KernelTimer kt(K_VEC, "gmres_rotate_kernel");
    if (!cuda_ok(launch_gmres_rotate(i, m + 1, d_hcol, d_R, d_cs, d_sn, d_g, d_res, ctx().stream), "rotate")) rc = 1;
  }
  {
    KernelTimer kt(K_VEC, "scale_rsqrt_kernel");
    if (!cuda_ok(launch_scale_rsqrt(w->ownedSize(), w->owned(), d_hcol + i + 1, 1.0, ctx().num_sms, ctx().stream),
                 "normalise")) rc = 1;
  }
  l2_window(nullptr, 0);
  return rc;
}

// The first use of an iteration index runs directly (buffers sized on first use are allocated outside any capture),
// the second is captured into a graph, later ones replay it.
int GMRES::runIteration(int i) {
  Context &c = ctx();
  if (!use_graphs || monitor_time) return iterationBody(i);
  if (graphs[i]) {
    c.kernel_launches += graph_launches[i];
    return cuda_ok(cudaGraphLaunch(graphs[i], c.stream.s), "graph launch") ? 0 : 1;
  }
  if (graph_launches[i] == 0) {
    const long before = c.kernel_launches;
    const int rc = iterationBody(i);
    graph_launches[i] = (int)(c.kernel_launches - before);
    return rc;
  }
  cudaGraph_t graph = nullptr;
  if (cudaStreamBeginCapture(c.stream.s, cudaStreamCaptureModeRelaxed) != cudaSuccess) {
    cudaGetLastError();
    use_graphs = false;
    return iterationBody(i);
  }
  const long before = c.kernel_launches;
  const int rc_body = iterationBody(i);
  const cudaError_t end = cudaStreamEndCapture(c.stream.s, &graph);
  c.kernel_launches = before;
  if (rc_body || end != cudaSuccess || !graph ||
      cudaGraphInstantiate(&graphs[i], graph, nullptr, nullptr, 0) != cudaSuccess) {
    // capture is not available here (e.g. a collective that cannot be captured): launch kernel by kernel
    cudaGetLastError();
    if (graph) cudaGraphDestroy(graph);
    graphs[i] = nullptr;
    use_graphs = false;
    fprintf(stderr, "tacs_b200: GMRES iteration graph capture failed; continuing with direct launches\n");
    return iterationBody(i);
  }
  cudaGraphDestroy(graph);
  c.kernel_launches += graph_launches[i];
  return cuda_ok(cudaGraphLaunch(graphs[i], c.stream.s), "graph launch") ? 0 : 1;
}

int GMRES::solve(TACSBVec *b, TACSBVec *x, int zero_guess) {
  NvtxRange nvtx_range("tacs_b200::GMRES::solve");
  Context &c = ctx();
  if (!d_hcol || !h_res || !dot_buffers()) return 0;
  const auto t_begin = std::chrono::steady_clock::now();
  t_pc = t_ortho = t_total = 0.0;
  l2_carveout(true);
  iters = 0;
  int rc = 0;
  bool converged = false;
  double rhs_norm = 0.0;
  const long n = x->ownedSize();
  for (int cycle = 0; cycle <= nrestart && !converged; cycle++) {
    // direction of the residual into W[0]: b for a zero first guess, otherwise A x - b normalised with a minus sign
    double sign = 1.0;
    if (zero_guess && cycle == 0) {
      x->zeroEntries();
      W[0]->copyValues(b);
    } else {
      rc |= mat->multFused(x, W[0], 1.0, -1.0, b);
      sign = -1.0;
    }
    rc |= orth_step(W[0], nullptr, nullptr, nullptr, d_ticket.ptr, d_sumsq);
    rc |= orth_finish(d_sumsq);
    {
      KernelTimer kt(K_VEC, "gmres_start_kernel");
      if (!cuda_ok(launch_gmres_start(d_sumsq, d_g, m, c.stream), "gmres start")) rc = 1;
    }
    {
      KernelTimer kt(K_VEC, "scale_rsqrt_kernel");
      if (!cuda_ok(launch_scale_rsqrt(n, W[0]->owned(), d_sumsq, sign, c.num_sms, c.stream), "normalise")) rc = 1;
    }
    if (!cuda_ok(cudaMemcpyAsync(h_res, d_g, sizeof(double), cudaMemcpyDeviceToHost, c.stream), "D2H") ||
        !cuda_ok(cudaStreamSynchronize(c.stream.s), "sync")) {
      l2_carveout(false);
      return 0;
    }
    const double beta0 = h_res[0];
    if (cycle == 0) {
      rhs_norm = beta0;
      resnorm = beta0;
      if (monitor_freq > 0 && c.rank == 0) printf("%s[%3d]: %15.8e\n", monitor_name.c_str(), 0, beta0);
    }
    if (beta0 < atol) {
      converged = true;
      break;
    }
    int k = 0;  // columns used by the update
    for (int i0 = 0; i0 < m && !converged; i0 += check_every) {
      const int i1 = std::min(m, i0 + check_every);
      for (int i = i0; i < i1; i++) rc |= runIteration(i);
      if (!cuda_ok(cudaMemcpyAsync(h_res + i0, d_res + i0, (size_t)(i1 - i0) * sizeof(double), cudaMemcpyDeviceToHost,
                                   c.stream), "D2H") ||
          !cuda_ok(cudaStreamSynchronize(c.stream.s), "sync")) {
        l2_carveout(false);
        return 0;
      }
      for (int i = i0; i < i1; i++) {
        k = i + 1;
        resnorm = h_res[i];
        if (monitor_freq > 0 && c.rank == 0 && (iters + k) % monitor_freq == 0)
          printf("%s[%3d]: %15.8e\n", monitor_name.c_str(), iters + k, resnorm);
        if (resnorm < atol || resnorm < rtol * rhs_norm) {
          converged = true;  // iterations k .. i1-1 of this batch are discarded
          break;
        }
      }
    }
    iters += k;
    // x += V_k y with R(0:k,0:k) y = g(0:k)
    {
      KernelTimer kt(K_VEC, "gmres_backsolve_kernel");
      if (!cuda_ok(launch_gmres_backsolve(k, m + 1, d_R, d_g, d_y, c.stream), "back substitution")) rc = 1;
    }
    TACSBVec *target = x;
    const std::vector<TACSBVec *> &basis = flexible ? Z : W;
    if (!flexible && pc) {
      work->zeroEntries();
      target = work;
    }
    for (int j0 = 0; j0 < k; j0 += 8) {
      const int nv = std::min(8, k - j0);
      const double *ptrs[8];
      for (int v = 0; v < nv; v++) ptrs[v] = basis[j0 + v]->owned();
      KernelTimer kt(K_VEC, "multi_axpy_kernel");
      if (!cuda_ok(launch_multi_axpy(n, target->owned(), nv, ptrs, d_y + j0, 1.0, c.num_sms, c.stream), "update"))
        rc = 1;
    }
    if (!flexible && pc) {
      rc |= pc->applyFactor(work, W[0]);  // M^{-1} of the combination; W[0] enters as the smoother's initial guess
      x->axpy(1.0, W[0]);
    }
  }
  cudaStreamSynchronize(c.stream.s);
  l2_carveout(false);
  t_total = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
  if (monitor_time && c.rank == 0)
    printf("GMRES time monitor: total %.3f ms, %d iterations, %.3f ms per iteration\n", t_total, iters,
           iters ? t_total / iters : 0.0);
  return (converged && !rc) ? 1 : 0;
}

}  // namespace tb2
