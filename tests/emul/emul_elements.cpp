// CPU replay of the device element phases (tacs_b200/csrc/elem_phases.cuh) for tests.
//
// The element mathematics of the CUDA kernels is written as host/device task functions; this
// harness executes the same phases in the same order, one task after another, so the math and
// the indexing of the kernels can be checked against the oracle on a machine without a GPU.
// It is test scaffolding: it is compiled only by tests/ and is not part of the product library.
#include <cmath>
#include <cstring>
#include <vector>

#include "../../tacs_b200/csrc/elem_phases.cuh"

using namespace tb2;

// same rule as shell_desc_uncoupled (tacs_b200/csrc/tb2_host.h)
static bool desc_uncoupled(const double *desc) {
  double amax = 0.0, dmax = 0.0, bmax = 0.0;
  for (int i = 0; i < 6; i++) {
    amax = std::fmax(amax, std::fabs(desc[i]));
    bmax = std::fmax(bmax, std::fabs(desc[6 + i]));
    dmax = std::fmax(dmax, std::fabs(desc[12 + i]));
  }
  return bmax <= 1e-14 * std::sqrt(amax * dmax);
}

template <int O, class WK>
static void shell_outputs(WK *w, const ShellTables<O> &tab, const double *desc, double alpha, double gamma,
                          bool inertia, std::vector<double> &acc, double *res, double *mat) {
  constexpr int n = WK::n, nd = WK::nd;
  double *rpart = w->rpart();
  for (int t = 0; t < WK::ntiles; t++)
    shell_p6_finish<O>(t, *w, tab, desc, alpha, gamma, inertia, &acc[36 * t], rpart + 6 * t);
  for (int t = 0; t < WK::ntiles; t++) {
    int i = t / n, j = t % n;
    for (int a = 0; a < 6; a++)
      for (int b = 0; b < 6; b++) mat[nd * (6 * i + a) + 6 * j + b] = acc[36 * t + 6 * a + b];
  }
  for (int k = 0; k < nd; k++) {
    int i = k / 6, a = k % 6;
    double s = 0.0;
    for (int j = 0; j < n; j++) s += rpart[(i * n + j) * 6 + a];
    res[k] = s;
  }
}

// general kernel flow (shell_element_kernel<O>)
template <int O, int QC>
static void run_shell(const double *Xpts, const double *vars, const double *ddvars, const double *desc,
                      double alpha, double gamma, double *res, double *mat) {
  using WK = ShellWork<O, QC>;
  constexpr int n = WK::n, nd = WK::nd, nq = WK::nq, nty = WK::nty;
  static ShellTables<O> tab;
  build_shell_tables<O>(tab);
  WK *w = new WK;
  for (int k = 0; k < 3 * n; k++) w->X()[k] = Xpts[k];
  for (int k = 0; k < nd; k++) { w->u[k] = vars[k]; w->acc[k] = ddvars ? ddvars[k] : 0.0; }
  const bool inertia = (gamma != 0.0) || (ddvars != nullptr);
  for (int i = 0; i < n; i++) shell_p1_node<O>(i, *w, tab, desc);
  for (int t = 0; t < nty; t++) shell_p2_tying<O>(t, *w, tab);
  for (int q = 0; q < nq; q++) shell_p2_qgeom<O>(q, *w, tab, desc);
  std::vector<double> acc((size_t)WK::ntiles * 36, 0.0);
  for (int q0 = 0; q0 < nq; q0 += QC) {
    for (int t = 0; t < QC * nty; t++) shell_p3_weights<O, QC>(t, q0, *w, tab);
    for (int t = 0; t < QC * 22; t++) shell_p3_cw<O, QC>(t, q0, *w, desc);
    for (int t = 0; t < QC * n * 3; t++) shell_p3_columns<O, QC>(t, q0, *w, tab);
    for (int t = 0; t < WK::ntiles; t++)
      tile_accumulate<QC * 9, nd, 6, 6>(&w->B[0][0][0], &w->CB[0][0][0], 6 * (t / n), 6 * (t % n), &acc[36 * t]);
  }
  shell_outputs<O>(w, tab, desc, alpha, gamma, inertia, acc, res, mat);
  delete w;
}

// Tensor-core kernels (shell4_mma_kernel / shell9_mma_kernel): the scalar phases are the device task functions; the
// two MMA products are replayed as plain loops over the same shared-memory panels (the lane <-> fragment mapping
// itself is only exercised on the GPU)
template <int O>
static void run_shell_mma(const double *Xpts, const double *vars, const double *ddvars, const double *desc,
                          double alpha, double gamma, double *res, double *mat) {
  using WK = ShellMmaWork<O>;
  constexpr int n = WK::n, nd = WK::nd, nq = WK::nq, nty = WK::nty, KS = WK::KS, LDP = WK::LDP;
  static ShellTables<O> tab;
  build_shell_tables<O>(tab);
  WK *w = new WK;
  std::memset(w, 0, sizeof(WK));  // the kernel zeroes the work area once
  for (int k = 0; k < WK::SCR; k++) w->scr[k] = 1e300;  // stale scratch must not matter
  for (int k = 0; k < 3 * n; k++) w->X()[k] = Xpts[k];
  const bool inertia = (gamma != 0.0) || (ddvars != nullptr);
  for (int i = 0; i < n; i++) shell_p1_node<O>(i, *w, tab, desc);
  for (int t = 0; t < nty; t++) shell_p2_tying<O>(t, *w, tab);
  for (int q = 0; q < nq; q++) shell_unc_qgeom<O>(q, *w, tab, desc);
  for (int t = 0; t < 5 * nq; t++) shell_unc_G<O>(t, *w, desc);
  for (int k = 0; k < nty * (nty + 1) / 2; k++) shell_unc_S_entry<O>(shell_unc_tri<O>(k), *w, tab);
  for (int t = 0; t < (WK::LDS_ - nty) * nty; t++)  // k padding of the A operand S
    w->scr[WK::oS + (t / (WK::LDS_ - nty)) * WK::LDS_ + nty + t % (WK::LDS_ - nty)] = 0.0;
  // Rty = S Bty (rows 0..nty-1), pad rows zero
  {
    std::vector<double> R((size_t)KS * LDP, 0.0);
    for (int ty = 0; ty < 4 * KS; ty++)
      for (int col = 0; col < nd; col++) {
        double sum = 0.0;
        if (ty < nty)
          for (int t = 0; t < 4 * KS; t++)
            sum += w->scr[WK::oS + ty * WK::LDS_ + t] * w->Lty[t >> 2][col * 4 + (t & 3)];
        R[(ty >> 2) * LDP + col * 4 + (ty & 3)] = sum;
      }
    for (int k = 0; k < KS * LDP; k++) w->scr[WK::oRty + k] = R[k];
  }
  for (int t = 0; t < n * 3; t++) shell_unc_rows<O>(t, 0, *w, tab, desc, w->buf(0));
  std::vector<double> K((size_t)nd * nd, 0.0);
  for (int i = 0; i < nd; i++)
    for (int j = 0; j < nd; j++)
      for (int k = 0; k < 4 * KS; k++)
        K[i * nd + j] += w->Lty[k >> 2][i * 4 + (k & 3)] * w->scr[WK::oRty + (k >> 2) * LDP + j * 4 + (k & 3)];
  for (int q = 0; q < nq; q++) {
    if (q + 1 < nq) {
      for (int t = 0; t < n * 3; t++) shell_unc_rows<O>(t, q + 1, *w, tab, desc, w->buf((q + 1) & 1));
    } else {
      for (int k = 0; k < nd; k++) { w->uvec()[k] = vars[k]; w->avec()[k] = ddvars ? ddvars[k] : 0.0; }
    }
    const double *L = w->buf(q & 1);
    for (int i = 0; i < nd; i++)
      for (int j = 0; j < nd; j++)
        for (int r = 0; r < 4; r++)
          K[i * nd + j] += L[(r >> 1) * WK::HS + 2 * i + (r & 1)] * L[WK::LPAN + (r >> 1) * WK::HS + 2 * j + (r & 1)];
  }
  for (int i = 0; i < nd; i++) {
    double r = 0.0;
    for (int j = 0; j < nd; j++) r += K[i * nd + j] * w->uvec()[j];
    w->scr[WK::oRes + i] = r;
    for (int j = 0; j < nd; j++) mat[i * nd + j] = alpha * K[i * nd + j];
  }
  if (inertia) {
    for (int t = 0; t < n * n; t++) {
      double M[36];
      shell_mass_tile<O>(t, *w, tab, desc, M);
      const int i = t / n, j = t % n;
      for (int a = 0; a < 6; a++) {
        double sacc = 0.0;
        for (int b = 0; b < 6; b++) {
          sacc += M[6 * a + b] * w->avec()[6 * j + b];
          mat[nd * (6 * i + a) + 6 * j + b] += gamma * M[6 * a + b];
        }
        w->rpart()[6 * t + a] = sacc;
      }
    }
  }
  for (int k = 0; k < nd; k++) {
    double sres = w->scr[WK::oRes + k];
    if (inertia)
      for (int j = 0; j < n; j++) sres += w->rpart()[((k / 6) * n + j) * 6 + k % 6];
    res[k] = sres;
  }
  delete w;
}

// residual-only branch of shell4_mma_kernel (assembleRes, no inertia)
template <int O>
static void run_shell_mma_residual(const double *Xpts, const double *vars, const double *desc, double *res) {
  using WK = ShellMmaWork<O>;
  constexpr int n = WK::n, nd = WK::nd, nq = WK::nq, nty = WK::nty;
  static ShellTables<O> tab;
  build_shell_tables<O>(tab);
  WK *w = new WK;
  std::memset(w, 0, sizeof(WK));
  for (int k = 0; k < WK::SCR; k++) w->scr[k] = 1e300;
  for (int k = 0; k < 3 * n; k++) w->X()[k] = Xpts[k];
  double *sc = w->scr;
  for (int k = 0; k < nd; k++) sc[WK::oRu + k] = vars[k];
  for (int i = 0; i < n; i++) shell_p1_node<O>(i, *w, tab, desc);
  for (int t = 0; t < nty; t++) shell_p2_tying<O>(t, *w, tab);
  for (int q = 0; q < nq; q++) shell_unc_qgeom<O>(q, *w, tab, desc);
  for (int t = 0; t < nty; t++) shell_unc_res_tying<O>(t, *w, sc + WK::oRu, sc + WK::oRt);
  for (int q = 0; q < nq; q++) shell_unc_res_point<O>(q, *w, tab, desc, sc + WK::oRt, sc + WK::oRs5);
  for (int t = 0; t < nty; t++) shell_unc_res_back<O>(t, *w, tab, sc + WK::oRs5, sc + WK::oRsty);
  for (int k = 0; k < nd; k++) {
    res[k] = 0.0;
    for (int ty = 0; ty < nty; ty++) res[k] += w->bty(ty, k) * sc[WK::oRsty + ty];
  }
  for (int q = 0; q < nq; q++) {
    for (int t = 0; t < 3 * n; t++) shell_unc_rows<O, WK, true>(t, q, *w, tab, desc, w->buf(0));
    {
      // the kernel's butterfly: parts (0+1)+(2+3) for Quad4, ((0+1)+(2+3))+((4+5)+(6+7)) for Quad9
      constexpr int PARTS = (O == 2) ? 4 : 8;
      for (int r = 0; r < 4; r++) {
        double p[PARTS];
        for (int k = 0; k < PARTS; k++)
          p[k] = shell_unc_res_rowstrain<O, WK, PARTS>(r * PARTS + k, *w, w->buf(0), sc + WK::oRu);
        for (int stride = 1; stride < PARTS; stride *= 2)
          for (int k = 0; k < PARTS; k += 2 * stride) p[k] += p[k + stride];
        sc[WK::oRt4 + r] = p[0];
      }
    }
    for (int k = 0; k < nd; k++) res[k] += shell_unc_res_rowback<O>(k, q, *w, desc, w->buf(0), sc + WK::oRt4);
  }
  delete w;
}

template <int O, int QC>
static void run_solid(const double *Xpts, const double *vars, const double *ddvars, const double *desc,
                      double alpha, double gamma, double *res, double *mat) {
  using WK = SolidWork<O, QC>;
  constexpr int n = WK::n, nd = WK::nd, nq = WK::nq, TR = WK::TR, TC = WK::TC, ntc = nd / TC;
  static SolidTables<O> tab;
  build_solid_tables<O>(tab);
  WK *w = new WK;
  for (int k = 0; k < 3 * n; k++) w->X[k] = Xpts[k];
  for (int k = 0; k < nd; k++) { w->u[k] = vars[k]; w->acc[k] = ddvars ? ddvars[k] : 0.0; }
  for (int k = 0; k < kDescStride; k++) w->desc[k] = desc[k];
  for (int q = 0; q < nq; q++) solid_p1_qgeom<O, QC>(q, *w, tab);
  std::vector<double> acc((size_t)WK::ntiles * TR * TC, 0.0);
  for (int q0 = 0; q0 < nq; q0 += QC) {
    for (int t = 0; t < QC * n; t++) solid_p3_bcols<O, QC>(t, q0, *w, tab);
    for (int t = 0; t < WK::ntiles; t++)
      solid_tile_accumulate<QC, nd, TR, TC>(&w->G[0][0], &w->CB[0][0], TR * (t / ntc), TC * (t % ntc),
                                            &acc[(size_t)TR * TC * t]);
  }
  for (int t = 0; t < WK::ntiles; t++) solid_p6_finish<O, QC>(t, *w, tab, alpha, gamma, (gamma != 0.0) || (ddvars != nullptr), &acc[(size_t)TR * TC * t]);
  for (int t = 0; t < WK::ntiles; t++) {
    int r0 = TR * (t / ntc), c0 = TC * (t % ntc);
    for (int a = 0; a < TR; a++)
      for (int b = 0; b < TC; b++) mat[nd * (r0 + a) + c0 + b] = acc[(size_t)TR * TC * t + a * TC + b];
  }
  for (int k = 0; k < nd; k++) {
    int ti = k / TR, a = k % TR;
    double s = 0.0;
    for (int tj = 0; tj < ntc; tj++) s += w->rpart[ti * ntc + tj][a];
    res[k] = s;
  }
  delete w;
}

extern "C" {
// residual-only fast path (Quad4, uncoupled descriptor, no inertia): 0 when handled
int emul_residual(int kind, const double *Xpts, const double *vars, const double *desc, double *res) {
  if (kind > 2 || !desc_uncoupled(desc)) return 1;
  if (kind == 1) run_shell_mma_residual<2>(Xpts, vars, desc, res);
  else run_shell_mma_residual<3>(Xpts, vars, desc, res);
  return 0;
}

// kind: 1 Quad4, 2 Quad9, 3 hex8, 4 hex27; desc = one 32-double descriptor row
int emul_element(int kind, const double *Xpts, const double *vars, const double *ddvars, const double *desc,
                 double alpha, double gamma, double *res, double *mat) {
  switch (kind) {
    case 1:
      if (desc_uncoupled(desc)) run_shell_mma<2>(Xpts, vars, ddvars, desc, alpha, gamma, res, mat);
      else run_shell<2, 1>(Xpts, vars, ddvars, desc, alpha, gamma, res, mat);
      return 0;
    case 2:
      if (desc_uncoupled(desc)) run_shell_mma<3>(Xpts, vars, ddvars, desc, alpha, gamma, res, mat);
      else run_shell<3, 3>(Xpts, vars, ddvars, desc, alpha, gamma, res, mat);
      return 0;
    case 3: run_solid<2, 4>(Xpts, vars, ddvars, desc, alpha, gamma, res, mat); return 0;
    case 4: run_solid<3, 3>(Xpts, vars, ddvars, desc, alpha, gamma, res, mat); return 0;
  }
  return 1;
}
}
