// Element mathematics of the assembly hot path, written as *team phases*.
//
// One element is evaluated by a team of threads that share a work area in shared memory. The
// evaluation is a fixed sequence of phases; inside a phase every task is independent, tasks are
// dealt round-robin to the team's threads, and a team barrier separates phases. Every function
// here is a single task of a phase (TB2_HD: compiled for the device by nvcc, and for the host by
// the CPU emulation harness in tests/emul/ that replays the same phases sequentially).
//
// What is computed (citations into /root/reference/src/elements):
//   shell  TACSShellElement<...>::addJacobian / addResidual   shell/TACSShellElement.h:294-641
//          node normals  TacsShellComputeNodeNormals             shell/TACSShellUtilities.h:301-342
//          drill strain  TacsShellComputeDrillStrain             shell/TACSShellUtilities.h:649-693
//          director      TACSLinearizedRotation                  shell/TACSDirector.h:14-610
//          tying strain  TACSShellLinearModel::computeTyingStrain  shell/TACSShellElementModel.h:28-73
//          disp. gradient  TacsShellComputeDispGrad              shell/TACSShellUtilities.h:361-421
//          strain / stress  evalStrain Model.h:717-732, computeStress constitutive/TACSShellConstitutive.h:133-155
//          transforms    shell/TACSShellElementTransform.h:21-215
//   solid  TACSElement3D::addJacobian / addResidual           TACSElement3D.cpp:139-224
//          TACSLinearElasticity3D (linear strain)               TACSLinearElasticity.cpp:875-947, 1158-1342
//          getFieldGradient                                     basis/TACSElementBasis.cpp:266-326
//
// Formulation. Both models are linear, so the element tangent is state independent and equals
// K = sum_q w_q det_q B_q^T C B_q with B_q the strain-displacement rows at quadrature point q
// (for the shell: MITC-interpolated membrane/shear rows, bending rows through the director
// d = q x n, and the nodal drill rows interpolated to q). The residual is B^T C B u plus the
// inertial terms; the Jacobian is alpha*K + gamma*M. This is the reference's bilinear form
// evaluated directly instead of through its first/second-derivative back-propagation chain.
#pragma once

#include <math.h>

#include "elem_tables.h"

namespace tb2 {

// ------------------------------------------------------------------------------------------
// small dense helpers
// ------------------------------------------------------------------------------------------
TB2_HD void cross3(const double *x, const double *y, double *o) {
  o[0] = x[1] * y[2] - x[2] * y[1];
  o[1] = x[2] * y[0] - x[0] * y[2];
  o[2] = x[0] * y[1] - x[1] * y[0];
}

TB2_HD double inv3x3(const double *A, double *Ai) {
  double det = (A[8] * (A[0] * A[4] - A[3] * A[1]) - A[7] * (A[0] * A[5] - A[3] * A[2]) +
                A[6] * (A[1] * A[5] - A[2] * A[4]));
  double di = 1.0 / det;
  Ai[0] = (A[4] * A[8] - A[5] * A[7]) * di;
  Ai[1] = -(A[1] * A[8] - A[2] * A[7]) * di;
  Ai[2] = (A[1] * A[5] - A[2] * A[4]) * di;
  Ai[3] = -(A[3] * A[8] - A[5] * A[6]) * di;
  Ai[4] = (A[0] * A[8] - A[2] * A[6]) * di;
  Ai[5] = -(A[0] * A[5] - A[2] * A[3]) * di;
  Ai[6] = (A[3] * A[7] - A[4] * A[6]) * di;
  Ai[7] = -(A[0] * A[7] - A[1] * A[6]) * di;
  Ai[8] = (A[0] * A[4] - A[1] * A[3]) * di;
  return det;
}

TB2_HD void mat3mul(const double *A, const double *B, double *C) {
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++)
      C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}

// six consecutive doubles from a 16-byte aligned shared-memory address: three 128-bit loads on the device
TB2_HD void load6(const double *p, double *o) {
#if defined(__CUDA_ARCH__)
  const double2 a = reinterpret_cast<const double2 *>(p)[0], b = reinterpret_cast<const double2 *>(p)[1],
                c = reinterpret_cast<const double2 *>(p)[2];
  o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y; o[4] = c.x; o[5] = c.y;
#else
  for (int k = 0; k < 6; k++) o[k] = p[k];
#endif
}

// Local shell frame T = [t1 t2 n] (columns). kind 0: natural transform, which removes the normal
// component from t1[0] only (the reference repeats one line three times, Transform.h:42-44);
// kind 1: reference-axis transform with a pre-normalised axis.
TB2_HD void shell_frame(int kind, const double *axis, const double *Xxi, const double *n0, double *T) {
  double n[3] = {n0[0], n0[1], n0[2]};
  double inv = 1.0 / sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
  n[0] *= inv;
  n[1] *= inv;
  n[2] *= inv;
  double t1[3], t2[3];
  if (kind == 0) {
    t1[0] = Xxi[0];
    t1[1] = Xxi[2];
    t1[2] = Xxi[4];
    double d = n[0] * t1[0] + n[1] * t1[1] + n[2] * t1[2];
    t1[0] = t1[0] - d * n[0];
    t1[0] = t1[0] - d * n[0];
    t1[0] = t1[0] - d * n[0];
  } else {
    double an = axis[0] * n[0] + axis[1] * n[1] + axis[2] * n[2];
    t1[0] = axis[0] - an * n[0];
    t1[1] = axis[1] - an * n[1];
    t1[2] = axis[2] - an * n[2];
  }
  inv = 1.0 / sqrt(t1[0] * t1[0] + t1[1] * t1[1] + t1[2] * t1[2]);
  t1[0] *= inv;
  t1[1] *= inv;
  t1[2] *= inv;
  cross3(n, t1, t2);
  T[0] = t1[0]; T[3] = t1[1]; T[6] = t1[2];
  T[1] = t2[0]; T[4] = t2[1]; T[7] = t2[2];
  T[2] = n[0];  T[5] = n[1];  T[8] = n[2];
}

// Per-descriptor constants of an element (one row of the device descriptor table, 32 doubles):
//   shell: [0..21] tangent stiffness A(6) B(6) D(6) As(3) drill, [22..24] mass moments,
//          [25] transform kind, [26..28] reference axis
//   solid: [0..20] C (upper triangle by rows), [21] density
static constexpr int kDescStride = 32;

// ------------------------------------------------------------------------------------------
// shell work area and phases
// ------------------------------------------------------------------------------------------
template <int O, int QC>
struct alignas(16) ShellWork {
  using D = ShellDims<O>;
  static constexpr int n = D::n, nd = D::nd, nq = D::nq, nty = D::nty;
  static constexpr int NS = 9;          // strain rows per quadrature point
  static constexpr int TR = 6, TC = 6;  // K tile of one thread = one node pair
  static constexpr int ntiles = n * n;
  static_assert(nq % QC == 0, "quadrature points must split evenly into chunks");
  static_assert(ntiles * 6 <= QC * NS * nd, "residual partials are staged in the B rows");
  double u[nd];
  double acc[nd];  // second time derivative of the state
  double desc[kDescStride];  // descriptor row of this element (constitutive constants, transform)
  double fn[3 * n];
  static constexpr int LDT = nd + 2;  // padded row stride (16-byte aligned rows, lanes spread over banks)
  alignas(16) double Bdr[n][LDT];
  alignas(16) double Bty[nty][LDT];
  double T[nq][9], A[nq][9], Az[nq][9];
  double wdet[nq];
  alignas(16) double W[QC][nty][6];  // tying-point -> strain-row weights of the current chunk (m padded to 6)
  double Cw[QC][24];      // w det C of the current chunk (22 used)
  alignas(16) double B[QC][NS][nd];  // strain rows of the current chunk; also holds X (start), rpart (end)
  alignas(16) double CB[QC][NS][nd];
  TB2_HD double *X() { return &B[0][0][0]; }
  TB2_HD double *rpart() { return &B[0][0][0]; }
  TB2_HD double *uvec() { return u; }
  TB2_HD double *avec() { return acc; }
  TB2_HD double &bty(int ty, int col) { return Bty[ty][col]; }
  // nodal drill row of node i: displacement / rotation columns of node j (rotation part non-zero for j == i only)
  static constexpr bool kFullBdr = true;
  TB2_HD double &bdr_u(int i, int j, int c) { return Bdr[i][6 * j + c]; }
  TB2_HD double &bdr_q(int i, int c) { return Bdr[i][6 * i + 3 + c]; }
};

// phase 1, task i in [0,n): node normal, nodal frame, nodal drill-strain row
template <int O, class WK>
TB2_HD void shell_p1_node(int i, WK &w, const ShellTables<O> &tab, const double *desc) {
  constexpr int n = ShellDims<O>::n;
  const double *X = w.X();
  double Xxi[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  for (int j = 0; j < n; j++) {
    const double d0 = tab.dNn_T[j][0][i], d1 = tab.dNn_T[j][1][i];
#pragma unroll
    for (int c = 0; c < 3; c++) {
      Xxi[2 * c] += d0 * X[3 * j + c];
      Xxi[2 * c + 1] += d1 * X[3 * j + c];
    }
  }
  double a[3] = {Xxi[0], Xxi[2], Xxi[4]}, b[3] = {Xxi[1], Xxi[3], Xxi[5]}, f[3];
  cross3(a, b, f);
  double nrm = sqrt(f[0] * f[0] + f[1] * f[1] + f[2] * f[2]);
  if (nrm != 0.0) {
    double inv = 1.0 / nrm;
    f[0] *= inv;
    f[1] *= inv;
    f[2] *= inv;
  }
  w.fn[3 * i] = f[0];
  w.fn[3 * i + 1] = f[1];
  w.fn[3 * i + 2] = f[2];
  double Xd[9], T[9], Xdinv[9], XdinvT[9];
#pragma unroll
  for (int c = 0; c < 3; c++) {
    Xd[3 * c] = Xxi[2 * c];
    Xd[3 * c + 1] = Xxi[2 * c + 1];
    Xd[3 * c + 2] = f[c];
  }
  shell_frame((int)desc[25], &desc[26], Xxi, f, T);
  inv3x3(Xd, Xdinv);
  mat3mul(Xdinv, T, XdinvT);
  for (int j = 0; j < n; j++) {
    const double d0 = tab.dNn_T[j][0][i], d1 = tab.dNn_T[j][1][i];
    const double g0 = d0 * XdinvT[0] + d1 * XdinvT[3];
    const double g1 = d0 * XdinvT[1] + d1 * XdinvT[4];
#pragma unroll
    for (int c = 0; c < 3; c++) {
      w.bdr_u(i, j, c) = 0.5 * (T[3 * c + 1] * g0 - T[3 * c] * g1);
      if (WK::kFullBdr) w.bdr_u(i, j, 3 + c) = 0.0;
    }
  }
  double t1[3] = {T[0], T[3], T[6]}, t2[3] = {T[1], T[4], T[7]}, t12[3];
  cross3(t1, t2, t12);
  w.bdr_q(i, 0) = -t12[0];
  w.bdr_q(i, 1) = -t12[1];
  w.bdr_q(i, 2) = -t12[2];
}

// phase 2, task ty in [0,nty): tying-strain row (reads fn of every node: runs after phase 1).
// The five strain definitions (Model.h:46-70) are evaluated by one branch-free expression with
// per-field 0 / 1 / 0.5 coefficients so that the tasks of a team do not diverge:
//   row_u[c] = p00 d0 X,1[c] + p01 d0 X,2[c] + p10 d1 X,1[c] + p11 d1 X,2[c] + n0[c] (s0 d0 + s1 d1)
//   row_d[c] = N (t0 X,1[c] + t1 X,2[c])          (d0 = dN/dxi1, d1 = dN/dxi2, X,k = dX/dxi_k)
// Multiplying by an exact 0 or 1 does not change the rounded value of the surviving term.
template <int O, class WK>
TB2_HD void shell_p2_tying(int ty, WK &w, const ShellTables<O> &tab) {
  constexpr int n = ShellDims<O>::n;
  const double *X = w.X();
  const int field = shell_ty_field<O>(ty);
  const double p00 = (field == 0) ? 1.0 : 0.0, p11 = (field == 1) ? 1.0 : 0.0;
  const double p01 = (field == 2) ? 0.5 : 0.0, p10 = p01;
  const double s0 = (field == 4) ? 0.5 : 0.0, s1 = (field == 3) ? 0.5 : 0.0;
  double Xxi[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0}, n0[3] = {0.0, 0.0, 0.0};
  for (int j = 0; j < n; j++) {
    const double d0 = tab.dNt_T[j][0][ty], d1 = tab.dNt_T[j][1][ty], N = tab.Nt_T[j][ty];
#pragma unroll
    for (int c = 0; c < 3; c++) {
      Xxi[2 * c] += d0 * X[3 * j + c];
      Xxi[2 * c + 1] += d1 * X[3 * j + c];
      n0[c] += N * w.fn[3 * j + c];
    }
  }
  for (int j = 0; j < n; j++) {
    const double d0 = tab.dNt_T[j][0][ty], d1 = tab.dNt_T[j][1][ty], N = tab.Nt_T[j][ty];
    const double a1 = p00 * d0 + p10 * d1, a2 = p01 * d0 + p11 * d1, an = s0 * d0 + s1 * d1;
    const double b1 = s0 * N, b2 = s1 * N;  // t0 == s0, t1 == s1
    double du[3], dd[3], dq[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
      du[c] = a1 * Xxi[2 * c] + a2 * Xxi[2 * c + 1] + an * n0[c];
      dd[c] = b1 * Xxi[2 * c] + b2 * Xxi[2 * c + 1];
    }
    cross3(&w.fn[3 * j], dd, dq);
#pragma unroll
    for (int c = 0; c < 3; c++) {
      w.bty(ty, 6 * j + c) = du[c];
      w.bty(ty, 6 * j + 3 + c) = dq[c];
    }
  }
}

// phase 2 (same barrier interval), task q in [0,nq): frame, inverse Jacobian products, weighted determinant
template <int O, class WK>
TB2_HD void shell_p2_qgeom(int q, WK &w, const ShellTables<O> &tab, const double *desc) {
  constexpr int n = ShellDims<O>::n;
  const double *X = w.X();
  double Xxi[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0}, n0[3] = {0.0, 0.0, 0.0};
  double nxi[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  for (int j = 0; j < n; j++) {
    const double d0 = tab.dNq_T[j][0][q], d1 = tab.dNq_T[j][1][q], N = tab.Nq_T[j][q];
#pragma unroll
    for (int c = 0; c < 3; c++) {
      Xxi[2 * c] += d0 * X[3 * j + c];
      Xxi[2 * c + 1] += d1 * X[3 * j + c];
      n0[c] += N * w.fn[3 * j + c];
      nxi[2 * c] += d0 * w.fn[3 * j + c];
      nxi[2 * c + 1] += d1 * w.fn[3 * j + c];
    }
  }
  double T[9], Xd[9], Xdz[9], Xdinv[9], A[9], Az[9], tmp[9];
  shell_frame((int)desc[25], &desc[26], Xxi, n0, T);
#pragma unroll
  for (int c = 0; c < 3; c++) {
    Xd[3 * c] = Xxi[2 * c];
    Xd[3 * c + 1] = Xxi[2 * c + 1];
    Xd[3 * c + 2] = n0[c];
    Xdz[3 * c] = nxi[2 * c];
    Xdz[3 * c + 1] = nxi[2 * c + 1];
    Xdz[3 * c + 2] = 0.0;
  }
  const double det = inv3x3(Xd, Xdinv);
  mat3mul(Xdinv, Xdz, tmp);
#pragma unroll
  for (int k = 0; k < 9; k++) tmp[k] = -tmp[k];
  mat3mul(Xdinv, T, A);
  mat3mul(tmp, A, Az);
  w.wdet[q] = det * tab.wq[q];
#pragma unroll
  for (int k = 0; k < 9; k++) {
    w.T[q][k] = T[k];
    w.A[q][k] = A[k];
    w.Az[q][k] = Az[k];
  }
}

// phase 3a, task (ql, ty): weights that turn the tying-point strains into the membrane /
// transverse-shear strain rows at quadrature point q0+ql
//   e0ty(a,b) = sum_cd A(c,a) G(c,d) A(d,b); strain rows fed by e0ty:
//   m=0: e0 = e0ty(0,0)  m=1: e1 = e0ty(1,1)  m=2: e2 = 2 e0ty(0,1)  m=3: e6 = 2 e0ty(1,2)  m=4: e7 = 2 e0ty(0,2)
template <int O, int QC>
TB2_HD void shell_p3_weights(int task, int q0, ShellWork<O, QC> &w, const ShellTables<O> &tab) {
  constexpr int nty = ShellDims<O>::nty;
  const int ql = task / nty, ty = task % nty, q = q0 + ql;
  const double *A = w.A[q];
  const int f = shell_ty_field<O>(ty);
  const int c = (f == 1 || f == 3) ? 1 : 0;  // g11,g12,g13 -> 0 ; g22,g23 -> 1
  const int d = (f == 0) ? 0 : ((f == 1 || f == 2) ? 1 : 2);
  const double off = (c == d) ? 0.0 : 1.0;   // off-diagonal tensor components appear twice
  const double Nt = tab.Ntq[q][ty];
  const double Ac0 = A[3 * c], Ac1 = A[3 * c + 1], Ac2 = A[3 * c + 2];
  const double Ad0 = A[3 * d], Ad1 = A[3 * d + 1], Ad2 = A[3 * d + 2];
  w.W[ql][ty][0] = Nt * (Ac0 * Ad0 + off * (Ad0 * Ac0));
  w.W[ql][ty][1] = Nt * (Ac1 * Ad1 + off * (Ad1 * Ac1));
  w.W[ql][ty][2] = 2.0 * Nt * (Ac0 * Ad1 + off * (Ad0 * Ac1));
  w.W[ql][ty][3] = 2.0 * Nt * (Ac1 * Ad2 + off * (Ad1 * Ac2));
  w.W[ql][ty][4] = 2.0 * Nt * (Ac0 * Ad2 + off * (Ad0 * Ac2));
  w.W[ql][ty][5] = 0.0;
}

// phase 3a (same barrier interval), task (ql, k < 22): w det C
template <int O, int QC>
TB2_HD void shell_p3_cw(int task, int q0, ShellWork<O, QC> &w, const double *desc) {
  const int ql = task / 22, k = task % 22;
  w.Cw[ql][k] = w.wdet[q0 + ql] * desc[k];
}

// phase 3b, task (ql, j, c) with c in {0,1,2}: the two columns 6j+c (displacement) and 6j+3+c (rotation)
// of the nine strain rows B and of CB = w det C B at quadrature point q0+ql
template <int O, int QC, bool WITH_CB = true>
TB2_HD void shell_p3_columns(int task, int q0, ShellWork<O, QC> &w, const ShellTables<O> &tab) {
  constexpr int n = ShellDims<O>::n, nty = ShellDims<O>::nty;
  const int c = task % 3, j = (task / 3) % n, ql = task / (3 * n);
  const int q = q0 + ql;
  const int cu = 6 * j + c, cq = cu + 3;
  double bu[9], bq[9];
  // rows 0,1,2,6,7: tying-strain rows combined with the weights of this quadrature point
  {
    double su[5] = {0.0, 0.0, 0.0, 0.0, 0.0}, sq[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    // the g11, g22, g12 tying strains do not involve the director: their rotation columns are exactly zero
    constexpr int nmem = ShellDims<O>::c11 + ShellDims<O>::c22 + ShellDims<O>::c12;
    for (int ty = 0; ty < nmem; ty++) {
      const double tu = w.Bty[ty][cu];
      double wt[6];
      load6(&w.W[ql][ty][0], wt);
#pragma unroll
      for (int m = 0; m < 5; m++) su[m] += wt[m] * tu;
    }
    for (int ty = nmem; ty < nty; ty++) {
      const double tu = w.Bty[ty][cu], tq = w.Bty[ty][cq];
      double wt[6];
      load6(&w.W[ql][ty][0], wt);
#pragma unroll
      for (int m = 0; m < 5; m++) {
        su[m] += wt[m] * tu;
        sq[m] += wt[m] * tq;
      }
    }
    bu[0] = su[0]; bu[1] = su[1]; bu[2] = su[2]; bu[6] = su[3]; bu[7] = su[4];
    bq[0] = sq[0]; bq[1] = sq[1]; bq[2] = sq[2]; bq[6] = sq[3]; bq[7] = sq[4];
  }
  // rows 3,4,5: bending strains from u1x = T^T (u1d XdinvT + u0d XdinvzT); rotation part through d = q x n
  {
    const double d0 = tab.dNq[q][j][0], d1 = tab.dNq[q][j][1], N = tab.Nq[q][j];
    const double *T = w.T[q], *A = w.A[q], *Az = w.Az[q];
    const double hz0 = d0 * Az[0] + d1 * Az[3], hz1 = d0 * Az[1] + d1 * Az[4];
    const double h0 = d0 * A[0] + d1 * A[3] + N * Az[6], h1 = d0 * A[1] + d1 * A[4] + N * Az[7];
    bu[3] = T[3 * c] * hz0;
    bu[4] = T[3 * c + 1] * hz1;
    bu[5] = T[3 * c] * hz1 + T[3 * c + 1] * hz0;
    const int c1 = (c + 1) % 3, c2 = (c + 2) % 3;
    const double f1 = w.fn[3 * j + c1], f2 = w.fn[3 * j + c2];
    // (fn x rd)_c = fn[c1] rd[c2] - fn[c2] rd[c1]
    bq[3] = f1 * (T[3 * c2] * h0) - f2 * (T[3 * c1] * h0);
    bq[4] = f1 * (T[3 * c2 + 1] * h1) - f2 * (T[3 * c1 + 1] * h1);
    bq[5] = f1 * (T[3 * c2] * h1 + T[3 * c2 + 1] * h0) - f2 * (T[3 * c1] * h1 + T[3 * c1 + 1] * h0);
  }
  // row 8: drill strain, nodal values interpolated with the nodal shape functions
  {
    double su = 0.0, sq = 0.0;
    for (int i = 0; i < n; i++) {
      const double N = tab.Nq[q][i];
      su += N * w.Bdr[i][cu];
      sq += N * w.Bdr[i][cq];
    }
    bu[8] = su;
    bq[8] = sq;
  }
  if (!WITH_CB) {
#pragma unroll
    for (int r = 0; r < 9; r++) {
      w.B[ql][r][cu] = bu[r];
      w.B[ql][r][cq] = bq[r];
    }
    return;
  }
  // CB = (w det C) B, C = [A B 0; B D 0; 0 0 As | drill], symmetric 3x3 blocks packed [0 1 2; 1 3 4; 2 4 5]
  const double *C = w.Cw[ql];
  double cbu[9], cbq[9];
#pragma unroll
  for (int r = 0; r < 3; r++) {
    const int i0 = (r == 0) ? 0 : ((r == 1) ? 1 : 2), i1 = (r == 0) ? 1 : ((r == 1) ? 3 : 4),
              i2 = (r == 0) ? 2 : ((r == 1) ? 4 : 5);
    cbu[r] = C[i0] * bu[0] + C[i1] * bu[1] + C[i2] * bu[2] + C[6 + i0] * bu[3] + C[6 + i1] * bu[4] + C[6 + i2] * bu[5];
    cbq[r] = C[i0] * bq[0] + C[i1] * bq[1] + C[i2] * bq[2] + C[6 + i0] * bq[3] + C[6 + i1] * bq[4] + C[6 + i2] * bq[5];
    cbu[3 + r] = C[6 + i0] * bu[0] + C[6 + i1] * bu[1] + C[6 + i2] * bu[2] + C[12 + i0] * bu[3] + C[12 + i1] * bu[4] +
                 C[12 + i2] * bu[5];
    cbq[3 + r] = C[6 + i0] * bq[0] + C[6 + i1] * bq[1] + C[6 + i2] * bq[2] + C[12 + i0] * bq[3] + C[12 + i1] * bq[4] +
                 C[12 + i2] * bq[5];
  }
  cbu[6] = C[18] * bu[6] + C[19] * bu[7];
  cbq[6] = C[18] * bq[6] + C[19] * bq[7];
  cbu[7] = C[19] * bu[6] + C[20] * bu[7];
  cbq[7] = C[19] * bq[6] + C[20] * bq[7];
  cbu[8] = C[21] * bu[8];
  cbq[8] = C[21] * bq[8];
#pragma unroll
  for (int r = 0; r < 9; r++) {
    w.B[ql][r][cu] = bu[r];
    w.B[ql][r][cq] = bq[r];
    w.CB[ql][r][cu] = cbu[r];
    w.CB[ql][r][cq] = cbq[r];
  }
}

// ------------------------------------------------------------------------------------------
// membrane/bending-uncoupled shells (B block of the constitutive matrix == 0, e.g. isotropic plates without
// offset and symmetric laminates): the tying-strain part and the drill part of K do not need per-point rows,
//   K = Bty^T S Bty + Bdr^T Sd Bdr + sum_q Bb_q^T (w det D) Bb_q,
//   S  = sum_q w det W_q^T [A 0; 0 As] W_q   (nty x nty),   Sd = drill * sum_q w det N_q N_q^T   (n x n),
// so the tile loop runs over nty + n + 3 nq rows instead of 9 nq and the per-point column work shrinks to
// the three bending rows. The tying weights factor as W_q[ty][m] = Ntq[q][ty] * P_q[field(ty)][m] (the five
// tying fields share the frame products P_q), hence
//   S[t1][t2] = sum_q Ntq[q][t1] Ntq[q][t2] G_q[f1][f2],   G_q = w det P_q [A 0; 0 As] P_q^T   (5 x 5 per point),
// which moves all the S arithmetic out of the quadrature loop.
// ------------------------------------------------------------------------------------------
// Work area of the tensor-core (DMMA m8n8k4) kernels (Quad4: shell4_mma_kernel, Quad9: shell9_mma_kernel). Every
// operand of a matrix product is kept as "k-step panels" X[kstep][col][4]: the four rows of one k-step are contiguous
// per column, so that the A / B fragment of lane l for tile t of a panel is the single double at panel[32 t + l] (a
// warp reads 256 contiguous bytes, conflict free) and the producers of the rows store 128-bit pieces. Columns are
// padded to a whole number of 8-wide tiles (Quad9: 54 -> 56; what the products put in the pad columns is never
// stored), tying rows to a whole number of k-steps (Quad4: 9 -> 12, the pad rows stay exactly zero).
template <int O>
struct alignas(16) ShellMmaWork {
  using D = ShellDims<O>;
  static constexpr int n = D::n, nd = D::nd, nq = D::nq, nty = D::nty;
  static constexpr int ntiles = n * n;
  static constexpr int NT = (nd + 7) / 8, NDP = 8 * NT;   // 8x8 tiles per side, padded columns
  static constexpr int KS = (nty + 3) / 4;                // k-steps of the tying rows
  static constexpr int LDP = 4 * NDP + 4;  // panel stride (+4: the C-fragment stores of two panels hit distinct banks)
  static constexpr int LDS_ = 4 * KS;      // row stride of S (A operand of S * Bty, k padded with zeros)
  static constexpr int SROWS = 8 * ((nty + 7) / 8);  // rows >= nty are never written: their products are dropped
  // Row buffers (one k-step each, rewritten per quadrature point) split a panel into its row pairs,
  // X[k >> 1][col][k & 1] with the two halves HS doubles apart: the producers (one lane per column pair) then store
  // 16-byte pieces at consecutive addresses instead of 32-byte strides, and the fragment of lane (gq,tq) for tile t
  // sits at (tq >> 1) HS + 16 t + 2 gq + (tq & 1)  (HS = 8 mod 16 keeps a half-warp conflict free).
  static constexpr int HS = 2 * NDP + 8, LPAN = 2 * HS;
  static constexpr int LBUF = 2 * LPAN;  // row buffer: L panel then R panel
  static constexpr int even(int x) { return x + (x & 1); }
  // scratch layout (doubles). Lifetimes: X [load .. p2]; P [p2 .. G]; G [G .. S]; S [S .. SB product];
  // Rty = S Bty [SB product .. tying contraction]; row buffer 0 [after SB product ..] overlays S, row buffer 1
  // overlays Rty [after the tying contraction ..]; u, acc, residual, Rp [last loop interval .. finish] in the
  // buffer the last interval does not read.
  static constexpr int oS = 0;
  static constexpr int oG = SROWS * LDS_;            // G[nq][26]
  static constexpr int oP = oG + 26 * nq;            // P[nq][5][6]
  static constexpr int oX = oP + 30 * nq;            // X[3n]
  static constexpr int oRty = (LBUF > SROWS * LDS_) ? LBUF : SROWS * LDS_;  // Rty[KS][LDP]: clear of buffer 0 and S
  static constexpr int SCR = oRty + KS * LDP;
  static constexpr int oU = (nq & 1) * oRty, oAcc = oU + nd, oRes = oAcc + nd, oRp = oRes + NDP;
  // residual-only path (no tangent): state, tying strains / stresses and row strains beyond X; only buffer 0 is used
  static constexpr int oRu = oX + even(3 * n), oRt = oRu + nd, oRs5 = oRt + LDS_, oRsty = oRs5 + 6 * nq,
                       oRt4 = oRsty + LDS_;
  double fn[3 * n];
  double Bdu[n][3 * n];   // nodal drill rows, displacement columns [node i][3 j + c]
  double Bdq[n][4];       // nodal drill rows, rotation columns of the own node [node i][c]
  alignas(16) double Lty[KS][LDP];
  alignas(16) double geo[nq][18];
  double wdet[nq + (nq & 1)];
  alignas(16) double scr[SCR];
  TB2_HD double *X() { return scr + oX; }
  TB2_HD double *rpart() { return scr + oRp; }
  TB2_HD double *uvec() { return scr + oU; }
  TB2_HD double *avec() { return scr + oAcc; }
  TB2_HD double *buf(int k) { return scr + k * oRty; }
  TB2_HD double &bty(int ty, int col) { return Lty[ty >> 2][col * 4 + (ty & 3)]; }
  static constexpr bool kFullBdr = false;
  TB2_HD double &bdr_u(int i, int j, int c) { return Bdu[i][3 * j + c]; }
  TB2_HD double &bdr_q(int i, int c) { return Bdq[i][c]; }
  static_assert(oX + even(3 * n) <= SCR && oRt4 + 4 <= SCR && oRu >= LBUF, "setup / residual-only scratch");
  static_assert(oRp + 6 * ntiles <= SCR && (oU != 0 || oRp + 6 * ntiles <= LBUF), "finish data");
  static_assert(HS % 16 == 8 && LDP % 16 == 4, "bank spreading of the panels");
};
using ShellQ4MmaWork = ShellMmaWork<2>;
using ShellQ9MmaWork = ShellMmaWork<3>;
static_assert(ShellQ4MmaWork::LDP == 100 && ShellQ4MmaWork::oRty == 224 && ShellQ4MmaWork::SCR == 524 &&
                  ShellQ4MmaWork::oG == 192 && ShellQ4MmaWork::oRu == 428 && ShellQ9MmaWork::SCR == 2492,
              "work-area layouts");

// phase 2 (same barrier interval as shell_p2_tying), task q: frame, inverse Jacobian products, weighted
// determinant (as shell_p2_qgeom) and the frame products of the five tying fields,
// P[f][m] with (c,d) = (0,0) (1,1) (0,1) (1,2) (0,2) -- the expressions of shell_p3_weights without Ntq
template <int O, class WK>
TB2_HD void shell_unc_qgeom(int q, WK &w, const ShellTables<O> &tab, const double *desc) {
  constexpr int n = WK::n;
  const double *X = w.X();
  double Xxi[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0}, n0[3] = {0.0, 0.0, 0.0};
  double nxi[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  for (int j = 0; j < n; j++) {
    const double d0 = tab.dNq_T[j][0][q], d1 = tab.dNq_T[j][1][q], N = tab.Nq_T[j][q];
#pragma unroll
    for (int c = 0; c < 3; c++) {
      Xxi[2 * c] += d0 * X[3 * j + c];
      Xxi[2 * c + 1] += d1 * X[3 * j + c];
      n0[c] += N * w.fn[3 * j + c];
      nxi[2 * c] += d0 * w.fn[3 * j + c];
      nxi[2 * c + 1] += d1 * w.fn[3 * j + c];
    }
  }
  double T[9], Xd[9], Xdz[9], Xdinv[9], A[9], Az[9], tmp[9];
  shell_frame((int)desc[25], &desc[26], Xxi, n0, T);
#pragma unroll
  for (int c = 0; c < 3; c++) {
    Xd[3 * c] = Xxi[2 * c];
    Xd[3 * c + 1] = Xxi[2 * c + 1];
    Xd[3 * c + 2] = n0[c];
    Xdz[3 * c] = nxi[2 * c];
    Xdz[3 * c + 1] = nxi[2 * c + 1];
    Xdz[3 * c + 2] = 0.0;
  }
  const double det = inv3x3(Xd, Xdinv);
  mat3mul(Xdinv, Xdz, tmp);
#pragma unroll
  for (int k = 0; k < 9; k++) tmp[k] = -tmp[k];
  mat3mul(Xdinv, T, A);
  mat3mul(tmp, A, Az);
  w.wdet[q] = det * tab.wq[q];
  double *geo = w.geo[q];
  geo[0] = T[0]; geo[1] = T[1]; geo[2] = T[3]; geo[3] = T[4]; geo[4] = T[6]; geo[5] = T[7];
  geo[6] = A[0]; geo[7] = A[1]; geo[8] = A[3]; geo[9] = A[4];
  geo[10] = Az[0]; geo[11] = Az[1]; geo[12] = Az[3]; geo[13] = Az[4]; geo[14] = Az[6]; geo[15] = Az[7];
  double *P = w.scr + WK::oP + 30 * q;
#pragma unroll
  for (int f = 0; f < 5; f++) {
    const int c = (f == 1 || f == 3) ? 1 : 0;
    const int d = (f == 0) ? 0 : ((f == 1 || f == 2) ? 1 : 2);
    const double off = (c == d) ? 0.0 : 1.0;
    const double Ac0 = A[3 * c], Ac1 = A[3 * c + 1], Ac2 = A[3 * c + 2];
    const double Ad0 = A[3 * d], Ad1 = A[3 * d + 1], Ad2 = A[3 * d + 2];
    P[6 * f + 0] = Ac0 * Ad0 + off * (Ad0 * Ac0);
    P[6 * f + 1] = Ac1 * Ad1 + off * (Ad1 * Ac1);
    P[6 * f + 2] = 2.0 * (Ac0 * Ad1 + off * (Ad0 * Ac1));
    P[6 * f + 3] = 2.0 * (Ac1 * Ad2 + off * (Ad1 * Ac2));
    P[6 * f + 4] = 2.0 * (Ac0 * Ad2 + off * (Ad0 * Ac2));
    P[6 * f + 5] = 0.0;
  }
}

// G phase, task (q, f2): column f2 of G_q = P_q (w det C_TT) P_q^T, C_TT = [A 0; 0 As] on (e0,e1,e2 | e6,e7)
template <int O, class WK>
TB2_HD void shell_unc_G(int task, WK &w, const double *desc) {
  const int q = task / 5, f2 = task % 5;
  const double *P = w.scr + WK::oP + 30 * q;
  const double wd = w.wdet[q];
  double b[6], v[5];
  load6(P + 6 * f2, b);
  v[0] = wd * (desc[0] * b[0] + desc[1] * b[1] + desc[2] * b[2]);
  v[1] = wd * (desc[1] * b[0] + desc[3] * b[1] + desc[4] * b[2]);
  v[2] = wd * (desc[2] * b[0] + desc[4] * b[1] + desc[5] * b[2]);
  v[3] = wd * (desc[18] * b[3] + desc[19] * b[4]);
  v[4] = wd * (desc[19] * b[3] + desc[20] * b[4]);
  double *G = w.scr + WK::oG + 26 * q;
#pragma unroll
  for (int f1 = 0; f1 < 5; f1++) {
    double a[6];
    load6(P + 6 * f1, a);
    G[5 * f1 + f2] = a[0] * v[0] + a[1] * v[1] + a[2] * v[2] + a[3] * v[3] + a[4] * v[4];
  }
}

// entry k of the upper triangle (t1 <= t2) of the symmetric nty x nty matrix S, packed for the kernel as
// t1 | t2 << 8 | (5 f1 + f2) << 16 (decoded once per thread, outside the element loop)
template <int O>
TB2_HD int shell_unc_tri(int k) {
  constexpr int nty = ShellDims<O>::nty;
  int t1 = 0;
  while (k >= nty - t1) {
    k -= nty - t1;
    t1++;
  }
  const int t2 = t1 + k;
  return t1 | (t2 << 8) | ((5 * shell_ty_field<O>(t1) + shell_ty_field<O>(t2)) << 16);
}

// S phase, one packed entry: S[t1][t2] = S[t2][t1] = sum_q Ntq[q][t1] Ntq[q][t2] G_q[f1][f2]
template <int O, class WK>
TB2_HD void shell_unc_S_entry(int packed, WK &w, const ShellTables<O> &tab) {
  constexpr int nq = WK::nq;
  const int t1 = packed & 0xff, t2 = (packed >> 8) & 0xff, g = packed >> 16;
  const double *G = w.scr + WK::oG + g;
  double s = 0.0;
#pragma unroll
  for (int q = 0; q < nq; q++) s += (tab.Ntq[q][t1] * tab.Ntq[q][t2]) * G[26 * q];
  w.scr[WK::oS + t1 * WK::LDS_ + t2] = s;
  w.scr[WK::oS + t2 * WK::LDS_ + t1] = s;
}

// row buffer of quadrature point q, task (j, c): the columns 6j+c and 6j+3+c of
//   L rows 0..2: bending rows 3,4,5 of B (same expressions as shell_p3_columns);  L row 3: drill row
//   R rows 0..2: (w det D) L;                                                     R row 3: (w det drill) L3
template <int O, class WK, bool LONLY = false>
TB2_HD void shell_unc_rows(int task, int q, WK &w, const ShellTables<O> &tab, const double *desc, double *buf) {
  constexpr int n = ShellDims<O>::n;
  const int c = task % 3, j = (task / 3) % n;
  const int cu = 6 * j + c, cq = cu + 3;
  const double d0 = tab.dNq[q][j][0], d1 = tab.dNq[q][j][1], N = tab.Nq[q][j];
  const double *geo = w.geo[q];
  const double *T = geo, *A = geo + 6, *Az = geo + 10;  // T(r,0..1) at T[2r..], A(0..1,0..1) at A[2r..], Az rows 0..2
  const double hz0 = d0 * Az[0] + d1 * Az[2], hz1 = d0 * Az[1] + d1 * Az[3];
  const double h0 = d0 * A[0] + d1 * A[2] + N * Az[4], h1 = d0 * A[1] + d1 * A[3] + N * Az[5];
  double bu[4], bq[4];
  bu[0] = T[2 * c] * hz0;
  bu[1] = T[2 * c + 1] * hz1;
  bu[2] = T[2 * c] * hz1 + T[2 * c + 1] * hz0;
  const int c1 = (c + 1) % 3, c2 = (c + 2) % 3;
  const double f1 = w.fn[3 * j + c1], f2 = w.fn[3 * j + c2];
  bq[0] = f1 * (T[2 * c2] * h0) - f2 * (T[2 * c1] * h0);
  bq[1] = f1 * (T[2 * c2 + 1] * h1) - f2 * (T[2 * c1 + 1] * h1);
  bq[2] = f1 * (T[2 * c2] * h1 + T[2 * c2 + 1] * h0) - f2 * (T[2 * c1] * h1 + T[2 * c1 + 1] * h0);
  // drill strain: nodal rows interpolated with the nodal shape functions; the rotation columns of node i's
  // row are non-zero at node i only
  {
    double su = 0.0;
    for (int i = 0; i < n; i++) su += tab.Nq[q][i] * w.bdr_u(i, j, c);
    bu[3] = su;
    bq[3] = N * w.bdr_q(j, c);
  }
  // D block of the descriptor at [12..17], packed [0 1 2; 1 3 4; 2 4 5]
  const double *Dm = desc + 12;
  const double wd = w.wdet[q];
  double ru[4], rq[4];
#pragma unroll
  for (int r = 0; r < 3; r++) {
    const int i0 = (r == 0) ? 0 : ((r == 1) ? 1 : 2), i1 = (r == 0) ? 1 : ((r == 1) ? 3 : 4),
              i2 = (r == 0) ? 2 : ((r == 1) ? 4 : 5);
    ru[r] = wd * (Dm[i0] * bu[0] + Dm[i1] * bu[1] + Dm[i2] * bu[2]);
    rq[r] = wd * (Dm[i0] * bq[0] + Dm[i1] * bq[1] + Dm[i2] * bq[2]);
  }
  const double wdr = wd * desc[21];
  ru[3] = wdr * bu[3];
  rq[3] = wdr * bq[3];
  // half-split panels X[k >> 1][col][k & 1] (ShellMmaWork): four 16-byte pieces per column
  double *L = buf, *R = buf + WK::LPAN;
#if defined(__CUDA_ARCH__)
  *reinterpret_cast<double2 *>(L + 2 * cu) = make_double2(bu[0], bu[1]);
  *reinterpret_cast<double2 *>(L + WK::HS + 2 * cu) = make_double2(bu[2], bu[3]);
  *reinterpret_cast<double2 *>(L + 2 * cq) = make_double2(bq[0], bq[1]);
  *reinterpret_cast<double2 *>(L + WK::HS + 2 * cq) = make_double2(bq[2], bq[3]);
  if (LONLY) return;
  *reinterpret_cast<double2 *>(R + 2 * cu) = make_double2(ru[0], ru[1]);
  *reinterpret_cast<double2 *>(R + WK::HS + 2 * cu) = make_double2(ru[2], ru[3]);
  *reinterpret_cast<double2 *>(R + 2 * cq) = make_double2(rq[0], rq[1]);
  *reinterpret_cast<double2 *>(R + WK::HS + 2 * cq) = make_double2(rq[2], rq[3]);
#else
  for (int r = 0; r < 4; r++) {
    L[(r >> 1) * WK::HS + 2 * cu + (r & 1)] = bu[r]; L[(r >> 1) * WK::HS + 2 * cq + (r & 1)] = bq[r];
    if (!LONLY) { R[(r >> 1) * WK::HS + 2 * cu + (r & 1)] = ru[r]; R[(r >> 1) * WK::HS + 2 * cq + (r & 1)] = rq[r]; }
  }
#endif
}

// ------------------------------------------------------------------------------------------
// residual-only path of the uncoupled shells (assembleRes): no tangent, no S; the state is pushed through the
// tying space,  res = Bty^T r_ty + sum_q L_q^T (w det [D 0; 0 drill]) L_q u,
//   t = Bty u;  e_q = W_q t;  s_q = w det [A 0; 0 As] e_q;  r_ty = sum_q W_q^T s_q;   W_q[ty][m] = Ntq[q][ty] P_q[f(ty)][m]
// ------------------------------------------------------------------------------------------
// task ty: tying strain of the state
template <int O, class WK>
TB2_HD void shell_unc_res_tying(int ty, WK &w, const double *u, double *tvec) {
  constexpr int nd = ShellDims<O>::nd;
  double s = 0.0;
  for (int col = 0; col < nd; col++) s += w.bty(ty, col) * u[col];
  tvec[ty] = s;
}

// task q: membrane / transverse-shear strains and stresses at quadrature point q -> s5[6 q + m]
template <int O, class WK>
TB2_HD void shell_unc_res_point(int q, WK &w, const ShellTables<O> &tab, const double *desc, const double *tvec,
                                double *s5) {
  constexpr int nty = ShellDims<O>::nty;
  const double *P = w.scr + WK::oP + 30 * q;
  double e[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
  for (int ty = 0; ty < nty; ty++) {
    const double a = tab.Ntq[q][ty] * tvec[ty];
    double p[6];
    load6(P + 6 * shell_ty_field<O>(ty), p);
#pragma unroll
    for (int m = 0; m < 5; m++) e[m] += a * p[m];
  }
  const double wd = w.wdet[q];
  double *so = s5 + 6 * q;
  so[0] = wd * (desc[0] * e[0] + desc[1] * e[1] + desc[2] * e[2]);
  so[1] = wd * (desc[1] * e[0] + desc[3] * e[1] + desc[4] * e[2]);
  so[2] = wd * (desc[2] * e[0] + desc[4] * e[1] + desc[5] * e[2]);
  so[3] = wd * (desc[18] * e[3] + desc[19] * e[4]);
  so[4] = wd * (desc[19] * e[3] + desc[20] * e[4]);
  so[5] = 0.0;
}

// task ty: stresses projected back on the tying point
template <int O, class WK>
TB2_HD void shell_unc_res_back(int ty, WK &w, const ShellTables<O> &tab, const double *s5, double *sty) {
  constexpr int nq = ShellDims<O>::nq;
  const int f = shell_ty_field<O>(ty);
  double sum = 0.0;
  for (int q = 0; q < nq; q++) {
    double p[6], sv[6];
    load6(w.scr + WK::oP + 30 * q + 6 * f, p);
    load6(s5 + 6 * q, sv);
    sum += tab.Ntq[q][ty] * (p[0] * sv[0] + p[1] * sv[1] + p[2] * sv[2] + p[3] * sv[3] + p[4] * sv[4]);
  }
  sty[ty] = sum;
}

// task (r, part) with r in [0,4), part in [0,PARTS): partial strain r of the row buffer (bending 0..2, drill 3) for
// the state over the columns part, part + PARTS, ...; the kernel sums the PARTS partials of a row with shuffles
// (tasks of one row sit in adjacent lanes), the host replay adds them in the same order
template <int O, class WK, int PARTS>
TB2_HD double shell_unc_res_rowstrain(int task, WK &w, const double *L, const double *u) {
  constexpr int nd = ShellDims<O>::nd;
  const int r = task / PARTS, part = task % PARTS;
  const double *row = L + (r >> 1) * WK::HS + (r & 1);
  double s = 0.0;
  for (int col = part; col < nd; col += PARTS) s += row[2 * col] * u[col];
  return s;
}

// dof `col`: contribution of the row buffer of point q, L^T (w det [D 0; 0 drill]) t4
template <int O, class WK>
TB2_HD double shell_unc_res_rowback(int col, int q, WK &w, const double *desc, const double *L, const double *t4) {
  const double wd = w.wdet[q];
  const double *Dm = desc + 12;
  const double v0 = wd * (Dm[0] * t4[0] + Dm[1] * t4[1] + Dm[2] * t4[2]);
  const double v1 = wd * (Dm[1] * t4[0] + Dm[3] * t4[1] + Dm[4] * t4[2]);
  const double v2 = wd * (Dm[2] * t4[0] + Dm[4] * t4[1] + Dm[5] * t4[2]);
  const double v3 = (wd * desc[21]) * t4[3];
  return L[2 * col] * v0 + L[2 * col + 1] * v1 + L[WK::HS + 2 * col] * v2 + L[WK::HS + 2 * col + 1] * v3;
}

// residual-only path (assembleRes): strains of the chunk, task (ql, r): e = B u, kept in the unused CB rows
template <int O, int QC>
TB2_HD void shell_res_strain(int task, ShellWork<O, QC> &w) {
  constexpr int nd = ShellDims<O>::nd;
  const int ql = task / 9, r = task % 9;
  double e = 0.0;
  for (int k = 0; k < nd; k++) e += w.B[ql][r][k] * w.u[k];
  w.CB[ql][0][r] = e;
}

// residual-only path, task dof k: returns sum_ql sum_r B[ql][r][k] * (w det C e)[r]
template <int O, int QC>
TB2_HD double shell_res_accumulate(int k, ShellWork<O, QC> &w) {
  double out = 0.0;
  for (int ql = 0; ql < QC; ql++) {
    const double *C = w.Cw[ql], *e = &w.CB[ql][0][0];
    double s[9];
#pragma unroll
    for (int r = 0; r < 3; r++) {
      const int i0 = (r == 0) ? 0 : ((r == 1) ? 1 : 2), i1 = (r == 0) ? 1 : ((r == 1) ? 3 : 4),
                i2 = (r == 0) ? 2 : ((r == 1) ? 4 : 5);
      s[r] = C[i0] * e[0] + C[i1] * e[1] + C[i2] * e[2] + C[6 + i0] * e[3] + C[6 + i1] * e[4] + C[6 + i2] * e[5];
      s[3 + r] = C[6 + i0] * e[0] + C[6 + i1] * e[1] + C[6 + i2] * e[2] + C[12 + i0] * e[3] + C[12 + i1] * e[4] +
                 C[12 + i2] * e[5];
    }
    s[6] = C[18] * e[6] + C[19] * e[7];
    s[7] = C[19] * e[6] + C[20] * e[7];
    s[8] = C[21] * e[8];
#pragma unroll
    for (int r = 0; r < 9; r++) out += w.B[ql][r][k] * s[r];
  }
  return out;
}

// phase 5 (all families): one TRxTC tile of K accumulates B^T (CB) over the rows of this chunk
template <int NROWS, int LD, int TR, int TC, int LDC = LD>
TB2_HD void tile_accumulate(const double *B, const double *CB, int row0, int col0, double *acc) {
#pragma unroll 3
  for (int r = 0; r < NROWS; r++) {
    double bi[TR], cj[TC];
    if (TR == 6 && TC == 6 && LD % 2 == 0 && LDC % 2 == 0) {
      // rows are 16-byte aligned and the tile starts at a multiple of 6 doubles: 128-bit loads
      load6(&B[r * LD + row0], bi);
      load6(&CB[r * LDC + col0], cj);
    } else {
#pragma unroll
      for (int a = 0; a < TR; a++) bi[a] = B[r * LD + row0 + a];
#pragma unroll
      for (int b = 0; b < TC; b++) cj[b] = CB[r * LDC + col0 + b];
    }
#pragma unroll
    for (int a = 0; a < TR; a++)
#pragma unroll
      for (int b = 0; b < TC; b++) acc[a * TC + b] += bi[a] * cj[b];
  }
}

// inertial block of node pair (i,j): M = int rho-moments N_i N_j [I, D_j; D_i^T, D_i^T D_j] (director d = q x t)
template <int O, class WK>
TB2_HD void shell_mass_tile(int tile, WK &w, const ShellTables<O> &tab, const double *desc, double *M) {
  constexpr int n = ShellDims<O>::n, nq = ShellDims<O>::nq;
  const int i = tile / n, j = tile % n;
  double S = 0.0;
  for (int q = 0; q < nq; q++) S += w.wdet[q] * tab.Nq[q][i] * tab.Nq[q][j];
  const double m0 = desc[22], m1 = desc[23], m2 = desc[24];
  // d = D q with D(c,e) = eps_{cef} t_f
  const double *ti = &w.fn[3 * i], *tj = &w.fn[3 * j];
  const double Di[9] = {0.0, ti[2], -ti[1], -ti[2], 0.0, ti[0], ti[1], -ti[0], 0.0};
  const double Dj[9] = {0.0, tj[2], -tj[1], -tj[2], 0.0, tj[0], tj[1], -tj[0], 0.0};
#pragma unroll
  for (int k = 0; k < 36; k++) M[k] = 0.0;
#pragma unroll
  for (int c = 0; c < 3; c++) {
    M[6 * c + c] = S * m0;
#pragma unroll
    for (int e = 0; e < 3; e++) {
      M[6 * c + 3 + e] = S * m1 * Dj[3 * c + e];
      M[6 * (3 + e) + c] = S * m1 * Di[3 * c + e];
    }
  }
#pragma unroll
  for (int e = 0; e < 3; e++)
#pragma unroll
    for (int f = 0; f < 3; f++) {
      double v = 0.0;
#pragma unroll
      for (int c = 0; c < 3; c++) v += Di[3 * c + e] * Dj[3 * c + f];
      M[6 * (3 + e) + 3 + f] = S * m2 * v;
    }
}

// phase 6, task tile (i,j): inertial block (only when `inertia`), residual partials; on return acc holds
// alpha*K_tile + gamma*M_tile. Runs after the last chunk barrier: the partials reuse the B rows.
template <int O, class WK>
TB2_HD void shell_p6_finish(int tile, WK &w, const ShellTables<O> &tab, const double *desc,
                            double alpha, double gamma, bool inertia, double *acc, double *rp) {
  constexpr int n = ShellDims<O>::n;
  const int j = tile % n;
#pragma unroll
  for (int a = 0; a < 6; a++) {
    double s = 0.0;
#pragma unroll
    for (int b = 0; b < 6; b++) s += acc[6 * a + b] * w.uvec()[6 * j + b];
    rp[a] = s;
  }
#pragma unroll
  for (int k = 0; k < 36; k++) acc[k] *= alpha;
  if (inertia) {
    double M[36];
    shell_mass_tile<O>(tile, w, tab, desc, M);
#pragma unroll
    for (int a = 0; a < 6; a++) {
      double s = 0.0;
#pragma unroll
      for (int b = 0; b < 6; b++) s += M[6 * a + b] * w.avec()[6 * j + b];
      rp[a] += s;
    }
#pragma unroll
    for (int k = 0; k < 36; k++) acc[k] += gamma * M[k];
  }
}

// ------------------------------------------------------------------------------------------
// solid work area and phases
// ------------------------------------------------------------------------------------------
template <int O, int QC>
struct SolidWork {
  using D = SolidDims<O>;
  static constexpr int n = D::n, nd = D::nd, nq = D::nq;
  static constexpr int NS = 6;
  static constexpr int TR = (O == 2) ? 6 : 9, TC = (O == 2) ? 6 : 3;  // hex27: 9x3 tiles, 243 per element (8 warps)
  static constexpr int ntiles = (nd / TR) * (nd / TC);
  static constexpr int LD = nd;
  double X[3 * n];
  double u[nd];
  double acc[nd];
  double desc[kDescStride];
  double J[nq][9];
  double wdet[nq];
  alignas(16) double G[QC][nd];  // physical shape-function gradients: G[ql][3a+dir] = dN_a/dx_dir
  // w det C B; the strain rows B themselves are implied by G (3 non-zeros per column). Each quadrature slab is
  // padded by 8 doubles so that the lanes of two slabs (one lane per node and slab) store to distinct banks.
  static constexpr int CBQ = NS * nd + 8;
  double CB[QC][CBQ];
  double rpart[ntiles][TR];
};

// phase 1, task q: Jacobian inverse and weighted determinant
template <int O, int QC>
TB2_HD void solid_p1_qgeom(int q, SolidWork<O, QC> &w, const SolidTables<O> &tab) {
  constexpr int n = SolidDims<O>::n;
  double Xd[9] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0}, J[9];
  for (int a = 0; a < n; a++) {
#pragma unroll
    for (int c = 0; c < 3; c++)
#pragma unroll
      for (int k = 0; k < 3; k++) Xd[3 * c + k] += tab.dNq[q][a][k] * w.X[3 * a + c];
  }
  const double det = inv3x3(Xd, J);
  w.wdet[q] = det * tab.wq[q];
#pragma unroll
  for (int k = 0; k < 9; k++) w.J[q][k] = J[k];
}

// phase 3, task (ql, a): strain rows and C-weighted rows for the three dofs of node a
template <int O, int QC>
TB2_HD void solid_p3_bcols(int task, int q0, SolidWork<O, QC> &w, const SolidTables<O> &tab) {
  constexpr int n = SolidDims<O>::n;
  const int a = task % n, ql = task / n, q = q0 + ql;
  const double *J = w.J[q];
  const double x0 = tab.dNq[q][a][0], x1 = tab.dNq[q][a][1], x2 = tab.dNq[q][a][2];
  const double gx = x0 * J[0] + x1 * J[3] + x2 * J[6];
  const double gy = x0 * J[1] + x1 * J[4] + x2 * J[7];
  const double gz = x0 * J[2] + x1 * J[5] + x2 * J[8];
  // strain order xx,yy,zz,yz,xz,xy with engineering shears (TACSLinearElasticity.cpp:1186-1192):
  //   column x: (gx,0,0,0,gz,gy)   column y: (0,gy,0,gz,0,gx)   column z: (0,0,gz,gy,gx,0)
  const double *C = w.desc;
  // symmetric 6x6 from the upper triangle stored by rows
  const int idx[6][6] = {{0, 1, 2, 3, 4, 5},     {1, 6, 7, 8, 9, 10},    {2, 7, 11, 12, 13, 14},
                         {3, 8, 12, 15, 16, 17}, {4, 9, 13, 16, 18, 19}, {5, 10, 14, 17, 19, 20}};
  const double wd = w.wdet[q];
  const int c0 = 3 * a;
  w.G[ql][c0] = gx;
  w.G[ql][c0 + 1] = gy;
  w.G[ql][c0 + 2] = gz;
#pragma unroll
  for (int r = 0; r < 6; r++) {
    // only the three structurally non-zero strain entries of each displacement column contribute
    double *cb = &w.CB[ql][r * SolidDims<O>::nd + c0];
    cb[0] = wd * (C[idx[r][0]] * gx + C[idx[r][4]] * gz + C[idx[r][5]] * gy);
    cb[1] = wd * (C[idx[r][1]] * gy + C[idx[r][3]] * gz + C[idx[r][5]] * gx);
    cb[2] = wd * (C[idx[r][2]] * gz + C[idx[r][3]] * gy + C[idx[r][4]] * gx);
  }
}

// residual-only path (assembleRes), task (ql, r): strain e_r = (B u)_r from the gradients, staged in rpart
// (free until phase 6). Strain order xx,yy,zz,yz,xz,xy with engineering shears.
template <int O, int QC>
TB2_HD void solid_res_strain(int task, SolidWork<O, QC> &w) {
  constexpr int n = SolidDims<O>::n;
  const int ql = task / 6, r = task % 6;
  // e_r = sum_a G[a][d1] u[a][c1] + G[a][d2] u[a][c2]   (normal strains use one term)
  const int c1 = (r < 3) ? r : ((r == 3) ? 1 : 0);
  const int d1 = (r < 3) ? r : ((r == 5) ? 1 : 2);
  const int c2 = (r == 5) ? 1 : 2;
  const int d2 = (r == 3) ? 1 : 0;
  const double two = (r < 3) ? 0.0 : 1.0;
  double e = 0.0;
  for (int a = 0; a < n; a++)
    e += w.G[ql][3 * a + d1] * w.u[3 * a + c1] + two * (w.G[ql][3 * a + d2] * w.u[3 * a + c2]);
  (&w.rpart[0][0])[6 * ql + r] = e;
}

// residual-only path, task dof k: sum_ql sum_r (w det C B)[ql][r][k] e[ql][r]   (C symmetric)
template <int O, int QC>
TB2_HD double solid_res_accumulate(int k, SolidWork<O, QC> &w) {
  const double *e = &w.rpart[0][0];
  double out = 0.0;
  for (int ql = 0; ql < QC; ql++)
#pragma unroll
    for (int r = 0; r < 6; r++) out += w.CB[ql][r * SolidDims<O>::nd + k] * e[6 * ql + r];
  return out;
}

// phase 5 for solids: a TRxTC tile of K accumulates B^T (CB) using the three structurally non-zero strain
// entries of every displacement column (x: rows 0,4,5 = gx,gz,gy; y: rows 1,3,5 = gy,gz,gx; z: rows 2,3,4 =
// gz,gy,gx) -- half the multiply-adds of the dense row product.
template <int QC, int ND, int TR, int TC>
TB2_HD void solid_tile_accumulate(const double *G, const double *CB, int row0, int col0, double *acc) {
  for (int ql = 0; ql < QC; ql++) {
    const double *g = G + ql * ND + row0;
    const double *cbq = CB + ql * (6 * ND + 8) + col0;  // SolidWork::CBQ
#pragma unroll
    for (int bn = 0; bn < TC / 3; bn++) {
      double cb[6][3];
#pragma unroll
      for (int r = 0; r < 6; r++)
#pragma unroll
        for (int b = 0; b < 3; b++) cb[r][b] = cbq[r * ND + 3 * bn + b];
#pragma unroll
      for (int an = 0; an < TR / 3; an++) {
        const double gx = g[3 * an], gy = g[3 * an + 1], gz = g[3 * an + 2];
#pragma unroll
        for (int b = 0; b < 3; b++) {
          double *ax = &acc[(3 * an) * TC + 3 * bn + b];
          ax[0] += gx * cb[0][b] + gz * cb[4][b] + gy * cb[5][b];
          ax[TC] += gy * cb[1][b] + gz * cb[3][b] + gx * cb[5][b];
          ax[2 * TC] += gz * cb[2][b] + gy * cb[3][b] + gx * cb[4][b];
        }
      }
    }
  }
}

// ---- geometric stiffness (TACSElement3D::getMatType(TACS_GEOMETRIC_STIFFNESS_MATRIX), TACSElement3D.cpp:316-360 with
// TACSLinearElasticity3D::evalWeakMatrix, TACSLinearElasticity.cpp:1704-1800): with the stress s = C e(u) of the
// current state, K_G[3a+i][3b+j] = delta_ij sum_q w det (grad N_a . S grad N_b), S the symmetric stress tensor.
// task (ql, a): gradients only
template <int O, int QC>
TB2_HD void solid_geo_grad(int task, int q0, SolidWork<O, QC> &w, const SolidTables<O> &tab) {
  constexpr int n = SolidDims<O>::n;
  const int a = task % n, ql = task / n, q = q0 + ql;
  const double *J = w.J[q];
  const double x0 = tab.dNq[q][a][0], x1 = tab.dNq[q][a][1], x2 = tab.dNq[q][a][2];
  w.G[ql][3 * a] = x0 * J[0] + x1 * J[3] + x2 * J[6];
  w.G[ql][3 * a + 1] = x0 * J[1] + x1 * J[4] + x2 * J[7];
  w.G[ql][3 * a + 2] = x0 * J[2] + x1 * J[5] + x2 * J[8];
}
// task ql: strains (left in rpart by solid_res_strain) -> stresses, in place
template <int O, int QC>
TB2_HD void solid_geo_stress(int ql, SolidWork<O, QC> &w) {
  double *e = &w.rpart[0][0] + 6 * ql;
  const double *C = w.desc;
  const int idx[6][6] = {{0, 1, 2, 3, 4, 5},     {1, 6, 7, 8, 9, 10},    {2, 7, 11, 12, 13, 14},
                         {3, 8, 12, 15, 16, 17}, {4, 9, 13, 16, 18, 19}, {5, 10, 14, 17, 19, 20}};
  double s[6];
#pragma unroll
  for (int r = 0; r < 6; r++) {
    s[r] = C[idx[r][0]] * e[0] + C[idx[r][1]] * e[1] + C[idx[r][2]] * e[2] + C[idx[r][3]] * e[3] + C[idx[r][4]] * e[4] +
           C[idx[r][5]] * e[5];
  }
#pragma unroll
  for (int r = 0; r < 6; r++) e[r] = s[r];
}
// task (ql, b): T_b = w det S grad N_b into the first three rows' worth of the CB slab, CB[ql][3b + dir]
template <int O, int QC>
TB2_HD void solid_geo_sgrad(int task, int q0, SolidWork<O, QC> &w) {
  constexpr int n = SolidDims<O>::n;
  const int b = task % n, ql = task / n;
  const double *s = &w.rpart[0][0] + 6 * ql;  // xx yy zz yz xz xy
  const double gx = w.G[ql][3 * b], gy = w.G[ql][3 * b + 1], gz = w.G[ql][3 * b + 2], wd = w.wdet[q0 + ql];
  double *t = &w.CB[ql][3 * b];
  t[0] = wd * (s[0] * gx + s[5] * gy + s[4] * gz);
  t[1] = wd * (s[5] * gx + s[1] * gy + s[3] * gz);
  t[2] = wd * (s[4] * gx + s[3] * gy + s[2] * gz);
}
// tile accumulation: the same scalar on the three diagonal entries of every node pair of the tile
template <int QC, int ND, int TR, int TC>
TB2_HD void solid_geo_accumulate(const double *G, const double *CB, int row0, int col0, double *acc) {
  for (int ql = 0; ql < QC; ql++) {
    const double *g = G + ql * ND + row0;
    const double *t = CB + ql * (6 * ND + 8) + col0;  // SolidWork::CBQ
#pragma unroll
    for (int an = 0; an < TR / 3; an++)
#pragma unroll
      for (int bn = 0; bn < TC / 3; bn++) {
        const double k = g[3 * an] * t[3 * bn] + g[3 * an + 1] * t[3 * bn + 1] + g[3 * an + 2] * t[3 * bn + 2];
#pragma unroll
        for (int c = 0; c < 3; c++) acc[(3 * an + c) * TC + 3 * bn + c] += k;
      }
  }
}

// phase 6, task tile: residual partials, consistent mass block (only when `inertia`);
// acc <- alpha*acc + gamma*M
template <int O, int QC>
TB2_HD void solid_p6_finish(int tile, SolidWork<O, QC> &w, const SolidTables<O> &tab, double alpha,
                            double gamma, bool inertia, double *acc) {
  using WK = SolidWork<O, QC>;
  constexpr int TR = WK::TR, TC = WK::TC, nq = WK::nq, ntc = WK::nd / TC;
  const int row0 = (tile / ntc) * TR, col0 = (tile % ntc) * TC;
#pragma unroll
  for (int a = 0; a < TR; a++) {
    double rp = 0.0;
#pragma unroll
    for (int b = 0; b < TC; b++) rp += acc[a * TC + b] * w.u[col0 + b];
    w.rpart[tile][a] = rp;
  }
#pragma unroll
  for (int k = 0; k < TR * TC; k++) acc[k] *= alpha;
  if (inertia) {
    const double rho = w.desc[21];
#pragma unroll
    for (int an = 0; an < TR / 3; an++)
#pragma unroll
      for (int bn = 0; bn < TC / 3; bn++) {
        const int na = row0 / 3 + an, nb = col0 / 3 + bn;
        double S = 0.0;
        for (int q = 0; q < nq; q++) S += w.wdet[q] * tab.Nq[q][na] * tab.Nq[q][nb];
        const double m = rho * S;
#pragma unroll
        for (int c = 0; c < 3; c++) {
          w.rpart[tile][3 * an + c] += m * w.acc[col0 + 3 * bn + c];
          acc[(3 * an + c) * TC + 3 * bn + c] += gamma * m;
        }
      }
  }
}

}  // namespace tb2
