// Shape-function / quadrature / MITC-tying tables shared by every element of a family.
//
// The tables are evaluated once on the host with the reference's own expressions and literals
// (15-digit Gauss points, product-form Lagrange polynomials) so that device arithmetic starts
// from bit-identical constants:
//   shell basis   /root/reference/src/elements/shell/TACSShellElementQuadBasis.h:17-30, 62-114, 125-150, 486-616
//   shell quadrature  .../shell/TACSShellElementQuadrature.h:13-27, 62-75
//   hexa basis    /root/reference/src/elements/basis/TACSHexaBasis.cpp:103-110, 137-190, 344-351, 378-448
//   Gauss literals  .../basis/TACSGaussQuadrature.h:23-30
// They are copied to device memory once per element family and staged into shared memory by each
// CTA with a bulk async copy (cp.async.bulk) before the first element is processed.
#pragma once

#if defined(__CUDACC__)
#define TB2_HD __host__ __device__ __forceinline__
#else
#define TB2_HD inline
#endif

namespace tb2 {

template <int O>
struct ShellDims {
  static constexpr int order = O;
  static constexpr int n = O * O;        // nodes
  static constexpr int nd = 6 * O * O;   // dofs
  static constexpr int nq = O * O;       // quadrature points
  static constexpr int c11 = O * (O - 1), c22 = O * (O - 1), c12 = (O - 1) * (O - 1);
  static constexpr int c23 = O * (O - 1), c13 = O * (O - 1);
  static constexpr int nty = c11 + c22 + c12 + c23 + c13;  // 9 (Quad4) / 28 (Quad9)
};

// field of a tying index in storage order g11,g22,g12,g23,g13
template <int O>
TB2_HD int shell_ty_field(int ty) {
  using D = ShellDims<O>;
  if (ty < D::c11) return 0;
  if (ty < D::c11 + D::c22) return 1;
  if (ty < D::c11 + D::c22 + D::c12) return 2;
  if (ty < D::c11 + D::c22 + D::c12 + D::c23) return 3;
  return 4;
}

template <int O>
struct alignas(16) ShellTables {
  using D = ShellDims<O>;
  double Nq[D::nq][D::n];        // shape functions at quadrature points
  double dNq[D::nq][D::n][2];    // parametric derivatives at quadrature points
  double wq[D::nq];              // quadrature weights
  double Ntq[D::nq][D::nty];     // tying-strain interpolation evaluated at quadrature points
  // tables of the phases whose tasks are "one lane per tying point / node / quadrature point" are stored with the
  // lane index as the fastest dimension, so a team's loads hit consecutive shared-memory banks:
  double Nt_T[D::n][D::nty], dNt_T[D::n][2][D::nty];  // shape functions / derivatives at tying points, [j][(k)][ty]
  double dNn_T[D::n][2][D::n];   // derivatives at the node parametric points, [j][k][i] = dN_j/dxi_k at node i
  double Nq_T[D::n][D::nq], dNq_T[D::n][2][D::nq];
};

template <int O>
struct SolidDims {
  static constexpr int order = O;
  static constexpr int n = O * O * O;
  static constexpr int nd = 3 * O * O * O;
  static constexpr int nq = O * O * O;
};

template <int O>
struct alignas(16) SolidTables {
  using D = SolidDims<O>;
  double Nq[D::nq][D::n];
  double dNq[D::nq][D::n][3];
  double wq[D::nq];
};


namespace tables_detail {
static const double GP1[1] = {0.0};
static const double GP2[2] = {-0.577350269189626, 0.577350269189626};
static const double GW2[2] = {1.0, 1.0};
static const double GP3[3] = {-0.774596669241483, 0.0, 0.774596669241483};
static const double GW3[3] = {5.0 / 9.0, 8.0 / 9.0, 5.0 / 9.0};
static const double LIN_TY[2] = {-1.0, 1.0};

inline void shape1d(int order, double u, double *N, double *dN) {
  if (order == 2) {
    N[0] = 0.5 * (1.0 - u);
    N[1] = 0.5 * (1.0 + u);
    dN[0] = -0.5;
    dN[1] = 0.5;
  } else {
    N[0] = -0.5 * u * (1.0 - u);
    N[1] = (1.0 - u) * (1.0 + u);
    N[2] = 0.5 * (1.0 + u) * u;
    dN[0] = -0.5 + u;
    dN[1] = -2.0 * u;
    dN[2] = 0.5 + u;
  }
}

inline void lagrange(int n, double u, const double *knots, double *N) {
  for (int i = 0; i < n; i++) {
    N[i] = 1.0;
    for (int j = 0; j < n; j++) {
      if (i != j) {
        double d = 1.0 / (knots[i] - knots[j]);
        N[i] *= (u - knots[j]) * d;
      }
    }
  }
}

template <int O>
inline void shape2d(const double pt[2], double *N, double (*dN)[2]) {
  double na[3], dna[3], nb[3], dnb[3];
  shape1d(O, pt[0], na, dna);
  shape1d(O, pt[1], nb, dnb);
  for (int j = 0; j < O; j++)
    for (int i = 0; i < O; i++) {
      int k = i + O * j;
      if (N) N[k] = na[i] * nb[j];
      dN[k][0] = dna[i] * nb[j];
      dN[k][1] = na[i] * dnb[j];
    }
}
}  // namespace tables_detail

template <int O>
inline void build_shell_tables(ShellTables<O> &t) {
  using namespace tables_detail;
  using D = ShellDims<O>;
  const double *gp = (O == 2) ? GP2 : GP3, *gw = (O == 2) ? GW2 : GW3;
  const double *full = (O == 2) ? LIN_TY : GP3, *red = (O == 2) ? GP1 : GP2;
  for (int q = 0; q < D::nq; q++) {
    double pt[2] = {gp[q % O], gp[q / O]};
    t.wq[q] = gw[q % O] * gw[q / O];
    shape2d<O>(pt, t.Nq[q], t.dNq[q]);
    // evalTyingInterp
    double na[3], nb[3], nar[2], nbr[2];
    lagrange(O, pt[0], full, na);
    lagrange(O, pt[1], full, nb);
    lagrange(O - 1, pt[0], red, nar);
    lagrange(O - 1, pt[1], red, nbr);
    double *N = t.Ntq[q];
    for (int j = 0; j < O; j++) for (int i = 0; i < O - 1; i++) *N++ = nar[i] * nb[j];
    for (int j = 0; j < O - 1; j++) for (int i = 0; i < O; i++) *N++ = na[i] * nbr[j];
    for (int j = 0; j < O - 1; j++) for (int i = 0; i < O - 1; i++) *N++ = nar[i] * nbr[j];
    for (int j = 0; j < O - 1; j++) for (int i = 0; i < O; i++) *N++ = na[i] * nbr[j];
    for (int j = 0; j < O; j++) for (int i = 0; i < O - 1; i++) *N++ = nar[i] * nb[j];
  }
  static double Nt[D::nty][D::n], dNt[D::nty][D::n][2], dNn[D::n][D::n][2];  // host-side staging of the tables
  for (int i = 0; i < D::n; i++) {
    double pt[2] = {-1.0 + (2.0 / (O - 1)) * (i % O), -1.0 + (2.0 / (O - 1)) * (i / O)};
    shape2d<O>(pt, nullptr, dNn[i]);
  }
  const int cnt[5] = {D::c11, D::c22, D::c12, D::c23, D::c13};
  for (int index = 0; index < D::nty; index++) {
    int f = 0, ty = index;
    while (ty >= cnt[f]) { ty -= cnt[f]; f++; }
    double pt[2];
    if (f == 0 || f == 4) { pt[0] = red[ty % (O - 1)]; pt[1] = full[ty / (O - 1)]; }
    else if (f == 1 || f == 3) { pt[0] = full[ty % O]; pt[1] = red[ty / O]; }
    else { pt[0] = red[ty % (O - 1)]; pt[1] = red[ty / (O - 1)]; }
    shape2d<O>(pt, Nt[index], dNt[index]);
  }
  for (int j = 0; j < D::n; j++) {
    for (int ty = 0; ty < D::nty; ty++) {
      t.Nt_T[j][ty] = Nt[ty][j];
      t.dNt_T[j][0][ty] = dNt[ty][j][0];
      t.dNt_T[j][1][ty] = dNt[ty][j][1];
    }
    for (int i = 0; i < D::n; i++) {
      t.dNn_T[j][0][i] = dNn[i][j][0];
      t.dNn_T[j][1][i] = dNn[i][j][1];
    }
    for (int q = 0; q < D::nq; q++) {
      t.Nq_T[j][q] = t.Nq[q][j];
      t.dNq_T[j][0][q] = t.dNq[q][j][0];
      t.dNq_T[j][1][q] = t.dNq[q][j][1];
    }
  }
}

template <int O>
inline void build_solid_tables(SolidTables<O> &t) {
  using namespace tables_detail;
  using D = SolidDims<O>;
  const double *gp = (O == 2) ? GP2 : GP3, *gw = (O == 2) ? GW2 : GW3;
  const int o2 = O * O;
  for (int q = 0; q < D::nq; q++) {
    double pt[3] = {gp[q % O], gp[(q % o2) / O], gp[q / o2]};
    t.wq[q] = gw[q % O] * gw[(q % o2) / O] * gw[q / o2];
    double n1[3], d1[3], n2[3], d2[3], n3[3], d3[3];
    shape1d(O, pt[0], n1, d1);
    shape1d(O, pt[1], n2, d2);
    shape1d(O, pt[2], n3, d3);
    for (int k = 0, a = 0; k < O; k++)
      for (int j = 0; j < O; j++)
        for (int i = 0; i < O; i++, a++) {
          t.Nq[q][a] = n1[i] * n2[j] * n3[k];
          t.dNq[q][a][0] = d1[i] * n2[j] * n3[k];
          t.dNq[q][a][1] = n1[i] * d2[j] * n3[k];
          t.dNq[q][a][2] = n1[i] * n2[j] * d3[k];
        }
  }
}


}  // namespace tb2
