/*
 * oracle/tacs_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked into tacs_b200/).
 *
 * Plain-C, single-threaded CPU restatement of the reference's assembly + SpMV
 * hot path, used as the checker for the CUDA kernels:
 *
 *   shell elements  TACSShellElement<...>::addJacobian/addResidual
 *                   (/root/reference/src/elements/shell/TACSShellElement.h:294-641)
 *   solid elements  TACSElement3D::addJacobian/addResidual
 *                   (/root/reference/src/elements/TACSElement3D.cpp:139-224)
 *   constitutive    TACSIsoShell / TACSCompositeShell / TACSSolid evalTangentStiffness
 *   integer pipeline  TACSCreator::partitionMesh first-touch numbering, createTACS
 *                   element ordering, TACSAssembler::computeLocalNodeToNodeCSR
 *   scatter / BCs / SpMV  BCSRMat::addRowValues, zeroRow, BCSRMatVecMult{3,6}
 *
 * The element routines restate the reference's *mathematics* (strain definitions,
 * MITC tying interpolation, drill penalty, director parametrisation, constitutive
 * blocks, quadrature literals) as explicit strain-displacement rows B and
 * K = sum_q w det B^T C B rather than the reference's chain of first/second
 * derivative back-propagations; both are the same bilinear form for the linear
 * shell/solid models.  PARITY PIN: this file is validated against the compiled
 * reference itself (oracle/_ref/libtacs_ref.so, built by oracle/Makefile from the
 * unmodified sources) in tests/test_oracle_vs_reference.py and against the
 * committed fixtures tests/golden/ generated from it.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define MAXN 27   /* max nodes per element */
#define MAXTY 28  /* max tying points (Quad9) */

/* TacsGaussQuadrature.h:23-30 -- 15-digit literals, not computed values */
static const double GP1[1] = {0.0};
static const double GP2[2] = {-0.577350269189626, 0.577350269189626};
static const double GW2[2] = {1.0, 1.0};
static const double GP3[3] = {-0.774596669241483, 0.0, 0.774596669241483};
static const double GW3[3] = {5.0 / 9.0, 8.0 / 9.0, 5.0 / 9.0};
static const double LIN_TY[2] = {-1.0, 1.0}; /* TacsShellLinearTyingPoints, QuadBasis.h:116 */

/* ------------------------------------------------------------------ 3x3 helpers */
static void cross(const double x[3], const double y[3], double o[3]) {
  o[0] = x[1] * y[2] - x[2] * y[1];
  o[1] = x[2] * y[0] - x[0] * y[2];
  o[2] = x[0] * y[1] - x[1] * y[0];
}
static double dot3(const double x[3], const double y[3]) { return x[0] * y[0] + x[1] * y[1] + x[2] * y[2]; }

/* TACSElementAlgebra.h:1980 */
static double inv3(const double A[9], double Ai[9]) {
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
static void mm3(const double A[9], const double B[9], double C[9]) {
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}

/* ------------------------------------------------------------------ 1-D shape functions */
/* TACSShellElementQuadBasis.h:62-114, TACSHexaBasis.cpp:137-190,378-448 */
static void shape1d(int order, double u, double N[], double dN[]) {
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
/* TacsLagrangeShapeFunction, QuadBasis.h:17-30 */
static void lagrange(int n, double u, const double *knots, double N[]) {
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

/* 2-D tensor-product shape functions at pt: N[j], dN[2*j+k], node j = jx + order*jy */
static void shape2d(int order, const double pt[2], double N[], double dN[]) {
  double na[3], dna[3], nb[3], dnb[3];
  shape1d(order, pt[0], na, dna);
  shape1d(order, pt[1], nb, dnb);
  for (int j = 0; j < order; j++)
    for (int i = 0; i < order; i++) {
      int k = i + order * j;
      N[k] = na[i] * nb[j];
      dN[2 * k] = dna[i] * nb[j];
      dN[2 * k + 1] = na[i] * dnb[j];
    }
}

/* ------------------------------------------------------------------ shell transform */
/* kind 0: TACSShellNaturalTransform (TACSShellElementTransform.h:21-93, including the
   quirk that only t1[0] has the normal component removed, lines 42-44);
   kind 1: TACSShellRefAxisTransform (:95-215), axis already normalised by the ctor. */
static void shell_transform(int kind, const double axis[3], const double Xxi[6], const double n0[3],
                            double T[9]) {
  double n[3] = {n0[0], n0[1], n0[2]};
  double inv = 1.0 / sqrt(dot3(n, n));
  n[0] *= inv; n[1] *= inv; n[2] *= inv;
  double t1[3], t2[3];
  if (kind == 0) {
    t1[0] = Xxi[0]; t1[1] = Xxi[2]; t1[2] = Xxi[4];
    double d = dot3(n, t1);
    t1[0] = t1[0] - d * n[0];
    t1[0] = t1[0] - d * n[0];
    t1[0] = t1[0] - d * n[0];
  } else {
    double an = dot3(axis, n);
    t1[0] = axis[0] - an * n[0];
    t1[1] = axis[1] - an * n[1];
    t1[2] = axis[2] - an * n[2];
  }
  inv = 1.0 / sqrt(dot3(t1, t1));
  t1[0] *= inv; t1[1] *= inv; t1[2] *= inv;
  cross(n, t1, t2);
  T[0] = t1[0]; T[3] = t1[1]; T[6] = t1[2];
  T[1] = t2[0]; T[4] = t2[1]; T[7] = t2[2];
  T[2] = n[0];  T[5] = n[1];  T[8] = n[2];
}

/* ------------------------------------------------------------------ MITC tying scheme */
/* QuadBasis.h:125-142 (counts), :486-564 (field / point of a tying index) */
static int ty_counts(int order, int cnt[5]) {
  cnt[0] = order * (order - 1); /* g11 */
  cnt[1] = order * (order - 1); /* g22 */
  cnt[2] = (order - 1) * (order - 1); /* g12 */
  cnt[3] = order * (order - 1); /* g23 */
  cnt[4] = order * (order - 1); /* g13 */
  return cnt[0] + cnt[1] + cnt[2] + cnt[3] + cnt[4];
}
static void ty_knots(int order, const double **full, const double **red) {
  if (order == 2) { *full = LIN_TY; *red = GP1; }
  else { *full = GP3; *red = GP2; }
}
/* field id: 0 g11, 1 g22, 2 g12, 3 g23, 4 g13 (storage order of the reference) */
static void ty_point(int order, int index, int *field, double pt[2]) {
  int cnt[5];
  ty_counts(order, cnt);
  int f = 0, ty = index;
  while (ty >= cnt[f]) { ty -= cnt[f]; f++; }
  const double *full, *red;
  ty_knots(order, &full, &red);
  if (f == 0 || f == 4) { pt[0] = red[ty % (order - 1)]; pt[1] = full[ty / (order - 1)]; }
  else if (f == 1 || f == 3) { pt[0] = full[ty % order]; pt[1] = red[ty / order]; }
  else { pt[0] = red[ty % (order - 1)]; pt[1] = red[ty / (order - 1)]; }
  *field = f;
}
/* evalTyingInterp, QuadBasis.h:569-616 */
static void ty_interp(int order, const double pt[2], double N[]) {
  const double *full, *red;
  ty_knots(order, &full, &red);
  double na[3], nb[3], nar[2], nbr[2];
  lagrange(order, pt[0], full, na);
  lagrange(order, pt[1], full, nb);
  lagrange(order - 1, pt[0], red, nar);
  lagrange(order - 1, pt[1], red, nbr);
  for (int j = 0; j < order; j++) for (int i = 0; i < order - 1; i++) *N++ = nar[i] * nb[j];      /* g11 */
  for (int j = 0; j < order - 1; j++) for (int i = 0; i < order; i++) *N++ = na[i] * nbr[j];      /* g22 */
  for (int j = 0; j < order - 1; j++) for (int i = 0; i < order - 1; i++) *N++ = nar[i] * nbr[j]; /* g12 */
  for (int j = 0; j < order - 1; j++) for (int i = 0; i < order; i++) *N++ = na[i] * nbr[j];      /* g23 */
  for (int j = 0; j < order; j++) for (int i = 0; i < order - 1; i++) *N++ = nar[i] * nb[j];      /* g13 */
}

/* ------------------------------------------------------------------ shell element */
/* C22 = [A(6) B(6) D(6) As(3) drill]  (TACSShellConstitutive.cpp:36-55);
   stress = computeStress (TACSShellConstitutive.h:133-155) */
static void shell_stress(const double Cs[22], const double e[9], double s[9]) {
  const double *A = Cs, *B = Cs + 6, *D = Cs + 12, *As = Cs + 18;
  s[0] = A[0] * e[0] + A[1] * e[1] + A[2] * e[2] + B[0] * e[3] + B[1] * e[4] + B[2] * e[5];
  s[1] = A[1] * e[0] + A[3] * e[1] + A[4] * e[2] + B[1] * e[3] + B[3] * e[4] + B[4] * e[5];
  s[2] = A[2] * e[0] + A[4] * e[1] + A[5] * e[2] + B[2] * e[3] + B[4] * e[4] + B[5] * e[5];
  s[3] = B[0] * e[0] + B[1] * e[1] + B[2] * e[2] + D[0] * e[3] + D[1] * e[4] + D[2] * e[5];
  s[4] = B[1] * e[0] + B[3] * e[1] + B[4] * e[2] + D[1] * e[3] + D[3] * e[4] + D[4] * e[5];
  s[5] = B[2] * e[0] + B[4] * e[1] + B[5] * e[2] + D[2] * e[3] + D[4] * e[4] + D[5] * e[5];
  s[6] = As[0] * e[6] + As[1] * e[7];
  s[7] = As[1] * e[6] + As[2] * e[7];
  s[8] = Cs[21] * e[8];
}

/*
 * Quad4 (order 2) / Quad9 (order 3) MITC shell with linearised rotations.
 *   Xpts[3n], vars/ddvars[6n]; res[6n] += , mat[(6n)^2] += (row-major); either may be NULL.
 *   transform: 0 natural, 1 ref-axis (axis normalised by caller as the reference ctor does)
 *   Cs[22] tangent stiffness, moments[3] mass moments.
 */
void oracle_shell_element(int order, const double *Xpts, const double *vars, const double *ddvars,
                          int transform, const double *axis, const double *Cs, const double *moments,
                          double alpha, double beta, double gamma, double *res, double *mat) {
  (void)beta; /* no damping terms in the linear shell */
  const int n = order * order, nd = 6 * n;
  double fn[3 * 9], Xdn[9 * 9], Bdr[9][54], Bty[MAXTY][54];
  double N[9], dN[18];

  /* node normals: TacsShellComputeNodeNormals, TACSShellUtilities.h:301-342 */
  for (int i = 0; i < n; i++) {
    double pt[2] = {-1.0 + (2.0 / (order - 1)) * (i % order), -1.0 + (2.0 / (order - 1)) * (i / order)};
    shape2d(order, pt, N, dN);
    double Xxi[6] = {0, 0, 0, 0, 0, 0};
    for (int j = 0; j < n; j++)
      for (int c = 0; c < 3; c++) {
        Xxi[2 * c] += dN[2 * j] * Xpts[3 * j + c];
        Xxi[2 * c + 1] += dN[2 * j + 1] * Xpts[3 * j + c];
      }
    double a[3] = {Xxi[0], Xxi[2], Xxi[4]}, b[3] = {Xxi[1], Xxi[3], Xxi[5]};
    cross(a, b, &fn[3 * i]);
    double nrm = sqrt(dot3(&fn[3 * i], &fn[3 * i]));
    if (nrm != 0.0) {
      double inv = 1.0 / nrm;
      fn[3 * i] *= inv; fn[3 * i + 1] *= inv; fn[3 * i + 2] *= inv;
    }
    for (int c = 0; c < 3; c++) {
      Xdn[9 * i + 3 * c] = Xxi[2 * c];
      Xdn[9 * i + 3 * c + 1] = Xxi[2 * c + 1];
      Xdn[9 * i + 3 * c + 2] = fn[3 * i + c];
    }
  }

  /* drill strain rows at the nodes: TacsShellComputeDrillStrain (:649-693) with
     et = 0.5*(Ct[3] + u0x[3] - Ct[1] - u0x[1]) (TACSDirector.h:562-566) */
  for (int i = 0; i < n; i++) {
    double pt[2] = {-1.0 + (2.0 / (order - 1)) * (i % order), -1.0 + (2.0 / (order - 1)) * (i / order)};
    shape2d(order, pt, N, dN);
    double Xxi[6], T[9], Xdinv[9], XdinvT[9];
    for (int c = 0; c < 3; c++) { Xxi[2 * c] = Xdn[9 * i + 3 * c]; Xxi[2 * c + 1] = Xdn[9 * i + 3 * c + 1]; }
    shell_transform(transform, axis, Xxi, &fn[3 * i], T);
    inv3(&Xdn[9 * i], Xdinv);
    mm3(Xdinv, T, XdinvT);
    for (int k = 0; k < nd; k++) Bdr[i][k] = 0.0;
    for (int j = 0; j < n; j++) {
      double g0 = dN[2 * j] * XdinvT[0] + dN[2 * j + 1] * XdinvT[3];
      double g1 = dN[2 * j] * XdinvT[1] + dN[2 * j + 1] * XdinvT[4];
      for (int c = 0; c < 3; c++) Bdr[i][6 * j + c] = 0.5 * (T[3 * c + 1] * g0 - T[3 * c] * g1);
    }
    double t1[3] = {T[0], T[3], T[6]}, t2[3] = {T[1], T[4], T[7]}, t12[3];
    cross(t1, t2, t12);
    for (int e = 0; e < 3; e++) Bdr[i][6 * i + 3 + e] += -t12[e];
  }

  /* tying strain rows: TACSShellLinearModel::computeTyingStrain (Model.h:28-73);
     director d_j = q_j x fn_j (TACSDirector.h:244-267) => d(row)/dq_j = fn_j x d(row)/dd_j */
  int cnt[5];
  const int nty = ty_counts(order, cnt);
  for (int ty = 0; ty < nty; ty++) {
    int field;
    double pt[2];
    ty_point(order, ty, &field, pt);
    shape2d(order, pt, N, dN);
    double Xxi[6] = {0, 0, 0, 0, 0, 0}, n0[3] = {0, 0, 0};
    for (int j = 0; j < n; j++)
      for (int c = 0; c < 3; c++) {
        Xxi[2 * c] += dN[2 * j] * Xpts[3 * j + c];
        Xxi[2 * c + 1] += dN[2 * j + 1] * Xpts[3 * j + c];
        n0[c] += N[j] * fn[3 * j + c];
      }
    for (int j = 0; j < n; j++) {
      double du[3], dd[3] = {0, 0, 0};
      for (int c = 0; c < 3; c++) {
        if (field == 0) du[c] = dN[2 * j] * Xxi[2 * c];
        else if (field == 1) du[c] = dN[2 * j + 1] * Xxi[2 * c + 1];
        else if (field == 2) du[c] = 0.5 * (dN[2 * j] * Xxi[2 * c + 1] + dN[2 * j + 1] * Xxi[2 * c]);
        else if (field == 3) { du[c] = 0.5 * n0[c] * dN[2 * j + 1]; dd[c] = 0.5 * N[j] * Xxi[2 * c + 1]; }
        else { du[c] = 0.5 * n0[c] * dN[2 * j]; dd[c] = 0.5 * N[j] * Xxi[2 * c]; }
      }
      double dq[3];
      cross(&fn[3 * j], dd, dq);
      for (int c = 0; c < 3; c++) { Bty[ty][6 * j + c] = du[c]; Bty[ty][6 * j + 3 + c] = dq[c]; }
    }
  }

  /* dense 9x9 constitutive matrix */
  double Cm[81];
  for (int r = 0; r < 9; r++) {
    double e[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, s[9];
    e[r] = 1.0;
    shell_stress(Cs, e, s);
    for (int c = 0; c < 9; c++) Cm[9 * c + r] = s[c];
  }

  double Sab[81]; /* sum_q w det N_a N_b for the mass terms */
  for (int k = 0; k < n * n; k++) Sab[k] = 0.0;

  const int nq = order * order;
  const double *gp = (order == 2) ? GP2 : GP3, *gw = (order == 2) ? GW2 : GW3;
  for (int q = 0; q < nq; q++) {
    double pt[2] = {gp[q % order], gp[q / order]};
    double weight = gw[q % order] * gw[q / order];
    shape2d(order, pt, N, dN);
    double Xxi[6] = {0, 0, 0, 0, 0, 0}, n0[3] = {0, 0, 0}, nxi[6] = {0, 0, 0, 0, 0, 0};
    for (int j = 0; j < n; j++)
      for (int c = 0; c < 3; c++) {
        Xxi[2 * c] += dN[2 * j] * Xpts[3 * j + c];
        Xxi[2 * c + 1] += dN[2 * j + 1] * Xpts[3 * j + c];
        n0[c] += N[j] * fn[3 * j + c];
        nxi[2 * c] += dN[2 * j] * fn[3 * j + c];
        nxi[2 * c + 1] += dN[2 * j + 1] * fn[3 * j + c];
      }
    double T[9], Xd[9], Xdz[9], Xdinv[9], XdinvT[9], XdinvzT[9], tmp[9];
    shell_transform(transform, axis, Xxi, n0, T);
    for (int c = 0; c < 3; c++) {
      Xd[3 * c] = Xxi[2 * c]; Xd[3 * c + 1] = Xxi[2 * c + 1]; Xd[3 * c + 2] = n0[c];
      Xdz[3 * c] = nxi[2 * c]; Xdz[3 * c + 1] = nxi[2 * c + 1]; Xdz[3 * c + 2] = 0.0;
    }
    /* TacsShellComputeDispGrad, TACSShellUtilities.h:361-421 */
    double detXd = inv3(Xd, Xdinv) * weight;
    mm3(Xdinv, Xdz, tmp);
    for (int k = 0; k < 9; k++) tmp[k] *= -1.0;
    mm3(Xdinv, T, XdinvT);
    mm3(tmp, XdinvT, XdinvzT);

    double B[9][54];
    for (int r = 0; r < 9; r++) for (int k = 0; k < nd; k++) B[r][k] = 0.0;

    /* membrane + transverse shear rows from e0ty = XdinvT^T gty XdinvT
       (interpTyingStrain QuadBasis.h:651-672, mat3x3SymmTransformTranspose Algebra.h:1094,
       evalStrain Model.h:717-732) */
    double Nty[MAXTY];
    ty_interp(order, pt, Nty);
    static const int fc[5] = {0, 1, 0, 1, 0}, fd[5] = {0, 1, 1, 2, 2}; /* gty(c,d) of each field */
    static const int ea[5] = {0, 1, 0, 1, 0}, eb[5] = {0, 1, 1, 2, 2};   /* e0ty(a,b) feeding rows */
    static const int erow[5] = {0, 1, 2, 6, 7};
    static const double escale[5] = {1.0, 1.0, 2.0, 2.0, 2.0};
    int ty = 0;
    for (int f = 0; f < 5; f++) {
      int c = fc[f], d = fd[f];
      for (int k = 0; k < cnt[f]; k++, ty++) {
        for (int m = 0; m < 5; m++) {
          int a = ea[m], b = eb[m];
          double coef = (c == d) ? XdinvT[3 * c + a] * XdinvT[3 * c + b]
                                 : XdinvT[3 * c + a] * XdinvT[3 * d + b] + XdinvT[3 * d + a] * XdinvT[3 * c + b];
          coef *= escale[m] * Nty[ty];
          for (int kk = 0; kk < nd; kk++) B[erow[m]][kk] += coef * Bty[ty][kk];
        }
      }
    }

    /* bending rows from u1x = T^T (u1d XdinvT + u0d XdinvzT) */
    for (int j = 0; j < n; j++) {
      double hz[3], h[3];
      for (int b = 0; b < 3; b++) {
        hz[b] = dN[2 * j] * XdinvzT[b] + dN[2 * j + 1] * XdinvzT[3 + b];
        h[b] = dN[2 * j] * XdinvT[b] + dN[2 * j + 1] * XdinvT[3 + b] + N[j] * XdinvzT[6 + b];
      }
      double ru[3][3], rd[3][3];
      for (int c = 0; c < 3; c++) {
        ru[0][c] = T[3 * c] * hz[0];                          rd[0][c] = T[3 * c] * h[0];
        ru[1][c] = T[3 * c + 1] * hz[1];                      rd[1][c] = T[3 * c + 1] * h[1];
        ru[2][c] = T[3 * c] * hz[1] + T[3 * c + 1] * hz[0];   rd[2][c] = T[3 * c] * h[1] + T[3 * c + 1] * h[0];
      }
      for (int r = 0; r < 3; r++) {
        double dq[3];
        cross(&fn[3 * j], rd[r], dq);
        for (int c = 0; c < 3; c++) { B[3 + r][6 * j + c] += ru[r][c]; B[3 + r][6 * j + 3 + c] += dq[c]; }
      }
    }

    /* drill row: nodal drill strains interpolated with the nodal shape functions */
    for (int i = 0; i < n; i++)
      for (int k = 0; k < nd; k++) B[8][k] += N[i] * Bdr[i][k];

    /* residual: strain -> stress -> B^T s ; Jacobian: B^T C B */
    double CB[9][54];
    for (int r = 0; r < 9; r++)
      for (int k = 0; k < nd; k++) {
        double s = 0.0;
        for (int c = 0; c < 9; c++) s += Cm[9 * r + c] * B[c][k];
        CB[r][k] = s;
      }
    if (res) {
      double e[9], s[9];
      for (int r = 0; r < 9; r++) {
        e[r] = 0.0;
        for (int k = 0; k < nd; k++) e[r] += B[r][k] * vars[k];
      }
      shell_stress(Cs, e, s);
      for (int k = 0; k < nd; k++) {
        double v = 0.0;
        for (int r = 0; r < 9; r++) v += B[r][k] * s[r];
        res[k] += detXd * v;
      }
    }
    if (mat) {
      double sc = alpha * detXd;
      for (int i = 0; i < nd; i++)
        for (int j = 0; j < nd; j++) {
          double v = 0.0;
          for (int r = 0; r < 9; r++) v += B[r][i] * CB[r][j];
          mat[nd * i + j] += sc * v;
        }
    }
    for (int a = 0; a < n; a++)
      for (int b = 0; b < n; b++) Sab[n * a + b] += detXd * N[a] * N[b];
  }

  /* inertial terms (TACSShellElement.h:391-411, 580-617; addDirectorJacobian TACSDirector.h:368-488):
     generalised mass in (u,d) space [[m0 m1],[m1 m2]] N_a N_b, d_j = D_j q_j, D_j(c,e) = eps_{cef} fn_j[f] */
  double D[9][9];
  for (int j = 0; j < n; j++) {
    const double *t = &fn[3 * j];
    double Dj[9] = {0.0, t[2], -t[1], -t[2], 0.0, t[0], t[1], -t[0], 0.0};
    memcpy(D[j], Dj, sizeof(Dj));
  }
  for (int a = 0; a < n; a++)
    for (int b = 0; b < n; b++) {
      double S = Sab[n * a + b];
      double M[36];
      for (int k = 0; k < 36; k++) M[k] = 0.0;
      for (int c = 0; c < 3; c++) {
        M[6 * c + c] = S * moments[0];
        for (int e = 0; e < 3; e++) {
          M[6 * c + 3 + e] = S * moments[1] * D[b][3 * c + e];
          M[6 * (3 + e) + c] = S * moments[1] * D[a][3 * c + e];
        }
      }
      for (int e = 0; e < 3; e++)
        for (int f = 0; f < 3; f++) {
          double v = 0.0;
          for (int c = 0; c < 3; c++) v += D[a][3 * c + e] * D[b][3 * c + f];
          M[6 * (3 + e) + 3 + f] = S * moments[2] * v;
        }
      for (int r = 0; r < 6; r++)
        for (int c = 0; c < 6; c++) {
          if (mat) mat[nd * (6 * a + r) + 6 * b + c] += gamma * M[6 * r + c];
          if (res && ddvars) res[6 * a + r] += M[6 * r + c] * ddvars[6 * b + c];
        }
    }
}

/* ------------------------------------------------------------------ solid element */
/* 8-node (order 2) / 27-node (order 3) hexahedron, TACSLinearElasticity3D (linear strain).
   C21: upper triangle by rows (TACSSolidConstitutive.cpp:148-159); strain order
   (xx,yy,zz,yz,xz,xy) with engineering shears (TACSLinearElasticity.cpp:1186-1192). */
void oracle_solid_element(int order, const double *Xpts, const double *vars, const double *ddvars,
                          const double *C21, double rho, double alpha, double beta, double gamma,
                          double *res, double *mat) {
  (void)beta;
  const int n = order * order * order, nd = 3 * n;
  double Cm[36];
  {
    int k = 0;
    for (int i = 0; i < 6; i++)
      for (int j = i; j < 6; j++, k++) Cm[6 * i + j] = Cm[6 * j + i] = C21[k];
  }
  const double *gp = (order == 2) ? GP2 : GP3, *gw = (order == 2) ? GW2 : GW3;
  const int o2 = order * order;
  for (int q = 0; q < n; q++) {
    double pt[3] = {gp[q % order], gp[(q % o2) / order], gp[q / o2]};
    double weight = gw[q % order] * gw[(q % o2) / order] * gw[q / o2];
    double n1[3], d1[3], n2[3], d2[3], n3[3], d3[3];
    shape1d(order, pt[0], n1, d1);
    shape1d(order, pt[1], n2, d2);
    shape1d(order, pt[2], n3, d3);
    double N[MAXN], Nxi[3 * MAXN];
    for (int k = 0, a = 0; k < order; k++)
      for (int j = 0; j < order; j++)
        for (int i = 0; i < order; i++, a++) {
          N[a] = n1[i] * n2[j] * n3[k];
          Nxi[3 * a] = d1[i] * n2[j] * n3[k];
          Nxi[3 * a + 1] = n1[i] * d2[j] * n3[k];
          Nxi[3 * a + 2] = n1[i] * n2[j] * d3[k];
        }
    /* getFieldGradient, TACSElementBasis.cpp:266-326 */
    double Xd[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, J[9];
    for (int a = 0; a < n; a++)
      for (int c = 0; c < 3; c++)
        for (int k = 0; k < 3; k++) Xd[3 * c + k] += Nxi[3 * a + k] * Xpts[3 * a + c];
    double detXd = inv3(Xd, J) * weight;
    double G[3 * MAXN]; /* physical gradient of N_a */
    for (int a = 0; a < n; a++)
      for (int dir = 0; dir < 3; dir++)
        G[3 * a + dir] = Nxi[3 * a] * J[dir] + Nxi[3 * a + 1] * J[3 + dir] + Nxi[3 * a + 2] * J[6 + dir];

    /* B_a (6x3) */
    static double B[6][3 * MAXN], CB[6][3 * MAXN];
    for (int a = 0; a < n; a++) {
      double gx = G[3 * a], gy = G[3 * a + 1], gz = G[3 * a + 2];
      double Ba[6][3] = {{gx, 0, 0}, {0, gy, 0}, {0, 0, gz}, {0, gz, gy}, {gz, 0, gx}, {gy, gx, 0}};
      for (int r = 0; r < 6; r++)
        for (int c = 0; c < 3; c++) B[r][3 * a + c] = Ba[r][c];
    }
    for (int r = 0; r < 6; r++)
      for (int k = 0; k < nd; k++) {
        double s = 0.0;
        for (int c = 0; c < 6; c++) s += Cm[6 * r + c] * B[c][k];
        CB[r][k] = s;
      }
    if (res) {
      double e[6], s[6];
      for (int r = 0; r < 6; r++) {
        e[r] = 0.0;
        for (int k = 0; k < nd; k++) e[r] += B[r][k] * vars[k];
      }
      for (int r = 0; r < 6; r++) {
        s[r] = 0.0;
        for (int c = 0; c < 6; c++) s[r] += Cm[6 * r + c] * e[c];
      }
      for (int k = 0; k < nd; k++) {
        double v = 0.0;
        for (int r = 0; r < 6; r++) v += B[r][k] * s[r];
        res[k] += detXd * v;
      }
      if (ddvars)
        for (int a = 0; a < n; a++)
          for (int b = 0; b < n; b++)
            for (int c = 0; c < 3; c++) res[3 * a + c] += detXd * rho * N[a] * N[b] * ddvars[3 * b + c];
    }
    if (mat) {
      double sc = alpha * detXd;
      for (int i = 0; i < nd; i++)
        for (int j = 0; j < nd; j++) {
          double v = 0.0;
          for (int r = 0; r < 6; r++) v += B[r][i] * CB[r][j];
          mat[nd * i + j] += sc * v;
        }
      for (int a = 0; a < n; a++)
        for (int b = 0; b < n; b++)
          for (int c = 0; c < 3; c++) mat[nd * (3 * a + c) + 3 * b + c] += gamma * detXd * rho * N[a] * N[b];
    }
  }
}

/* ------------------------------------------------------------------ constitutive */
/* TACSMaterialProperties::evalTangentStiffness2D (.cpp:323-341), G = 0.5 E/(1+nu) (:31) */
/* TACSIsoShellConstitutive::evalTangentStiffness (.cpp:192-226), evalMassMoments (:120-129) */
void oracle_iso_shell_stiffness(double rho, double E, double nu, double t, double tOffset, double kcorr,
                                double kdrill, double *Cs, double *moments) {
  double *A = Cs, *B = Cs + 6, *D = Cs + 12, *As = Cs + 18;
  double G = 0.5 * E / (1.0 + nu);
  double Dm = E / (1.0 - nu * nu);
  A[0] = Dm; A[1] = nu * Dm; A[2] = 0.0; A[3] = Dm; A[4] = 0.0; A[5] = G;
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
  Cs[21] = 0.5 * kdrill * (As[0] + As[2]);
  moments[0] = rho * t;
  moments[1] = -rho * t * t * tOffset;
  moments[2] = rho * t * t * t * (tOffset * tOffset + 1.0 / 12.0);
}

/* TACSCompositeShellConstitutive::evalTangentStiffness (.cpp:249-301), evalMassMoments (:70-100),
   TACSOrthotropicPly ctor + calculateQbar/Abar (TACSMaterialProperties.cpp:523-545, 724-773).
   ply[7*k..] = rho, E1, E2, nu12, G12, G13, G23 */
void oracle_composite_shell_stiffness(int nplies, const double *ply, const double *thick, const double *angle,
                                      double kcorr, double tOffset, double kdrill, double *Cs, double *moments) {
  double *A = Cs, *B = Cs + 6, *D = Cs + 12, *As = Cs + 18;
  for (int k = 0; k < 6; k++) A[k] = B[k] = D[k] = 0.0;
  for (int k = 0; k < 3; k++) As[k] = 0.0;
  moments[0] = moments[1] = moments[2] = 0.0;
  double t = 0.0;
  for (int i = 0; i < nplies; i++) t += thick[i];
  double t0 = -(0.5 + tOffset) * t;
  for (int k = 0; k < nplies; k++) {
    const double *p = ply + 7 * k;
    double rho = p[0], E1 = p[1], E2 = p[2], nu12 = p[3], G12 = p[4], G13 = p[5], G23 = p[6];
    double nu21 = nu12 * E2 / E1;
    double Q11 = E1 / (1.0 - nu12 * nu21), Q22 = E2 / (1.0 - nu12 * nu21), Q12 = nu12 * E2 / (1.0 - nu12 * nu21);
    double Q44 = G23, Q55 = G13, Q66 = G12;
    double C12 = (Q11 + Q22 - 4.0 * Q66), C16 = (Q11 - Q12 - 2.0 * Q66), C26 = (Q12 - Q22 + 2.0 * Q66);
    double C66 = (Q11 + Q22 - 2.0 * Q12 - 2.0 * Q66);
    double cos1 = cos(angle[k]), sin1 = sin(angle[k]);
    double cos2 = cos1 * cos1, sin2 = sin1 * sin1, cos4 = cos2 * cos2, sin4 = sin2 * sin2;
    double Qbar[6], Abar[3];
    Qbar[0] = Q11 * cos4 + 2.0 * (Q12 + 2.0 * Q66) * sin2 * cos2 + Q22 * sin4;
    Qbar[1] = C12 * sin2 * cos2 + Q12 * (sin4 + cos4);
    Qbar[2] = C16 * sin1 * cos2 * cos1 + C26 * sin2 * sin1 * cos1;
    Qbar[3] = Q11 * sin4 + 2.0 * (Q12 + 2.0 * Q66) * sin2 * cos2 + Q22 * cos4;
    Qbar[4] = C16 * sin2 * sin1 * cos1 + C26 * sin1 * cos2 * cos1;
    Qbar[5] = C66 * sin2 * cos2 + Q66 * (sin4 + cos4);
    Abar[0] = cos2 * Q44 + sin2 * Q55;
    Abar[1] = cos1 * sin1 * (Q55 - Q44);
    Abar[2] = sin2 * Q44 + cos2 * Q55;
    double t1 = t0 + thick[k];
    double a = (t1 - t0), b = 0.5 * (t1 * t1 - t0 * t0), d = 1.0 / 3.0 * (t1 * t1 * t1 - t0 * t0 * t0);
    for (int i = 0; i < 6; i++) { A[i] += a * Qbar[i]; B[i] += b * Qbar[i]; D[i] += d * Qbar[i]; }
    for (int i = 0; i < 3; i++) As[i] += kcorr * a * Abar[i];
    moments[0] += a * rho; moments[1] += b * rho; moments[2] += d * rho;
    t0 = t1;
  }
  Cs[21] = 0.5 * kdrill * (As[0] + As[2]);
}

/* TACSSolidConstitutive::evalTangentStiffness (.cpp:166-178) over
   TACSMaterialProperties::evalTangentStiffness3D isotropic branch (.cpp:270-292) */
void oracle_solid_stiffness(double rho, double E, double nu, double t, double *C21, double *density) {
  double G = 0.5 * E / (1.0 + nu);
  double D = E / ((1.0 + nu) * (1.0 - 2.0 * nu));
  for (int i = 0; i < 21; i++) C21[i] = 0.0;
  C21[0] = (1.0 - nu) * D; C21[1] = nu * D; C21[2] = nu * D;
  C21[6] = (1.0 - nu) * D; C21[7] = nu * D;
  C21[11] = (1.0 - nu) * D;
  C21[15] = G; C21[18] = G; C21[20] = G;
  for (int i = 0; i < 21; i++) C21[i] *= t;
  *density = t * rho;
}

/* ------------------------------------------------------------------ integer pipeline */
static int cmp_int(const void *a, const void *b) {
  int x = *(const int *)a, y = *(const int *)b;
  return (x > y) - (x < y);
}
/* TacsUniqueSort (TacsUtilities.cpp:76-99): ascending, negatives dropped, duplicates removed */
int oracle_unique_sort(int len, int *a) {
  qsort(a, len, sizeof(int), cmp_int);
  int i = 0;
  while (i < len && a[i] < 0) i++;
  int n = 0;
  for (; i < len; i++)
    if (n == 0 || a[n - 1] != a[i]) a[n++] = a[i];
  return n;
}

/* TACSCreator::partitionMesh tail (TACSCreator.cpp:1141-1205): first-touch node renumbering
   given the element partition; returns owned_nodes / owned_elements per part. */
void oracle_first_touch_numbering(int num_nodes, int num_elements, const int *ptr, const int *conn,
                                  const int *partition, int nparts, int *new_nodes, int *owned_nodes,
                                  int *owned_elements) {
  for (int k = 0; k < num_nodes; k++) new_nodes[k] = 0;
  for (int k = 0; k < nparts; k++) owned_nodes[k] = owned_elements[k] = 0;
  for (int j = 0; j < num_elements; j++) {
    int owner = partition[j];
    owned_elements[owner]++;
    for (int i = ptr[j]; i < ptr[j + 1]; i++) {
      int node = conn[i];
      if (node >= 0 && !new_nodes[node]) { new_nodes[node] = 1; owned_nodes[owner]++; }
    }
  }
  for (int k = 0; k < num_nodes; k++) new_nodes[k] = -1;
  int *off = (int *)malloc(nparts * sizeof(int));
  off[0] = 0;
  for (int k = 1; k < nparts; k++) off[k] = off[k - 1] + owned_nodes[k - 1];
  for (int j = 0; j < num_elements; j++) {
    int owner = partition[j];
    for (int i = ptr[j]; i < ptr[j + 1]; i++) {
      int node = conn[i];
      if (node >= 0 && new_nodes[node] < 0) new_nodes[node] = off[owner]++;
    }
  }
  free(off);
}

/* Node->node block sparsity of one rank with no external nodes
   (TACSAssembler::computeLocalNodeToNodeCSR :1899-2053 + TacsSortAndUniquifyCSR, diagonal kept).
   conn is in the numbering of the rows. rowp has nnodes+1 entries; returns nnzb; cols may be NULL
   for the counting pass. */
int oracle_node_to_node_csr(int nnodes, int nelems, const int *ptr, const int *conn, int *rowp, int *cols) {
  int *cnt = (int *)calloc(nnodes + 1, sizeof(int));
  for (int e = 0; e < nelems; e++)
    for (int i = ptr[e]; i < ptr[e + 1]; i++) cnt[conn[i] + 1] += ptr[e + 1] - ptr[e];
  for (int i = 0; i < nnodes; i++) cnt[i + 1] += cnt[i];
  int *buf = (int *)malloc((cnt[nnodes] ? cnt[nnodes] : 1) * sizeof(int));
  int *pos = (int *)malloc((nnodes + 1) * sizeof(int));
  memcpy(pos, cnt, (nnodes + 1) * sizeof(int));
  for (int e = 0; e < nelems; e++)
    for (int i = ptr[e]; i < ptr[e + 1]; i++)
      for (int j = ptr[e]; j < ptr[e + 1]; j++) buf[pos[conn[i]]++] = conn[j];
  int nnz = 0;
  rowp[0] = 0;
  for (int r = 0; r < nnodes; r++) {
    int len = oracle_unique_sort(cnt[r + 1] - cnt[r], buf + cnt[r]);
    if (cols) memcpy(cols + nnz, buf + cnt[r], len * sizeof(int));
    nnz += len;
    rowp[r + 1] = nnz;
  }
  free(cnt); free(buf); free(pos);
  return nnz;
}

/* BCSRMat::addRowValues semantics for one element (BCSRMat.cpp:1800-1848 via
   TACSMatDistribute::addValues :712-846, single rank): binary search of each column. */
int oracle_bcsr_add_element(int bs, const int *rowp, const int *cols, double *A, int nn, const int *nodes,
                            const double *mat) {
  const int nv = bs * nn, b2 = bs * bs;
  for (int i = 0; i < nn; i++) {
    int row = nodes[i];
    for (int j = 0; j < nn; j++) {
      int col = nodes[j];
      int lo = rowp[row], hi = rowp[row + 1] - 1, k = -1;
      while (lo <= hi) {
        int mid = (lo + hi) / 2;
        if (cols[mid] == col) { k = mid; break; }
        if (cols[mid] < col) lo = mid + 1; else hi = mid - 1;
      }
      if (k < 0) return 1;
      double *a = A + (size_t)b2 * k;
      for (int ii = 0; ii < bs; ii++)
        for (int jj = 0; jj < bs; jj++) a[bs * ii + jj] += mat[nv * (bs * i + ii) + bs * j + jj];
    }
  }
  return 0;
}

/* TACSParallelMat::applyBCs -> BCSRMat::zeroRow (BCSRMat.cpp:2027-2053): zero the flagged
   rows of every block of the block row, unit diagonal. bc_vars is the TACSBcMap bit mask. */
void oracle_bcsr_apply_bcs(int bs, const int *rowp, const int *cols, double *A, int nbcs, const int *bc_nodes,
                           const int *bc_vars) {
  const int b2 = bs * bs;
  for (int i = 0; i < nbcs; i++) {
    int row = bc_nodes[i];
    for (int k = rowp[row]; k < rowp[row + 1]; k++) {
      double *a = A + (size_t)b2 * k;
      for (int ii = 0; ii < bs; ii++)
        if (bc_vars[i] & (1 << ii)) {
          for (int jj = 0; jj < bs; jj++) a[bs * ii + jj] = 0.0;
          if (cols[k] == row) a[bs * ii + ii] = 1.0;
        }
    }
  }
}

/* TACSBVec::applyBCs (TACSBVec.cpp:546-596): flagged dofs <- u - lambda*value when a state vector
   u is given (residual), else <- 0. bc_vals is [nbcs][bs] (TACSBcMap, KSM.cpp:110-131) or NULL (zeros). */
void oracle_vec_apply_bcs(int bs, double *x, int nbcs, const int *bc_nodes, const int *bc_vars,
                          const double *bc_vals, const double *u, double lambda) {
  for (int i = 0; i < nbcs; i++)
    for (int ii = 0; ii < bs; ii++)
      if (bc_vars[i] & (1 << ii)) {
        int k = bs * bc_nodes[i] + ii;
        x[k] = u ? u[k] - lambda * (bc_vals ? bc_vals[bs * i + ii] : 0.0) : 0.0;
      }
}

/* BCSRMatVecMult6 / BCSRMatVecMult3 (BCSRMatMult6.cpp:82-121, BCSRMatMult3.cpp:27-48): blocks in
   ascending column order; each row's bs-term product summed left to right, then added. */
void oracle_bcsr_mult(int bs, int nrows, const int *rowp, const int *cols, const double *A, const double *x,
                      double *y) {
  const int b2 = bs * bs;
  for (int i = 0; i < nrows; i++) {
    double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int k = rowp[i]; k < rowp[i + 1]; k++) {
      const double *a = A + (size_t)b2 * k, *xj = x + bs * cols[k];
      for (int ii = 0; ii < bs; ii++) {
        double s = a[bs * ii] * xj[0];
        for (int jj = 1; jj < bs; jj++) s += a[bs * ii + jj] * xj[jj];
        acc[ii] += s;
      }
    }
    for (int ii = 0; ii < bs; ii++) y[bs * i + ii] = acc[ii];
  }
}

/*
 * Whole single-rank assembly: TACSAssembler::assembleJacobian (TACSAssembler.cpp:4291-4406).
 * kind: 1 Quad4 shell, 2 Quad9 shell, 3 hex8, 4 hex27. conn in matrix-row numbering.
 * elem_data: per distinct element descriptor `elem_desc[e]`, 32 doubles:
 *   shell: [0..21] Cs, [22..24] moments, [25] transform kind, [26..28] axis
 *   solid: [0..20] C21, [21] density
 * res (bs*nnodes) and A (bs^2*nnzb) are overwritten; BCs applied last (res <- u - lambda*value on
 * constrained dofs; matrix rows zeroed with unit diagonal).
 */
int oracle_assemble_jacobian(int kind, int nnodes, int nelems, const int *conn, const int *elem_desc,
                             const double *elem_data, const double *Xpts, const double *vars,
                             const double *ddvars, double alpha, double beta, double gamma, const int *rowp,
                             const int *cols, int nbcs, const int *bc_nodes, const int *bc_vars,
                             const double *bc_vals, double lambda, double *res, double *A) {
  const int order = (kind == 1 || kind == 3) ? 2 : 3;
  const int shell = (kind <= 2);
  const int nn = shell ? order * order : order * order * order;
  const int bs = shell ? 6 : 3, nv = bs * nn;
  double *ex = (double *)malloc(3 * nn * sizeof(double));
  double *ev = (double *)malloc(nv * sizeof(double)), *ea = (double *)malloc(nv * sizeof(double));
  double *er = (double *)malloc(nv * sizeof(double)), *em = (double *)malloc((size_t)nv * nv * sizeof(double));
  if (res) memset(res, 0, (size_t)bs * nnodes * sizeof(double));
  if (A) memset(A, 0, (size_t)bs * bs * rowp[nnodes] * sizeof(double));
  int fail = 0;
  for (int e = 0; e < nelems && !fail; e++) {
    const int *nodes = conn + (size_t)nn * e;
    for (int i = 0; i < nn; i++) {
      for (int c = 0; c < 3; c++) ex[3 * i + c] = Xpts[3 * (size_t)nodes[i] + c];
      for (int c = 0; c < bs; c++) {
        ev[bs * i + c] = vars ? vars[bs * (size_t)nodes[i] + c] : 0.0;
        ea[bs * i + c] = ddvars ? ddvars[bs * (size_t)nodes[i] + c] : 0.0;
      }
    }
    memset(er, 0, nv * sizeof(double));
    memset(em, 0, (size_t)nv * nv * sizeof(double));
    const double *d = elem_data + 32 * (size_t)elem_desc[e];
    if (shell)
      oracle_shell_element(order, ex, ev, ea, (int)d[25], d + 26, d, d + 22, alpha, beta, gamma, er, A ? em : NULL);
    else
      oracle_solid_element(order, ex, ev, ea, d, d[21], alpha, beta, gamma, er, A ? em : NULL);
    if (res)
      for (int i = 0; i < nn; i++)
        for (int c = 0; c < bs; c++) res[bs * (size_t)nodes[i] + c] += er[bs * i + c];
    if (A) fail = oracle_bcsr_add_element(bs, rowp, cols, A, nn, nodes, em);
  }
  if (res) {
    /* with no state vector the reference's varsVec is all zeros: res <- 0 - lambda*value */
    if (vars) oracle_vec_apply_bcs(bs, res, nbcs, bc_nodes, bc_vars, bc_vals, vars, lambda);
    else {
      double *z = (double *)calloc((size_t)bs * nnodes, sizeof(double));
      oracle_vec_apply_bcs(bs, res, nbcs, bc_nodes, bc_vars, bc_vals, z, lambda);
      free(z);
    }
  }
  if (A) oracle_bcsr_apply_bcs(bs, rowp, cols, A, nbcs, bc_nodes, bc_vars);
  free(ex); free(ev); free(ea); free(er); free(em);
  return fail;
}
