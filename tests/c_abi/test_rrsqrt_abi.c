/* The C ABI of liboak_b200.so exercised without Python: the known-answer case of the reference's own
 * test/test_rrsqrt.F90:283-322 (m = 5, n = 10, N = 12, Ef = sin(3 i^2), H = reshape(1..50), y = 1..5, R = 2 I):
 *   global scheme   xa = xf + Pf H' (H Pf H' + R)^-1 (y - H xf)         test_rrsqrt.F90:57-74, tol 1e-8
 *   local scheme with one zone holding every point and no cut-off = the global result   test_rrsqrt.F90:142-159
 * The check is computed here with a small Gaussian elimination (no LAPACK, nothing from oracle/).
 *   gcc -O1 -I include tests/c_abi/test_rrsqrt_abi.c -L oak_b200 -loak_b200 -Wl,-rpath,$PWD/oak_b200 -lm */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "oak_b200.h"

#define M 5
#define NN 10
#define NE 12

static void solve(int n, double *A, double *b) { /* A x = b, partial pivoting, A row-major n x n, b overwritten */
  for (int k = 0; k < n; k++) {
    int p = k;
    for (int i = k + 1; i < n; i++) if (fabs(A[i * n + k]) > fabs(A[p * n + k])) p = i;
    for (int j = 0; j < n; j++) { double t = A[k * n + j]; A[k * n + j] = A[p * n + j]; A[p * n + j] = t; }
    { double t = b[k]; b[k] = b[p]; b[p] = t; }
    for (int i = k + 1; i < n; i++) {
      const double f = A[i * n + k] / A[k * n + k];
      for (int j = k; j < n; j++) A[i * n + j] -= f * A[k * n + j];
      b[i] -= f * b[k];
    }
  }
  for (int k = n - 1; k >= 0; k--) {
    for (int j = k + 1; j < n; j++) b[k] -= A[k * n + j] * b[j];
    b[k] /= A[k * n + k];
  }
}

#define CHECK(call)                                                                  \
  do {                                                                               \
    int rc_ = (call);                                                                \
    if (rc_ != 0) { fprintf(stderr, "%s -> %d: %s\n", #call, rc_, oakb200_last_error()); return 2; } \
  } while (0)

int main(void) {
  static double Ef[NN * NE], H[M * NN], y[M], xf[NN], Sf[NN * NE], HSf[M * NE], Hxf[M], var[M];
  static double xa[NN], Sa[NN * NE], xa2[NN], Sa2[NN * NE];
  for (int i = 0; i < NN * NE; i++) { const double q = i + 1.; Ef[i] = sin(3. * q * q); }   /* column-major n x N */
  for (int i = 0; i < M * NN; i++) H[i] = i + 1.;                                            /* column-major m x n */
  for (int i = 0; i < M; i++) { y[i] = i + 1.; var[i] = 2.; }
  for (int i = 0; i < NN; i++) {
    xf[i] = 0.;
    for (int k = 0; k < NE; k++) xf[i] += Ef[i + NN * k];
    xf[i] /= NE;
    for (int k = 0; k < NE; k++) Sf[i + NN * k] = (Ef[i + NN * k] - xf[i]) / sqrt(NE - 1.);
  }
  for (int l = 0; l < M; l++) {
    Hxf[l] = 0.;
    for (int i = 0; i < NN; i++) Hxf[l] += H[l + M * i] * xf[i];
    for (int k = 0; k < NE; k++) {
      HSf[l + M * k] = 0.;
      for (int i = 0; i < NN; i++) HSf[l + M * k] += H[l + M * i] * Sf[i + NN * k];
    }
  }
  /* known answer: xa = xf + Sf HSf' (HSf HSf' + R)^-1 (y - Hxf) */
  double A[M * M], b[M], xcheck[NN];
  for (int p = 0; p < M; p++) {
    b[p] = y[p] - Hxf[p];
    for (int q = 0; q < M; q++) {
      A[p * M + q] = (p == q) ? var[p] : 0.;
      for (int k = 0; k < NE; k++) A[p * M + q] += HSf[p + M * k] * HSf[q + M * k];
    }
  }
  solve(M, A, b);
  for (int i = 0; i < NN; i++) {
    xcheck[i] = xf[i];
    for (int k = 0; k < NE; k++) {
      double t = 0.;
      for (int p = 0; p < M; p++) t += HSf[p + M * k] * b[p];
      xcheck[i] += Sf[i + NN * k] * t;
    }
  }

  oakb200_handle *h = NULL;
  oakb200_stats st;
  CHECK(oakb200_create(0, &h));
  /* global scheme */
  CHECK(oakb200_global_analysis(h, NN, NE, M, xf, Hxf, y, Sf, NN, HSf, M, var, NULL, xa, Sa, NN, NULL, &st));
  /* local scheme: one zone with every row, weight function "none" (every observation, weight 1) */
  const int32_t zoneSize[1] = {NN};
  const double zx[1] = {0.}, zy[1] = {0.}, corr[1] = {1.}, maxl[1] = {1e30};
  double ox[M], oy[M];
  for (int l = 0; l < M; l++) { ox[l] = l / (M - 1.); oy[l] = 0.; }
  CHECK(oakb200_set_zones(h, 1, zoneSize, zx, zy, NULL, NULL, corr, maxl, OAKB200_LOC_HORIZONTAL,
                          OAKB200_METRIC_CARTESIAN, OAKB200_WEIGHT_UNIFORM));
  CHECK(oakb200_set_observations(h, M, ox, oy, NULL, NULL));
  CHECK(oakb200_local_analysis(h, NN, NE, M, xf, Hxf, y, Sf, NN, HSf, M, var, NULL, xa2, Sa2, NN, NULL, &st));
  double e1 = 0., e2 = 0., e3 = 0., colsum = 0.;
  for (int i = 0; i < NN; i++) {
    e1 = fmax(e1, fabs(xa[i] - xcheck[i]));
    e2 = fmax(e2, fabs(xa2[i] - xcheck[i]));
    double s = 0.;
    for (int k = 0; k < NE; k++) { e3 = fmax(e3, fabs(Sa[i + NN * k] - Sa2[i + NN * k])); s += Sa[i + NN * k]; }
    colsum = fmax(colsum, fabs(s));
  }
  /* an argument error is a status, not a crash */
  const int bad = oakb200_local_analysis(h, NN + 1, NE, M, xf, Hxf, y, Sf, NN, HSf, M, var, NULL, xa2, Sa2, NN, NULL, &st);
  CHECK(oakb200_destroy(h));
  printf("global |xa - check| %.3e  local(1 zone) |xa - check| %.3e  |Sa_global - Sa_local| %.3e  |sum_k Sa| %.3e  bad-arg status %d\n",
         e1, e2, e3, colsum, bad);
  if (!(e1 < 1e-8 && e2 < 1e-8 && e3 < 1e-9 && colsum < 1e-12 && bad != 0)) { printf("FAILED\n"); return 1; }
  printf("OK\n");
  return 0;
}
