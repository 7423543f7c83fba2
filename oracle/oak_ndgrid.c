/* oak_ndgrid.c — CPU oracle, part 2: OAK's n-dimensional grid interpolation, the arithmetic behind the
 * observation operator (genObservationOper, assimilation.F90:2471-2656 -> cinterp, ndgrid.F90:1183-1257).
 *
 * TEST INFRASTRUCTURE ONLY (see oak_oracle.c).  A restatement, routine by routine, of
 *   split               ndgrid.F90:357-435    the n! 2^(n-1) simplices of a cell, all sharing the cell centre
 *   interp_tetrahedron  ndgrid.F90:464-629    barycentric coordinates in one simplex (inverse + determinant,
 *                                             tol 1e-8; degenerate simplices through the SVD, :499-627)
 *   interp_cube         ndgrid.F90:636-665    first simplex that contains the point
 *   InCube              ndgrid_inc.F90:304-367
 *   databox tree        ndgrid_inc.F90:534-608,:760-1001  lazily split bounding-box tree, search order of the sub-boxes
 *   getCoord0           ndgrid.F90:1128-1157
 *   cinterp             ndgrid.F90:1183-1257  corner indices (1-based), 2^n coefficients, nbp
 * The reference inverts with LAPACK dgetrf/dgetri (matoper_inc.F90:569-602) and decomposes with dgesvd (:695-714), both
 * unpinned (-llapack); here: Gaussian elimination with dgetrf's pivoting rule and a solve, and dgesvd from the
 * scipy-bundled OpenBLAS.  Pinned on the reference's own test, test/test_ndgrid.F90:11-32 (analytic linear fields on 1- to
 * 5-D grids including singleton dimensions, tol 1e-6): tests/test_ndgrid_oracle.py.
 *
 * Coordinates are stored for every grid point and every dimension (coord[d*total + linear index]): the explicit-
 * coordinate variant of the reference with `dependence` all ones; a separable or regular grid is the same thing filled
 * by broadcasting.
 */
#include <dlfcn.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define NDMAX 5
#define TWONMAX 32

typedef void (*dgesvd_fn)(const char *, const char *, const int *, const int *, double *, const int *, double *,
                          double *, const int *, double *, const int *, double *, const int *, int *);
static dgesvd_fn p_dgesvd;

int oracle_ndgrid_init(const char *blas_path) {
  if (p_dgesvd) return 0;
  void *h = dlopen(blas_path, RTLD_NOW | RTLD_LOCAL);
  if (!h) return -1;
  p_dgesvd = (dgesvd_fn)dlsym(h, "scipy_dgesvd_");
  return p_dgesvd ? 0 : -2;
}

static int factorial(int n) {  /* ndgrid.F90:443-455 */
  int f = 1;
  for (int i = 2; i <= n; i++) f *= i;
  return f;
}

/* tet[(l*(n+1) + j)*twon + i] = tetrahedron(i+1, j+1, l+1).  split writes columns 0..subn of simplices l0.. */
static void split(int n, int subn, const int *fixeddim, const int *selection, double *tet, int l0) {
  const int twon = 1 << n;
  if (subn == 1) { /* a straight line: its two end points (:373-391) */
    for (int j = 0; j < 2; j++)
      for (int i = 0; i < twon; i++) tet[((size_t)l0 * (n + 1) + j) * twon + i] = 0.;
    int j = 0;
    for (int i = 0; i < twon; i++)
      if (selection[i]) { tet[((size_t)l0 * (n + 1) + j) * twon + i] = 1.; j++; }
    return;
  }
  int cnt = 0;
  for (int i = 0; i < twon; i++) cnt += selection[i] != 0;
  const int nbth = factorial(subn) * (1 << (subn - 1)), subnbth = factorial(subn - 1) * (1 << (subn - 2));
  for (int l = 0; l < nbth; l++) /* the middle point is the last vertex of every simplex (:399-412) */
    for (int i = 0; i < twon; i++) tet[((size_t)(l0 + l) * (n + 1) + subn) * twon + i] = selection[i] ? 1. / cnt : 0.;
  int m = 0;
  for (int i = 0; i < n; i++) {
    if (fixeddim[i]) continue;
    int fd[NDMAX];
    memcpy(fd, fixeddim, sizeof(int) * n);
    fd[i] = 1;
    for (int j = 0; j < 2; j++) { /* the two faces across dimension i (:418-434) */
      int sub[TWONMAX];
      for (int k = 0; k < twon; k++) sub[k] = selection[k] && (((k >> i) & 1) == j);
      split(n, subn - 1, fd, sub, tet, l0 + m);
      m += subnbth;
    }
  }
}

int oracle_ndgrid_nsimplex(int n) { return factorial(n) * (1 << (n - 1)); }

/* tetrahedron(2^n, n+1, nbth) of init_basegrid / initgrid_nd (ndgrid.F90:888,:978) */
void oracle_ndgrid_tetrahedra(int n, double *tet) {
  const int twon = 1 << n, nbth = oracle_ndgrid_nsimplex(n);
  memset(tet, 0, sizeof(double) * twon * (n + 1) * nbth);
  int fixed[NDMAX] = {0, 0, 0, 0, 0}, sel[TWONMAX];
  for (int i = 0; i < twon; i++) sel[i] = 1;
  split(n, n, fixed, sel, tet, 0);
}

/* inv(M, det) d with dgetrf's pivoting (largest modulus, first on ties); M is k x k column-major, destroyed */
static double lu_solve(int k, double *M, const double *d, double *c) {
  int piv[NDMAX + 1];
  double det = 1.;
  for (int j = 0; j < k; j++) {
    int p = j;
    for (int i = j + 1; i < k; i++)
      if (fabs(M[i + k * j]) > fabs(M[p + k * j])) p = i;
    piv[j] = p;
    if (p != j)
      for (int q = 0; q < k; q++) { const double t = M[j + k * q]; M[j + k * q] = M[p + k * q]; M[p + k * q] = t; }
    det = (p != j) ? -det * M[j + k * j] : det * M[j + k * j];
    if (M[j + k * j] != 0.)
      for (int i = j + 1; i < k; i++) {
        M[i + k * j] /= M[j + k * j];
        for (int q = j + 1; q < k; q++) M[i + k * q] -= M[i + k * j] * M[j + k * q];
      }
  }
  for (int i = 0; i < k; i++) c[i] = d[i];
  for (int j = 0; j < k; j++)
    if (piv[j] != j) { const double t = c[j]; c[j] = c[piv[j]]; c[piv[j]] = t; }
  for (int j = 0; j < k; j++)
    for (int i = j + 1; i < k; i++) c[i] -= M[i + k * j] * c[j];
  for (int j = k - 1; j >= 0; j--) {
    c[j] /= M[j + k * j];
    for (int i = 0; i < j; i++) c[i] -= M[i + k * j] * c[j];
  }
  return det;
}

/* interp_tetrahedron (ndgrid.F90:464-629): X is n x (n+1) column-major (vertex j at X + n*j) */
static void interp_tetrahedron(int n, const double *X, const double *xi, int *out, double *coeff) {
  const double tol = 1e-8;
  const int k = n + 1;
  double M[(NDMAX + 1) * (NDMAX + 1)], Mc[(NDMAX + 1) * (NDMAX + 1)], d[NDMAX + 1], c[NDMAX + 1], xc[NDMAX];
  for (int i = 0; i < n; i++) { /* relative to the average of the vertices (:509-512) */
    double s = 0.;
    for (int j = 0; j < k; j++) s += X[i + n * j];
    xc[i] = s / k;
  }
  for (int j = 0; j < k; j++) {
    M[0 + k * j] = 1.;
    for (int i = 0; i < n; i++) M[1 + i + k * j] = X[i + n * j] - xc[i];
  }
  d[0] = 1.;
  for (int i = 0; i < n; i++) d[1 + i] = xi[i] - xc[i];
  memcpy(Mc, M, sizeof(double) * k * k);
  const double determ = lu_solve(k, Mc, d, c);
  if (fabs(determ) > tol) { /* :516-520 */
    int in = 1;
    for (int j = 0; j < k; j++) in = in && (0. - tol <= c[j] && c[j] <= 1. + tol);
    *out = !in;
    for (int j = 0; j < k; j++) coeff[j] = c[j];
    return;
  }
  /* degenerate simplex (:527-627): drop the redundant constraints through the SVD, force coefficients to zero one
   * combination after the other, keep the admissible solution with the smallest |det| */
  double A[(NDMAX + 1) * (NDMAX + 1)], U[(NDMAX + 1) * (NDMAX + 1)], VT[(NDMAX + 1) * (NDMAX + 1)], S[NDMAX + 1], work[256];
  memcpy(A, M, sizeof(double) * k * k);
  int lwork = 256, info = 0;
  p_dgesvd("A", "A", &k, &k, A, &k, S, U, &k, VT, &k, work, &lwork, &info);
  int nzidx[NDMAX + 1], nnz = 0;
  for (int i = 0; i < k; i++)
    if (S[i] > tol) nzidx[nnz++] = i;
  const int nuncon = k - nnz;
  for (int j = 0; j < k; j++) coeff[j] = 0.;
  int cinit = 0;
  double best = 0.;
  *out = 1;
  long ncomb = 1;
  for (int j = 0; j < nuncon; j++) ncomb *= k;
  for (long it = 0; it < ncomb; it++) {
    int ind[NDMAX + 1];
    long tmp = it;
    for (int j = 0; j < nuncon; j++) { ind[j] = (int)(tmp % k); tmp /= k; }
    double M2[(NDMAX + 1) * (NDMAX + 1)], d2[NDMAX + 1], testc[NDMAX + 1];
    memset(M2, 0, sizeof M2);
    for (int j = 0; j < nuncon; j++) M2[j + k * ind[j]] = 1.;
    /* `any(sum(M2(1:nuncon,:),2) == 0)` (:575) can never hold (every row has its one); distinctness of the forced
     * coefficients is enforced by the determinant test below, as in the reference */
    for (int r = 0; r < nnz; r++)
      for (int q = 0; q < k; q++) M2[nuncon + r + k * q] = S[nzidx[r]] * VT[nzidx[r] + k * q]; /* diag(S) V' */
    for (int j = 0; j < k; j++) d2[j] = 0.;
    for (int r = 0; r < nnz; r++) {
      double s = 0.;
      for (int q = 0; q < k; q++) s += U[q + k * nzidx[r]] * d[q];
      d2[nuncon + r] = s;
    }
    double M2c[(NDMAX + 1) * (NDMAX + 1)];
    memcpy(M2c, M2, sizeof M2c);
    const double detM2 = lu_solve(k, M2c, d2, testc);
    if (fabs(detM2) < tol) continue;
    int in = 1;
    for (int j = 0; j < k; j++) in = in && (0 - tol <= testc[j] && testc[j] <= 1 + tol);
    if (!in) continue;
    double err = 0.;
    for (int i = 0; i < k; i++) {
      double s = 0.;
      for (int q = 0; q < k; q++) s += M[i + k * q] * testc[q];
      err = fmax(err, fabs(s - d[i]));
    }
    if (err < tol && (fabs(detM2) < best || !cinit)) {
      for (int j = 0; j < k; j++) coeff[j] = testc[j];
      cinit = 1;
      best = fabs(detM2);
      *out = 0;
    }
  }
}

/* interp_cube (ndgrid.F90:636-665): x is n x 2^n column-major; c (2^n) only written when a simplex is found */
static void interp_cube(int n, const double *tet, const double *x, const double *xi, int *out, double *c) {
  const int twon = 1 << n, nbth = oracle_ndgrid_nsimplex(n);
  *out = 1;
  for (int l = 0; l < nbth; l++) {
    const double *T = tet + (size_t)l * (n + 1) * twon;
    double X[NDMAX * (NDMAX + 1)], coeff[NDMAX + 1];
    for (int j = 0; j <= n; j++)
      for (int i = 0; i < n; i++) { /* matmul(x, tetrahedron(:,:,l)) */
        double s = 0.;
        for (int q = 0; q < twon; q++) s += x[i + n * q] * T[q + twon * j];
        X[i + n * j] = s;
      }
    interp_tetrahedron(n, X, xi, out, coeff);
    if (!*out) {
      if (c)
        for (int q = 0; q < twon; q++) {
          double s = 0.;
          for (int j = 0; j <= n; j++) s += T[q + twon * j] * coeff[j];
          c[q] = s;
        }
      return;
    }
  }
}

typedef struct databox {
  int imin[NDMAX], imax[NDMAX];
  double xmin[NDMAX], xmax[NDMAX];
  int type; /* 0 not splitted, 1 splitted, 2 cell */
  int nsub;
  struct databox *sub;
} databox;

typedef struct {
  int n, gshape[NDMAX], ioffset[NDMAX];
  int64_t total;
  const double *coord;   /* [n][total] */
  const uint8_t *masked; /* [total] or NULL */
  double *tet;
  databox root;
} ndgrid;

static void getcoord0(const ndgrid *g, const int *ind, double *x) { /* ndgrid.F90:1128-1157, dependence all ones */
  int64_t lin = 0;
  for (int d = 0; d < g->n; d++) lin += (int64_t)ind[d] * g->ioffset[d];
  for (int d = 0; d < g->n; d++) x[d] = g->coord[(int64_t)d * g->total + lin];
}

static void search_boundarybox(const ndgrid *g, databox *db) { /* ndgrid_inc.F90:760-840 */
  const int n = g->n;
  int ext[NDMAX], ind[NDMAX];
  int64_t tot = 1;
  for (int d = 0; d < n; d++) { ext[d] = db->imax[d] - db->imin[d] + 1; tot *= ext[d]; db->xmin[d] = HUGE_VAL; db->xmax[d] = -HUGE_VAL; }
  for (int64_t m = 0; m < tot; m++) {
    int64_t r = m;
    for (int d = 0; d < n; d++) { ind[d] = (int)(r % ext[d]) + db->imin[d]; r /= ext[d]; }
    double xc[NDMAX];
    getcoord0(g, ind, xc);
    for (int d = 0; d < n; d++) {
      if (xc[d] < db->xmin[d]) db->xmin[d] = xc[d];
      if (xc[d] > db->xmax[d]) db->xmax[d] = xc[d];
    }
  }
}

static void split_databox(const ndgrid *g, databox *db) { /* ndgrid_inc.F90:846-926 */
  const int n = g->n, twon = 1 << n;
  search_boundarybox(g, db);
  int any = 0;
  for (int d = 0; d < n; d++) any = any || (db->imax[d] - db->imin[d] > 1);
  if (!any) { db->type = 2; return; }
  db->type = 1;
  int im[NDMAX];
  for (int d = 0; d < n; d++) im[d] = (db->imin[d] + db->imax[d]) / 2;
  db->sub = (databox *)calloc(twon, sizeof(databox));
  db->nsub = 0;
  for (int m = 0; m < twon; m++) {
    databox s;
    memset(&s, 0, sizeof s);
    int keep = 0;
    for (int d = 0; d < n; d++) {
      if ((m >> d) & 1) { s.imin[d] = db->imin[d]; s.imax[d] = im[d]; }
      else { s.imin[d] = im[d]; s.imax[d] = db->imax[d]; }
      keep = keep || (s.imax[d] - s.imin[d] >= 1); /* subs = any(subimax-subimin >= 1, 1) */
    }
    if (keep) db->sub[db->nsub++] = s; /* define_databox: not splitted yet */
  }
}

static int incube(const ndgrid *g, const double *xi, const int *ind) { /* ndgrid_inc.F90:304-367 */
  const int n = g->n, twon = 1 << n;
  double px[NDMAX * TWONMAX];
  for (int j = 0; j < twon; j++) {
    int pind[NDMAX];
    for (int k = 0; k < n; k++) {
      const int up = ind[k] + 1 < g->gshape[k] - 1 ? ind[k] + 1 : g->gshape[k] - 1;
      pind[k] = ((j >> k) & 1) ? up : ind[k];
    }
    getcoord0(g, pind, px + n * j);
  }
  int out;
  interp_cube(n, g->tet, px, xi, &out, NULL);
  return !out;
}

static void locate_databox(const ndgrid *g, databox *db, const double *xi, int *ind, int *out) { /* :928-1001 */
  const int n = g->n;
  if (db->type == 0) split_databox(g, db);
  int outside = 0;
  for (int d = 0; d < n; d++) outside = outside || xi[d] < db->xmin[d] || xi[d] > db->xmax[d];
  if (outside) { *out = 1; return; }
  if (db->type != 2) {
    *out = 1;
    for (int p = 0; p < db->nsub; p++) {
      locate_databox(g, &db->sub[p], xi, ind, out);
      if (!*out) break;
    }
  } else {
    for (int d = 0; d < n; d++) ind[d] = db->imin[d];
    *out = !incube(g, xi, ind);
  }
}

static void free_databox(databox *db) {
  for (int p = 0; p < db->nsub; p++) free_databox(&db->sub[p]);
  free(db->sub);
}

void *oracle_ndgrid_create(int n, const int32_t *gshape, const double *coord, const uint8_t *masked) {
  if (n < 1 || n > NDMAX) return NULL;
  ndgrid *g = (ndgrid *)calloc(1, sizeof(ndgrid));
  g->n = n;
  g->total = 1;
  for (int d = 0; d < n; d++) { g->gshape[d] = gshape[d]; g->ioffset[d] = (int)g->total; g->total *= gshape[d]; }
  g->coord = coord;
  g->masked = masked;
  g->tet = (double *)malloc(sizeof(double) * (1 << n) * (n + 1) * oracle_ndgrid_nsimplex(n));
  oracle_ndgrid_tetrahedra(n, g->tet);
  for (int d = 0; d < n; d++) { g->root.imin[d] = 0; g->root.imax[d] = gshape[d] - 1; } /* init_databox :566 */
  return g;
}

void oracle_ndgrid_destroy(void *h) {
  ndgrid *g = (ndgrid *)h;
  if (!g) return;
  free_databox(&g->root);
  free(g->tet);
  free(g);
}

/* cinterp (ndgrid.F90:1183-1257) for m points: xi[m][n] row-major; indexes[m][2^n][n] 1-based, coeff[m][2^n], nbp[m] */
void oracle_cinterp(void *h, int m, const double *xi, int32_t *indexes, double *coeff, int32_t *nbp) {
  ndgrid *g = (ndgrid *)h;
  const int n = g->n, twon = 1 << n;
  for (int p = 0; p < m; p++) {
    int ind[NDMAX], out;
    int32_t *ix = indexes + (size_t)p * twon * n;
    double *cf = coeff + (size_t)p * twon;
    nbp[p] = 0;
    for (int q = 0; q < twon * n; q++) ix[q] = 0;
    for (int q = 0; q < twon; q++) cf[q] = 0.;
    locate_databox(g, &g->root, xi + (size_t)p * n, ind, &out);
    if (out) continue; /* the reference leaves indexes undefined (+1) here; callers only look at nbp */
    double px[NDMAX * TWONMAX];
    int anymasked = 0;
    for (int j = 0; j < twon; j++) {
      int cidx[NDMAX];
      int64_t lin = 0;
      for (int k = 0; k < n; k++) {
        cidx[k] = (((j >> k) & 1) && g->gshape[k] > 1) ? ind[k] + 1 : ind[k];
        lin += (int64_t)cidx[k] * g->ioffset[k];
        ix[j * n + k] = cidx[k] + 1;
      }
      if (g->masked && g->masked[lin]) anymasked = 1;
      getcoord0(g, cidx, px + n * j);
    }
    if (anymasked) continue;
    int o2;
    interp_cube(n, g->tet, px, xi + (size_t)p * n, &o2, cf);
    nbp[p] = twon;
  }
}
