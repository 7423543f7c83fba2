/*
 * oak_oracle.c — CPU restatement of OAK's local ensemble analysis.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (oak_b200/, include/) may
 * import, link or call this file; it is used by tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs, as the checker and as the
 * reported CPU baseline.
 *
 * Parity status: PINNED on the reference's own known-answer tests
 *   test/test_rrsqrt.F90 (global + local analysis, tol 1e-8),
 *   test/test_covariance.F90:612-617 (locfun known answers),
 *   test/test_assim.F90 (3x3x2 state, 1 obs, tol 1e-5),
 *   test/test_cellgrid.F90 (neighbour search completeness)
 * all regenerated from their closed-form inputs in tests/test_oracle_golden.py.
 * The reference itself (Fortran 2003 + NetCDF + LAPACK) cannot be compiled in
 * this image (no Fortran compiler), so oracle/_ref does not exist.
 *
 * Third-party arithmetic: the reference links an unpinned LAPACK/BLAS
 * (Compilers/libs.mk:75,80).  dsyev / dgemm / dgemv are taken here from the
 * OpenBLAS 0.3.x bundled with scipy (symbols scipy_dsyev_, scipy_dgemm_,
 * scipy_dgemv_), resolved at run time by oracle_init_blas().
 *
 * All citations are file:line under /root/reference.
 * All arrays are column-major with explicit leading dimensions, indices 0-based
 * at this C boundary (the Fortran code is 1-based).
 * Compile with -O2 -ffp-contract=off (strict IEEE: the selection predicate must
 * not be contracted into FMAs) and -fopenmp.
 */
#include <dlfcn.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/oak_b200_math.h" /* portable sin/cos/acos shared with the product (bit-exact metric) */

typedef void (*dsyev_fn)(const char *, const char *, const int *, double *, const int *, double *,
                         double *, const int *, int *);
typedef void (*dgemm_fn)(const char *, const char *, const int *, const int *, const int *,
                         const double *, const double *, const int *, const double *, const int *,
                         const double *, double *, const int *);
typedef void (*dgemv_fn)(const char *, const int *, const int *, const double *, const double *,
                         const int *, const double *, const int *, const double *, double *,
                         const int *);
typedef void (*setthr_fn)(int);

static dsyev_fn p_dsyev;
static dgemm_fn p_dgemm;
static dgemv_fn p_dgemv;
static void *blas_handle;

/* Resolve LAPACK/BLAS from the scipy-bundled OpenBLAS. Returns 0 on success. */
int oracle_init_blas(const char *path) {
  if (p_dsyev) return 0;
  blas_handle = dlopen(path, RTLD_NOW | RTLD_LOCAL);
  if (!blas_handle) {
    fprintf(stderr, "oracle: dlopen(%s) failed: %s\n", path, dlerror());
    return -1;
  }
  p_dsyev = (dsyev_fn)dlsym(blas_handle, "scipy_dsyev_");
  p_dgemm = (dgemm_fn)dlsym(blas_handle, "scipy_dgemm_");
  p_dgemv = (dgemv_fn)dlsym(blas_handle, "scipy_dgemv_");
  setthr_fn st = (setthr_fn)dlsym(blas_handle, "scipy_openblas_set_num_threads");
  if (!p_dsyev || !p_dgemm || !p_dgemv) return -2;
  /* one BLAS thread per zone; zones are parallelised by OpenMP (rrsqrt.F90:357) */
  if (st) st(1);
  return 0;
}

int oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* The launcher of a multi-process run (torch.distributed.run) exports OMP_NUM_THREADS=1; the CPU baseline of the
 * bench sets the zone-loop thread count explicitly (and reports it).  n <= 0: every online processor. */
int oracle_set_threads(int n) {
#ifdef _OPENMP
  if (n <= 0) n = omp_get_num_procs();
  omp_set_num_threads(n);
  return omp_get_max_threads();
#else
  (void)n;
  return 1;
#endif
}

/* ---------------------------------------------------------------------------
 * locfun — Gaspari–Cohn 5th-order piecewise rational, Horner form
 * covariance.F90:645-667
 * ------------------------------------------------------------------------- */
double oracle_locfun(double r) {
  if (r <= 1.) {
    return (((-r / 4. + 1. / 2.) * r + 5. / 8.) * r - 5. / 3.) * (r * r) + 1.;
  } else if (r <= 2.) {
    return ((((r / 12. - 1. / 2.) * r + 5. / 8.) * r + 5. / 3.) * r - 5.) * r + 4 - 2. / (3 * r);
  }
  return 0;
}

/* ---------------------------------------------------------------------------
 * distance(p0,p1) — assimilation.F90:3635-3672
 * p0 = observation (x0,y0), p1 = zone position (x1,y1).
 * metrictype 0 Cartesian, 1 Spherical (default), 2 SphericalApprox
 * (assimilation.F90:108-112).  trig = 0: libm sin/cos/acos (what gfortran calls);
 * trig = 1: the portable correctly-ordered implementations in
 * include/oak_b200_math.h, bit-identical to the device code.
 * ------------------------------------------------------------------------- */
#define ORACLE_PI 3.141592653589793238462643383279502884197
static const double EarthRadius = 6378137.;

double oracle_distance(int metrictype, int trig, double x0, double y0, double x1, double y1) {
  const double pi = ORACLE_PI;
  if (metrictype == 0) {
    double dx = x1 - x0, dy = y1 - y0;
    return sqrt(dx * dx + dy * dy); /* sqrt(sum((p1-p0)**2)) :3645 */
  } else if (metrictype == 2) {
    double coeff = pi * EarthRadius / (180.);
    double cc = trig ? oakm_cos((y0 + y1) * (pi / 360.)) : cos((y0 + y1) * (pi / 360.));
    double u = coeff * cc * (x1 - x0);
    double v = coeff * (y1 - y0);
    return sqrt(u * u + v * v); /* :3649-3651 */
  } else {
    double d2r = pi / 180.;
    double a = y0 * d2r, b = y1 * d2r, C = (x1 - x0) * d2r; /* :3654-3656 */
    double coeff;
    if (trig)
      coeff = oakm_sin(b) * oakm_sin(a) + oakm_cos(b) * oakm_cos(a) * oakm_cos(C);
    else
      coeff = sin(b) * sin(a) + cos(b) * cos(a) * cos(C); /* :3658 */
    coeff = fmax(fmin(coeff, 1.), -1.);                     /* :3659 */
    double d = trig ? oakm_acos(coeff) : acos(coeff);
    return EarthRadius * d; /* :3664 */
  }
}

/* ---------------------------------------------------------------------------
 * selectObservations — assimilation.F90:3683-3771 (production callback, Gaussian
 * weight, `<=` cut-off) and test/test_rrsqrt.F90:254-271 (Gaspari–Cohn callback,
 * relevant = c /= 0) and :239-247 (selectAllObservations).
 *
 * The zone position (zx,zy,zz,zt) is the coordinate of the zone's FIRST element
 * (rrsqrt.F90:368 passes startIndex(zi); assimilation.F90:3713-3740).
 * weightfun: 0 gaussian, 1 gaspari_cohn, 2 uniform (all observations, c=1).
 * Returns the number of relevant observations.
 * ------------------------------------------------------------------------- */
typedef struct {
  int32_t m;
  const double *obsx, *obsy, *obsz, *obst;
  int32_t loctype;    /* 1 horizontal, 2 depth, 3 time  (assimilation.F90:204-208) */
  int32_t metrictype; /* 0,1,2 */
  int32_t weightfun;
  int32_t trig; /* 0 libm, 1 portable */
} oracle_obs_t;

int oracle_select_observations(const oracle_obs_t *o, double zx, double zy, double zz, double zt,
                               double corrLen, double maxLen, double *weight, uint8_t *relevant) {
  int count = 0;
  for (int l = 0; l < o->m; l++) {
    double d;
    if (o->loctype == 1)
      d = oracle_distance(o->metrictype, o->trig, o->obsx[l], o->obsy ? o->obsy[l] : 0., zx, zy);
    else if (o->loctype == 2)
      d = fabs(o->obsz[l] - zz);
    else
      d = fabs(o->obst[l] - zt);
    weight[l] = d;
    if (o->weightfun == 0) {
      relevant[l] = d <= maxLen; /* :3756 */
    } else if (o->weightfun == 1) {
      double c = oracle_locfun(d / corrLen); /* test_rrsqrt.F90:268 */
      weight[l] = c;
      relevant[l] = c != 0; /* :269 */
    } else {
      weight[l] = 1.;
      relevant[l] = 1;
    }
    count += relevant[l];
  }
  if (o->weightfun == 0 && count > 0) {
    for (int l = 0; l < o->m; l++) {
      double t = weight[l] / corrLen;
      weight[l] = exp(-(t * t)); /* :3767 */
    }
  }
  return count;
}

/* ---------------------------------------------------------------------------
 * perpSpace — matoper.F90:509-535 ; H is n x (n-1), column-major ld n
 * ------------------------------------------------------------------------- */
static void perp_space(int n, const double *w, double *H) {
  memset(H, 0, sizeof(double) * n * (n - 1));
  double alpha = -1 / (fabs(w[n - 1]) + 1);
  for (int j = 0; j < n - 1; j++)
    for (int i = 0; i < n - 1; i++) {
      H[i + (size_t)n * j] = alpha * w[i] * w[j];
      if (i == j) H[i + (size_t)n * j] += 1;
    }
  double sg = copysign(1., w[n - 1]); /* sign(1.,w(n)) */
  for (int j = 0; j < n - 1; j++) H[(n - 1) + (size_t)n * j] = alpha * (w[n - 1] + sg) * w[j];
}

/* RotateVector(w,v) = v w^T + perpSpace(v) perpSpace(w)^T — rrsqrt.F90:737-744 */
void oracle_rotate_vector(int n, const double *w, const double *v, double *Omega) {
  double *Hv = malloc(sizeof(double) * n * (n - 1));
  double *Hw = malloc(sizeof(double) * n * (n - 1));
  perp_space(n, v, Hv);
  perp_space(n, w, Hw);
  for (int j = 0; j < n; j++)
    for (int i = 0; i < n; i++) {
      double s = 0;
      for (int k = 0; k < n - 1; k++) s += Hv[i + (size_t)n * k] * Hw[j + (size_t)n * k];
      Omega[i + (size_t)n * j] = v[i] * w[j] + s;
    }
  free(Hv);
  free(Hw);
}

/* ---------------------------------------------------------------------------
 * analysisIncrement — rrsqrt.F90:100-190 with ROTATE_ENSEMBLE (rrsqrt.F90:24)
 *
 *   R^-1 is the DCDCovar(weight, R) of rrsqrt.F90:388-397 with R = DiagCovar(var)
 *   optionally wrapped by DCDCovar(d01, .) for excluded observations
 *   (assimilation.F90:3086-3092):  R_loc^-1 x = w * ( e * ((e * (w*x)) / var) )
 *   (covariance.F90:612-619, :425-431), evaluated in that order.
 *
 * m local observations; Sf is nrow x N (ld ldS); HSf is m x N (ld ldH).
 * Outputs xa_xf[nrow], Sa (nrow x N, ld ldSa), ampl[N] (optional).
 * Returns 0, or 1 if ampl contains NaN (rrsqrt.F90:145-149 aborts), <0 LAPACK.
 * ------------------------------------------------------------------------- */
static inline double rinv_apply(double x, double w, double e, double var) {
  double t = w * x;          /* DCD outer: D*x                   covariance.F90:618 */
  t = e * ((e * t) / var);   /* inner DCD(d01, Diag): D*((D*x)/var)  :618, :431    */
  return w * t;              /* D * (...)                                            */
}

int oracle_analysis_increment(int m, int nrow, int N, const double *Hxf, const double *yo,
                              const double *Sf, int ldS, const double *HSf, int ldH,
                              const double *w, const double *e01, const double *var,
                              double *xa_xf, double *Sa, int ldSa, double *ampl_out) {
  const double one = 1., zero = 0.;
  const int ione = 1;
  int info = 0;
  double *RiH = malloc(sizeof(double) * (size_t)(m > 0 ? m : 1) * N);
  double *U = malloc(sizeof(double) * (size_t)N * N);
  double *lambda = calloc(N, sizeof(double));
  double *sq = malloc(sizeof(double) * N);
  double *tv = malloc(sizeof(double) * (size_t)(m > 0 ? m : 1));
  double *t1 = malloc(sizeof(double) * N), *t2 = malloc(sizeof(double) * N);
  double *ampl = malloc(sizeof(double) * N);

  /* temp = HSf .tx. (R%mldivide(HSf))   :135 ; Covar_mldivide_mat column by column (covariance.F90:291-302) */
  for (int k = 0; k < N; k++)
    for (int l = 0; l < m; l++)
      RiH[l + (size_t)m * k] = rinv_apply(HSf[l + (size_t)ldH * k], w[l], e01 ? e01[l] : 1., var[l]);
  /* dgemm('t','n') matoper_inc.F90:435 */
  if (m > 0)
    p_dgemm("t", "n", &N, &N, &m, &one, HSf, &ldH, RiH, &m, &zero, U, &N);
  else
    memset(U, 0, sizeof(double) * N * N);
  /* symeig -> dsyev('V','L') with workspace query  matoper_inc.F90:991-995 */
  {
    double rl;
    int lw = -1;
    p_dsyev("V", "L", &N, U, &N, lambda, &rl, &lw, &info);
    lw = (int)lround(rl);
    double *work = malloc(sizeof(double) * lw);
    p_dsyev("V", "L", &N, U, &N, lambda, work, &lw, &info);
    free(work);
    if (info != 0) goto done;
  }
  for (int i = 0; i < N; i++) {
    if (lambda[i] < 0) lambda[i] = 0; /* :137 */
    lambda[i] = 1. / (1 + lambda[i]); /* :140 */
  }
  /* ampl = U.x.(lambda.dx.(U.tx.(HSf.tx.(R%mldivide(yo-Hxf)))))   :142 */
  for (int l = 0; l < m; l++) tv[l] = rinv_apply(yo[l] - Hxf[l], w[l], e01 ? e01[l] : 1., var[l]);
  if (m > 0)
    p_dgemv("t", &m, &N, &one, HSf, &ldH, tv, &ione, &zero, t1, &ione);
  else
    memset(t1, 0, sizeof(double) * N);
  p_dgemv("t", &N, &N, &one, U, &N, t1, &ione, &zero, t2, &ione);
  for (int i = 0; i < N; i++) t2[i] = lambda[i] * t2[i];
  p_dgemv("n", &N, &N, &one, U, &N, t2, &ione, &zero, ampl, &ione);
  for (int i = 0; i < N; i++)
    if (ampl[i] != ampl[i]) { info = 1; goto done; } /* :145-149 */
  /* xa_xf = Sf.x.ampl   :151 */
  p_dgemv("n", &nrow, &N, &one, Sf, &ldS, ampl, &ione, &zero, xa_xf, &ione);

  if (Sa) {
    for (int i = 0; i < N; i++) sq[i] = sqrt(lambda[i]); /* :162 */
    /* w = 1/sqrt(N) ; v = U.x.(sum(U,1)/sqrt_lambda) ; v = normate(v)   :176-178 */
    double *wv = malloc(sizeof(double) * N), *v = malloc(sizeof(double) * N);
    double *Om = malloc(sizeof(double) * (size_t)N * N), *T1 = malloc(sizeof(double) * (size_t)N * N),
           *T2 = malloc(sizeof(double) * (size_t)N * N);
    for (int i = 0; i < N; i++) {
      wv[i] = 1. / sqrt(1. * N);
      double s = 0;
      for (int j = 0; j < N; j++) s += U[j + (size_t)N * i]; /* sum(U,1) */
      t1[i] = s / sq[i];
    }
    p_dgemv("n", &N, &N, &one, U, &N, t1, &ione, &zero, v, &ione);
    double nn = 0;
    for (int i = 0; i < N; i++) nn += v[i] * v[i];
    nn = sqrt(nn);
    for (int i = 0; i < N; i++) v[i] = v[i] / nn; /* normate :750-756 */
    oracle_rotate_vector(N, wv, v, Om);
    /* Sa = Sf.x.(U.x.(sqrt_lambda.dx.(U.tx.RotateVector(w,v))))   :182 */
    p_dgemm("t", "n", &N, &N, &N, &one, U, &N, Om, &N, &zero, T1, &N);
    for (int j = 0; j < N; j++)
      for (int i = 0; i < N; i++) T1[i + (size_t)N * j] *= sq[i];
    p_dgemm("n", "n", &N, &N, &N, &one, U, &N, T1, &N, &zero, T2, &N);
    p_dgemm("n", "n", &nrow, &N, &N, &one, Sf, &ldS, T2, &N, &zero, Sa, &ldSa);
    free(wv); free(v); free(Om); free(T1); free(T2);
  }
  if (ampl_out) memcpy(ampl_out, ampl, sizeof(double) * N);
done:
  free(RiH); free(U); free(lambda); free(sq); free(tv); free(t1); free(t2); free(ampl);
  return info;
}

/* ---------------------------------------------------------------------------
 * analysis (global scheme) — rrsqrt.F90:196-208: xa = xf + increment
 * ------------------------------------------------------------------------- */
int oracle_analysis(int m, int n, int N, const double *xf, const double *Hxf, const double *yo,
                    const double *Sf, int ldS, const double *HSf, int ldH, const double *var,
                    double *xa, double *Sa, int ldSa, double *ampl) {
  double *w = malloc(sizeof(double) * (m > 0 ? m : 1));
  for (int l = 0; l < m; l++) w[l] = 1.;
  int info = oracle_analysis_increment(m, n, N, Hxf, yo, Sf, ldS, HSf, ldH, w, NULL, var, xa, Sa,
                                       ldSa, ampl);
  for (int i = 0; i < n; i++) xa[i] = xf[i] + xa[i];
  free(w);
  return info;
}

/* ---------------------------------------------------------------------------
 * locAnalysis / locAnalysisIncrement — rrsqrt.F90:433-466, :258-426
 *
 * zoneSize[nzones]; zone z owns rows [start_z, start_z+zoneSize_z) of the
 * zone-permuted state (prefix sums :328-335).  Zone positions and localisation
 * lengths are given per zone (they are per-element arrays indexed by the zone's
 * first element in the reference: assimilation.F90:3713, :3756, :3767).
 *
 * local_obs != 0 (the default, rrsqrt.F90:318-320): only relevant observations
 * are packed (:395-404).  local_obs == 0: all observations with weight (:374-384);
 * amplitudes(:,zi) is then filled, otherwise it stays 0 (:324).
 *
 * zone_list/nlist: optional subset of zones to analyse (bench sampling); other
 * zones keep Sa=Sf, xa=xf.  mloc_out[nzones] optional: relevant-obs count.
 * Returns 0 or the first non-zero analysisIncrement status.
 * ------------------------------------------------------------------------- */
int oracle_loc_analysis(int nzones, const int32_t *zoneSize, const double *zx, const double *zy,
                        const double *zz, const double *zt, const double *corrLen,
                        const double *maxLen, const oracle_obs_t *obs, int local_obs, int n, int N,
                        const double *xf, const double *Hxf, const double *yo, const double *Sf,
                        int ldS, const double *HSf, int ldH, const double *var, const double *e01,
                        double *xa, double *Sa, int ldSa, double *amplitudes,
                        const int32_t *zone_list, int nlist, int32_t *mloc_out) {
  const int m = obs->m;
  int64_t *start = malloc(sizeof(int64_t) * (nzones + 1));
  start[0] = 0;
  for (int z = 0; z < nzones; z++) start[z + 1] = start[z] + zoneSize[z];
  if (start[nzones] != n) { free(start); return -10; }
  /* :322-326  xa_xf = 0 ; amplitudes = 0 ; Sa = Sf */
  for (int i = 0; i < n; i++) xa[i] = 0;
  if (amplitudes) memset(amplitudes, 0, sizeof(double) * (size_t)N * nzones);
  for (int k = 0; k < N; k++) memcpy(Sa + (size_t)ldSa * k, Sf + (size_t)ldS * k, sizeof(double) * n);
  if (mloc_out) memset(mloc_out, 0, sizeof(int32_t) * nzones);
  int status = 0;
  const int niter = zone_list ? nlist : nzones;
#pragma omp parallel
  {
    double *weight = malloc(sizeof(double) * (m > 0 ? m : 1));
    uint8_t *rel = malloc(m > 0 ? m : 1);
    double *yoz = malloc(sizeof(double) * (m > 0 ? m : 1));
    double *Hxfz = malloc(sizeof(double) * (m > 0 ? m : 1));
    double *wz = malloc(sizeof(double) * (m > 0 ? m : 1));
    double *varz = malloc(sizeof(double) * (m > 0 ? m : 1));
    double *ez = malloc(sizeof(double) * (m > 0 ? m : 1));
    double *HSfz = malloc(sizeof(double) * (size_t)(m > 0 ? m : 1) * N);
#pragma omp for schedule(dynamic) /* rrsqrt.F90:357 */
    for (int it = 0; it < niter; it++) {
      int zi = zone_list ? zone_list[it] : it;
      int64_t i1 = start[zi];
      int nrow = zoneSize[zi];
      /* callback with the zone's first element :368 */
      int nb = oracle_select_observations(obs, zx ? zx[zi] : 0., zy ? zy[zi] : 0., zz ? zz[zi] : 0.,
                                          zt ? zt[zi] : 0., corrLen[zi], maxLen[zi], weight, rel);
      if (mloc_out) mloc_out[zi] = nb;
      if (nb == 0) continue; /* :370-371 */
      int info;
      if (!local_obs) {
        info = oracle_analysis_increment(m, nrow, N, Hxf, yo, Sf + i1, ldS, HSf, ldH, weight, e01,
                                         var, xa + i1, Sa + i1, ldSa,
                                         amplitudes ? amplitudes + (size_t)N * zi : NULL);
      } else if (nb == m) { /* :388-391 */
        info = oracle_analysis_increment(m, nrow, N, Hxf, yo, Sf + i1, ldS, HSf, ldH, weight, e01,
                                         var, xa + i1, Sa + i1, ldSa, NULL);
      } else { /* :395-404 pack relevant observations */
        int nObs = 0;
        for (int j = 0; j < m; j++)
          if (rel[j]) {
            yoz[nObs] = yo[j];
            Hxfz[nObs] = Hxf[j];
            wz[nObs] = weight[j];
            varz[nObs] = var[j];
            ez[nObs] = e01 ? e01[j] : 1.;
            for (int k = 0; k < N; k++) HSfz[nObs + (size_t)nb * k] = HSf[j + (size_t)ldH * k];
            nObs++;
          }
        info = oracle_analysis_increment(nObs, nrow, N, Hxfz, yoz, Sf + i1, ldS, HSfz, nb, wz, ez,
                                         varz, xa + i1, Sa + i1, ldSa, NULL);
      }
      if (info != 0) {
#pragma omp critical
        if (!status) status = info;
      }
    }
    free(weight); free(rel); free(yoz); free(Hxfz); free(wz); free(varz); free(ez); free(HSfz);
  }
  /* locAnalysis: xa = xf + xa   :462 */
  for (int i = 0; i < n; i++) xa[i] = xf[i] + xa[i];
  free(start);
  return status;
}

/* ---------------------------------------------------------------------------
 * The same local analysis with a CPU cell grid in front of the exact predicate instead of the O(m) scan of
 * assimilation.F90:3745-3757 per zone: the "fair" CPU baseline of SURVEY.md 8(d)(ii) (what the reference's own
 * CELLGRID_SEARCH variant, ndgrid.F90:1489-1691, is for).  The predicate, the weights, the pack order
 * (increasing observation number, rrsqrt.F90:395-404) and analysisIncrement are the ones above, so the result is
 * the one of oracle_loc_analysis.  Only the case the benchmark uses: horizontal localisation, Cartesian metric,
 * Gaussian weights with a finite cut-off (returns -11 otherwise); local_obs branch only.
 * ------------------------------------------------------------------------- */
static int cmp_i32(const void *a, const void *b) {
  const int32_t x = *(const int32_t *)a, y = *(const int32_t *)b;
  return (x > y) - (x < y);
}

int oracle_loc_analysis_cellgrid(int nzones, const int32_t *zoneSize, const double *zx, const double *zy,
                                 const double *corrLen, const double *maxLen, const oracle_obs_t *obs, int n,
                                 int N, const double *xf, const double *Hxf, const double *yo, const double *Sf,
                                 int ldS, const double *HSf, int ldH, const double *var, const double *e01,
                                 double *xa, double *Sa, int ldSa, int32_t *mloc_out) {
  const int m = obs->m;
  if (obs->loctype != 1 || obs->metrictype != 0 || obs->weightfun != 0 || !zy || !obs->obsy) return -11;
  double rmax = 0;
  for (int z = 0; z < nzones; z++) {
    if (!(maxLen[z] < 1e300)) return -11;
    if (maxLen[z] > rmax) rmax = maxLen[z];
  }
  int64_t *start = malloc(sizeof(int64_t) * (nzones + 1));
  start[0] = 0;
  for (int z = 0; z < nzones; z++) start[z + 1] = start[z] + zoneSize[z];
  if (start[nzones] != n) { free(start); return -10; }
  for (int i = 0; i < n; i++) xa[i] = 0;
  for (int k = 0; k < N; k++) memcpy(Sa + (size_t)ldSa * k, Sf + (size_t)ldS * k, sizeof(double) * n);
  if (mloc_out) memset(mloc_out, 0, sizeof(int32_t) * nzones);
  /* buckets of edge rmax (at most ~4 m cells), observation numbers ascending inside a bucket */
  double x0 = 0, x1 = 0, y0 = 0, y1 = 0;
  for (int l = 0; l < m; l++) {
    const double x = obs->obsx[l], y = obs->obsy[l];
    if (l == 0) { x0 = x1 = x; y0 = y1 = y; }
    x0 = fmin(x0, x); x1 = fmax(x1, x);
    y0 = fmin(y0, y); y1 = fmax(y1, y);
  }
  double cs = rmax > 0 ? rmax : 1.;
  while (((x1 - x0) / cs + 1.) * ((y1 - y0) / cs + 1.) > 4. * (m > 1024 ? m : 1024)) cs *= 2.;
  const int ncx = (int)floor((x1 - x0) / cs) + 1, ncy = (int)floor((y1 - y0) / cs) + 1;
  int32_t *cstart = calloc((size_t)ncx * ncy + 1, sizeof(int32_t));
  int32_t *cell = malloc(sizeof(int32_t) * (m > 0 ? m : 1));
  int32_t *perm = malloc(sizeof(int32_t) * (m > 0 ? m : 1));
  for (int l = 0; l < m; l++) {
    int cx = (int)floor((obs->obsx[l] - x0) / cs), cy = (int)floor((obs->obsy[l] - y0) / cs);
    cx = cx < 0 ? 0 : (cx >= ncx ? ncx - 1 : cx);
    cy = cy < 0 ? 0 : (cy >= ncy ? ncy - 1 : cy);
    cell[l] = cy * ncx + cx;
    cstart[cell[l] + 1]++;
  }
  for (int c = 0; c < ncx * ncy; c++) cstart[c + 1] += cstart[c];
  {
    int32_t *fill = malloc(sizeof(int32_t) * ((size_t)ncx * ncy + 1));
    memcpy(fill, cstart, sizeof(int32_t) * ((size_t)ncx * ncy + 1));
    for (int l = 0; l < m; l++) perm[fill[cell[l]]++] = l;
    free(fill);
  }
  int status = 0;
#pragma omp parallel
  {
    int32_t *sel = malloc(sizeof(int32_t) * (m > 0 ? m : 1));
    double *yoz = malloc(sizeof(double) * (m > 0 ? m : 1));
    double *Hxfz = malloc(sizeof(double) * (m > 0 ? m : 1));
    double *wz = malloc(sizeof(double) * (m > 0 ? m : 1));
    double *varz = malloc(sizeof(double) * (m > 0 ? m : 1));
    double *ez = malloc(sizeof(double) * (m > 0 ? m : 1));
    double *HSfz = NULL;
    size_t HSfz_cap = 0;
#pragma omp for schedule(dynamic)
    for (int zi = 0; zi < nzones; zi++) {
      const double px = zx[zi], py = zy[zi], R = maxLen[zi];
      const double slack = R * 1e-9 + 1e-12 * (fabs(px) + fabs(py) + fabs(x0) + fabs(y0));
      int cx0 = (int)floor((px - R - slack - x0) / cs), cx1 = (int)floor((px + R + slack - x0) / cs);
      int cy0 = (int)floor((py - R - slack - y0) / cs), cy1 = (int)floor((py + R + slack - y0) / cs);
      cx0 = cx0 < 0 ? 0 : cx0;
      cy0 = cy0 < 0 ? 0 : cy0;
      cx1 = cx1 >= ncx ? ncx - 1 : cx1;
      cy1 = cy1 >= ncy ? ncy - 1 : cy1;
      int nb = 0;
      for (int cy = cy0; cy <= cy1; cy++)
        for (int q = cstart[cy * ncx + cx0]; cx0 <= cx1 && q < cstart[cy * ncx + cx1 + 1]; q++) {
          const int l = perm[q];
          const double d = oracle_distance(0, obs->trig, obs->obsx[l], obs->obsy[l], px, py);
          if (d <= R) sel[nb++] = l; /* assimilation.F90:3756 */
        }
      if (mloc_out) mloc_out[zi] = nb;
      if (nb == 0) continue; /* rrsqrt.F90:370-371 */
      qsort(sel, nb, sizeof(int32_t), cmp_i32); /* pack() keeps the observation order :395-404 */
      if ((size_t)nb * N > HSfz_cap) {
        free(HSfz);
        HSfz_cap = (size_t)nb * N * 2;
        HSfz = malloc(sizeof(double) * HSfz_cap);
      }
      for (int q = 0; q < nb; q++) {
        const int j = sel[q];
        const double d = oracle_distance(0, obs->trig, obs->obsx[j], obs->obsy[j], px, py);
        const double t = d / corrLen[zi];
        yoz[q] = yo[j];
        Hxfz[q] = Hxf[j];
        wz[q] = exp(-(t * t)); /* assimilation.F90:3767 */
        varz[q] = var[j];
        ez[q] = e01 ? e01[j] : 1.;
        for (int k = 0; k < N; k++) HSfz[q + (size_t)nb * k] = HSf[j + (size_t)ldH * k];
      }
      const int64_t i1 = start[zi];
      const int info = oracle_analysis_increment(nb, zoneSize[zi], N, Hxfz, yoz, Sf + i1, ldS, HSfz, nb, wz, ez,
                                                 varz, xa + i1, Sa + i1, ldSa, NULL);
      if (info != 0) {
#pragma omp critical
        if (!status) status = info;
      }
    }
    free(sel); free(yoz); free(Hxfz); free(wz); free(varz); free(ez); free(HSfz);
  }
  for (int i = 0; i < n; i++) xa[i] = xf[i] + xa[i];
  free(start); free(cstart); free(cell); free(perm);
  return status;
}

/* ---------------------------------------------------------------------------
 * obsoper: COO SpMV  Hx(i(k)) += s(k) * x(j(k)) then + Hshift
 * assimilation.F90:2836-2898 ; matoper_inc.F90:220-242.  Indices 1-based as
 * stored by the Fortran code; entries with j<=0 (out-of-grid obs,
 * assimilation.F90:2597-2611) are skipped only if their coefficient is 0.
 * ------------------------------------------------------------------------- */
void oracle_obsoper(int m, int64_t nnz, const int32_t *Hi, const int32_t *Hj, const double *Hs,
                    const double *Hshift, const double *x, double *Hx) {
  for (int i = 0; i < m; i++) Hx[i] = 0;
  for (int64_t k = 0; k < nnz; k++) {
    if (Hj[k] <= 0) continue;
    Hx[Hi[k] - 1] += Hs[k] * x[Hj[k] - 1];
  }
  if (Hshift)
    for (int i = 0; i < m; i++) Hx[i] += Hshift[i];
}

/* interp1 — anamorphosis.F90:304-339: first bracket x(kp) <= xi < x(kp+1), linear blend
 * (1-alpha) y(k) + alpha y(k+1); outside every bracket: y(1) if xi < x(1), else y(end); *out = no bracket. */
double oracle_interp1(int K, const double *x, const double *y, double xi, int *out) {
  int k = -1;
  for (int kp = 0; kp < K - 1; kp++)
    if (x[kp] <= xi && xi < x[kp + 1]) { k = kp; break; }
  double yi;
  if (k != -1) {
    const double alpha = (xi - x[k]) / (x[k + 1] - x[k]);
    yi = (1 - alpha) * y[k] + alpha * y[k + 1];
  } else {
    yi = (xi < x[0]) ? y[0] : y[K - 1];
  }
  if (out) *out = (k == -1);
  return yi;
}

/* anamtransform, one element — assimilation.F90:4516-4576.  Types: 1 identity, 2 log/exp, 3 tabulated
 * (transform(:,1) physical values, transform(:,2) transformed values; tab = K x 2 column-major).  For type 3 an
 * extrapolated value is overwritten by the end of the table's INPUT-side column (:4560-4567): the already
 * interpolated (output-side) value is compared with transform(1,ti) — reproduced as is. */
double oracle_anam(int type, int forward, int K, const double *tab, double x) {
  if (type == 2) return forward ? log(x) : exp(x);
  if (type == 3) {
    const double *ti = forward ? tab : tab + K, *tj = forward ? tab + K : tab;
    int out;
    double v = oracle_interp1(K, ti, tj, x, &out);
    if (out) v = (v < ti[0]) ? ti[0] : ti[K - 1];
    return v;
  }
  return x;
}
static const double *g_anamtab = 0;  /* table of the tabulated anamorphosis (set by oracle_set_anam_table) */
static int g_anamK = 0;
void oracle_set_anam_table(int K, const double *tab) { g_anamtab = tab; g_anamK = K; }
/* per-variable transforms (anamtype 0): AnamTrans%anam(v) looked up for every element through ind2submv
 * (assimilation.F90:4531-4535): rowvar[i] = 0-based variable of (permuted) row i, vtype[v] its type, the K_v x 2
 * table of a tabulated one at vtab + voff[v] */
static const int32_t *g_rowvar = 0, *g_vtype = 0, *g_voff = 0, *g_vK = 0;
static const double *g_vtab = 0;
void oracle_set_anam_vars(const int32_t *rowvar, const int32_t *vtype, const int32_t *voff, const int32_t *vK,
                          const double *vtab) {
  g_rowvar = rowvar; g_vtype = vtype; g_voff = voff; g_vK = vK; g_vtab = vtab;
}
static inline double anam_el(int type, int forward, int row, double x) {
  if (type == 0) {
    const int v = g_rowvar[row];
    return oracle_anam(g_vtype[v], forward, g_vK[v], g_vtab + g_voff[v], x);
  }
  return oracle_anam(type, forward, g_anamK, g_anamtab, x);
}
#define anam_fwd(type, x) anam_el(type, 1, i, x)
#define anam_inv(type, x) anam_el(type, 0, i, x)

/* ---------------------------------------------------------------------------
 * Ensemble branch of Assim with the local scheme:
 *   prologue  assimilation.F90:3083, :3106-3134
 *   locanalysis call :3235-3236
 *   epilogue  :3301-3312 (inflation, maxCorrection), :3318-3357, :3558-3562
 * E is n x N (ld ldE) zone-permuted; Ea output n x N (ld ldEa).
 * scaling = sqrt(N-1) (ppdef.h:52).
 * ------------------------------------------------------------------------- */
int oracle_assim_ensemble(int nzones, const int32_t *zoneSize, const double *zx, const double *zy,
                          const double *zz, const double *zt, const double *corrLen,
                          const double *maxLen, const oracle_obs_t *obs, int n, int N,
                          const double *E, int ldE, int64_t nnz, const int32_t *Hi,
                          const int32_t *Hj, const double *Hs, const double *Hshift,
                          const double *yo, const double *var, const double *e01,
                          int anamtype, double inflation, const double *maxCorrection, double *Ea,
                          int ldEa, double *xf_out, double *xa_out) {
  const int m = obs->m;
  const double scaling = sqrt(N - 1.);
  double *Sf = malloc(sizeof(double) * (size_t)n * N);
  double *Sa = malloc(sizeof(double) * (size_t)n * N);
  double *HSf = malloc(sizeof(double) * (size_t)(m > 0 ? m : 1) * N);
  double *xf = calloc(n, sizeof(double)), *xa = calloc(n, sizeof(double));
  double *Hxf = calloc(m > 0 ? m : 1, sizeof(double));
  /* HSf(:,k) = obsoper(H,Sf(:,k)) + Hshift  :3112-3114 (on the untransformed state) */
  for (int k = 0; k < N; k++)
    oracle_obsoper(m, nnz, Hi, Hj, Hs, Hshift, E + (size_t)ldE * k, HSf + (size_t)m * k);
  /* anamtransform(.true.) :3123-3125 */
  for (int k = 0; k < N; k++)
    for (int i = 0; i < n; i++) Sf[i + (size_t)n * k] = anam_fwd(anamtype, E[i + (size_t)ldE * k]);
  /* xf = sum(Sf,2)/N ; Hxf = sum(HSf,2)/N ; anomalies / scaling  :3127-3134 */
  for (int k = 0; k < N; k++)
    for (int i = 0; i < n; i++) xf[i] += Sf[i + (size_t)n * k];
  for (int i = 0; i < n; i++) xf[i] /= N;
  for (int k = 0; k < N; k++)
    for (int i = 0; i < m; i++) Hxf[i] += HSf[i + (size_t)m * k];
  for (int i = 0; i < m; i++) Hxf[i] /= N;
  for (int k = 0; k < N; k++) {
    for (int i = 0; i < n; i++) Sf[i + (size_t)n * k] = (Sf[i + (size_t)n * k] - xf[i]) / scaling;
    for (int i = 0; i < m; i++) HSf[i + (size_t)m * k] = (HSf[i + (size_t)m * k] - Hxf[i]) / scaling;
  }
  int info = oracle_loc_analysis(nzones, zoneSize, zx, zy, zz, zt, corrLen, maxLen, obs, 1, n, N, xf,
                                 Hxf, yo, Sf, n, HSf, m, var, e01, xa, Sa, n, NULL, NULL, 0, NULL);
  /* Sa = inflation * Sa :3301-3304 */
  if (inflation != 1.)
    for (size_t i = 0; i < (size_t)n * N; i++) Sa[i] *= inflation;
  /* saturate correction :3308-3312 */
  if (maxCorrection)
    for (int i = 0; i < n; i++) {
      /* where (xa-maxCorrection.gt.xf) xa=xf+maxCorrection ; where (xa.lt.xf-maxCorrection) xa=xf-maxCorrection */
      if (xa[i] - maxCorrection[i] > xf[i]) xa[i] = xf[i] + maxCorrection[i];
      if (xa[i] < xf[i] - maxCorrection[i]) xa[i] = xf[i] - maxCorrection[i];
    }
  /* Ea(:,k) = xa + scaling*Sa(:,k) ; inverse anamorphosis  :3318-3326, :3558-3562 */
  for (int k = 0; k < N; k++)
    for (int i = 0; i < n; i++)
      Ea[i + (size_t)ldEa * k] = anam_inv(anamtype, xa[i] + Sa[i + (size_t)n * k] * scaling);
  if (xf_out) memcpy(xf_out, xf, sizeof(double) * n);
  if (xa_out) {   /* xa = sum(Sa,2)/size(Sa,2) of the back-transformed ensemble :3343-3349 */
    for (int i = 0; i < n; i++) xa_out[i] = 0.;
    for (int k = 0; k < N; k++)
      for (int i = 0; i < n; i++) xa_out[i] += Ea[i + (size_t)ldEa * k];
    for (int i = 0; i < n; i++) xa_out[i] /= N;
  }
  free(Sf); free(Sa); free(HSf); free(xf); free(xa); free(Hxf);
  return info;
}

/* ---------------------------------------------------------------------------
 * initPartition — assimilation.F90:578-641: stable counting sort of partition
 * labels (already gap-free 1..nzones) -> zoneSize, zoneIndex (1-based packed
 * index stored at each permuted position), invZoneIndex.
 * ------------------------------------------------------------------------- */
void oracle_init_partition(int n, const int32_t *partition, int nzones, int32_t *zoneSize,
                           int32_t *zoneIndex, int32_t *invZoneIndex) {
  int64_t *cursor = calloc(nzones + 1, sizeof(int64_t));
  for (int z = 0; z < nzones; z++) zoneSize[z] = 0;
  for (int i = 0; i < n; i++) zoneSize[partition[i] - 1]++;
  for (int z = 0; z < nzones; z++) cursor[z + 1] = cursor[z] + zoneSize[z];
  for (int i = 0; i < n; i++) {
    int z = partition[i] - 1;
    zoneIndex[cursor[z]] = i + 1;
    invZoneIndex[i] = (int32_t)cursor[z] + 1;
    cursor[z]++;
  }
  free(cursor);
}
