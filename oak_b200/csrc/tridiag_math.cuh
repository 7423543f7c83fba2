// tridiag_math.cuh — scalar (one thread) routines on a symmetric tridiagonal matrix T = tridiag(e, d, e)
// used by the tridiagonal route of the per-zone transform (eig_tridiag.cu):
//   tql_eigenvalues : all eigenvalues by the implicit QL iteration (no vectors), ascending
//   twisted_vector  : the eigenvector for one eigenvalue by the double (twisted) factorisation of T - lambda I
// They replace, together with the Householder reduction, LAPACK dsyev of matoper_inc.F90:991-995
// (rrsqrt.F90:136).  Arrays are strided so that each thread of a warp can own one problem with
// conflict-free shared-memory accesses.  OAK_HD lets tools/test_tridiag_host.cpp compile the same code with
// g++ and check it on the CPU (no GPU in the build container).
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define OAK_HD __host__ __device__ __forceinline__
#else
#define OAK_HD inline
#endif

#define OAK_DBL_EPS 2.220446049250313e-16
#ifndef OAK_RCP_NEWTON
#define OAK_RCP_NEWTON 0
#endif

// 1/x for |x| in the normal range: hardware seed (MUFU.RCP64H, ~20 bits) + one cubically convergent step
// y0 (1 + r + r^2), r = 1 - x y0: 3 dependent FMAs instead of the ~25 instructions of the IEEE division.
// Relative error a few ulp, far below what the consumers need (residual test of the eigenvectors 1e-12,
// eigenvalues to eps |T|); -DOAK_RCP_NEWTON=1 adds a Newton step (measured: whole step 2.3 % slower, same
// parity results)
OAK_HD double oak_rcp(double x) {
#ifdef __CUDA_ARCH__
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double r = fma(-x, y, 1.);
  y = fma(y, fma(r, r, r), y);
#if OAK_RCP_NEWTON
  r = fma(-x, y, 1.);
  y = fma(y, r, y);
#endif
  return y;
#else
  return 1. / x;
#endif
}

OAK_HD double oak_rsqrt(double x) {
#ifdef __CUDA_ARCH__
  return rsqrt(x);
#else
  return 1. / sqrt(x);
#endif
}

// Implicit QL (EISPACK tql1 organisation) on d[0..n-1], e[0..n-2] (strided by s; e[n-1] is scratch).
// On return d holds the eigenvalues in ascending order.  tn = max(|d|,|e|) sets the absolute deflation
// threshold (eigenvalues are needed to eps*|T| only: everything downstream is a function of 1 + lambda).
// Returns the number of rotations, or -1 if an eigenvalue needed more than 60 iterations.
OAK_HD int tql_eigenvalues(int n, double *d, double *e, int s, double tn) {
  int rot = 0;
  const double abstol = 0.5 * OAK_DBL_EPS * tn;
  if (n > 0) e[(n - 1) * s] = 0.;
  for (int l = 0; l < n; l++) {
    int iter = 0;
    for (;;) {
      int m = l;
      for (; m < n - 1; m++) {
        const double em = fabs(e[m * s]);
        if (em <= abstol || em <= OAK_DBL_EPS * (fabs(d[m * s]) + fabs(d[(m + 1) * s]))) break;
      }
      if (m == l) break;
      if (++iter > 60) return -1;
      const double el = e[l * s], dl = d[l * s];
      double g = (d[(l + 1) * s] - dl) * 0.5 * oak_rcp(el);
      double r = sqrt(fma(g, g, 1.));
      g = d[m * s] - dl + el * oak_rcp(g + copysign(r, g));
      double sn = 1., cs = 1., p = 0.;
      int i = m - 1;
      bool under = false;
      for (; i >= l; i--) {
        rot++;
        const double ei = e[i * s];
        const double f = sn * ei, b = cs * ei;
        const double h = fma(f, f, g * g);
        if (h == 0.) {  // recover from underflow (tql1): deflate here and restart
          d[(i + 1) * s] -= p;
          e[m * s] = 0.;
          under = true;
          break;
        }
        const double ir = oak_rsqrt(h);
        e[(i + 1) * s] = h * ir;
        sn = f * ir;
        cs = g * ir;
        g = d[(i + 1) * s] - p;
        r = fma(d[i * s] - g, sn, 2. * cs * b);
        p = sn * r;
        d[(i + 1) * s] = g + p;
        g = fma(cs, r, -b);
      }
      if (under) continue;
      d[l * s] -= p;
      e[l * s] = g;
      e[m * s] = 0.;
    }
  }
  // ascending order (insertion sort: QL delivers them nearly sorted)
  for (int i = 1; i < n; i++) {
    const double v = d[i * s];
    int j = i - 1;
    while (j >= 0 && d[j * s] > v) { d[(j + 1) * s] = d[j * s]; j--; }
    d[(j + 1) * s] = v;
  }
  return rot;
}

// The same eigenvalues by the square-root-free QL variant of Pal, Walker and Kahan (the organisation of
// LAPACK dsterf): works on e_i^2, one reciprocal chain per rotation instead of an inverse square root plus
// the longer dependent chain of the plain QL step, i.e. about half the latency per rotation, which is
// what bounds k_tql (one thread per zone).  Same interface as tql_eigenvalues; e is overwritten by squares.
OAK_HD int pwk_eigenvalues(int n, double *d, double *e, int s, double tn) {
  int rot = 0;
  const double eps2 = OAK_DBL_EPS * OAK_DBL_EPS;
  const double abstol2 = 0.25 * eps2 * tn * tn;
  for (int i = 0; i < n - 1; i++) { const double ei = e[i * s]; e[i * s] = ei * ei; }
  if (n > 0) e[(n - 1) * s] = 0.;
  // [l, m] is the current unreduced block: e_m is negligible.  m is found by a scan only when a new block
  // starts; inside a block the sweep itself notices the off-diagonals it makes negligible (a scan per
  // iteration, as in dsterf, would cost as much as the sweep here).
  int m = -1;
  for (int l = 0; l < n; l++) {
    int iter = 0;
    for (;;) {
      if (m < l) {
        for (m = l; m < n - 1; m++) {
          const double em = e[m * s];
          if (em <= abstol2 || em <= eps2 * fabs(d[m * s] * d[(m + 1) * s])) break;
        }
      } else {
        const double el = e[l * s];
        if (el <= abstol2 || el <= eps2 * fabs(d[l * s] * d[(l + 1) * s])) m = l;
      }
      if (m == l) break;
      if (++iter > 60) return -1;
      const double rte = sqrt(e[l * s]);
      double p = d[l * s];
      double sigma = (d[(l + 1) * s] - p) * 0.5 * oak_rcp(rte);
      const double r0 = sqrt(fma(sigma, sigma, 1.));
      sigma = p - rte * oak_rcp(sigma + copysign(r0, sigma));
      double c = 1., sn = 0., gamma = d[m * s] - sigma;
      p = gamma * gamma;
      int msplit = m;          // lowest index whose new off-diagonal is negligible
      double dnext = 0., enew = 0.;  // d_{i+2} (final) and the new e_{i+1} of the previous trip
      for (int i = m - 1; i >= l; i--) {
        rot++;
        const double bb = e[i * s];
        const double r = p + bb;
        if (i != m - 1) { enew = sn * r; e[(i + 1) * s] = enew; }
        const double oldc = c;
        const double ir = oak_rcp(r);
        const double ip = oak_rcp(p);  // independent of ir: the two reciprocals overlap
        c = p * ir;
        sn = bb * ir;
        const double oldgam = gamma, alpha = d[i * s];
        gamma = fma(c, alpha - sigma, -sn * oldgam);
        const double dn = oldgam + (alpha - gamma);
        d[(i + 1) * s] = dn;
        if (i != m - 1 && (enew <= abstol2 || enew <= eps2 * fabs(dn * dnext))) msplit = i + 1;
        dnext = dn;
        p = (c != 0.) ? gamma * gamma * (r * ip) : oldc * bb;
      }
      e[l * s] = sn * p;
      d[l * s] = sigma + gamma;
      m = msplit;
    }
  }
  for (int i = 1; i < n; i++) {
    const double v = d[i * s];
    int j = i - 1;
    while (j >= 0 && d[j * s] > v) { d[(j + 1) * s] = d[j * s]; j--; }
    d[(j + 1) * s] = v;
  }
  return rot;
}

// Eigenvector of T for the eigenvalue lam by the twisted factorisation (Fernando; Parlett & Dhillon):
//   forward pivots  p_0 = d_0 - lam , p_{i+1} = (d_{i+1} - lam) - e_i^2 / p_i
//   backward pivots q_{n-1} = d_{n-1} - lam , q_i = (d_i - lam) - e_i^2 / q_{i+1}
//   gamma_i = p_i - e_i^2 / q_{i+1} ; r = argmin |gamma_i| ; z_r = 1 ,
//   z_{i-1} = -e_{i-1} z_i / p_{i-1} (i <= r) , z_{i+1} = -e_i z_i / q_{i+1} (i >= r)
// so that (T - lam) z = gamma_r e_r.  d, e are read with stride sd (broadcast arrays: sd = 1); the vector is
// written to w[i*sw] (unnormalised; also used as the only work array).  Returns |z|^2; gamma_r in *gam.
OAK_HD double twisted_vector(int n, const double *d, const double *e, int sd, double lam, double pivmin,
                             double *w, int sw, double *gam) {
  double p = d[0] - lam;
  for (int i = 0; i < n - 1; i++) {
    if (fabs(p) < pivmin) p = -pivmin;
    w[i * sw] = p;
    const double ei = e[i * sd];
    p = fma(-ei * ei, oak_rcp(p), d[(i + 1) * sd] - lam);
  }
  if (fabs(p) < pivmin) p = -pivmin;
  w[(n - 1) * sw] = p;
  // backward pivots, gamma, position of the twist
  double q = d[(n - 1) * sd] - lam;
  double best = fabs(p), gbest = p;
  int r = n - 1;
  for (int i = n - 2; i >= 0; i--) {
    if (fabs(q) < pivmin) q = -pivmin;
    const double ei = e[i * sd];
    const double t = ei * ei * oak_rcp(q);
    const double gm = w[i * sw] - t;
    q = (d[i * sd] - lam) - t;
    if (fabs(gm) < best) { best = fabs(gm); gbest = gm; r = i; }
  }
  // backward pivots again, kept for i > r (the forward ones are no longer needed there)
  q = d[(n - 1) * sd] - lam;
  for (int i = n - 1; i > r; i--) {
    if (fabs(q) < pivmin) q = -pivmin;
    w[i * sw] = q;
    const double ei = e[(i - 1) * sd];
    q = fma(-ei * ei, oak_rcp(q), d[(i - 1) * sd] - lam);
  }
  double zz = 1., z = 1.;
  for (int i = r; i > 0; i--) {
    z = -e[(i - 1) * sd] * z * oak_rcp(w[(i - 1) * sw]);
    w[(i - 1) * sw] = z;
    zz = fma(z, z, zz);
  }
  z = 1.;
  for (int i = r; i < n - 1; i++) {
    z = -e[i * sd] * z * oak_rcp(w[(i + 1) * sw]);
    w[(i + 1) * sw] = z;
    zz = fma(z, z, zz);
  }
  w[r * sw] = 1.;
  *gam = gbest;
  return zz;
}

// twisted_vector2 — the same vector with about a third of the dependent reciprocal chain (prepared for k_tvec,
// -DTVEC_TWISTED2=1; checked on the host against twisted_vector, tests/test_tridiag_host.py).
//   phase 1  forward pivots p_0..p_{h-1} and backward pivots q_{n-1}..q_h run in ONE loop (two independent chains);
//            their RECIPROCALS are stored in w (forward ones below h = n/2, backward ones from h up)
//   phase 2  both recurrences continue into the other half, where the stored reciprocals give gamma_i on the fly
//            (gamma_i = p_i - e_i^2 / q_{i+1} = q_i - e_{i-1}^2 / p_{i-1}); nothing is stored; r = argmin |gamma_i|
//   phase 3  the |r - h| reciprocals still missing on the far side of the twist are recomputed from the pivot
//            saved at the crossing
//   phase 4  z_{i-1} = -e_{i-1} z_i (1/p_{i-1}) upwards and z_{i+1} = -e_i z_i (1/q_{i+1}) downwards: one
//            dependent multiplication per step, the two directions interleaved
// Chain: n/2 + n/2 + |r-h| reciprocal steps (two chains wide in phases 1-2) instead of 2n + (n-r) + n.
OAK_HD double twisted_vector2(int n, const double *d, const double *e, int sd, double lam, double pivmin,
                              double *w, int sw, double *gam) {
  if (n == 1) { w[0] = 1.; *gam = d[0] - lam; return 1.; }
  const int h = n / 2;  // forward reciprocals live in w[0..h), backward ones in w[h..n)
  // ---- phase 1 ----
  double p = d[0] - lam, q = d[(n - 1) * sd] - lam;
  {
    int i = 0, j = n - 1;
    while (i < h || j >= h) {
      if (i < h) {
        if (fabs(p) < pivmin) p = -pivmin;
        const double ip = oak_rcp(p);
        w[i * sw] = ip;
        const double ei = e[i * sd];
        p = fma(-ei * ei, ip, d[(i + 1) * sd] - lam);  // p_{i+1}; i+1 <= h <= n-1
        i++;
      }
      if (j >= h) {
        if (fabs(q) < pivmin) q = -pivmin;
        const double iq = oak_rcp(q);
        w[j * sw] = iq;
        if (j > 0) {
          const double ej = e[(j - 1) * sd];
          q = fma(-ej * ej, iq, d[(j - 1) * sd] - lam);  // q_{j-1}
        }
        j--;
      }
    }
  }
  const double p_h = p, q_hm1 = q;  // p_h and q_{h-1} (q_{h-1} only meaningful for h >= 1, true for n >= 2)
  // ---- phase 2 ----
  double best = 1e300, gbest = 0.;
  int r = 0;
  {
    int i = h, j = h - 1;
    while (i < n || j >= 0) {
      if (i < n) {
        if (fabs(p) < pivmin) p = -pivmin;
        double gm = p;
        if (i < n - 1) {
          const double ei = e[i * sd];
          gm = fma(-ei * ei, w[(i + 1) * sw], p);          // gamma_i = p_i - e_i^2 / q_{i+1}
          p = fma(-ei * ei, oak_rcp(p), d[(i + 1) * sd] - lam);
        }
        if (fabs(gm) < best) { best = fabs(gm); gbest = gm; r = i; }
        i++;
      }
      if (j >= 0) {
        if (fabs(q) < pivmin) q = -pivmin;
        double gm = q;
        if (j > 0) {
          const double ej = e[(j - 1) * sd];
          gm = fma(-ej * ej, w[(j - 1) * sw], q);          // gamma_j = q_j - e_{j-1}^2 / p_{j-1}
          q = fma(-ej * ej, oak_rcp(q), d[(j - 1) * sd] - lam);
        }
        if (fabs(gm) < best) { best = fabs(gm); gbest = gm; r = j; }
        j--;
      }
    }
  }
  // ---- phase 3 ----
  if (r >= h) {           // forward reciprocals for h <= i < r
    p = p_h;
    for (int i = h; i < r; i++) {
      if (fabs(p) < pivmin) p = -pivmin;
      const double ip = oak_rcp(p);
      w[i * sw] = ip;
      const double ei = e[i * sd];
      p = fma(-ei * ei, ip, d[(i + 1) * sd] - lam);
    }
  } else {                // backward reciprocals for r < j <= h-1
    q = q_hm1;
    for (int j = h - 1; j > r; j--) {
      if (fabs(q) < pivmin) q = -pivmin;
      const double iq = oak_rcp(q);
      w[j * sw] = iq;
      const double ej = e[(j - 1) * sd];
      q = fma(-ej * ej, iq, d[(j - 1) * sd] - lam);
    }
  }
  // ---- phase 4 ----
  double zz = 1., zu = 1., zd = 1.;
  int iu = r, id = r;
  while (iu > 0 || id < n - 1) {
    if (iu > 0) {
      zu = -e[(iu - 1) * sd] * w[(iu - 1) * sw] * zu;
      w[(iu - 1) * sw] = zu;
      zz = fma(zu, zu, zz);
      iu--;
    }
    if (id < n - 1) {
      zd = -e[id * sd] * w[(id + 1) * sw] * zd;
      w[(id + 1) * sw] = zd;
      zz = fma(zd, zd, zz);
      id++;
    }
  }
  w[r * sw] = 1.;
  *gam = gbest;
  return zz;
}
