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
#ifndef PWK_SHORT_CHAIN
#define PWK_SHORT_CHAIN 0   // 1: measured equal on a B200 (14.35 vs 14.26 ms per 90 k zones): the lanes of a warp wait for each other, not for the chain
#endif
#ifndef PWK_FLAT
#define PWK_FLAT 1   // pwk_eigenvalues: 1 = one loop over sweeps instead of a loop nest over (l, sweeps at l), see there
#endif
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
// PF > 0 (k_tql with d, e in global memory): the entries a rotation reads, (d_i, e_i), are loaded PF rotations ahead
// into a register queue, so that the L1 / L2 latency of the loads stays off the dependent chain of the rotations; a
// sweep only writes entries above the ones it still has to read, so the queue never holds a stale value.  Same
// operations on the same values as PF = 0.
template <int PF>
OAK_HD int pwk_eigenvalues_t(int n, double *d, double *e, int s, double tn) {
  int rot = 0;
  const double eps2 = OAK_DBL_EPS * OAK_DBL_EPS;
  const double abstol2 = 0.25 * eps2 * tn * tn;
  for (int i = 0; i < n - 1; i++) { const double ei = e[i * s]; e[i * s] = ei * ei; }
  if (n > 0) e[(n - 1) * s] = 0.;
  // [l, mb] is the current unreduced block: e_mb is negligible.  A scan for the end of a block costs as much as a
  // sweep, and rescanning at every new l (as dsterf does) was 24 % of k_tql's instructions and 35 % of its stall
  // samples (ncu, profiles/r2_ncu_lines_k_tql.txt), so block ends are TRACKED instead: the sweep notices every
  // off-diagonal it makes negligible (all entries it rewrites are tested with their final neighbours; e_l is tested at
  // the head of the next trip), untouched entries cannot change status, hence after d_l has converged the block simply
  // goes on as [l+1, mb]; a split at msplit < mb leaves [msplit+1, mb] for later (mhi remembers that end).  Only when
  // more than two ends would have to be remembered (two splits in one sweep, a split inside a split) the bookkeeping
  // is declared dirty and the next block end is found by the scan again.  The sequence of (l, block end, shift) is
  // the one of the rescanning form, so the eigenvalues are bit-identical to it.
  int mb = -1, mhi = -1;
  bool dirty = false;
#define PWK_BLOCK_END()                                                                          \
  if (mb < l) {                                                                                  \
    if (!dirty && mhi >= l) mb = mhi;                                                            \
    else {                                                                                       \
      for (mb = l; mb < n - 1; mb++) {                                                           \
        const double em = e[mb * s];                                                             \
        if (em <= abstol2 || em <= eps2 * fabs(d[mb * s] * d[(mb + 1) * s])) break;              \
      }                                                                                          \
      mhi = mb; dirty = false;                                                                   \
    }                                                                                            \
  }
#if PWK_FLAT
  // ONE loop over sweeps, l advanced inside it: the 32 zones of a warp then only wait for each other's sweep lengths,
  // not for the zone that needs the most sweeps at every single l (the loop nest costs a warp sum_l max_lanes(sweeps at
  // l) ~ 2.1 n sweeps where a lane needs ~1.9 n: 7349 -> 4706 rotation trips per warp on C3-like spectra,
  // tools/sim_tql_divergence.py).  Measured on a B200 (k_tql per C3 step): rescanning form 69.4 ms (loop nest) / 75.0 ms
  // (flat: the desynchronised scans collide on the shared-memory banks); tracked block ends 47.1 ms (loop nest) /
  // 40.7 ms (flat) -> C3 249.2 -> 234.4 ms per step.
  int l = 0, iter = 0;
  for (;;) {
    for (; l < n; l++, iter = 0) {
      PWK_BLOCK_END()
      if (l == mb) continue;
      const double el = e[l * s];
      if (el <= abstol2 || el <= eps2 * fabs(d[l * s] * d[(l + 1) * s])) continue;   // d_l converged, the block goes on
      break;
    }
    if (l >= n) break;
    {
      if (++iter > 60) return -1;
#else
  for (int l = 0; l < n; l++) {
    int iter = 0;
    for (;;) {
      PWK_BLOCK_END()
      if (l == mb) break;
      {
        const double el = e[l * s];
        if (el <= abstol2 || el <= eps2 * fabs(d[l * s] * d[(l + 1) * s])) break;   // d_l converged, the block goes on
      }
      if (++iter > 60) return -1;
#endif
      // the loads of a sweep's set-up, issued together (one latency, not four, when d, e live in global memory)
      const double el_ = e[l * s], dl_ = d[l * s], dl1_ = d[(l + 1) * s], dmb_ = d[mb * s];
      double qd[PF > 0 ? PF : 1], qe[PF > 0 ? PF : 1];
      if (PF > 0) {
#pragma unroll
        for (int t = 0; t < PF; t++) {
          const int ip = mb - 1 - t;
          if (ip >= l) { qd[t] = d[ip * s]; qe[t] = e[ip * s]; }
        }
      }
      const double rte = sqrt(el_);
      double p = dl_;
      double sigma = (dl1_ - p) * 0.5 * oak_rcp(rte);
      const double r0 = sqrt(fma(sigma, sigma, 1.));
      sigma = p - rte * oak_rcp(sigma + copysign(r0, sigma));
      double c = 1., sn = 0., gamma = dmb_ - sigma;
      p = gamma * gamma;
      int msplit = mb, nsplit = 0;   // lowest index whose new off-diagonal is negligible, number of such indices
      double dnext = 0., enew = 0.;  // d_{i+2} (final) and the new e_{i+1} of the previous trip
      for (int i0 = mb - 1; i0 >= l; i0 -= (PF > 0 ? PF : 1)) {
#pragma unroll
       for (int t = 0; t < (PF > 0 ? PF : 1); t++) {
        const int i = i0 - t;
        if (i < l) break;
        rot++;
        double bb, alpha;
        if (PF > 0) {
          bb = qe[t]; alpha = qd[t];
          const int ipf = i - PF;
          if (ipf >= l) { qd[t] = d[ipf * s]; qe[t] = e[ipf * s]; }
        } else {
          bb = e[i * s]; alpha = d[i * s];
        }
        const double r = p + bb;
        if (i != mb - 1) { enew = sn * r; e[(i + 1) * s] = enew; }
        const double oldc = c;
#if PWK_SHORT_CHAIN
        // gamma = c (alpha - sigma) - sn oldgam = num / r with num = p (alpha - sigma) - bb oldgam, and the next
        // p = gamma^2 r / p = num^2 / (r p): the loop-carried chain p -> p' is add, mul, reciprocal, mul (7 dependent
        // operations with the 4 of the reciprocal) instead of 10; gamma, c, sn hang off a second reciprocal
        const double oldgam = gamma;
        const double num = fma(p, alpha - sigma, -(bb * oldgam));
        const double irp = oak_rcp(r * p);
        const double ir = oak_rcp(r);
        c = p * ir;
        sn = bb * ir;
        gamma = num * ir;
        const double dn = oldgam + (alpha - gamma);
        d[(i + 1) * s] = dn;
        if (i != mb - 1 && (enew <= abstol2 || enew <= eps2 * fabs(dn * dnext))) { msplit = i + 1; nsplit++; }
        dnext = dn;
        p = (p != 0.) ? (num * num) * irp : oldc * bb;
#else
        const double ir = oak_rcp(r);
        const double ip = oak_rcp(p);  // independent of ir: the two reciprocals overlap
        c = p * ir;
        sn = bb * ir;
        const double oldgam = gamma;
        gamma = fma(c, alpha - sigma, -sn * oldgam);
        const double dn = oldgam + (alpha - gamma);
        d[(i + 1) * s] = dn;
        if (i != mb - 1 && (enew <= abstol2 || enew <= eps2 * fabs(dn * dnext))) { msplit = i + 1; nsplit++; }
        dnext = dn;
        p = (c != 0.) ? gamma * gamma * (r * ip) : oldc * bb;
#endif
       }
      }
      e[l * s] = sn * p;
      d[l * s] = sigma + gamma;
      if (nsplit != 0) {
        if (nsplit > 1 || mhi != mb) dirty = true;   // more block ends than the two that are remembered
        mb = msplit;
      }
    }
  }
#undef PWK_BLOCK_END
  for (int i = 1; i < n; i++) {
    const double v = d[i * s];
    int j = i - 1;
    while (j >= 0 && d[j * s] > v) { d[(j + 1) * s] = d[j * s]; j--; }
    d[(j + 1) * s] = v;
  }
  return rot;
}

OAK_HD int pwk_eigenvalues(int n, double *d, double *e, int s, double tn) { return pwk_eigenvalues_t<0>(n, d, e, s, tn); }

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

// twisted_vector3 — the same vector WITHOUT a division in any recurrence (default in k_tvec since round 2; the
// pivot form above stays as the fallback for the cases below).  With delta_i = s (d_i - lam), eps_i = s e_i (s a
// power of two that makes |delta| < 1/2, |eps| < 1/8) the pivots are ratios of the Sturm polynomials,
//   forward   F_0 = 1, F_1 = delta_0,       F_{i+1} = delta_i F_i - eps_{i-1}^2 F_{i-1}     (F_i = P_{i-1}, p_i = F_{i+1}/F_i)
//   backward  B_{n-1} = 1, B_{n-2} = delta_{n-1}, B_{i-1} = delta_i B_i - eps_i^2 B_{i+1}   (B_i = Q_{i+1}, q_i = B_{i-1}/B_i)
// and gamma_i = det(T - lam) / (F_i B_i) (1/gamma_i is the i-th diagonal entry of the inverse), so that the twist is
// r = argmax |F_i B_i| and, with z_r = 1,
//   z_i = (F_i / F_r) prod_{j=i}^{r-1} (-eps_j)  (i < r) ,   z_i = (B_i / B_r) prod_{j=r}^{i-1} (-eps_j)  (i > r) ,
//   gamma_r = (F_{r+1} B_r - eps_r^2 F_r B_{r+1}) / (F_r B_r) / s .
// Every recurrence step is ONE dependent FMA (the second product only needs the value of two steps ago) instead of
// the reciprocal + FMA (~70 cycles) of a pivot step, and forward and backward chains run interleaved: the serial
// chain of a vector is ~2.5 n FMAs instead of ~3.5 n reciprocal steps.  Rounding: each step perturbs delta_i and
// eps_{i-1}^2 by a few ulp, exactly as a pivot step does.  The scaled polynomials only ever decay (|F_{i+1}| <=
// |F_i|/2 + |F_{i-1}|/64), so nothing overflows; if F_r B_r underflows (a long run of diagonal entries within
// ~1e-5 |T| of lam ...) the function reports failure (returns -1) and the caller uses the pivot form.
//   phase 1  F_i for i < h = n/2 and B_i for i >= h, stored in w
//   phase 2  both recurrences continue into the other half; |F_i B_i| on the fly, r = argmax
//   phase 3  the |r - h| values of the far side of the twist that are still missing are recomputed
//   phase 4  the products towards both ends, z, |z|^2
// scale = s (the same for every vector of the matrix: exp2(-(ilogb(tn) + 4)), tn = max(|d|, |e|)).
OAK_HD double twisted_vector3(int n, const double *ds /* s d_i */, const double *e2 /* (s e_i)^2 */,
                              const double *en /* -s e_i */, double sl /* s lam */, double inv_scale /* 1/s */,
                              double *w, int sw, double *gam) {
  // ds, e2, en are per-matrix arrays (unit stride, shared by all the vectors of the matrix); e2[n-1] = en[n-1] = 0
  if (n == 1) { w[0] = 1.; *gam = (ds[0] - sl) * inv_scale; return 1.; }
  const int h = n / 2;
  // ---- phase 1 ----  (uniform trip count: no divergence between the vectors of a warp)
  double f1 = 1., f0 = 0.;     // F_i, F_{i-1}
  double b1 = 1., b0 = 0.;     // B_j, B_{j+1}
  {
    double fe = 0., be = 0.;   // eps_{i-1}^2 F_{i-1} ; eps_j^2 B_{j+1}   (one step ahead of the chain)
    double *wf = w, *wb = w + (n - 1) * sw;
    const double *df = ds, *db = ds + (n - 1), *ef = e2, *eb = e2 + (n - 2);
#pragma unroll 4
    for (int i = 0; i < h; i++) {
      *wf = f1; wf += sw;
      const double fn = fma(*df - sl, f1, -fe);
      fe = *ef * f1; f0 = f1; f1 = fn; df++; ef++;
      *wb = b1; wb -= sw;
      const double bn = fma(*db - sl, b1, -be);
      be = *eb * b1; b0 = b1; b1 = bn; db--; eb--;
    }
    if (n & 1) {               // the backward half is one longer
      *wb = b1;
      const double bn = fma(*db - sl, b1, -be);
      b0 = b1; b1 = bn;
    }
  }
  // now f1 = F_h, f0 = F_{h-1}; b1 = B_{h-1}, b0 = B_h
  const double fh1 = f1, fh0 = f0, bh1 = b1, bh0 = b0;
  // ---- phase 2 ----
  double best = -1.;
  int r = 0;
  {
    const double *wf = w + h * sw, *wb = w + (h - 1) * sw;
    const double *df = ds + h, *db = ds + (h - 1), *ef = e2 + (h - 1), *eb = e2 + (h - 1);
#pragma unroll 4
    for (int t = 0; t < h; t++) {   // forward at i = h + t, backward at j = h - 1 - t
      const double kf = fabs(f1 * *wf), kb = fabs(b1 * *wb);
      if (kf > best) { best = kf; r = h + t; }
      if (kb > best) { best = kb; r = h - 1 - t; }
      const double fn = fma(*df - sl, f1, -(*ef * f0));
      f0 = f1; f1 = fn; wf += sw; df++; ef++;
      const double bn = fma(*db - sl, b1, -(*eb * b0));
      b0 = b1; b1 = bn; wb -= sw; db--; eb--;
    }
    if (n & 1) {                    // forward at i = n - 1
      const double kf = fabs(f1 * *wf);
      if (kf > best) { best = kf; r = n - 1; }
    }
  }
  if (!(best > 1e-250)) return -1.;   // also keeps det = gamma_r F_r B_r (~1e-16 smaller) out of the denormals
  // ---- phase 3 ----  one loop for both directions: x1, x0 = the recurrence pair at idx, moving by step towards r
  double fr, fr1, br, br1;
  {
    const bool up = r >= h;
    const int step = up ? 1 : -1, cnt = up ? r - h : h - 1 - r;
    int idx = up ? h : h - 1;
    double x1 = up ? fh1 : bh1, x0 = up ? fh0 : bh0;
    const double *eo = e2 + (up ? -1 : 0);   // forward steps use eps_{idx-1}^2, backward ones eps_idx^2
    for (int t = 0; t < cnt; t++) {
      w[idx * sw] = x1;
      const double xn = fma(ds[idx] - sl, x1, -(eo[idx] * x0));
      x0 = x1; x1 = xn; idx += step;
    }
    // idx == r : x1 = F_r (B_r), x0 = F_{r-1} (B_{r+1})
    if (up) {
      fr = x1;
      fr1 = fma(ds[r] - sl, x1, -((r >= 1 ? e2[r - 1] : 0.) * x0));
      br = w[r * sw];
      br1 = r + 1 < n ? w[(r + 1) * sw] : 0.;
    } else {
      br = x1; br1 = x0;
      fr = w[r * sw];
      fr1 = fma(ds[r] - sl, fr, -((r >= 1 ? e2[r - 1] * w[(r - 1) * sw] : 0.)));
    }
  }
  const double invF = oak_rcp(fr), invB = oak_rcp(br);
  {
    const double D = fma(fr1, br, -(e2[r] * fr) * br1);   // e2[n-1] = 0
    *gam = D * invF * invB * inv_scale;
  }
  // ---- phase 4 ----
  double zz = 1.;
  {
    double pu = invF, pd = invB;
    const int nu = r, nd = n - 1 - r, nt = nu > nd ? nu : nd;
    double *wu = w + (r - 1) * sw, *wd = w + (r + 1) * sw;
    const double *eu = en + (r - 1), *ed = en + r;
    for (int t = 0; t < nt; t++) {
      if (t < nu) {
        pu *= *eu;
        const double z = *wu * pu;
        *wu = z;
        zz = fma(z, z, zz);
      }
      if (t < nd) {
        pd *= *ed;
        const double z = *wd * pd;
        *wd = z;
        zz = fma(z, z, zz);
      }
      wu -= sw; eu--; wd += sw; ed++;
    }
  }
  w[r * sw] = 1.;
  return zz;
}
