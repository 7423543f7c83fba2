// eig_common.cuh — rotation parameters shared by the two eig kernels.
#pragma once

// rotations whose cosine is below this leave the columns unchanged to fp64 rounding
#define JACOBI_SKIP 1e-13f

// One-sided Jacobi rotation for the column pair (x,y) with a = |x|^2, b = |y|^2, g = x.y:
//   x' = c x - s y ,  y' = s x + c y   with  x'.y' = 0.
// t = tan(theta) is the smaller root of t^2 + 2 zeta t - 1 = 0, zeta = (b-a)/(2g), written without
// the division by g:  t = 2g / (d + sign(d) sqrt(d^2 + 4 g^2)),  d = b - a.
// Returns |cos(x,y)| (float is enough: it only steers skipping and the convergence test).
__device__ __forceinline__ float jacobi_params(double a, double b, double g, double &c, double &s, double &t) {
  const float cosang = fabsf((float)g) * rsqrtf((float)a) * rsqrtf((float)b);
  const double d = b - a, g2 = g + g;
  const double h = sqrt(fma(d, d, g2 * g2));
  t = g2 / (d + copysign(h, d));
  c = rsqrt(fma(t, t, 1.));
  s = t * c;
  return cosang;
}

__device__ __forceinline__ float jacobi_params(double a, double b, double g, double &c, double &s) {
  double t;
  return jacobi_params(a, b, g, c, s, t);
}

// Stopping rule.  After a sweep whose rotated pairs had cosines <= mx and tangents |t| <= mt, the
// remaining non-orthogonality is bounded by about mx * min(1, mt): for well separated singular values
// t ~ cos / relative gap is tiny and the convergence is quadratic; inside clusters (rank-deficient G:
// many sigma = 1) rotations have large angles, re-mix the cosines they touch and the convergence is
// only linear, so the sweep maximum itself must fall below the tolerance.
__device__ __forceinline__ bool jacobi_converged(float mx, float mt, float tol) {
  return mx * fminf(1.f, mt) < tol;
}
