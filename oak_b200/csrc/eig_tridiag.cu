// eig_tridiag.cu — the per-zone transform (ampl, T) from (G, c) by the tridiagonal route
// (option "eig_kernel" = 4; NP <= 64).  Replaces dsyev of matoper_inc.F90:991-995 as called by
// analysisIncrement (rrsqrt.F90:136) with ~7 N^3 flops instead of the ~25 N^3 of the Jacobi kernel:
//
//   k_tridiag  G = Q T Q^T, Householder reduction; one CTA per zone, thread t owns row t in registers
//   k_tql      eigenvalues of T by implicit QL; one THREAD per zone (the iteration is a scalar chain)
//   k_tvec     eigenvectors of T by the twisted factorisation (thread j owns eigenpair j), Gram-Schmidt
//              inside groups of close eigenvalues, U = Q W (thread j owns column j in registers), then
//                  (I+G)^-1/2 = I - Y Y^T ,  Y = U diag((1 - (1+lambda_j)^-1/2)^1/2)
//                  ampl = c - U diag(lambda/(1+lambda)) U^T c                       rrsqrt.F90:137-142
//                  v ~ (I+G)^1/2 1 = 1 + U diag((1+lambda)^1/2 - 1) U^T 1           rrsqrt.F90:176-178
//                  T = (I+G)^-1/2 Omega , Omega = RotateVector(w, v) in closed form  rrsqrt.F90:737-744
//              (same closed form of Omega as eig_simple.cu).
//
// The "I + sum_j (f(lambda_j) - f(0)) u_j u_j^T" form makes eigenvectors of (numerically) zero eigenvalues
// unnecessary: the null space of rank-deficient G (few observations, padding) is never computed.
// lambda <- max(lambda, 0) as rrsqrt.F90:137.
//
// Robustness: the vectors of T come from independent factorisations, so their orthogonality is
// residual / gap (residual ~ eps |T|).  Neighbouring pairs whose bound |g| (res_j + res_j-1) / gap exceeds
// TRI_ORTHTOL are orthogonalised explicitly (harmless for a matrix
// function: mixing inside a tight group changes f(A) by f' eps |T| only).  Zones where that is not enough
// (vectors of a group nearly parallel, groups larger than TRI_MAXGROUP, residual test failed, QL not
// converged) are flagged and recomputed by the Jacobi kernel (eig_fast.cu), which has no such cases.
#include <algorithm>
#include <cstdlib>
#include "common.cuh"
#include "tridiag_math.cuh"

#ifndef TRI_MINB
#define TRI_MINB 5   // CTAs per SM requested from the compiler for k_tridiag(_tile) (register cap 204: 8.24 -> 7.52 ms per 90 k zones)
#endif
#ifndef TRI_TILE
#define TRI_TILE 1   // 1 = k_tridiag_tile (8 x 8 cyclic register tiles), 0 = k_tridiag (row per thread)
#endif
#ifndef TRI_WARP
#define TRI_WARP 1   // NP = 64: 1 = k_tridiag_warp (one warp per zone, lower block triangle), 0 = k_tridiag_tile
#endif
#ifndef EIG_HALVES
#define EIG_HALVES 1   // 2: a batch goes through tridiag / QL / eigenvectors in two software-pipelined halves
#endif
#ifndef TVEC_DOT8
#define TVEC_DOT8 0     // 1: 8 instead of 4 accumulation chains in the dot products of the back-transformation and 4 partial sums in the
                        // N-long sums behind it: measured equal (k_tvec 75.9 -> 76.4 ms per C3 step), the chains are not what bounds it
#endif
#ifndef TQL_GLOBAL
#define TQL_GLOBAL 0   // k_tql: 1 = d, e of the warp's 32 zones transposed into a GLOBAL scratch array (L1 / L2 resident) and read
                       // through a register prefetch queue, no shared memory; 0 = transposed into 33 KB of shared memory per warp
#endif
#ifndef TQL_PF
#define TQL_PF 6       // TQL_GLOBAL: rotations of read-ahead
#endif
#ifndef TQL_PWK
#define TQL_PWK 1    // k_tql: 1 = square-root-free QL (Pal-Walker-Kahan), 0 = plain implicit QL
#endif
#ifndef TQL_LOCAL
#define TQL_LOCAL 0    // k_tql: 1 = d, e in per-thread local memory (L1) instead of 33 KB of shared memory per warp, so
#endif                 //    that its CTAs fit next to any other kernel's - prepared, emulation-checked, not yet measured
#ifndef TVEC_TWISTED2
#define TVEC_TWISTED2 0   // 1: twisted_vector2 (interleaved pivot recurrences, stored reciprocals) - prepared, host-tested,
#endif                    //    not yet measured on the GPU
#ifndef TVEC_PRODUCT
#define TVEC_PRODUCT 1    // 1: twisted_vector3 (division-free Sturm-product recurrences), pivot form as fallback
#endif
#ifndef TVEC_MMA_T
#define TVEC_MMA_T 1      // NP = 64: the product M = I - Y Y^T on mma.m8n8k4 tiles (36 lower tiles, mirrored on store)
#endif
#ifndef TVEC_MINB
#define TVEC_MINB 1
#endif

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int BW = 16;            // live-range granularity of the unrolled row/column loops
#define TRI_NULL 1e-13            // |(1+lambda)^-1/2 - 1| below this: eigenpair contributes nothing
#define TRI_ORTHTOL 1e-11         // accepted loss of orthogonality (times |g|) between neighbouring eigenvectors
#define TRI_RESTOL 1e-12
#define TRI_MAXGROUP 6
#define TRI_MINREM 0.03

// workspace per zone: d[NP] e[NP] tau[NP] lam[NP]
__device__ __forceinline__ double *ws_zone(double *ws, int NP, int zl) { return ws + (int64_t)zl * 4 * NP; }

template <int NW>
__device__ __forceinline__ double block_sum(double x, double *sred, int warp, int lane) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(FULL, x, o);
  if (NW == 1) return x;
  if (lane == 0) sred[warp] = x;
  __syncthreads();
  double r = sred[0];
#pragma unroll
  for (int w = 1; w < NW; w++) r += sred[w];
  return r;
}

// ---------------------------------------------------------------------------------------------------
// k_tridiag : thread t owns row t of the (symmetric, fully stored) matrix in registers.
//
// One Householder step per loop trip, organised so that a step has ONE block reduction and TWO barriers:
// the matrix-vector product of step k+1 is accumulated inside the rank-2 update loop of step k.
// State entering step k (column k is being reduced):
//   x_t = A[t][k] (t > k), y_t = sum_j A[t][j] x_j (the product with the UNSCALED column, so it does not
//   wait for the norm), row k+1 of A published in shared memory (sx2).
// With beta = -sign(alpha)|x|, v = (x - beta e_{k+1}) scale, v_{k+1} = 1:
//   A v = scale (y - beta A[:,k+1]) ,  v^T A v = scale^2 (x^T y - 2 beta y_{k+1} + beta^2 A[k+1][k+1]),
// so sigma^2 = sum x_t^2 and x^T y are reduced together; p = tau A v, w = p - (tau/2)(p^T v) v follow
// without another reduction, and the next column x'_t = A[t][k+1] - v_t w_{k+1} - w_t is known BEFORE the
// update A -= v w^T + w v^T, which therefore accumulates y' = A_new x' on the fly.
// ---------------------------------------------------------------------------------------------------
template <int NP, int OFF>
struct TriSteps {
  // steps k = OFF .. min(OFF+BW, N-2)-1; columns < OFF are finished, so every loop over the row held in
  // registers runs over the static range [OFF, NP)
  static __device__ __forceinline__ void run(double (&a)[NP], double &x, double &y, int N, int t, int warp,
                                             int lane, double *sx2, double2 *svw, double *sxn, double *sred,
                                             double *Vz, double *wd, double *we, double *wtau) {
    constexpr int NW = NP / 32;
    const int kend = min(OFF + BW, N - 2);
    for (int k = OFF; k < kend; k++) {
      double s1 = (t > k + 1) ? x * x : 0.;
      double s2 = (t > k) ? x * y : 0.;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(FULL, s1, o);
        s2 += __shfl_xor_sync(FULL, s2, o);
      }
      double *sr = sred + 8 * (k & 1);
      if (lane == 0) { sr[warp] = s1; sr[2 + warp] = s2; }
      if (t == k + 1) { sr[6] = x; sr[7] = y; }
      __syncthreads();
      if (NW == 2) { s1 = sr[0] + sr[1]; s2 = sr[2] + sr[3]; }
      const double alpha = sr[6], yk1 = sr[7];
      const double c1 = sx2[t];       // A[t][k+1]
      const double akk = sx2[k + 1];  // A[k+1][k+1]
      double tau = 0., beta = alpha, scale = 0.;
      if (s1 != 0.) {
        // beta = -sign(alpha)|x| ; tau = (beta-alpha)/beta = 1 + |alpha|/|x| ; 1/(alpha-beta) = sign(alpha)/(|alpha|+|x|)
        const double n2 = fma(alpha, alpha, s1);
        const double inrm = rsqrt(n2), nrm = n2 * inrm, aa = fabs(alpha);
        beta = -copysign(nrm, alpha);
        tau = fma(aa, inrm, 1.);
        scale = copysign(oak_rcp(aa + nrm), alpha);
      }
      const double ts = tau * scale;
      const double vAv = scale * scale * fma(beta, fma(beta, akk, -2. * yk1), s2);
      const double hpv = 0.5 * tau * tau * vAv;  // (tau/2) p^T v
      const double vt = (t > k + 1) ? x * scale : (t == k + 1 ? 1. : 0.);
      const double pt = (t > k) ? ts * fma(-beta, c1, y) : 0.;
      const double wt = fma(-hpv, vt, pt);
      const double wk1 = fma(ts, fma(-beta, akk, yk1), -hpv);  // w_{k+1}
      const double xn = (t > k + 1) ? c1 - fma(vt, wk1, wt) : 0.;  // new A[t][k+1]
      svw[t] = make_double2(vt, wt);
      sxn[t] = xn;
      Vz[k * NP + t] = vt;
      if (t == k + 1) { we[k] = beta; wtau[k] = tau; wd[k + 1] = fma(-2., wk1, akk); }
      __syncthreads();
      double y0 = 0., y1 = 0., y2 = 0., y3 = 0., y4 = 0., y5 = 0., y6 = 0., y7 = 0.;
      if (32 * warp + 31 > k + 1) {  // a warp whose rows are all finished only takes part in the barriers
#pragma unroll
        for (int j = OFF; j < NP; j += 8) {
#pragma unroll
          for (int jj = 0; jj < 8; jj++) {
            const double2 q = svw[j + jj];
            a[j + jj] = fma(-vt, q.y, fma(-wt, q.x, a[j + jj]));
          }
          const double2 n01 = *reinterpret_cast<const double2 *>(sxn + j);
          const double2 n23 = *reinterpret_cast<const double2 *>(sxn + j + 2);
          const double2 n45 = *reinterpret_cast<const double2 *>(sxn + j + 4);
          const double2 n67 = *reinterpret_cast<const double2 *>(sxn + j + 6);
          y0 = fma(a[j], n01.x, y0);
          y1 = fma(a[j + 1], n01.y, y1);
          y2 = fma(a[j + 2], n23.x, y2);
          y3 = fma(a[j + 3], n23.y, y3);
          y4 = fma(a[j + 4], n45.x, y4);
          y5 = fma(a[j + 5], n45.y, y5);
          y6 = fma(a[j + 6], n67.x, y6);
          y7 = fma(a[j + 7], n67.y, y7);
        }
        if (t == k + 2) {
#pragma unroll
          for (int j = OFF; j < NP; j += 2) *reinterpret_cast<double2 *>(sx2 + j) = make_double2(a[j], a[j + 1]);
        }
      }
      x = xn;
      y = ((y0 + y1) + (y2 + y3)) + ((y4 + y5) + (y6 + y7));
    }
    if constexpr (OFF + BW < NP) {
      if (N - 2 > OFF + BW) TriSteps<NP, OFF + BW>::run(a, x, y, N, t, warp, lane, sx2, svw, sxn, sred, Vz, wd, we, wtau);
    }
  }
};

template <int NP>
__global__ void __launch_bounds__(NP, TRI_MINB) k_tridiag(int N, const int32_t *__restrict__ mloc,
                                                 const double *__restrict__ G, double *__restrict__ V,
                                                 double *__restrict__ ws) {
  __shared__ __align__(16) double sx2[NP];
  __shared__ __align__(16) double sxn[NP];
  __shared__ __align__(16) double2 svw[NP];
  __shared__ double sred[16];
  const int zl = blockIdx.x;
  if (mloc[zl] == 0) return;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const double *Gz = G + (int64_t)zl * NP * NP;
  double *Vz = V + (int64_t)zl * NP * NP;
  double *wz = ws_zone(ws, NP, zl);
  double *wd = wz, *we = wz + NP, *wtau = wz + 2 * NP;

  double a[NP];
#pragma unroll
  for (int j = 0; j < NP; j++) a[j] = Gz[j * NP + t];  // G is symmetric: row t read as column t (coalesced)
  double x = (t >= 1) ? a[0] : 0.;
  sxn[t] = x;
  if (t == 1) {
#pragma unroll
    for (int j = 0; j < NP; j++) sx2[j] = a[j];
  }
  if (t == 0) wd[0] = a[0];
  if (t >= N) { wd[t] = 0.; we[t] = 0.; }
  wtau[t] = 0.;
  __syncthreads();
  double y;
  {
    double y0 = 0., y1 = 0., y2 = 0., y3 = 0.;
#pragma unroll
    for (int j = 0; j < NP; j += 4) {
      y0 = fma(a[j], sxn[j], y0);
      y1 = fma(a[j + 1], sxn[j + 1], y1);
      y2 = fma(a[j + 2], sxn[j + 2], y2);
      y3 = fma(a[j + 3], sxn[j + 3], y3);
    }
    y = (y0 + y1) + (y2 + y3);
  }
  TriSteps<NP, 0>::run(a, x, y, N, t, warp, lane, sx2, svw, sxn, sred, Vz, wd, we, wtau);
  // x is now column N-2 (its only entry below the diagonal is e_{N-2}); sx2 holds row N-1
  __syncthreads();
  if (t == N - 1) { we[N - 2] = x; we[N - 1] = 0.; wd[N - 1] = sx2[N - 1]; }
}

// ---------------------------------------------------------------------------------------------------
// k_tridiag_tile : the same fused Householder step on a 2-D cyclic register tiling.
//
// The row-per-thread kernel above is bound by the shared-memory return path, not by the fp64 pipe: every
// DFMA of the rank-2 update needs its own 8 bytes of (v_j, w_j, x'_j) from shared memory (a broadcast
// LDS.128 still moves 512 B per warp), i.e. 1 FMA per double loaded against the ~4 the SM can sustain.
// Here 64 threads form an 8 x 8 grid; thread (rg, cg) holds A[8a + rg][8b + cg], a, b < NP/8 (cyclic in both
// directions), so a loaded column triple serves NP/8 rows and a loaded row pair NP/8 columns (8 FMA per
// double at NP = 64).  The cyclic distribution also makes the finished rows AND columns drop out of every
// thread's static loop ranges every 8 steps (exact triangular work instead of 1.25 x full rows), and the row
// that has to be published next (k+2) sits in a register slot that is static within such a block.
// Row sums are combined over the 8 column-group lanes by a transpose-reduction (7 shuffles) that leaves the
// complete sum of row 8 cg + rg in lane cg: that thread is the row's owner for the scalar part of the step.
// ---------------------------------------------------------------------------------------------------
template <int RA>
__device__ __forceinline__ double transpose_reduce8(double (&yp)[RA], int cg) {
  // lanes cg = 0..7 (consecutive) each hold RA partial sums; returns in lane cg the total of slot cg (RA = 8)
  // or of slot cg & 3 (RA = 4: lanes cg and cg ^ 4 both get it)
  double v4[4];
  if constexpr (RA == 8) {
    const bool up = cg & 4;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const double send = up ? yp[i] : yp[i + 4];
      const double keep = up ? yp[i + 4] : yp[i];
      v4[i] = keep + __shfl_xor_sync(FULL, send, 4);
    }
  } else {
#pragma unroll
    for (int i = 0; i < 4; i++) v4[i] = yp[i] + __shfl_xor_sync(FULL, yp[i], 4);
  }
  double v2[2];
  {
    const bool up = cg & 2;
#pragma unroll
    for (int i = 0; i < 2; i++) {
      const double send = up ? v4[i] : v4[i + 2];
      const double keep = up ? v4[i + 2] : v4[i];
      v2[i] = keep + __shfl_xor_sync(FULL, send, 2);
    }
  }
  const bool up = cg & 1;
  const double send = up ? v2[0] : v2[1];
  const double keep = up ? v2[1] : v2[0];
  return keep + __shfl_xor_sync(FULL, send, 1);
}

template <int NP, int KB>
struct TileSteps {
  static constexpr int RA = NP / 8;
  // steps k = 8 KB .. min(8 KB + 8, N-2) - 1 ; row / column slots < KB are finished
  static __device__ __forceinline__ void run(double (&a)[NP / 8][NP / 8], double &x, double &y, int N, int tid,
                                             int rg, int cg, int t, bool owner, double *sx2, double2 *svw,
                                             double *sxn, double *sred, double *Vz, double *wd, double *we,
                                             double *wtau) {
    const int warp = tid >> 5, lane = tid & 31;
    const int kend = min(8 * KB + 8, N - 2);
    for (int k = 8 * KB; k < kend; k++) {
      double s1 = (owner && t > k + 1) ? x * x : 0.;
      double s2 = (owner && t > k) ? x * y : 0.;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(FULL, s1, o);
        s2 += __shfl_xor_sync(FULL, s2, o);
      }
      double *sr = sred + 8 * (k & 1);
      if (lane == 0) { sr[warp] = s1; sr[2 + warp] = s2; }
      if (owner && t == k + 1) { sr[6] = x; sr[7] = y; }
      __syncthreads();
      s1 = sr[0] + sr[1]; s2 = sr[2] + sr[3];
      const double alpha = sr[6], yk1 = sr[7];
      const double c1 = sx2[t];       // A[t][k+1]
      const double akk = sx2[k + 1];  // A[k+1][k+1]
      double tau = 0., beta = alpha, scale = 0.;
      if (s1 != 0.) {
        const double n2 = fma(alpha, alpha, s1);
        const double inrm = rsqrt(n2), nrm = n2 * inrm, aa = fabs(alpha);
        beta = -copysign(nrm, alpha);
        tau = fma(aa, inrm, 1.);
        scale = copysign(oak_rcp(aa + nrm), alpha);
      }
      const double ts = tau * scale;
      const double vAv = scale * scale * fma(beta, fma(beta, akk, -2. * yk1), s2);
      const double hpv = 0.5 * tau * tau * vAv;  // (tau/2) p^T v
      const double vt = (t > k + 1) ? x * scale : (t == k + 1 ? 1. : 0.);
      const double pt = (t > k) ? ts * fma(-beta, c1, y) : 0.;
      const double wt = fma(-hpv, vt, pt);
      const double wk1 = fma(ts, fma(-beta, akk, yk1), -hpv);  // w_{k+1}
      const double xn = (t > k + 1) ? c1 - fma(vt, wk1, wt) : 0.;  // new A[t][k+1]
      if (owner) {
        svw[t] = make_double2(vt, wt);
        sxn[t] = xn;
        if (t == k + 1) { we[k] = beta; wtau[k] = tau; wd[k + 1] = fma(-2., wk1, akk); }
      }
      __syncthreads();
      if (tid < NP) Vz[k * NP + tid] = svw[tid].x;  // reflector k, coalesced
      double yp[RA];
#pragma unroll
      for (int i = 0; i < RA; i++) yp[i] = 0.;
      {
        double2 cvw[RA];
        double cxn[RA];
#pragma unroll
        for (int b = KB; b < RA; b++) { cvw[b] = svw[8 * b + cg]; cxn[b] = sxn[8 * b + cg]; }
#pragma unroll
        for (int i = KB; i < RA; i++) {
          const double2 r = svw[8 * i + rg];
#pragma unroll
          for (int b = KB; b < RA; b++) {
            a[i][b] = fma(-r.x, cvw[b].y, fma(-r.y, cvw[b].x, a[i][b]));
            yp[i] = fma(a[i][b], cxn[b], yp[i]);
          }
        }
      }
      // publish row k+2 (slot (k+2)/8 is KB or KB+1, register index static either way)
      {
        const int pr = k + 2;
        if (rg == (pr & 7)) {
          if ((pr >> 3) == KB) {
#pragma unroll
            for (int b = KB; b < RA; b++) sx2[8 * b + cg] = a[KB][b];
          } else if constexpr (KB + 1 < RA) {
#pragma unroll
            for (int b = KB; b < RA; b++) sx2[8 * b + cg] = a[KB + 1][b];
          }
        }
      }
      const double ysum = transpose_reduce8<RA>(yp, cg);
      x = xn;
      y = ysum;
    }
    if constexpr (KB + 1 < NP / 8) {
      if (N - 2 > 8 * KB + 8) TileSteps<NP, KB + 1>::run(a, x, y, N, tid, rg, cg, t, owner, sx2, svw, sxn, sred, Vz, wd, we, wtau);
    }
  }
};

template <int NP>
__global__ void __launch_bounds__(64, TRI_MINB) k_tridiag_tile(int N, const int32_t *__restrict__ mloc,
                                                                const double *__restrict__ G,
                                                                double *__restrict__ V, double *__restrict__ ws) {
  constexpr int RA = NP / 8;
  __shared__ __align__(16) double sx2[NP];
  __shared__ __align__(16) double sxn[NP];
  __shared__ __align__(16) double2 svw[NP];
  __shared__ double sred[16];
  const int zl = blockIdx.x;
  if (mloc[zl] == 0) return;
  const int tid = threadIdx.x, rg = tid >> 3, cg = tid & 7;
  const bool owner = cg < RA;          // this thread owns row t = 8 cg + rg for the per-row scalars
  const int t = owner ? 8 * cg + rg : NP - 1;
  const double *Gz = G + (int64_t)zl * NP * NP;
  double *Vz = V + (int64_t)zl * NP * NP;
  double *wz = ws_zone(ws, NP, zl);
  double *wd = wz, *we = wz + NP, *wtau = wz + 2 * NP;

  double a[RA][RA];
#pragma unroll
  for (int i = 0; i < RA; i++)
#pragma unroll
    for (int b = 0; b < RA; b++) a[i][b] = Gz[(8 * i + rg) * NP + 8 * b + cg];
  double x = 0.;
  if (owner) {
    x = (t >= 1) ? Gz[t] : 0.;  // column 0 (= row 0: G is symmetric)
    sxn[t] = x;
  }
  if (tid < NP) {
    sx2[tid] = Gz[NP + tid];    // row 1
    if (tid >= N) { wd[tid] = 0.; we[tid] = 0.; }
    wtau[tid] = 0.;
    if (tid == 0) wd[0] = Gz[0];
  }
  __syncthreads();
  double y;
  {
    double yp[RA];
#pragma unroll
    for (int i = 0; i < RA; i++) {
      yp[i] = 0.;
#pragma unroll
      for (int b = 0; b < RA; b++) yp[i] = fma(a[i][b], sxn[8 * b + cg], yp[i]);
    }
    y = transpose_reduce8<RA>(yp, cg);
  }
  TileSteps<NP, 0>::run(a, x, y, N, tid, rg, cg, t, owner, sx2, svw, sxn, sred, Vz, wd, we, wtau);
  // x of row N-1's owner is now e_{N-2}; sx2 holds row N-1
  __syncthreads();
  if (owner && t == N - 1) { we[N - 2] = x; we[N - 1] = 0.; wd[N - 1] = sx2[N - 1]; }
}

#include "tridiag_warp.cuh"

// ---------------------------------------------------------------------------------------------------
// k_tql : one thread per zone; d, e transposed into shared memory with an odd stride
// ---------------------------------------------------------------------------------------------------
#if TQL_GLOBAL
// k_tql without shared memory (TQL_GLOBAL = 1; measured, OFF).  Hypothesis: the ~20 ms per C3 step of k_tql that the overlap
// with the other stream slots never hides come from the 33.8 KB of shared memory every resident warp holds for ~0.7 ms
// (3.5 - 5 warps per SM: half of the SM's shared memory, taken from k_gram_mma / k_tvec / k_apply beside it).  Here the
// transposed copy [i][lane] of the warp's 32 problems lives in a global scratch array (32 KB per warp, L2 / L1 resident;
// lanes at the same i read one 256-byte run), and pwk_eigenvalues_t<PF> loads (d_i, e_i) PF rotations ahead so the load
// latency is off the dependent chain.  Result on a B200: bit-identical, but the kernel is slower (56.5 against 40.1 ms
// serial per C3 step) and the exposed share is unchanged (step 243.8 against 233.9 ms): refuted, see DESIGN.md 3.3c.
template <int NP>
__global__ void __launch_bounds__(32) k_tql(int N, int nz, const int32_t *__restrict__ mloc,
                                             double *__restrict__ ws, int32_t *__restrict__ flags, double *__restrict__ scr) {
  const int lane = threadIdx.x;
  const int z0 = blockIdx.x * 32;
  double *gd = scr + (size_t)blockIdx.x * 2 * NP * 32, *ge = gd + NP * 32;
  for (int z = 0; z < 32; z++) {
    if (z0 + z >= nz) break;
    const double *wz = ws_zone(ws, NP, z0 + z);
    for (int i = lane; i < NP; i += 32) {
      gd[i * 32 + z] = wz[i];
      ge[i * 32 + z] = wz[NP + i];
    }
  }
  __syncwarp();
  const int zl = z0 + lane;
  const bool active = zl < nz && mloc[zl] != 0;
  int rot = 0;
  if (active) {
    double tn = 0.;
    for (int i = 0; i < N; i++) tn = fmax(tn, fmax(fabs(gd[i * 32 + lane]), fabs(ge[i * 32 + lane])));
    rot = pwk_eigenvalues_t<TQL_PF>(N, gd + lane, ge + lane, 32, tn);
    // the reciprocals of the iteration are not guarded against denormals: verify instead (NaN-safe)
    for (int i = 0; i < N; i++)
      if (!(fabs(gd[i * 32 + lane]) <= 4. * tn)) rot = -1;
  }
  if (zl < nz) flags[zl] = (rot < 0) ? 1 : 0;
  __syncwarp();
  for (int z = 0; z < 32; z++) {
    if (z0 + z >= nz) break;
    double *wz = ws_zone(ws, NP, z0 + z);
    for (int i = lane; i < NP; i += 32) wz[3 * NP + i] = gd[i * 32 + z];
  }
}
#elif TQL_LOCAL
template <int NP>
__global__ void __launch_bounds__(32) k_tql(int N, int nz, const int32_t *__restrict__ mloc,
                                             double *__restrict__ ws, int32_t *__restrict__ flags, double *) {
  const int zl = blockIdx.x * 32 + threadIdx.x;
  if (zl >= nz) return;
  int rot = 0;
  if (mloc[zl] != 0) {
    double d[NP], e[NP];  // dynamically indexed: local memory, interleaved over the lanes (one line per warp access)
    double *wz = ws_zone(ws, NP, zl);
    double tn = 0.;
    for (int i = 0; i < NP; i++) {
      d[i] = wz[i];
      e[i] = wz[NP + i];
      if (i < N) tn = fmax(tn, fmax(fabs(d[i]), fabs(e[i])));
    }
#if TQL_PWK
    rot = pwk_eigenvalues(N, d, e, 1, tn);
#else
    rot = tql_eigenvalues(N, d, e, 1, tn);
#endif
    for (int i = 0; i < N; i++)
      if (!(fabs(d[i]) <= 4. * tn)) rot = -1;
    for (int i = 0; i < NP; i++) wz[3 * NP + i] = d[i];
  }
  flags[zl] = (rot < 0) ? 1 : 0;
}
#else
template <int NP>
__global__ void __launch_bounds__(32) k_tql(int N, int nz, const int32_t *__restrict__ mloc,
                                             double *__restrict__ ws, int32_t *__restrict__ flags, double *) {
  constexpr int S = 33;
  __shared__ double sd[NP * S], se[NP * S];
  const int lane = threadIdx.x;
  const int z0 = blockIdx.x * 32;
  for (int z = 0; z < 32; z++) {
    if (z0 + z >= nz) break;
    const double *wz = ws_zone(ws, NP, z0 + z);
    for (int i = lane; i < NP; i += 32) {
      sd[i * S + z] = wz[i];
      se[i * S + z] = wz[NP + i];
    }
  }
  __syncwarp();
  const int zl = z0 + lane;
  const bool active = zl < nz && mloc[zl] != 0;
  int rot = 0;
  if (active) {
    double tn = 0.;
    for (int i = 0; i < N; i++) tn = fmax(tn, fmax(fabs(sd[i * S + lane]), fabs(se[i * S + lane])));
#if TQL_PWK
    rot = pwk_eigenvalues(N, sd + lane, se + lane, S, tn);
#else
    rot = tql_eigenvalues(N, sd + lane, se + lane, S, tn);
#endif
    // the reciprocals of the iteration are not guarded against denormals: verify instead (NaN-safe)
    for (int i = 0; i < N; i++)
      if (!(fabs(sd[i * S + lane]) <= 4. * tn)) rot = -1;
  }
  if (zl < nz) flags[zl] = (rot < 0) ? 1 : 0;
  __syncwarp();
  for (int z = 0; z < 32; z++) {
    if (z0 + z >= nz) break;
    double *wz = ws_zone(ws, NP, z0 + z);
    for (int i = lane; i < NP; i += 32) wz[3 * NP + i] = sd[i * S + z];
  }
}
#endif  // TQL_LOCAL

// ---------------------------------------------------------------------------------------------------
// k_tvec
// ---------------------------------------------------------------------------------------------------
template <int NP, int OFF>
struct BackSteps {
  // reflectors k = min(OFF+BW, N-2)-1 .. OFF applied to the column u held in registers; rows < OFF of
  // these reflectors are zero, so the loops run over the static range [OFF, NP)
  static __device__ __forceinline__ void run(double (&u)[NP], int N, const double *Vs, const double *stau) {
    if constexpr (OFF + BW < NP) {
      if (N - 2 > OFF + BW) BackSteps<NP, OFF + BW>::run(u, N, Vs, stau);
    }
    for (int k = min(OFF + BW, N - 2) - 1; k >= OFF; k--) {
      const double tau = stau[k];
      if (tau == 0.) continue;
      const double *v = Vs + k * NP;
      double p0 = 0., p1 = 0., p2 = 0., p3 = 0., p4 = 0., p5 = 0., p6 = 0., p7 = 0.;
#pragma unroll
      for (int i = OFF; i < NP; i += 8) {
        const double2 v01 = *reinterpret_cast<const double2 *>(v + i);
        const double2 v23 = *reinterpret_cast<const double2 *>(v + i + 2);
        const double2 v45 = *reinterpret_cast<const double2 *>(v + i + 4);
        const double2 v67 = *reinterpret_cast<const double2 *>(v + i + 6);
        p0 = fma(u[i], v01.x, p0);
        p1 = fma(u[i + 1], v01.y, p1);
        p2 = fma(u[i + 2], v23.x, p2);
        p3 = fma(u[i + 3], v23.y, p3);
        p4 = fma(u[i + 4], v45.x, p4);
        p5 = fma(u[i + 5], v45.y, p5);
        p6 = fma(u[i + 6], v67.x, p6);
        p7 = fma(u[i + 7], v67.y, p7);
      }
      const double s = -tau * (((p0 + p1) + (p2 + p3)) + ((p4 + p5) + (p6 + p7)));
#pragma unroll
      for (int i = OFF; i < NP; i += 2) {
        const double2 v01 = *reinterpret_cast<const double2 *>(v + i);
        u[i] = fma(s, v01.x, u[i]);
        u[i + 1] = fma(s, v01.y, u[i + 1]);
      }
    }
  }
};

// Back-transformation on 2 x (NP/2) register tiles: thread (cp, rh) = (tid>>1, tid&1) holds columns
// 2cp, 2cp+1 of U and the rows 4m + 2rh + {0,1}, m < NP/4.  A reflector entry read from shared memory then
// serves two columns and both the dot product and the update (4 FMA per double instead of 1: the
// column-per-thread version above is bound by the shared-memory return path); the two partial dot products
// are completed with the neighbouring lane.
template <int NP, int OFF>
struct BackTile {
  static __device__ __forceinline__ void run(double (&u)[2][NP / 2], int N, int rh, const double *Vs,
                                             const double *stau) {
    if constexpr (OFF + BW < NP) {
      if (N - 2 > OFF + BW) BackTile<NP, OFF + BW>::run(u, N, rh, Vs, stau);
    }
    constexpr int M0 = OFF / 4, M1 = NP / 4;
    for (int k = min(OFF + BW, N - 2) - 1; k >= OFF; k--) {
      const double tau = stau[k];
      if (tau == 0.) continue;
      const double *v = Vs + k * NP + 2 * rh;
      double2 vv[M1];
#pragma unroll
      for (int m = M0; m < M1; m++) vv[m] = *reinterpret_cast<const double2 *>(v + 4 * m);
#if TVEC_DOT8
      // eight accumulation chains instead of four: a dependent DFMA costs ~25 cycles on this part, so four chains of up
      // to 16 links (400 cycles per reflector) were longer than the 128 issue cycles of the 64 FMAs of the dot products
      double a0 = 0., a1 = 0., b0 = 0., b1 = 0., a2 = 0., a3 = 0., b2 = 0., b3 = 0.;
#pragma unroll
      for (int m = M0; m < M1; m++) {
        if ((m - M0) & 1) {
          a2 = fma(u[0][2 * m], vv[m].x, a2);
          a3 = fma(u[0][2 * m + 1], vv[m].y, a3);
          b2 = fma(u[1][2 * m], vv[m].x, b2);
          b3 = fma(u[1][2 * m + 1], vv[m].y, b3);
        } else {
          a0 = fma(u[0][2 * m], vv[m].x, a0);
          a1 = fma(u[0][2 * m + 1], vv[m].y, a1);
          b0 = fma(u[1][2 * m], vv[m].x, b0);
          b1 = fma(u[1][2 * m + 1], vv[m].y, b1);
        }
      }
      double d0 = (a0 + a1) + (a2 + a3), d1 = (b0 + b1) + (b2 + b3);
#else
      double a0 = 0., a1 = 0., b0 = 0., b1 = 0.;
#pragma unroll
      for (int m = M0; m < M1; m++) {
        a0 = fma(u[0][2 * m], vv[m].x, a0);
        a1 = fma(u[0][2 * m + 1], vv[m].y, a1);
        b0 = fma(u[1][2 * m], vv[m].x, b0);
        b1 = fma(u[1][2 * m + 1], vv[m].y, b1);
      }
      double d0 = a0 + a1, d1 = b0 + b1;
#endif
      d0 += __shfl_xor_sync(FULL, d0, 1);
      d1 += __shfl_xor_sync(FULL, d1, 1);
      const double s0 = -tau * d0, s1 = -tau * d1;
#pragma unroll
      for (int m = M0; m < M1; m++) {
        u[0][2 * m] = fma(s0, vv[m].x, u[0][2 * m]);
        u[0][2 * m + 1] = fma(s0, vv[m].y, u[0][2 * m + 1]);
        u[1][2 * m] = fma(s1, vv[m].x, u[1][2 * m]);
        u[1][2 * m + 1] = fma(s1, vv[m].y, u[1][2 * m + 1]);
      }
    }
  }
};


// Tile map of the symmetric 64 x 64 product M = I - Y Y^T on two warps (the map of k_gram_mma's two-warp variant):
// warp 0 owns the LL triangle and the HL rows 4, 5 (18 tiles), warp 1 the HH triangle and the HL rows 6, 7 (18)
template <int W>
struct MTileMap {
  static constexpr int HB = 4;
  static __host__ __device__ constexpr bool owns(int bi, int bj) {
    if (bj > bi) return false;
    const bool LL = bi < HB, HH = bj >= HB, HL = !LL && !HH;
    return W == 0 ? (LL || (HL && bi < HB + 2)) : (HH || (HL && bi >= HB + 2));
  }
  static __host__ __device__ constexpr int slot(int bi, int bj) {
    const bool LL = bi < HB, HH = bj >= HB;
    if (LL) return bi * (bi + 1) / 2 + bj;
    if (HH) return (bi - HB) * (bi - HB + 1) / 2 + (bj - HB);
    return 10 + ((bi - HB) & 1) * HB + bj;
  }
  static __host__ __device__ constexpr bool needs_row(int b) {
    for (int jj = 0; jj <= b; jj++) if (owns(b, jj)) return true;
    return false;
  }
  static __host__ __device__ constexpr bool needs_col(int b) {
    for (int ii = b; ii < 2 * HB; ii++) if (owns(ii, b)) return true;
    return false;
  }
};

// ---------------------------------------------------------------------------------------------------
// Fused apply (option "fuse_apply"): the zone's rows are updated by k_tvec itself from the FACTORED transform,
//     Sa_z = ((Sf_z - (Sf_z Y) Y^T) - a1 u_v^T) D - a2 u_w^T ,  a1 = Sf_z g1 hv, a2 = Sf_z g2 hw ,  xa_z = xf_z + Sf_z ampl
// (the rows of T = (M - g1 hv u_v^T) D - g2 hw u_w^T, M = I - Y Y^T, are never formed).  For a zone of nr rows
// that is 4 nr N^2 flops instead of 2 N^3 + 2 nr N^2, no 8 N^2-byte round trip of T through HBM and no k_apply
// launch; it pays while nr < N (water columns: 30 rows, N = 64).  Both products run on the fp64 tensor cores
// (mma.m8n8k4): per chunk of 8 rows, P = S Y (A = S chunk, B = Y^T read from the transposed store Yt) with one
// extra tile whose B columns are (g1 hv, g2 hw, ampl), then Q = P Y^T.  Row strides of NP + 4 doubles make every
// fragment load conflict-free.  The chunk is staged completely before its results are stored: Sa may alias Sf.
// ---------------------------------------------------------------------------------------------------
template <int NP>
struct FusedApply {
  static constexpr int NW = NP / 32, NB = NP / 8, LD = NP + 4, RC = 8, TPW = NB / NW;
  static constexpr int SMEM_DOUBLES = 2 * RC * LD + 3 * RC;  // sS, sP, per-row sums
  static_assert(RC == 8, "row index = idx & 7");

  static __device__ __forceinline__ void run(int N, int tid, const double *Yt /* [NP][LD] */, double *sS, double *sP,
                                             double *s_row, const double *s_g1h, const double *s_g2h,
                                             const double *s_amp, const double *suv, const double *suw, double dNN,
                                             int64_t i1, int nrow, const double *__restrict__ xf, const double *Sf,
                                             int64_t ldS, double *__restrict__ xa, double *Sa, int64_t ldSa) {
    const int lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    const int NK = (N + 3) & ~3;  // members >= N are zero in S, Y and P
    for (int r0 = 0; r0 < nrow; r0 += RC) {
      const int rc = min(RC, nrow - r0);
      // stage S[r][k] = Sf(i1 + r0 + r, k), zero padded
      for (int idx = tid; idx < RC * NP; idx += NP) {
        const int r = idx & (RC - 1), k = idx >> 3;
        sS[r * LD + k] = (r < rc && k < N) ? Sf[i1 + r0 + r + ldS * k] : 0.;
      }
      __syncthreads();
      // P = S Y (+ the three row sums), tiles jb of this warp
      {
        double p[TPW][2], e0 = 0., e1 = 0.;
#pragma unroll
        for (int b = 0; b < TPW; b++) p[b][0] = p[b][1] = 0.;
#pragma unroll 2
        for (int i0 = 0; i0 < NK; i0 += 4) {
          const double a = sS[g * LD + i0 + t];
#pragma unroll
          for (int b = 0; b < TPW; b++)
            oak_dmma_m8n8k4(p[b][0], p[b][1], a, Yt[(8 * (warp * TPW + b) + g) * LD + i0 + t]);
          if (warp == 0) {
            const double bx = g == 0 ? s_g1h[i0 + t] : (g == 1 ? s_g2h[i0 + t] : (g == 2 ? s_amp[i0 + t] : 0.));
            oak_dmma_m8n8k4(e0, e1, a, bx);
          }
        }
#pragma unroll
        for (int b = 0; b < TPW; b++)
          *reinterpret_cast<double2 *>(sP + g * LD + 8 * (warp * TPW + b) + 2 * t) = make_double2(p[b][0], p[b][1]);
        if (warp == 0) {
          if (t == 0) { s_row[g] = e0; s_row[RC + g] = e1; }   // a1, a2
          if (t == 1) s_row[2 * RC + g] = e0;                  // Sf . ampl
        }
      }
      __syncthreads();
      // Q = P Y^T, tiles kb of this warp, and the finished rows in place of S
      {
        double q[TPW][2];
#pragma unroll
        for (int b = 0; b < TPW; b++) q[b][0] = q[b][1] = 0.;
#pragma unroll 2
        for (int j0 = 0; j0 < NK; j0 += 4) {
          const double a = sP[g * LD + j0 + t];
#pragma unroll
          for (int b = 0; b < TPW; b++)
            oak_dmma_m8n8k4(q[b][0], q[b][1], a, Yt[(j0 + t) * LD + 8 * (warp * TPW + b) + g]);
        }
        const double a1 = s_row[g], a2 = s_row[RC + g];
#pragma unroll
        for (int b = 0; b < TPW; b++)
#pragma unroll
          for (int e = 0; e < 2; e++) {
            const int k = 8 * (warp * TPW + b) + 2 * t + e;
            double v = (sS[g * LD + k] - q[b][e]) - a1 * suv[k];
            if (k == N - 1) v *= dNN;
            sS[g * LD + k] = v - a2 * suw[k];
          }
      }
      __syncthreads();
      for (int idx = tid; idx < RC * NP; idx += NP) {
        const int r = idx & (RC - 1), k = idx >> 3;
        if (r < rc && k < N) Sa[i1 + r0 + r + ldSa * k] = sS[r * LD + k];
      }
      if (tid < rc) xa[i1 + r0 + tid] = xf[i1 + r0 + tid] + s_row[2 * RC + tid];
      __syncthreads();
    }
  }
};

// FUSE: see FusedApply above (the zone geometry and the state arrays in `aa` are only used then)

// PART 0: the whole kernel.  Option "tvec_split": PART 1 = eigenvectors of T only (twisted factorisations and the
// grouping: serial reciprocal chains, few registers, so 4 x the zones in flight of the full kernel), W to global
// memory; PART 2 = everything from U = Q W on, starting from that W.
template <int NP, bool FUSE, int PART>
__global__ void __launch_bounds__(NP, TVEC_MINB) k_tvec(int N, const int32_t *__restrict__ mloc,
                                              const double *__restrict__ ws, const double *__restrict__ cin,
                                              double *__restrict__ Tout, double *__restrict__ ampl_out,
                                              int32_t *__restrict__ flags, DevCounters *ctr, double orthtol,
                                              int maxgroup, const FusedApplyArgs aa, double *__restrict__ Wg) {
  constexpr int LDW = PART == 1 ? NP : NP + 1;  // part 1 builds W directly in global memory (no shared-memory matrix)
  constexpr int LDY = NP + 4;  // Yt rows double as mma fragments (fused apply, tensor-core M product): conflict-free at NP + 4
  constexpr int NW = NP / 32;
  constexpr int TR = 8, TC = NP / 8;          // output tile of a thread: TR rows x TC columns
  constexpr int TJ = NP / TC;                 // thread grid: (NP/TR) x TJ = NP threads
  extern __shared__ __align__(16) double sm[];
  // [NP][LDW] : W[i*LDW + j] = element i of vector j ; later V, then Y
  double *W = PART == 1 ? Wg + (int64_t)blockIdx.x * NP * NP : sm;
  double *sd = PART == 1 ? sm : sm + NP * LDY;
  double *se = sd + NP, *slam = se + NP, *stau = slam + NP, *sc = stau + NP;
  double *sa = sc + NP, *sb = sa + NP, *suv = sb + NP, *sdw = suv + NP, *suw = sdw + NP;
  double *sg1 = suw + NP, *sg2 = sg1 + NP, *sgj = sg2 + NP, *sres = sgj + NP;
  double *sds = sres + NP, *se2 = sds + NP, *sen = se2 + NP;   // s d_i, (s e_i)^2, -s e_i of the Sturm-product recurrences
  constexpr int NVEC = 17;                                       // NP-vectors of shared memory after the matrix
  __shared__ double sred[8];
  __shared__ unsigned sclose[NW];
  __shared__ int sflag;

  const int zl = blockIdx.x;
  const int ml = mloc[zl];
  if (ml == 0) { if (threadIdx.x == 0) flags[zl] = 0; return; }
  const int j = threadIdx.x, warp = j >> 5, lane = j & 31;
  const double *wz = ws + (int64_t)zl * 4 * NP;
  if constexpr (PART == 2) {
    if (flags[zl] != 0) return;  // part 1 handed the zone to the Jacobi kernel
    stau[j] = wz[2 * NP + j];
    const double lam2 = wz[3 * NP + j];
    slam[j] = lam2;
    sc[j] = cin[(int64_t)zl * NP + j];
    sgj[j] = (j < N) ? 1. - 1. / sqrt(1. + fmax(lam2, 0.)) : 0.;
    const double *Wz = Wg + (int64_t)zl * NP * NP;
    for (int i = 0; i < NP; i++) W[i * LDW + j] = Wz[i * NP + j];
    __syncthreads();
  } else {
  sd[j] = wz[j];
  se[j] = wz[NP + j];
  stau[j] = wz[2 * NP + j];
  slam[j] = wz[3 * NP + j];
  sc[j] = cin[(int64_t)zl * NP + j];
  if (j == 0) { sflag = flags[zl]; if (sflag) atomicAdd(&ctr->fb_reason[0], 1ull); }
  __syncthreads();

  // ---- eigenvector j of T ----
  // tn = max |d_i|, |e_i| : one entry per thread and a block-wide maximum (every thread scanning all of d, e was 6 % of
  // the kernel's instructions, ncu source page; a maximum does not depend on the order, so tn is the same number)
  double tn = (j < N) ? fmax(fabs(sd[j]), fabs(se[j])) : 0.;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) tn = fmax(tn, __shfl_xor_sync(FULL, tn, o));
  if constexpr (NW > 1) {
    if (lane == 0) sred[warp] = tn;
    __syncthreads();
#pragma unroll
    for (int w = 0; w < NW; w++) tn = fmax(tn, sred[w]);
    __syncthreads();   // sred is reused by the block sums below
  }
  // power-of-two scale of the Sturm-product recurrences: |s (d - lam)| < 1/2, |s e| < 1/8
  const double tscale = tn > 0. ? scalbn(1., -(ilogb(tn) + 4)) : 1.;
#if TVEC_PRODUCT
  {
    const double es = (j < N - 1) ? tscale * se[j] : 0.;
    sds[j] = tscale * sd[j]; se2[j] = es * es; sen[j] = -es;
  }
  __syncthreads();
#endif
  const double lam = (j < N) ? slam[j] : 0.;
  const double lamc = fmax(lam, 0.);
  const double sig = sqrt(1. + lamc);
  const double gj = (j < N) ? 1. - 1. / sig : 0.;  // = -g_j >= 0
  const bool live = gj >= TRI_NULL;
  bool bad = false;
  double res = 0.;  // |(T - lam) w| of the normalised vector
  for (int i = N; i < NP; i++) W[i * LDW + j] = 0.;
  if (live) {
    const double pivmin = tn * 1e-150;
    double gam;
#if TVEC_TWISTED2
#define OAK_TWISTED twisted_vector2
#else
#define OAK_TWISTED twisted_vector
#endif
#if TVEC_PRODUCT
    // division-free (Sturm-product) form; the pivot form only where the products underflow
#define OAK_TWISTED_CALL(lam_)                                                  \
    zz = twisted_vector3(N, sds, se2, sen, tscale * (lam_), 1. / tscale, W + j, LDW, &gam); \
    if (zz < 0.) { atomicAdd(&ctr->tw_fallback, 1ull); zz = OAK_TWISTED(N, sd, se, 1, lam_, pivmin, W + j, LDW, &gam); }
#else
#define OAK_TWISTED_CALL(lam_) zz = OAK_TWISTED(N, sd, se, 1, lam_, pivmin, W + j, LDW, &gam);
#endif
    double zz;
    OAK_TWISTED_CALL(lam)
    if (!(fabs(gam) <= TRI_RESTOL * tn * sqrt(zz)) || !(zz < 1e300)) {  // one Rayleigh-quotient correction, then give up
      const double lam2 = lam + gam / zz;
      OAK_TWISTED_CALL(lam2)
      if (!(fabs(gam) <= TRI_RESTOL * tn * sqrt(zz)) || !(zz < 1e300)) bad = true;
    }
    const double sc_ = rsqrt(zz);
    res = fabs(gam) * sc_;
    for (int i = 0; i < N; i++) W[i * LDW + j] *= sc_;
  } else {
    for (int i = 0; i < N; i++) W[i * LDW + j] = 0.;
  }
  sgj[j] = gj;
  sres[j] = res;
  __syncthreads();
  // ---- groups of close eigenvalues ----
  bool close = false;
  if (j >= 1 && j < N && live && sgj[j - 1] >= TRI_NULL)
    close = fmax(gj, sgj[j - 1]) * (res + sres[j - 1] + 4. * OAK_DBL_EPS * tn) > orthtol * (lam - slam[j - 1]);
  const unsigned cm = __ballot_sync(FULL, close);
  if (lane == 0) sclose[warp] = cm;
  if (bad) { sflag = 1; atomicAdd(&ctr->fb_reason[1], 1ull); }
  __syncthreads();
  bool anyclose = false;
#pragma unroll
  for (int w = 0; w < NW; w++) anyclose |= (sclose[w] != 0);
  if (anyclose) {
    int gstart = 0;
    for (int k = 1; k < N; k++) {
      const bool ck = (sclose[k >> 5] >> (k & 31)) & 1;
      if (!ck) { gstart = k; continue; }
      if (k - gstart > maxgroup) { if (j == 0) { sflag = 1; atomicAdd(&ctr->fb_reason[2], 1ull); } continue; }
      if (j == 0) atomicAdd(&ctr->gs_pairs, (unsigned long long)(k - gstart));
      for (int i = gstart; i < k; i++) {
        const double prod = (j < N) ? W[j * LDW + i] * W[j * LDW + k] : 0.;
        __syncthreads();
        const double dot = block_sum<NW>(prod, sred + 4 * (i & 1), warp, lane);
        if (j < N) W[j * LDW + k] = fma(-dot, W[j * LDW + i], W[j * LDW + k]);
      }
      const double wk = (j < N) ? W[j * LDW + k] : 0.;
      __syncthreads();
      const double n2 = block_sum<NW>(wk * wk, sred + 4 * (k & 1), warp, lane);
      if (!(n2 >= TRI_MINREM * TRI_MINREM)) { if (j == 0) { sflag = 1; atomicAdd(&ctr->fb_reason[3], 1ull); } }
      else if (j < N) W[j * LDW + k] = wk * rsqrt(n2);
    }
    __syncthreads();
  }
  if (sflag) {  // recomputed by the Jacobi kernel (launched on flags[] after this kernel)
    if (j == 0) { flags[zl] = ml; atomicAdd(&ctr->fallback, 1ull); }
    return;
  }
  if constexpr (PART == 1) return;  // W is in place in Wg; flags[zl] is 0 (k_tql's verdict, unchanged)
  }  // PART != 2

  // ---- U = Q W : 2 columns x NP/2 rows per thread in registers, reflectors from shared memory ----
  const int cp = j >> 1, rh = j & 1;
  double u[2][NP / 2];
#pragma unroll
  for (int c = 0; c < 2; c++)
#pragma unroll
    for (int m = 0; m < NP / 4; m++) {
      u[c][2 * m] = W[(4 * m + 2 * rh) * LDW + 2 * cp + c];
      u[c][2 * m + 1] = W[(4 * m + 2 * rh + 1) * LDW + 2 * cp + c];
    }
  __syncthreads();
  {
    // the reflectors (up to 32 KB) come in by 16-byte asynchronous copies, all in flight at once: with plain loads
    // staged through registers this loop was 1 % of the instructions but 7.8 % of the stall samples (long scoreboard)
    const double2 *Vg = reinterpret_cast<const double2 *>(Tout + (int64_t)zl * NP * NP);
    double2 *Vs2 = reinterpret_cast<double2 *>(W);
    const int nv = max(N - 2, 0) * (NP / 2);
    for (int idx = j; idx < nv; idx += NP) cp_async16(Vs2 + idx, Vg + idx);
    cp_async_commit();
    cp_async_wait<0>();
  }
  __syncthreads();
  BackTile<NP, 0>::run(u, N, rh, W, stau);

  // ---- per-eigenpair coefficients (columns 2cp, 2cp+1; lane rh writes those of column 2cp+rh) ----
  double ysc[2];
  {
    double q[2] = {0., 0.}, s1[2] = {0., 0.};
#pragma unroll
    for (int c = 0; c < 2; c++) {
#pragma unroll
      for (int m = 0; m < NP / 4; m++) {
        q[c] = fma(u[c][2 * m], sc[4 * m + 2 * rh], q[c]);
        q[c] = fma(u[c][2 * m + 1], sc[4 * m + 2 * rh + 1], q[c]);
        s1[c] += u[c][2 * m] + u[c][2 * m + 1];
      }
      q[c] += __shfl_xor_sync(FULL, q[c], 1);
      s1[c] += __shfl_xor_sync(FULL, s1[c], 1);
      const double g = sgj[2 * cp + c];
      ysc[c] = (g >= TRI_NULL) ? sqrt(g) : 0.;
    }
    __syncthreads();  // every thread is done with the reflectors
    const int jc = 2 * cp + rh;                       // the column this lane writes the coefficients of
    const double lc = fmax(jc < N ? slam[jc] : 0., 0.);
    const double sg = sqrt(1. + lc);
    const double ysj = rh ? ysc[1] : ysc[0], iys = ysj > 0. ? 1. / ysj : 0.;
    sa[jc] = -(lc / (1. + lc)) * (rh ? q[1] : q[0]) * iys;
    sb[jc] = (sg - 1.) * (rh ? s1[1] : s1[0]) * iys;
  }
  // Y = U diag(ys), stored transposed: Yt[k*LDY + i] = Y[i][k] (row k = scaled eigenvector k, contiguous;
  // LDY = NP+2 keeps the 16-byte stores of the lanes of a quarter-warp in different banks)
  double *Yt = W;
#pragma unroll
  for (int c = 0; c < 2; c++)
#pragma unroll
    for (int m = 0; m < NP / 4; m++)
      *reinterpret_cast<double2 *>(Yt + (2 * cp + c) * LDY + 4 * m + 2 * rh) =
          make_double2(ysc[c] * u[c][2 * m], ysc[c] * u[c][2 * m + 1]);
  __syncthreads();
  // ---- ampl, v ----
  double vi = 0.;
  {
    double am = 0.;
    if (j < N) {
      am = sc[j]; vi = 1.;
#if TVEC_DOT8
      // four partial sums per result (see BackTile): chains of N/4 instead of N dependent FMAs
      double am1 = 0., am2 = 0., am3 = 0., vi1 = 0., vi2 = 0., vi3 = 0.;
      int k = 0;
      for (; k + 3 < N; k += 4) {
        const double y0 = Yt[k * LDY + j], y1 = Yt[(k + 1) * LDY + j], y2 = Yt[(k + 2) * LDY + j], y3 = Yt[(k + 3) * LDY + j];
        am = fma(y0, sa[k], am); am1 = fma(y1, sa[k + 1], am1); am2 = fma(y2, sa[k + 2], am2); am3 = fma(y3, sa[k + 3], am3);
        vi = fma(y0, sb[k], vi); vi1 = fma(y1, sb[k + 1], vi1); vi2 = fma(y2, sb[k + 2], vi2); vi3 = fma(y3, sb[k + 3], vi3);
      }
      for (; k < N; k++) {
        const double yv = Yt[k * LDY + j];
        am = fma(yv, sa[k], am);
        vi = fma(yv, sb[k], vi);
      }
      am = (am + am1) + (am2 + am3);
      vi = (vi + vi1) + (vi2 + vi3);
#else
      for (int k = 0; k < N; k++) {
        const double yv = Yt[k * LDY + j];
        am = fma(yv, sa[k], am);
        vi = fma(yv, sb[k], vi);
      }
#endif
    }
    if (am != am) atomicExch(&ctr->nan_flag, 1);
    ampl_out[(int64_t)zl * NP + j] = am;
    if constexpr (FUSE) sres[j] = am;  // sres is free after the grouping: ampl for the fused apply
  }
  const double vnorm = sqrt(block_sum<NW>(vi * vi, sred, warp, lane));
  const double wN = 1. / sqrt((double)N);
  sg1[j] = (j < N) ? vi / vnorm : 0.;  // v
  __syncthreads();
  const double vN = sg1[N - 1];
  const double sv_ = copysign(1., vN);
  const double dNN = sv_;
  const double hv = 1. / (1. + fabs(vN)), hw = 1. / (1. + fabs(wN));
  __syncthreads();
  {
    double uv = sg1[j];
    double uw = (j < N) ? wN : 0.;
    if (j == N - 1) { uv += sv_; uw += 1.; }
    suv[j] = uv;                               // u_v
    sdw[j] = (j == N - 1) ? dNN * uw : uw;     // D u_w
    suw[j] = uw;                               // u_w
  }
  __syncthreads();
  // g1 = M u_v, gm2 = M (D u_w), M = I - Y Y^T : first Y^T x from the columns still held in registers
  {
    double p1[2] = {0., 0.}, p2[2] = {0., 0.};
#pragma unroll
    for (int c = 0; c < 2; c++) {
#pragma unroll
      for (int m = 0; m < NP / 4; m++) {
        const int i0 = 4 * m + 2 * rh;
        p1[c] = fma(u[c][2 * m], suv[i0], p1[c]);
        p1[c] = fma(u[c][2 * m + 1], suv[i0 + 1], p1[c]);
        p2[c] = fma(u[c][2 * m], sdw[i0], p2[c]);
        p2[c] = fma(u[c][2 * m + 1], sdw[i0 + 1], p2[c]);
      }
      p1[c] += __shfl_xor_sync(FULL, p1[c], 1);
      p2[c] += __shfl_xor_sync(FULL, p2[c], 1);
    }
    sa[2 * cp + rh] = rh ? ysc[1] * p1[1] : ysc[0] * p1[0];
    sb[2 * cp + rh] = rh ? ysc[1] * p2[1] : ysc[0] * p2[0];
  }
  const double kappa = block_sum<NW>(suv[j] * hv * sdw[j], sred + 4, warp, lane);
  __syncthreads();
  {
    double g1 = suv[j], gm2 = sdw[j];
#if TVEC_DOT8
    double g11 = 0., g12 = 0., g13 = 0., g21 = 0., g22 = 0., g23 = 0.;
    int k = 0;
    for (; k + 3 < N; k += 4) {
      const double y0 = Yt[k * LDY + j], y1 = Yt[(k + 1) * LDY + j], y2 = Yt[(k + 2) * LDY + j], y3 = Yt[(k + 3) * LDY + j];
      g1 = fma(-y0, sa[k], g1); g11 = fma(-y1, sa[k + 1], g11); g12 = fma(-y2, sa[k + 2], g12); g13 = fma(-y3, sa[k + 3], g13);
      gm2 = fma(-y0, sb[k], gm2); g21 = fma(-y1, sb[k + 1], g21); g22 = fma(-y2, sb[k + 2], g22); g23 = fma(-y3, sb[k + 3], g23);
    }
    for (; k < N; k++) {
      const double yv = Yt[k * LDY + j];
      g1 = fma(-yv, sa[k], g1);
      gm2 = fma(-yv, sb[k], gm2);
    }
    g1 = (g1 + g11) + (g12 + g13);
    gm2 = (gm2 + g21) + (g22 + g23);
#else
    for (int k = 0; k < N; k++) {
      const double yv = Yt[k * LDY + j];
      g1 = fma(-yv, sa[k], g1);
      gm2 = fma(-yv, sb[k], gm2);
    }
#endif
    sg1[j] = g1;
    sg2[j] = gm2 - kappa * g1;
  }
  __syncthreads();
  if constexpr (FUSE) {
    // ---- rows of the zone from the factored transform; T itself is not formed ----
    sa[j] = sg1[j] * hv;   // sa, sb were last read before the barrier above
    sb[j] = sg2[j] * hw;
    __syncthreads();
    double *sS = sm + NP * LDY + NVEC * NP, *sP = sS + FusedApply<NP>::RC * FusedApply<NP>::LD;
    double *s_row = sP + FusedApply<NP>::RC * FusedApply<NP>::LD;
    const int64_t i1 = aa.zstart[zl] - aa.rowbase;
    const int nrow = (int)(aa.zstart[zl + 1] - aa.zstart[zl]);
    FusedApply<NP>::run(N, j, Yt, sS, sP, s_row, sa, sb, sres, suv, suw, dNN, i1, nrow, aa.xf, aa.Sf, aa.ldS, aa.xa,
                        aa.Sa, aa.ldSa);
  } else
  // ---- T[i][k] = (M[i][k] - g1[i] hv u_v[k]) D_k - g2[i] hw u_w[k] , row-major ----
#if TVEC_MMA_T
  // M = I - Y Y^T on the fp64 tensor cores: only the 36 tiles (8 x 8) on or below the diagonal are accumulated (Y Y^T
  // is symmetric), 18 per warp; a k-step of four eigenpairs needs one fragment load per block the warp touches
  // (lane 4g+t reads Yt[(k0+t)][8b+g]: A fragment of block row b and B fragment of block column b).  The rank-one
  // terms are applied to the accumulator fragments, the strictly lower tiles are written twice (as they are and
  // transposed).  ~580 DMMA + ~260 loads per zone instead of 8.2 k DFMA + 1 k loads, 36 instead of 64 accumulators.
  if constexpr (NP == 64) {
    const int gq = lane >> 2, tq = lane & 3;
    const int NK = (N + 3) & ~3;
    double *Tz = Tout + (int64_t)zl * NP * NP;
    auto run = [&](auto tm) {
      using TM = decltype(tm);
      double acc[18][2];
#pragma unroll
      for (int a_ = 0; a_ < 18; a_++) acc[a_][0] = acc[a_][1] = 0.;
#pragma unroll 2
      for (int k0 = 0; k0 < NK; k0 += 4) {
        const double *r = Yt + (k0 + tq) * LDY + gq;
        double v[8];
#pragma unroll
        for (int b = 0; b < 8; b++)
          if (TM::needs_row(b) || TM::needs_col(b)) v[b] = r[8 * b];
#pragma unroll
        for (int bi = 0; bi < 8; bi++)
#pragma unroll
          for (int bj = 0; bj <= bi; bj++)
            if (TM::owns(bi, bj)) oak_dmma_m8n8k4(acc[TM::slot(bi, bj)][0], acc[TM::slot(bi, bj)][1], v[bi], v[bj]);
      }
#pragma unroll
      for (int bi = 0; bi < 8; bi++)
#pragma unroll
        for (int bj = 0; bj <= bi; bj++)
          if (TM::owns(bi, bj)) {
            const int i = 8 * bi + gq, k = 8 * bj + 2 * tq;
            const double m0 = (i == k ? 1. : 0.) - acc[TM::slot(bi, bj)][0];
            const double m1 = (i == k + 1 ? 1. : 0.) - acc[TM::slot(bi, bj)][1];
            {
              const double g1i = sg1[i] * hv, g2i = sg2[i] * hw;
              double t0 = m0 - g1i * suv[k], t1 = m1 - g1i * suv[k + 1];
              if (k == N - 1) t0 *= dNN;
              if (k + 1 == N - 1) t1 *= dNN;
              t0 -= g2i * suw[k];
              t1 -= g2i * suw[k + 1];
              *reinterpret_cast<double2 *>(Tz + (int64_t)i * NP + k) = make_double2(t0, t1);
            }
            if (bi != bj) {   // the mirrored tile: T[k][i], T[k+1][i] from M[k][i] = M[i][k]
              const double uvi = suv[i], uwi = suw[i];
              double t0 = m0 - sg1[k] * hv * uvi, t1 = m1 - sg1[k + 1] * hv * uvi;
              if (i == N - 1) { t0 *= dNN; t1 *= dNN; }
              t0 -= sg2[k] * hw * uwi;
              t1 -= sg2[k + 1] * hw * uwi;
              Tz[(int64_t)k * NP + i] = t0;
              Tz[(int64_t)(k + 1) * NP + i] = t1;
            }
          }
    };
    if (warp == 0) run(MTileMap<0>{}); else run(MTileMap<1>{});
  } else
#endif
  {
    const int ti = j / TJ, tj = j % TJ;
    double acc[TR][TC];
#pragma unroll
    for (int a_ = 0; a_ < TR; a_++)
#pragma unroll
      for (int b = 0; b < TC; b++) acc[a_][b] = 0.;
    const double *yr = Yt + TR * ti, *yc = Yt + 2 * tj;
#pragma unroll 2
    for (int k = 0; k < N; k++) {
      double rv[TR], cv[TC];
#pragma unroll
      for (int a_ = 0; a_ < TR; a_ += 2) {
        const double2 r2 = *reinterpret_cast<const double2 *>(yr + k * LDY + a_);
        rv[a_] = r2.x; rv[a_ + 1] = r2.y;
      }
#pragma unroll
      for (int b = 0; b < TC; b += 2) {
        const double2 c2 = *reinterpret_cast<const double2 *>(yc + k * LDY + (2 * TJ) * (b >> 1));
        cv[b] = c2.x; cv[b + 1] = c2.y;
      }
#pragma unroll
      for (int a_ = 0; a_ < TR; a_++)
#pragma unroll
        for (int b = 0; b < TC; b++) acc[a_][b] = fma(-rv[a_], cv[b], acc[a_][b]);
    }
    double *Tz = Tout + (int64_t)zl * NP * NP;
#pragma unroll
    for (int a_ = 0; a_ < TR; a_++) {
      const int i = TR * ti + a_;
      const double g1i = sg1[i] * hv, g2i = sg2[i] * hw;
#pragma unroll
      for (int b = 0; b < TC; b += 2) {
        const int k = 2 * tj + (2 * TJ) * (b >> 1);
        double t0 = acc[a_][b] + (i == k ? 1. : 0.) - g1i * suv[k];
        double t1 = acc[a_][b + 1] + (i == k + 1 ? 1. : 0.) - g1i * suv[k + 1];
        if (k == N - 1) t0 *= dNN;
        if (k + 1 == N - 1) t1 *= dNN;
        t0 -= g2i * suw[k];
        t1 -= g2i * suw[k + 1];
        *reinterpret_cast<double2 *>(Tz + (int64_t)i * NP + k) = make_double2(t0, t1);
      }
    }
  }
  if (j == 0) flags[zl] = 0;
}

template <int NP, bool FUSE>
int launch_tvec(cudaStream_t st, int N, int nz, const int32_t *mloc, const double *c, double *T, double *ampl,
                double *ws, int32_t *flags, DevCounters *ctr, double orthtol, int maxgroup, const FusedApplyArgs &aa,
                double *Wg) {
  // experiment switches OAK_B200_TVEC_PAD / OAK_B200_TRI_PAD: extra (unused) dynamic shared memory per CTA, i.e. fewer
  // resident CTAs of this kernel per SM, so that CTAs of another stream's kernel fit beside them
  static const int tvec_pad = getenv("OAK_B200_TVEC_PAD") ? std::max(0, atoi(getenv("OAK_B200_TVEC_PAD"))) : 0;
  const size_t smem0 = sizeof(double) * 17 * NP;
  const size_t smem = sizeof(double) * (NP * (NP + 4) + 17 * NP + (FUSE ? FusedApply<NP>::SMEM_DOUBLES : 0)) + (size_t)tvec_pad;
  { int rc_ = oak_func_smem(k_tvec<NP, FUSE, 0>, (size_t)((int)smem)); if (rc_) return rc_; }
  { int rc_ = oak_func_smem(k_tvec<NP, false, 1>, (size_t)((int)smem0)); if (rc_) return rc_; }
  { int rc_ = oak_func_smem(k_tvec<NP, FUSE, 2>, (size_t)((int)smem)); if (rc_) return rc_; }
  const double ot = orthtol > 0. ? orthtol : TRI_ORTHTOL;
  const int mg = maxgroup >= 0 ? maxgroup : TRI_MAXGROUP;
  if (Wg) {
    k_tvec<NP, false, 1><<<nz, NP, smem0, st>>>(N, mloc, ws, c, T, ampl, flags, ctr, ot, mg, FusedApplyArgs{}, Wg);
    CUDA_TRY(cudaGetLastError());
    k_tvec<NP, FUSE, 2><<<nz, NP, smem, st>>>(N, mloc, ws, c, T, ampl, flags, ctr, ot, mg, aa, Wg);
  } else {
    k_tvec<NP, FUSE, 0><<<nz, NP, smem, st>>>(N, mloc, ws, c, T, ampl, flags, ctr, ot, mg, aa, nullptr);
  }
  CUDA_TRY(cudaGetLastError());
  return 0;
}

template <int NP>
int launch(cudaStream_t st, int N, int nz, const int32_t *mloc, const double *G, const double *c, double *T,
           double *ampl, double *ws, int32_t *flags, DevCounters *ctr, cudaEvent_t *ev, double orthtol, int maxgroup,
           const FusedApplyArgs *fuse, double *Wg, const TqlSide *side, double *scr) {
  // scratch of k_tql: 2 NP doubles per zone, addressed by groups of 32 zones (sub-ranges below start at multiples of 32)
  auto scr_at = [&](int z0) { return scr + (size_t)(z0 / 32) * 2 * NP * 32; };
#if TRI_TILE && TRI_WARP
  // Two halves of the batch, software-pipelined around the latency-bound QL kernel: while the eigenvalues of the first
  // half are computed on the side stream, the main stream reduces the second half; the eigenvectors of the first half
  // then overlap the QL of the second.  (The chain of a slot otherwise idles ~1.1 ms per batch inside k_tql.)
  static const int halves = getenv("OAK_B200_EIG_HALVES") ? atoi(getenv("OAK_B200_EIG_HALVES")) : EIG_HALVES;
  if (NP == 64 && side && !fuse && !Wg && !ev && halves == 2 && nz >= 4096) {
    const int h1 = (nz / 2 + 31) / 32 * 32;
    const cudaEvent_t evs[4] = {side->e0, side->e1, side->e2, side->e3};
    for (int part = 0; part < 2; part++) {
      const int z0 = part ? h1 : 0, n1 = part ? nz - h1 : h1;
      k_tridiag_warp<<<n1, 32, 0, st>>>(N, n1, mloc + z0, G + (size_t)z0 * NP * NP, T + (size_t)z0 * NP * NP, ws + (size_t)z0 * 4 * NP);
      CUDA_TRY(cudaGetLastError());
      CUDA_TRY(cudaEventRecord(evs[2 * part], st));
      CUDA_TRY(cudaStreamWaitEvent(side->qst, evs[2 * part], 0));
      k_tql<NP><<<(n1 + 31) / 32, 32, 0, side->qst>>>(N, n1, mloc + z0, ws + (size_t)z0 * 4 * NP, flags + z0, scr_at(z0));
      CUDA_TRY(cudaGetLastError());
      CUDA_TRY(cudaEventRecord(evs[2 * part + 1], side->qst));
    }
    for (int part = 0; part < 2; part++) {
      const int z0 = part ? h1 : 0, n1 = part ? nz - h1 : h1;
      CUDA_TRY(cudaStreamWaitEvent(st, evs[2 * part + 1], 0));
      int rc = launch_tvec<NP, false>(st, N, n1, mloc + z0, c + (size_t)z0 * NP, T + (size_t)z0 * NP * NP, ampl + (size_t)z0 * NP,
                                      ws + (size_t)z0 * 4 * NP, flags + z0, ctr, orthtol, maxgroup, FusedApplyArgs{}, nullptr);
      if (rc) return rc;
    }
    return 0;
  }
#endif
#if TRI_TILE
  static const int tri_warp = getenv("OAK_B200_TRI_WARP") ? atoi(getenv("OAK_B200_TRI_WARP")) : TRI_WARP;
  static const int tri_pad = getenv("OAK_B200_TRI_PAD") ? std::min(40000, std::max(0, atoi(getenv("OAK_B200_TRI_PAD")))) : 0;
  if (NP == 64 && tri_warp) k_tridiag_warp<<<nz, 32, tri_pad, st>>>(N, nz, mloc, G, T, ws);
  else k_tridiag_tile<NP><<<nz, 64, 0, st>>>(N, mloc, G, T, ws);
#else
  k_tridiag<NP><<<nz, NP, 0, st>>>(N, mloc, G, T, ws);
#endif
  CUDA_TRY(cudaGetLastError());
  if (ev) CUDA_TRY(cudaEventRecord(ev[0], st));
  if (side) {
    CUDA_TRY(cudaEventRecord(side->e0, st));
    CUDA_TRY(cudaStreamWaitEvent(side->qst, side->e0, 0));
    {
      // option: the batch in `tql_split` consecutive launches of the side stream (at most one 33 KB warp per SM at a
      // time then fits beside four k_tvec CTAs instead of displacing two of them)
      static const int split = getenv("OAK_B200_TQL_SPLIT") ? std::max(1, atoi(getenv("OAK_B200_TQL_SPLIT"))) : 1;
      const int per = ((nz + split - 1) / split + 31) / 32 * 32;
      for (int z0 = 0; z0 < nz; z0 += per) {
        const int n1 = std::min(per, nz - z0);
        k_tql<NP><<<(n1 + 31) / 32, 32, 0, side->qst>>>(N, n1, mloc + z0, ws + (size_t)z0 * 4 * NP, flags + z0, scr_at(z0));
      }
    }
    CUDA_TRY(cudaEventRecord(side->e1, side->qst));
    CUDA_TRY(cudaStreamWaitEvent(st, side->e1, 0));
  } else
  k_tql<NP><<<(nz + 31) / 32, 32, 0, st>>>(N, nz, mloc, ws, flags, scr);
  CUDA_TRY(cudaGetLastError());
  {  // timing experiment (OAK_B200_TQL_REPEAT=k): the kernel only reads d, e and writes lambda, so it can be repeated
    static const int rep = getenv("OAK_B200_TQL_REPEAT") ? atoi(getenv("OAK_B200_TQL_REPEAT")) : 0;
    for (int r = 0; r < rep; r++) k_tql<NP><<<(nz + 31) / 32, 32, 0, st>>>(N, nz, mloc, ws, flags, scr);
  }
  if (ev) CUDA_TRY(cudaEventRecord(ev[1], st));
  if (fuse) return launch_tvec<NP, true>(st, N, nz, mloc, c, T, ampl, ws, flags, ctr, orthtol, maxgroup, *fuse, Wg);
  return launch_tvec<NP, false>(st, N, nz, mloc, c, T, ampl, ws, flags, ctr, orthtol, maxgroup, FusedApplyArgs{}, Wg);
}

}  // namespace

// Per-zone workspace of the tridiagonal route: 4 NP doubles (d, e, tau, lambda) + one int32 flag + k_tql's scratch
// (transposed d, e: 2 NP doubles per zone, groups of 32 zones).
size_t oak_eig_tridiag_ws_bytes(int NP, int nz) {
  return sizeof(double) * 4 * (size_t)NP * nz + (sizeof(int32_t) * (size_t)nz + 255) / 256 * 256 +
         sizeof(double) * 2 * (size_t)NP * (((size_t)nz + 31) / 32 * 32) + 64;
}

// Enqueues k_tridiag, k_tql, k_tvec; on return (stream order) flags[zl] = mloc[zl] for the zones that must be
// recomputed by the Jacobi kernel and 0 for the others.  V (the reflectors) lives in T until k_tvec replaces it.
// fuse != NULL: k_tvec also updates the rows of its zones (FusedApply) and writes no T; the caller then applies
// only the zones k_tvec did not finish (not analysed, or flagged and recomputed by the Jacobi kernel).
int oak_launch_eig_tridiag(cudaStream_t st, int N, int NP, int nz, const int32_t *mloc, const double *G,
                           const double *c, double *T, double *ampl, void *ws, int32_t **flags_out,
                           DevCounters *ctr, cudaEvent_t *ev, double orthtol, int maxgroup,
                           const FusedApplyArgs *fuse, double *Wg, const TqlSide *side) {
  double *wsd = reinterpret_cast<double *>(ws);
  int32_t *flags = reinterpret_cast<int32_t *>(wsd + 4 * (size_t)NP * nz);
  *flags_out = flags;
  double *scr = reinterpret_cast<double *>(reinterpret_cast<char *>(flags) + (sizeof(int32_t) * (size_t)nz + 255) / 256 * 256);
  switch (NP) {
    case 32: return launch<32>(st, N, nz, mloc, G, c, T, ampl, wsd, flags, ctr, ev, orthtol, maxgroup, fuse, Wg, side, scr);
    case 64: return launch<64>(st, N, nz, mloc, G, c, T, ampl, wsd, flags, ctr, ev, orthtol, maxgroup, fuse, Wg, side, scr);
  }
  oak_set_error("eig_tridiag: unsupported padded ensemble size %d", NP);
  return OAK_ERR_UNSUPPORTED;
}
