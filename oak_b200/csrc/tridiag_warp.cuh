// tridiag_warp.cuh — k_tridiag_warp<64>: the fused Householder step of k_tridiag_tile with ONE WARP per zone and
// the matrix held as its lower block triangle.
//
// k_tridiag_tile (64 threads, full 64 x 64 matrix in registers) is bound by the latency of a step, not by any pipe:
// ~300 warp instructions per warp and step of which ~76 are the DFMAs that matter, two block barriers, a five-round
// block reduction and 6 zones in flight per SM (168 registers x 64 threads each).  Here
//   * a lane plays two cells (rho, gamma), rho = rho4 + 4 h, of the same 8 x 8 cyclic grid and keeps only the slots
//     [i][b], b <= i, of each: element (8 i + rho, 8 b + gamma).  Slots above the block diagonal live, transposed,
//     in cell (gamma, rho).  The rank-2 update touches 36 instead of 64 slots per cell; the matrix-vector product
//     uses every off-diagonal slot twice (row sum of 8 i + rho directly, row sum of 8 b + gamma for the mirrored
//     element), and the mirrored partial sums change cells through a small shared-memory exchange
//     (cell (rho, gamma) reads what cell (gamma, rho) wrote) before the usual transpose-reduction over gamma;
//   * there is no block barrier (__syncwarp only) and one packed 6-shuffle all-reduce for |x|^2 and x^T y;
//   * 8 zones are in flight per SM (one warp each at <= 255 registers) and a zone-step costs about half the
//     instructions of the two-warp kernel.
// Same outputs as k_tridiag_tile: reflectors V (row k = reflector k) in the T buffer, d, e, tau in the workspace.
#pragma once

namespace tw {

__host__ __device__ constexpr int tri(int i, int b) { return i * (i + 1) / 2 + b; }
// Row t of the published vectors lives at position t ^ (t >> 3): the owners' accesses (t = 8 gamma + rho over the
// lanes of a warp, a stride of 8 elements = every lane in the same banks without it), the column operands
// (8 b + gamma) and the row operands (8 i + rho) are then all conflict-free (ncu: 43.7 M conflicts per 7104 zones,
// shared-memory pipe 78 % busy, before).
__device__ __forceinline__ int sw(int t) { return t ^ (t >> 3); }

struct alignas(16) WarpSmem {
  double2 svw[64];      // (v_t, w_t) of the current step
  double sxn[64];       // next column x'
  double sx2[64];       // column k+1 of the matrix (row k+1 by symmetry)
  double ex[7 * 80];    // exchange of the mirrored partial sums: [b][10 rho + gamma]
};

// row sums of the trailing matrix times the published vector sxn: y (rows 8 gamma + rho_h in lane gamma)
template <int KB>
__device__ __forceinline__ void matvec(const double (&a)[2][36], double (&y)[2], int g, int r4, WarpSmem &S) {
  double cxn[8];
#pragma unroll
  for (int b = KB; b < 8; b++) cxn[b] = S.sxn[sw(8 * b + g)];
  double yd[2][8];
#pragma unroll
  for (int h = 0; h < 2; h++) {
    double ym[7];
#pragma unroll
    for (int i = 0; i < 8; i++) yd[h][i] = 0.;
#pragma unroll
    for (int b = 0; b < 7; b++) ym[b] = 0.;
#pragma unroll
    for (int i = KB; i < 8; i++) {
      const double rx = S.sxn[sw(8 * i + r4 + 4 * h)];
#pragma unroll
      for (int b = KB; b <= i; b++) {
        const double av = a[h][tri(i, b)];
        yd[h][i] = fma(av, cxn[b], yd[h][i]);
        if (b < i) ym[b] = fma(av, rx, ym[b]);
      }
    }
#pragma unroll
    for (int b = KB; b < 7; b++) S.ex[b * 80 + 10 * (r4 + 4 * h) + g] = ym[b];
  }
  __syncwarp();
#pragma unroll
  for (int h = 0; h < 2; h++) {
#pragma unroll
    for (int i = KB; i < 7; i++) yd[h][i] += S.ex[i * 80 + 10 * g + r4 + 4 * h];
    y[h] = transpose_reduce8<8>(yd[h], g);
  }
  __syncwarp();   // ex may be rewritten by the next step
}

template <int KB>
struct WarpSteps {
  static __device__ __forceinline__ void run(double (&a)[2][36], double (&x)[2], double (&y)[2], int N, int lane,
                                             int g, int r4, WarpSmem &S, double *Vz, double *wd, double *we,
                                             double *wtau) {
    const int t0 = 8 * g + r4, t1 = t0 + 4;
    const int kend = min(8 * KB + 8, N - 2);
    for (int k = 8 * KB; k < kend; k++) {
      double s1 = ((t0 > k + 1) ? x[0] * x[0] : 0.) + ((t1 > k + 1) ? x[1] * x[1] : 0.);
      double s2 = ((t0 > k) ? x[0] * y[0] : 0.) + ((t1 > k) ? x[1] * y[1] : 0.);
      {  // packed all-reduce: lanes < 16 finish s1, lanes >= 16 finish s2, then they swap
        const bool up = lane & 16;
        double v = up ? s2 : s1;
        const double o = up ? s1 : s2;
        v += __shfl_xor_sync(FULL, o, 16);
#pragma unroll
        for (int d = 8; d > 0; d >>= 1) v += __shfl_xor_sync(FULL, v, d);
        const double u = __shfl_xor_sync(FULL, v, 16);
        s1 = up ? u : v;
        s2 = up ? v : u;
      }
      const int kr = k + 1;
      const int Lk = ((kr & 3) << 3) | (kr >> 3);
      const bool hk = (kr >> 2) & 1;
      const double alpha = __shfl_sync(FULL, hk ? x[1] : x[0], Lk);
      const double yk1 = __shfl_sync(FULL, hk ? y[1] : y[0], Lk);
      const double c10 = S.sx2[sw(t0)], c11 = S.sx2[sw(t1)];   // A[t][k+1]
      const double akk = S.sx2[sw(kr)];                      // A[k+1][k+1]
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
      const double wk1 = fma(ts, fma(-beta, akk, yk1), -hpv);  // w_{k+1}
      const double vt0 = (t0 > kr) ? x[0] * scale : (t0 == kr ? 1. : 0.);
      const double vt1 = (t1 > kr) ? x[1] * scale : (t1 == kr ? 1. : 0.);
      const double pt0 = (t0 > k) ? ts * fma(-beta, c10, y[0]) : 0.;
      const double pt1 = (t1 > k) ? ts * fma(-beta, c11, y[1]) : 0.;
      const double wt0 = fma(-hpv, vt0, pt0), wt1 = fma(-hpv, vt1, pt1);
      const double xn0 = (t0 > kr) ? c10 - fma(vt0, wk1, wt0) : 0.;  // new A[t][k+1]
      const double xn1 = (t1 > kr) ? c11 - fma(vt1, wk1, wt1) : 0.;
      S.svw[sw(t0)] = make_double2(vt0, wt0);
      S.svw[sw(t1)] = make_double2(vt1, wt1);
      S.sxn[sw(t0)] = xn0;
      S.sxn[sw(t1)] = xn1;
      if (lane == Lk) { we[k] = beta; wtau[k] = tau; wd[kr] = fma(-2., wk1, akk); }
      __syncwarp();
      Vz[k * 64 + lane] = S.svw[sw(lane)].x;            // reflector k, coalesced
      Vz[k * 64 + 32 + lane] = S.svw[sw(32 + lane)].x;
      {  // rank-2 update of the stored slots
        double2 cvw[8];
#pragma unroll
        for (int b = KB; b < 8; b++) cvw[b] = S.svw[sw(8 * b + g)];
#pragma unroll
        for (int h = 0; h < 2; h++)
#pragma unroll
          for (int i = KB; i < 8; i++) {
            const double2 r = S.svw[sw(8 * i + r4 + 4 * h)];
#pragma unroll
            for (int b = KB; b <= i; b++)
              a[h][tri(i, b)] = fma(-r.x, cvw[b].y, fma(-r.y, cvw[b].x, a[h][tri(i, b)]));
          }
      }
      // publish column k+2 (block (k+2)/8 is KB or KB+1: static register indices either way)
      {
        const int pr = k + 2;
        if (g == (pr & 7)) {
          if ((pr >> 3) == KB) {
#pragma unroll
            for (int h = 0; h < 2; h++)
#pragma unroll
              for (int i = KB; i < 8; i++) S.sx2[sw(8 * i + r4 + 4 * h)] = a[h][tri(i, KB)];
          } else if constexpr (KB + 1 < 8) {
#pragma unroll
            for (int h = 0; h < 2; h++)
#pragma unroll
              for (int i = KB + 1; i < 8; i++) S.sx2[sw(8 * i + r4 + 4 * h)] = a[h][tri(i, KB + 1)];
          }
        }
      }
      matvec<KB>(a, y, g, r4, S);   // y' = A_new x' (ends with __syncwarp: sx2, svw, sxn are free to be rewritten)
      x[0] = xn0;
      x[1] = xn1;
    }
    if constexpr (KB + 1 < 8) {
      if (N - 2 > 8 * KB + 8) WarpSteps<KB + 1>::run(a, x, y, N, lane, g, r4, S, Vz, wd, we, wtau);
    }
  }
};

}  // namespace tw

#ifndef TRIW_MINB
#define TRIW_MINB 8
#endif
__global__ void __launch_bounds__(32, TRIW_MINB) k_tridiag_warp(int N, int nz, const int32_t *__restrict__ mloc,
                                                         const double *__restrict__ G, double *__restrict__ V,
                                                         double *__restrict__ ws) {
  constexpr int NP = 64;
  __shared__ tw::WarpSmem S;
  const int zl = blockIdx.x;
  if (zl >= nz || mloc[zl] == 0) return;
  const int lane = threadIdx.x, g = lane & 7, r4 = lane >> 3;
  const double *Gz = G + (int64_t)zl * NP * NP;
  double *Vz = V + (int64_t)zl * NP * NP;
  double *wz = ws_zone(ws, NP, zl);
  double *wd = wz, *we = wz + NP, *wtau = wz + 2 * NP;

  double a[2][36];
#pragma unroll
  for (int h = 0; h < 2; h++)
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
      for (int b = 0; b <= i; b++) a[h][tw::tri(i, b)] = Gz[(8 * i + r4 + 4 * h) * NP + 8 * b + g];
  const int t0 = 8 * g + r4, t1 = t0 + 4;
  double x[2], y[2];
  x[0] = (t0 >= 1) ? Gz[t0] : 0.;   // column 0 (= row 0: G is symmetric)
  x[1] = Gz[t1];
  S.sxn[tw::sw(t0)] = x[0];
  S.sxn[tw::sw(t1)] = x[1];
#pragma unroll
  for (int q = 0; q < 2; q++) {
    const int t = lane + 32 * q;
    S.sx2[tw::sw(t)] = Gz[NP + t];           // row 1
    if (t >= N) { wd[t] = 0.; we[t] = 0.; }
    wtau[t] = 0.;
  }
  if (lane == 0) wd[0] = Gz[0];
  __syncwarp();
  tw::matvec<0>(a, y, g, r4, S);
  tw::WarpSteps<0>::run(a, x, y, N, lane, g, r4, S, Vz, wd, we, wtau);
  // x of row N-1's owner is now e_{N-2}; sx2 holds column N-1
  __syncwarp();
  {
    const int t = N - 1;
    const int L = ((t & 3) << 3) | (t >> 3);
    if (lane == L) { we[N - 2] = ((t >> 2) & 1) ? x[1] : x[0]; we[N - 1] = 0.; wd[N - 1] = S.sx2[tw::sw(N - 1)]; }
  }
}
