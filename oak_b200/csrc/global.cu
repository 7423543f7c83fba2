// global.cu — the global scheme (schemetype = 0): `analysis` (rrsqrt.F90:196-208), i.e. analysisIncrement
// (rrsqrt.F90:100-190) with every observation and weight 1, on the primitives of the local path:
//     G = HSf^T R^-1 HSf ,  c = HSf^T R^-1 (yo - Hxf)        one tall-skinny contraction over the m observations
//     (ampl, T) from (G, c)                                 the transform kernels of the local path, one "zone"
//     Sa = Sf T ,  xa = xf + Sf ampl                        k_apply over row blocks that all use the same T
// R = DiagCovar, optionally wrapped by DCDCovar (0/1 vector), as on the local path (covariance.F90:425-431,:612-619).
//
// k_global_gram: a CTA takes a contiguous range of observations, stages 32 rows at a time (coalesced along the
// observation index, the fast index of the column-major HSf) and accumulates a register tile of G; the partial
// matrices are summed in a fixed order by k_global_reduce, so the result does not depend on the schedule.
#include <algorithm>

#include "common.cuh"

namespace {

constexpr int GG_ROWS = 32;   // observations staged per step
constexpr int GG_NT = 256;    // 16 x 16 thread grid

template <int NP>
__global__ void __launch_bounds__(GG_NT) k_global_gram(int m, int N, const double *__restrict__ HSf, int64_t ldH,
                                                       const double *__restrict__ yo, const double *__restrict__ Hxf,
                                                       const double *__restrict__ Rdiag, const double *__restrict__ d01,
                                                       int obs_per_cta, double *__restrict__ Gpart,
                                                       double *__restrict__ cpart) {
  constexpr int TT = NP / 16;        // tile edge of a thread
  constexpr int LDA = NP + 1;
  __shared__ double sA[GG_ROWS * LDA];
  __shared__ double s_coef[GG_ROWS], s_cd[GG_ROWS];
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int l0 = blockIdx.x * obs_per_cta, l1 = min(m, l0 + obs_per_cta);
  double acc[TT][TT];
#pragma unroll
  for (int a = 0; a < TT; a++)
#pragma unroll
    for (int b = 0; b < TT; b++) acc[a][b] = 0.;
  double cacc = 0.;
  for (int lb = l0; lb < l1; lb += GG_ROWS) {
    __syncthreads();
    for (int idx = tid; idx < GG_ROWS * NP; idx += GG_NT) {
      const int r = idx & (GG_ROWS - 1), i = idx >> 5;
      const int l = lb + r;
      sA[r * LDA + i] = (l < l1 && i < N) ? HSf[l + ldH * (int64_t)i] : 0.;
    }
    if (tid < GG_ROWS) {
      const int l = lb + tid;
      double coef = 0., cd = 0.;
      if (l < l1) {
        const double e = d01 ? d01[l] : 1.;
        coef = (e * e) / Rdiag[l];
        cd = coef * (yo[l] - Hxf[l]);
      }
      s_coef[tid] = coef;
      s_cd[tid] = cd;
    }
    __syncthreads();
#pragma unroll 4
    for (int r = 0; r < GG_ROWS; r++) {
      const double coef = s_coef[r];
      double rv[TT], cv[TT];
#pragma unroll
      for (int a = 0; a < TT; a++) rv[a] = sA[r * LDA + TT * ty + a] * coef;
#pragma unroll
      for (int b = 0; b < TT; b++) cv[b] = sA[r * LDA + TT * tx + b];
#pragma unroll
      for (int a = 0; a < TT; a++)
#pragma unroll
        for (int b = 0; b < TT; b++) acc[a][b] = fma(rv[a], cv[b], acc[a][b]);
      if (tid < NP) cacc = fma(s_cd[r], sA[r * LDA + tid], cacc);
    }
  }
  double *Gp = Gpart + (int64_t)blockIdx.x * NP * NP;
#pragma unroll
  for (int a = 0; a < TT; a++)
#pragma unroll
    for (int b = 0; b < TT; b++) Gp[(TT * ty + a) + NP * (TT * tx + b)] = acc[a][b];
  if (tid < NP) cpart[(int64_t)blockIdx.x * NP + tid] = cacc;
}

// G = sum of the partial matrices in CTA order; mloc[0] = m (the "zone" of the global scheme sees every observation)
__global__ void __launch_bounds__(256) k_global_reduce(int nparts, int NP, int m, const double *__restrict__ Gpart,
                                                       const double *__restrict__ cpart, double *__restrict__ G,
                                                       double *__restrict__ c, int32_t *__restrict__ mloc) {
  const int e = blockIdx.x * 256 + threadIdx.x;
  const int nn = NP * NP;
  if (e < nn) {
    double s = 0.;
    for (int p = 0; p < nparts; p++) s += Gpart[(int64_t)p * nn + e];
    G[e] = s;
  } else if (e < nn + NP) {
    const int i = e - nn;
    double s = 0.;
    for (int p = 0; p < nparts; p++) s += cpart[(int64_t)p * NP + i];
    c[i] = s;
  }
  if (e == 0) mloc[0] = m;
}

// prefix sums of the row blocks the apply kernel treats as zones
__global__ void k_block_starts(int64_t n, int rows_per_block, int nblocks, int64_t *zstart) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b <= nblocks) zstart[b] = min(n, (int64_t)b * rows_per_block);
}

}  // namespace

size_t oak_global_gram_ws_bytes(int NP, int nparts) { return sizeof(double) * (size_t)nparts * ((size_t)NP * NP + NP); }

// number of partial Gram matrices: enough CTAs for two waves, at least 32 * 8 observations each
int oak_global_gram_parts(int m) {
  const int per = GG_ROWS * 8;
  return std::max(1, std::min(2 * 148, (m + per - 1) / per));
}

int oak_launch_global_gram(cudaStream_t st, int m, int N, int NP, const double *HSf, int64_t ldH, const double *yo,
                           const double *Hxf, const double *Rdiag, const double *d01, void *ws, int nparts, double *G,
                           double *c, int32_t *mloc) {
  double *Gpart = reinterpret_cast<double *>(ws);
  double *cpart = Gpart + (size_t)nparts * NP * NP;
  int per = (m + nparts - 1) / nparts;
  per = std::max(GG_ROWS, ((per + GG_ROWS - 1) / GG_ROWS) * GG_ROWS);
  switch (NP) {
    case 32: k_global_gram<32><<<nparts, GG_NT, 0, st>>>(m, N, HSf, ldH, yo, Hxf, Rdiag, d01, per, Gpart, cpart); break;
    case 64: k_global_gram<64><<<nparts, GG_NT, 0, st>>>(m, N, HSf, ldH, yo, Hxf, Rdiag, d01, per, Gpart, cpart); break;
    case 128: k_global_gram<128><<<nparts, GG_NT, 0, st>>>(m, N, HSf, ldH, yo, Hxf, Rdiag, d01, per, Gpart, cpart); break;
    default: oak_set_error("global gram: unsupported padded ensemble size %d", NP); return OAK_ERR_UNSUPPORTED;
  }
  CUDA_TRY(cudaGetLastError());
  k_global_reduce<<<(NP * NP + NP + 255) / 256, 256, 0, st>>>(nparts, NP, m, Gpart, cpart, G, c, mloc);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int oak_launch_block_starts(cudaStream_t st, int64_t n, int rows_per_block, int nblocks, int64_t *zstart) {
  k_block_starts<<<(nblocks + 1 + 255) / 256, 256, 0, st>>>(n, rows_per_block, nblocks, zstart);
  CUDA_TRY(cudaGetLastError());
  return 0;
}
