// apply_mma.cu — k_apply on the fp64 tensor cores (option "apply_kernel" = 1; NP = 64; NOT the default):
//     Sa(I_z,:) = Sf(I_z,:) T ,  xa(I_z) = xf(I_z) + Sf(I_z,:) ampl          rrsqrt.F90:151,:185,:462
// Same contract as k_apply (apply.cu), used where the transform is applied as a matrix: zones the fused transform
// kernel did not finish, zones with more rows than members, and the row blocks of the global scheme (global.cu),
// where one T serves 512 rows and the kernel is the whole cost of the analysis.
//
// One CTA (4 warps) per zone.  T is copied once into shared memory with a row stride of NP + 4 doubles (16-byte
// cp.async), rows are staged 32 at a time as S[r][i] with the same stride: lane 4g+t then reads the A fragment
// S[8 rb + g][i0 + t] and the B fragment T[i0 + t][8 kb + g] without bank conflicts.  Warp w owns the column blocks
// kb = 2w, 2w+1 of all four row blocks (8 accumulator tiles) plus one tile whose B operand has ampl in column 0:
// the mean update of row block w.  Results go straight from the accumulator fragments to global memory (8
// consecutive rows of a member = one 64-byte segment); a chunk is staged completely before any of its rows is
// stored, so Sa may alias Sf.
#include "common.cuh"

namespace {

constexpr int RC = 32;  // rows per chunk

template <int NP>
__global__ void __launch_bounds__(128) k_apply_mma(int N, ZoneGeom zg, int zone0, int64_t rowbase,
                                                   const int32_t *__restrict__ mloc, const double *__restrict__ T,
                                                   const double *__restrict__ ampl, const double *__restrict__ xf,
                                                   const double *Sf, int64_t ldS, double *__restrict__ xa, double *Sa,
                                                   int64_t ldSa, const int32_t *__restrict__ only_flagged,
                                                   int64_t tstride, int astride) {
  static_assert(NP == 64, "warp / tile assignment is written for 8 column blocks and 4 warps");
  constexpr int LD = NP + 4, NB = NP / 8;
  extern __shared__ __align__(16) double sm[];
  double *sT = sm;               // [NP][LD]  T[i][k]
  double *sS = sm + NP * LD;     // [RC][LD]  S[r][i]
  double *s_ampl = sS + RC * LD; // [NP]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const int zl = blockIdx.x;
  const int zone = zone0 + zl;
  const int64_t i1 = zg.zstart[zone] - rowbase;
  const int nrow = (int)(zg.zstart[zone + 1] - zg.zstart[zone]);
  const bool analysed = !mloc || mloc[zone] != 0;
  if (nrow <= 0) return;
  if (only_flagged && analysed && only_flagged[zl] == 0) return;  // already updated by the fused transform kernel

  if (!analysed) {  // zone keeps the forecast (rrsqrt.F90:322-326,:370-371)
    for (int r = tid; r < nrow; r += 128) xa[i1 + r] = xf[i1 + r];
    if (Sa != Sf) {
      for (int k = warp; k < N; k += 4)
        for (int r = lane; r < nrow; r += 32) Sa[i1 + r + ldSa * k] = Sf[i1 + r + ldS * k];
    }
    return;
  }

  {
    const double *Tz = T + (int64_t)zl * tstride;
    for (int c = tid; c < NP * (NP / 2); c += 128) {
      const int i = c / (NP / 2), k2 = c % (NP / 2);
      cp_async16(sT + i * LD + 2 * k2, Tz + i * NP + 2 * k2);
    }
    cp_async_commit();
    if (tid < NP) s_ampl[tid] = ampl[(int64_t)zl * astride + tid];
  }

  const int NK = (N + 3) & ~3;  // members >= N: S is staged as zero there
  for (int r0 = 0; r0 < nrow; r0 += RC) {
    const int rc = min(RC, nrow - r0);
    __syncthreads();  // previous chunk's fragments are out of shared memory
    // stage: one warp instruction = 8 consecutive rows of 4 members
    {
      const int rr = lane & 7, ii = lane >> 3;
      for (int q = warp; q < (RC / 8) * (NP / 4); q += 4) {
        const int a = q & (RC / 8 - 1), i = 4 * (q / (RC / 8)) + ii;
        const int r = 8 * a + rr;
        sS[r * LD + i] = (r < rc && i < N) ? Sf[i1 + r0 + r + ldS * i] : 0.;
      }
    }
    cp_async_wait<0>();
    __syncthreads();

    double acc[4][2][2], e0 = 0., e1 = 0.;  // [rb][kb - 2 warp][col]
#pragma unroll
    for (int rb = 0; rb < 4; rb++)
#pragma unroll
      for (int b = 0; b < 2; b++) acc[rb][b][0] = acc[rb][b][1] = 0.;
#pragma unroll 2
    for (int i0 = 0; i0 < NK; i0 += 4) {
      double av[4];
#pragma unroll
      for (int rb = 0; rb < 4; rb++) av[rb] = sS[(8 * rb + g) * LD + i0 + t];
      const double b0 = sT[(i0 + t) * LD + 8 * (2 * warp) + g];
      const double b1 = sT[(i0 + t) * LD + 8 * (2 * warp + 1) + g];
      const double ba = g == 0 ? s_ampl[i0 + t] : 0.;
#pragma unroll
      for (int rb = 0; rb < 4; rb++) {
        oak_dmma_m8n8k4(acc[rb][0][0], acc[rb][0][1], av[rb], b0);
        oak_dmma_m8n8k4(acc[rb][1][0], acc[rb][1][1], av[rb], b1);
      }
      // mean update of row block `warp`: av[warp] with a static register index
      const double aw = warp == 0 ? av[0] : (warp == 1 ? av[1] : (warp == 2 ? av[2] : av[3]));
      oak_dmma_m8n8k4(e0, e1, aw, ba);
    }
    (void)e1;
    // store: fragment (rb, kb) holds rows 8 rb + g, members 8 kb + 2t, 2t + 1
#pragma unroll
    for (int rb = 0; rb < 4; rb++) {
      const int r = 8 * rb + g;
      if (r < rc) {
#pragma unroll
        for (int b = 0; b < 2; b++)
#pragma unroll
          for (int e = 0; e < 2; e++) {
            const int k = 8 * (2 * warp + b) + 2 * t + e;
            if (k < N) Sa[i1 + r0 + r + ldSa * k] = acc[rb][b][e];
          }
      }
    }
    if (t == 0) {
      const int r = 8 * warp + g;
      if (r < rc) xa[i1 + r0 + r] = xf[i1 + r0 + r] + e0;
    }
  }
  (void)NB;
}

}  // namespace

// Same arguments as oak_launch_apply without the peer destinations.  NP = 64 only.
int oak_launch_apply_mma(cudaStream_t st, int N, int NP, const ZoneGeom &zg, int zone0, int nz, int64_t rowbase,
                         const int32_t *mloc, const double *T, const double *ampl, const double *xf, const double *Sf,
                         int64_t ldS, double *xa, double *Sa, int64_t ldSa, const int32_t *only_flagged,
                         bool shared_transform) {
  if (nz <= 0) return 0;
  if (NP != 64) { oak_set_error("apply_mma: padded ensemble size %d (only 64)", NP); return OAK_ERR_UNSUPPORTED; }
  constexpr int NPc = 64;
  const size_t smem = sizeof(double) * (NPc * (NPc + 4) + RC * (NPc + 4) + NPc);
  { int rc_ = oak_func_smem(k_apply_mma<NPc>, (size_t)((int)smem)); if (rc_) return rc_; }
  k_apply_mma<NPc><<<nz, 128, smem, st>>>(N, zg, zone0, rowbase, mloc, T, ampl, xf, Sf, ldS, xa, Sa, ldSa, only_flagged,
                                          shared_transform ? 0 : (int64_t)NPc * NPc, shared_transform ? 0 : NPc);
  CUDA_TRY(cudaGetLastError());
  return 0;
}
