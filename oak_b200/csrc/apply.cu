// apply.cu — applies the per-zone transform to the zone's rows of the ensemble anomalies and mean:
//     xa(I_z)   = xf(I_z) + Sf(I_z,:) ampl          rrsqrt.F90:151, :462
//     Sa(I_z,:) = Sf(I_z,:) T                        rrsqrt.F90:185
// Zones without relevant observations keep Sa = Sf, xa = xf (rrsqrt.F90:322-326,:370-371).
//
// Layout: Sf/Sa are n x N column-major (member-major, leading dimension ld); the rows of a zone are
// contiguous inside every member column, so each warp-wide load/store covers one contiguous run of
// the zone's rows of one member.  T (row-major, NP x NP, written by the eig kernel) is pulled into
// shared memory with one bulk asynchronous copy (cp.async.bulk, TMA engine) signalled on an mbarrier
// while the threads stage the first chunk of rows.  Rows are processed in chunks of 32; a chunk is
// fully staged in shared memory before its results are stored, so Sa may alias Sf.
#include "common.cuh"

namespace {

#ifdef OAK_CUEMU  // functional CPU emulation for tests (tools/cuemu): the bulk copy completes at once
__device__ __forceinline__ void mbar_init(uint64_t *bar, int) { *bar = 0; }
__device__ __forceinline__ void mbar_fence_init() {}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *, uint32_t) {}
__device__ __forceinline__ void mbar_wait(uint64_t *, uint32_t) {}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *) { memcpy(dst, src, bytes); }
#else
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t phase) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(phase)
        : "memory");
  }
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
#endif

constexpr int RC = 32;  // rows per chunk

template <int NP>
__global__ void __launch_bounds__(128) k_apply(int N, ZoneGeom zg, int zone0, int64_t rowbase,
                                               const int32_t *__restrict__ mloc,
                                               const double *__restrict__ T, const double *__restrict__ ampl,
                                               const double *__restrict__ xf, const double *Sf, int64_t ldS,
                                               double *__restrict__ xa, double *Sa, int64_t ldSa,
                                               const PeerOut P, const int32_t *__restrict__ only_flagged,
                                               int64_t tstride, int astride) {
  extern __shared__ __align__(128) double sm[];
  double *sT = sm;                   // [NP][NP] row-major: sT[k*NP + k']
  double *sS = sm + NP * NP;         // [NP][RC+?] member-major chunk: sS[k*LDS + r]
  constexpr int LDS = RC + 2;
  double *s_ampl = sS + NP * LDS;    // [NP]
  __shared__ __align__(8) uint64_t bar;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int zl = blockIdx.x;
  const int zone = zone0 + zl;
  const int64_t i1 = zg.zstart[zone] - rowbase;
  const int nrow = (int)(zg.zstart[zone + 1] - zg.zstart[zone]);
  const bool analysed = !mloc || mloc[zone] != 0;  // mloc == NULL: row blocks of the global scheme, all analysed
  if (nrow <= 0) return;
  if (only_flagged && analysed && only_flagged[zl] == 0) return;  // already updated by the fused transform kernel
  const int64_t ip = P.row0 + zg.zstart[zone];  // first row of the zone in the peers' (global) arrays

  if (!analysed) {
    // zone keeps the forecast
    for (int r = tid; r < nrow; r += 128) {
      const double v = xf[i1 + r];
      xa[i1 + r] = v;
      for (int d = 0; d < P.n; d++) P.xa[d][ip + r] = v;
    }
    if (Sa != Sf || P.n > 0) {
      for (int k = warp; k < N; k += 4)
        for (int r = lane; r < nrow; r += 32) {
          const double v = Sf[i1 + r + ldS * k];
          if (Sa != Sf) Sa[i1 + r + ldSa * k] = v;
          for (int d = 0; d < P.n; d++) P.Sa[d][ip + r + P.ld * k] = v;
        }
    }
    return;
  }

  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (tid == 0) {
    constexpr uint32_t bytes = NP * NP * sizeof(double);
    mbar_expect_tx(&bar, bytes);
    bulk_g2s(sT, T + (int64_t)zl * tstride, bytes, &bar);
  }
  if (tid < NP) s_ampl[tid] = ampl[(int64_t)zl * astride + tid];

  // thread tile of the chunk product: rows 4*ty.., columns 2*tx + 32*b (+1)
  const int ty = tid >> 4, tx = tid & 15;  // 8 x 16
  constexpr int CB = NP / 32;              // column pairs per thread
  bool t_ready = false;

  for (int r0 = 0; r0 < nrow; r0 += RC) {
    const int rc = min(RC, nrow - r0);
    __syncthreads();  // previous chunk fully consumed
    for (int k = warp; k < NP; k += 4) {
      double v = 0.;
      if (k < N && lane < rc) v = Sf[i1 + r0 + lane + ldS * k];
      sS[k * LDS + lane] = v;
    }
    __syncthreads();
    if (!t_ready) { mbar_wait(&bar, 0); t_ready = true; }

    double acc[4][2 * CB];
    double dm[4] = {0., 0., 0., 0.};  // Sf(rows,:) . ampl for the thread's 4 rows (the mean update rides along)
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
      for (int b = 0; b < 2 * CB; b++) acc[a][b] = 0.;
#pragma unroll 4
    for (int k = 0; k < NP; k++) {
      const double2 s01 = *reinterpret_cast<const double2 *>(sS + k * LDS + 4 * ty);
      const double2 s23 = *reinterpret_cast<const double2 *>(sS + k * LDS + 4 * ty + 2);
      const double sv[4] = {s01.x, s01.y, s23.x, s23.y};
      const double ak = s_ampl[k];
#pragma unroll
      for (int a = 0; a < 4; a++) dm[a] = fma(sv[a], ak, dm[a]);
#pragma unroll
      for (int b = 0; b < CB; b++) {
        const double2 t = *reinterpret_cast<const double2 *>(sT + k * NP + 2 * tx + 32 * b);
#pragma unroll
        for (int a = 0; a < 4; a++) {
          acc[a][2 * b] = fma(sv[a], t.x, acc[a][2 * b]);
          acc[a][2 * b + 1] = fma(sv[a], t.y, acc[a][2 * b + 1]);
        }
      }
    }
    // mean update for the chunk: the threads of column group 0 own 4 rows each
    if (tx == 0) {
#pragma unroll
      for (int a = 0; a < 4; a++) {
        const int r = 4 * ty + a;
        if (r < rc) {
          const double v = xf[i1 + r0 + r] + dm[a];
          xa[i1 + r0 + r] = v;
          for (int dd = 0; dd < P.n; dd++) P.xa[dd][ip + r0 + r] = v;
        }
      }
    }
    __syncthreads();  // all reads of sS done: reuse it to transpose the results for coalesced stores
#pragma unroll
    for (int b = 0; b < CB; b++)
#pragma unroll
      for (int e = 0; e < 2; e++) {
        const int kk = 2 * tx + 32 * b + e;
        *reinterpret_cast<double2 *>(sS + kk * LDS + 4 * ty) = make_double2(acc[0][2 * b + e], acc[1][2 * b + e]);
        *reinterpret_cast<double2 *>(sS + kk * LDS + 4 * ty + 2) = make_double2(acc[2][2 * b + e], acc[3][2 * b + e]);
      }
    __syncthreads();
    for (int k = warp; k < N; k += 4)
      if (lane < rc) {
        const double v = sS[k * LDS + lane];
        Sa[i1 + r0 + lane + ldSa * k] = v;
        // fused all-gather: the same run of rows into every peer's array (NVLink stores)
        for (int d = 0; d < P.n; d++) P.Sa[d][ip + r0 + lane + P.ld * k] = v;
      }
  }
}

template <int NP>
int launch(cudaStream_t st, int N, const ZoneGeom &zg, int zone0, int nz, int64_t rowbase, const int32_t *mloc,
           const double *T, const double *ampl, const double *xf, const double *Sf, int64_t ldS, double *xa,
           double *Sa, int64_t ldSa, const PeerOut &peers, const int32_t *only_flagged, bool shared_transform) {
  const size_t smem = sizeof(double) * (NP * NP + NP * (RC + 2) + NP);
  { int rc_ = oak_func_smem(k_apply<NP>, (size_t)((int)smem)); if (rc_) return rc_; }
  k_apply<NP><<<nz, 128, smem, st>>>(N, zg, zone0, rowbase, mloc, T, ampl, xf, Sf, ldS, xa, Sa, ldSa, peers, only_flagged,
                                       shared_transform ? 0 : (int64_t)NP * NP, shared_transform ? 0 : NP);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

}  // namespace

int oak_launch_apply(cudaStream_t st, int N, int NP, const ZoneGeom &zg, int zone0, int nz, int64_t rowbase,
                     const int32_t *mloc, const double *T, const double *ampl, const double *xf,
                     const double *Sf, int64_t ldS, double *xa, double *Sa, int64_t ldSa, const PeerOut &peers,
                     const int32_t *only_flagged, bool shared_transform) {
  if (nz <= 0) return 0;
  switch (NP) {
    case 32: return launch<32>(st, N, zg, zone0, nz, rowbase, mloc, T, ampl, xf, Sf, ldS, xa, Sa, ldSa, peers, only_flagged, shared_transform);
    case 64: return launch<64>(st, N, zg, zone0, nz, rowbase, mloc, T, ampl, xf, Sf, ldS, xa, Sa, ldSa, peers, only_flagged, shared_transform);
    case 128: return launch<128>(st, N, zg, zone0, nz, rowbase, mloc, T, ampl, xf, Sf, ldS, xa, Sa, ldSa, peers, only_flagged, shared_transform);
  }
  oak_set_error("apply: unsupported padded ensemble size %d", NP);
  return OAK_ERR_UNSUPPORTED;
}
