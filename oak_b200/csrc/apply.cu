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
#ifndef OAK_CUEMU
#include <cuda.h>          // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint)
#endif

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

#ifndef OAK_CUEMU
// 2-D tiled tensor copies (TMA): box = (rows of the zone) x (members), see k_apply_tma
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *map, int c0, int c1, const void *src) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];"
               ::"l"(map), "r"(c0), "r"(c1), "r"(smem_u32(src)) : "memory");
}
__device__ __forceinline__ void tma_store_commit_wait() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
#endif

constexpr int RC = 32;  // rows per chunk

// ---- ensemble prologue / epilogue on a staged chunk (EnsFuse, common.cuh): sS[k*lds + r], r < rc rows, k < N members.
// Same operations, in the same order, as k_mean_anom / k_epilogue (ensemble.cu), which remain the unfused form.
// All threads of the block call both (they contain block barriers).
template <int NT>
__device__ __forceinline__ void ens_prologue(const EnsFuse &F, double *sS, int lds, int N, int rc, int64_t grow0,
                                             double *s_mean, int tid) {
  const int r = tid & 31;   // a chunk has at most 32 rows: lane = row, warps stride over the members
  if (F.anamtype != 1) {
    if (r < rc)
      for (int k = tid >> 5; k < N; k += NT / 32)
        sS[k * lds + r] = oak_anam_row(F.anamtype, true, F.at, grow0 + r, sS[k * lds + r]);
    __syncthreads();
  }
  if (tid < rc) {
    double s = 0.;
    for (int k = 0; k < N; k++) s = __dadd_rn(s, sS[k * lds + tid]);   // sum(Sf,2)/N  assimilation.F90:3127
    s = s / (double)N;
    s_mean[tid] = s;
    F.xf_out[grow0 + tid] = s;
  }
  __syncthreads();
  if (r < rc) {
    const double mu = s_mean[r];
    for (int k = tid >> 5; k < N; k += NT / 32) sS[k * lds + r] = __ddiv_rn(__dsub_rn(sS[k * lds + r], mu), F.scaling);   // :3130
  }
  __syncthreads();
}

// s_x[r]: analysed mean of row r on entry (xf + Sf ampl); on exit the mean of the back-transformed members
template <int NT>
__device__ __forceinline__ void ens_epilogue(const EnsFuse &F, double *sS, int lds, int N, int rc, int64_t grow0,
                                             const double *s_mean, double *s_x, int tid) {
  const int r = tid & 31;
  if (tid < rc && F.maxCorr) {   // the two `where` statements of assimilation.F90:3311-3312
    double x = s_x[tid];
    const double mc = F.maxCorr[grow0 + tid], f = s_mean[tid];
    if (__dsub_rn(x, mc) > f) x = __dadd_rn(f, mc);
    if (x < __dsub_rn(f, mc)) x = __dsub_rn(f, mc);
    s_x[tid] = x;
  }
  __syncthreads();
  if (r < rc) {
    const double x = s_x[r];
    for (int k = tid >> 5; k < N; k += NT / 32) {
      double s = sS[k * lds + r];
      if (F.inflation != 1.) s = __dmul_rn(s, F.inflation);                                     // :3301-3304
      sS[k * lds + r] = oak_anam_row(F.anamtype, false, F.at, grow0 + r, __dadd_rn(x, __dmul_rn(s, F.scaling)));  // :3318-3326
    }
  }
  __syncthreads();
  if (tid < rc) {
    double sum = 0.;
    for (int k = 0; k < N; k++) sum = __dadd_rn(sum, sS[k * lds + tid]);
    s_x[tid] = sum / (double)N;   // xa = sum(Ea,2)/N of the back-transformed ensemble (:3343-3349)
  }
  __syncthreads();
}

template <int NP>
__global__ void __launch_bounds__(128) k_apply(int N, ZoneGeom zg, int zone0, int64_t rowbase,
                                               const int32_t *__restrict__ mloc,
                                               const double *__restrict__ T, const double *__restrict__ ampl,
                                               const double *__restrict__ xf, const double *Sf, int64_t ldS,
                                               double *__restrict__ xa, double *Sa, int64_t ldSa,
                                               const PeerOut P, const int32_t *__restrict__ only_flagged,
                                               int64_t tstride, int astride, const EnsFuse F) {
  extern __shared__ __align__(128) double sm[];
  double *sT = sm;                   // [NP][NP] row-major: sT[k*NP + k']
  double *sS = sm + NP * NP;         // [NP][RC+?] member-major chunk: sS[k*LDS + r]
  constexpr int LDS = RC + 2;
  double *s_ampl = sS + NP * LDS;    // [NP]
  __shared__ __align__(8) uint64_t bar;
  __shared__ double s_mean[RC], s_x[RC];   // ensemble form (F.on): forecast mean and analysed mean of the chunk's rows

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int zl = blockIdx.x;
  const int zone = zone0 + zl;
  const int64_t i1 = zg.zstart[zone] - rowbase;
  const int nrow = (int)(zg.zstart[zone + 1] - zg.zstart[zone]);
  const bool analysed = !mloc || mloc[zone] != 0;  // mloc == NULL: row blocks of the global scheme, all analysed
  if (nrow <= 0) return;
  if (only_flagged && analysed && only_flagged[zl] == 0) return;  // already updated by the fused transform kernel
  const int64_t ip = P.row0 + zg.zstart[zone];  // first row of the zone in the peers' (global) arrays

  if (!analysed && !F.on) {
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
  if (tid == 0 && analysed) {
    constexpr uint32_t bytes = NP * NP * sizeof(double);
    mbar_expect_tx(&bar, bytes);
    bulk_g2s(sT, T + (int64_t)zl * tstride, bytes, &bar);
  }
  if (tid < NP && analysed) s_ampl[tid] = ampl[(int64_t)zl * astride + tid];
  const int64_t grow = rowbase + i1;   // global (zone-permuted) index of the zone's first row: per-row arrays of the ensemble form

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
    if (F.on) ens_prologue<128>(F, sS, LDS, N, rc, grow + r0, s_mean, tid);
    if (!analysed) {   // ensemble form only: the zone keeps the forecast (Sa = Sf, xa = xf) and still goes through the epilogue
      if (tid < rc) s_x[tid] = s_mean[tid];
      __syncthreads();
    } else {
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
          if (F.on) { s_x[r] = s_mean[r] + dm[a]; continue; }
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
    }  // analysed
    if (F.on) {
      ens_epilogue<128>(F, sS, LDS, N, rc, grow + r0, s_mean, s_x, tid);
      if (tid < rc) {
        const double v = s_x[tid];
        xa[i1 + r0 + tid] = v;
        for (int dd = 0; dd < P.n; dd++) P.xa[dd][ip + r0 + tid] = v;
      }
    }
    for (int k = warp; k < N; k += 4)
      if (lane < rc) {
        const double v = sS[k * LDS + lane];
        Sa[i1 + r0 + lane + ldSa * k] = v;
        // fused all-gather: the same run of rows into every peer's array (NVLink stores)
        for (int d = 0; d < P.n; d++) P.Sa[d][ip + r0 + lane + P.ld * k] = v;
      }
  }
}


#ifndef OAK_CUEMU
// ---------------------------------------------------------------------------------------------------------
// k_apply_tma : the same update with the zone's rows staged by the TMA engine.  A zone of `nrow` rows is a box of
// nrow x N doubles of the member-major state (row stride ld*8 bytes): ONE cp.async.bulk.tensor.2d brings it into
// shared memory ([member][row], dense) and one brings the result back, both issued by a single thread next to the
// bulk copy of T, all three signalled on / fenced by the same mbarrier and bulk group.  The per-thread 8-byte
// loads and stores of k_apply (30 of 32 lanes active on 240-byte runs that straddle sectors; long-scoreboard
// stalls 35 % in ncu) disappear.  Conditions (checked by the launcher, otherwise k_apply): every zone has the same
// even number of rows <= 32, ld even, 16-byte aligned arrays, no peer stores.
// ---------------------------------------------------------------------------------------------------------
template <int NP>
__global__ void __launch_bounds__(128) k_apply_tma(int N, int nrow, ZoneGeom zg, int zone0, int64_t rowbase,
                                                   const int32_t *__restrict__ mloc, const double *__restrict__ T,
                                                   const double *__restrict__ ampl, const double *__restrict__ xf,
                                                   double *__restrict__ xa, const __grid_constant__ CUtensorMap mapS,
                                                   const __grid_constant__ CUtensorMap mapA, int in_place,
                                                   const int32_t *__restrict__ only_flagged, const EnsFuse F) {
  extern __shared__ __align__(128) double sm[];
  double *sT = sm;                   // [NP][NP] row-major
  double *sS = sm + NP * NP;         // [N][nrow] dense box (+ slack for the tile reads of rows >= nrow)
  double *s_ampl = sS + NP * RC + 8;
  __shared__ __align__(8) uint64_t bar;
  __shared__ double s_mean[RC], s_x[RC];   // ensemble form (F.on), as in k_apply

  const int tid = threadIdx.x;
  const int zl = blockIdx.x;
  const int zone = zone0 + zl;
  const int64_t i1 = zg.zstart[zone] - rowbase;
  const bool analysed = mloc[zone] != 0;
  if (only_flagged && analysed && only_flagged[zl] == 0) return;
  if (!analysed && in_place && !F.on) {        // the zone keeps the forecast: only the mean has to be copied
    for (int r = tid; r < nrow; r += 128) xa[i1 + r] = xf[i1 + r];
    return;
  }
  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (tid == 0) {
    const uint32_t tbytes = analysed ? NP * NP * sizeof(double) : 0u;
    mbar_expect_tx(&bar, tbytes + (uint32_t)(nrow * N * sizeof(double)));
    if (analysed) bulk_g2s(sT, T + (int64_t)zl * NP * NP, tbytes, &bar);
    tma_load_2d(sS, &mapS, (int)i1, 0, &bar);
  }
  if (tid < NP) s_ampl[tid] = analysed ? ampl[(int64_t)zl * NP + tid] : 0.;
  __syncthreads();                    // s_ampl
  mbar_wait(&bar, 0);
  const int ty = tid >> 4, tx = tid & 15;  // 8 x 16: rows 4 ty .., columns 2 tx + 32 b (+1)
  constexpr int CB = NP / 32;
  const int64_t grow = rowbase + i1;
  if (F.on) ens_prologue<128>(F, sS, nrow, N, nrow, grow, s_mean, tid);
  if (analysed) {
    double acc[4][2 * CB];
    double dm[4] = {0., 0., 0., 0.};
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
      for (int b = 0; b < 2 * CB; b++) acc[a][b] = 0.;
#pragma unroll 4
    for (int k = 0; k < N; k++) {
      const double2 s01 = *reinterpret_cast<const double2 *>(sS + k * nrow + 4 * ty);
      const double2 s23 = *reinterpret_cast<const double2 *>(sS + k * nrow + 4 * ty + 2);
      const double sv[4] = {s01.x, s01.y, s23.x, s23.y};   // rows >= nrow: the next member's values, never stored
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
    if (tx == 0) {
#pragma unroll
      for (int a = 0; a < 4; a++) {
        const int r = 4 * ty + a;
        if (r < nrow) {
          if (F.on) s_x[r] = s_mean[r] + dm[a];
          else xa[i1 + r] = xf[i1 + r] + dm[a];
        }
      }
    }
    __syncthreads();                  // all reads of the box done: it now receives the results
#pragma unroll
    for (int b = 0; b < CB; b++)
#pragma unroll
      for (int e = 0; e < 2; e++) {
        const int kk = 2 * tx + 32 * b + e;
        if (kk < N) {
#pragma unroll
          for (int a = 0; a < 4; a++)
            if (4 * ty + a < nrow) sS[kk * nrow + 4 * ty + a] = acc[a][2 * b + e];
        }
      }
  } else if (F.on) {
    if (tid < nrow) s_x[tid] = s_mean[tid];
  } else {
    for (int r = tid; r < nrow; r += 128) xa[i1 + r] = xf[i1 + r];
  }
  if (F.on) {
    __syncthreads();                  // results / s_x of all threads
    ens_epilogue<128>(F, sS, nrow, N, nrow, grow, s_mean, s_x, tid);
    if (tid < nrow) xa[i1 + tid] = s_x[tid];
  }
  fence_proxy_async();                // generic-proxy writes of the box before the async-proxy (TMA) read
  __syncthreads();
  if (tid == 0) {
    tma_store_2d(&mapA, (int)i1, 0, sS);
    tma_store_commit_wait();          // the box must stay valid until the engine has read it
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess) { cudaGetLastError(); return nullptr; }
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// rows x N doubles, row stride ld; box nrow x N
bool make_state_map(CUtensorMap *map, const double *base, int64_t rows, int64_t ld, int N, int nrow) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)rows, (cuuint64_t)N};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(double)};
  const cuuint32_t box[2] = {(cuuint32_t)nrow, (cuuint32_t)N};
  const cuuint32_t estr[2] = {1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double *>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int NP>
int launch_tma(cudaStream_t st, int N, int nrow, int64_t rows_in_buffers, const ZoneGeom &zg, int zone0, int nz,
               int64_t rowbase, const int32_t *mloc, const double *T, const double *ampl, const double *xf,
               const double *Sf, int64_t ldS, double *xa, double *Sa, int64_t ldSa, const int32_t *only_flagged,
               const EnsFuse &ens, bool *done) {
  *done = false;
  CUtensorMap mS, mA;
  if (!make_state_map(&mS, Sf, rows_in_buffers, ldS, N, nrow) || !make_state_map(&mA, Sa, rows_in_buffers, ldSa, N, nrow))
    return 0;   // the caller falls back to k_apply
  const size_t smem = sizeof(double) * (NP * NP + NP * RC + 8 + NP);
  { int rc_ = oak_func_smem(k_apply_tma<NP>, smem); if (rc_) return rc_; }
  k_apply_tma<NP><<<nz, 128, smem, st>>>(N, nrow, zg, zone0, rowbase, mloc, T, ampl, xf, xa, mS, mA, Sa == Sf ? 1 : 0,
                                        only_flagged, ens);
  CUDA_TRY(cudaGetLastError());
  *done = true;
  return 0;
}
#endif  // OAK_CUEMU

template <int NP>
int launch(cudaStream_t st, int N, const ZoneGeom &zg, int zone0, int nz, int64_t rowbase, const int32_t *mloc,
           const double *T, const double *ampl, const double *xf, const double *Sf, int64_t ldS, double *xa,
           double *Sa, int64_t ldSa, const PeerOut &peers, const int32_t *only_flagged, bool shared_transform,
           const EnsFuse &ens) {
  const size_t smem = sizeof(double) * (NP * NP + NP * (RC + 2) + NP);
  { int rc_ = oak_func_smem(k_apply<NP>, (size_t)((int)smem)); if (rc_) return rc_; }
  k_apply<NP><<<nz, 128, smem, st>>>(N, zg, zone0, rowbase, mloc, T, ampl, xf, Sf, ldS, xa, Sa, ldSa, peers, only_flagged,
                                       shared_transform ? 0 : (int64_t)NP * NP, shared_transform ? 0 : NP, ens);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

}  // namespace

int oak_launch_apply(cudaStream_t st, int N, int NP, const ZoneGeom &zg, int zone0, int nz, int64_t rowbase,
                     const int32_t *mloc, const double *T, const double *ampl, const double *xf,
                     const double *Sf, int64_t ldS, double *xa, double *Sa, int64_t ldSa, const PeerOut &peers,
                     const int32_t *only_flagged, bool shared_transform, int uniform_rows, int64_t rows_in_buffers,
                     const EnsFuse *ens_in) {
  if (nz <= 0) return 0;
  EnsFuse ens{};
  if (ens_in) ens = *ens_in;
  if (ens.on && (shared_transform || !mloc)) { oak_set_error("apply: the ensemble form is for the local scheme"); return OAK_ERR_ARG; }
#ifndef OAK_CUEMU
  // TMA-staged state (k_apply_tma) where a zone is one box: equal, even zone sizes <= 32, even leading dimensions,
  // 16-byte aligned arrays, local scheme, no stores to peers from the kernel
  if (uniform_rows > 0 && uniform_rows <= RC && (uniform_rows & 1) == 0 && !shared_transform && peers.n == 0 && mloc &&
      (ldS & 1) == 0 && (ldSa & 1) == 0 && ((uintptr_t)Sf & 15) == 0 && ((uintptr_t)Sa & 15) == 0 && N >= 2 &&
      rows_in_buffers > 0 && rows_in_buffers < 0x7fffffffll && (NP == 32 || NP == 64)) {
    bool done = false;
    int rc = NP == 64 ? launch_tma<64>(st, N, uniform_rows, rows_in_buffers, zg, zone0, nz, rowbase, mloc, T, ampl, xf, Sf, ldS, xa, Sa, ldSa, only_flagged, ens, &done)
                      : launch_tma<32>(st, N, uniform_rows, rows_in_buffers, zg, zone0, nz, rowbase, mloc, T, ampl, xf, Sf, ldS, xa, Sa, ldSa, only_flagged, ens, &done);
    if (rc || done) return rc;
  }
#endif
  switch (NP) {
    case 32: return launch<32>(st, N, zg, zone0, nz, rowbase, mloc, T, ampl, xf, Sf, ldS, xa, Sa, ldSa, peers, only_flagged, shared_transform, ens);
    case 64: return launch<64>(st, N, zg, zone0, nz, rowbase, mloc, T, ampl, xf, Sf, ldS, xa, Sa, ldSa, peers, only_flagged, shared_transform, ens);
    case 128: return launch<128>(st, N, zg, zone0, nz, rowbase, mloc, T, ampl, xf, Sf, ldS, xa, Sa, ldSa, peers, only_flagged, shared_transform, ens);
  }
  oak_set_error("apply: unsupported padded ensemble size %d", NP);
  return OAK_ERR_UNSUPPORTED;
}
