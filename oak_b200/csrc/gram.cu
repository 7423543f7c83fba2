// gram.cu — per-zone observation selection fused with the weighted Gram matrix.
//
// For zone z (one CTA):   L = { l : relevant_l }               assimilation.F90:3745-3757
//                         G = sum_{l in L} coef_l a_l a_l^T    rrsqrt.F90:135  (HSf^T R_loc^-1 HSf)
//                         c = sum_{l in L} coef_l delta_l a_l  rrsqrt.F90:142  (HSf^T R_loc^-1 (yo-Hxf))
// with a_l = HSf(l,:), coef_l = w_l^2 d01_l^2 / R_ll  (covariance.F90:612-619, :425-431) and
// delta_l = yo_l - Hxf_l.  The candidates come from the cell grid (obsgrid.cu); the exact predicate
// decides, so the set L is the reference's.  The relevant rows are staged in shared memory
// (coalesced 16-byte loads of the row-major packed rows) and accumulated into a register tile:
// thread (ty,tx) of a 16 x TX grid owns rows RT*ty.. and the column pairs 2tx+2TX*b (conflict-free
// 16-byte shared loads).
#include "common.cuh"

#ifndef GRAM_SYM
#define GRAM_SYM 1   // NP = 64: accumulate only the blocks LL, HL, HH of the symmetric G (96 threads)
#endif

namespace {

constexpr int GRAM_CH = 64;    // candidates examined per chunk (warps 0 and 1, one per lane)
constexpr int GRAM_MAXR = 64;  // cell ranges per row group

template <int K>
__device__ __forceinline__ void lds_vec(double *dst, const double *src) {
  if constexpr (K % 2 == 0) {
#pragma unroll
    for (int i = 0; i < K / 2; i++) {
      const double2 v = reinterpret_cast<const double2 *>(src)[i];
      dst[2 * i] = v.x;
      dst[2 * i + 1] = v.y;
    }
  } else {
#pragma unroll
    for (int i = 0; i < K; i++) dst[i] = src[i];
  }
}

// Software pipeline over chunks of GRAM_CH candidates (per group of cell rows):
//   E(c)  warps 0,1 evaluate the exact predicate on 32 candidates each and compact the relevant ones
//         (per-warp segment of the list: position, coef = w^2 d01^2/R, coef*delta)
//   L(c)  all warps copy the relevant rows (row-major, NP doubles) into a shared row buffer with
//         16-byte cp.async (LDGSTS), one warp instruction per 512 bytes of a row
//   F(c)  every thread accumulates its register tile over the staged rows
// E runs two chunks ahead (3 list buffers), L one chunk ahead (2 row buffers), so the global-memory
// latency of the rows of chunk c+1 is hidden behind F(c).  Two barriers per chunk.
// SYM (NP = 64, 96 threads): G is symmetric, so only the blocks LL, HL, HH of the 2 x 2 partition into 32 x 32
// blocks are accumulated (one warp each, 4 x 8 register tiles on an 8 x 4 thread grid) and HL is mirrored
// into LH when G is written: 25 % fewer DFMA and shared-memory reads than the full product.
template <int NP, int NT, bool SYM>
__global__ void __launch_bounds__(NT) k_gram(ZoneGeom zg, ObsGrid og, ObsRows orows, int zone0, int nz,
                                             double *__restrict__ G, double *__restrict__ cvec,
                                             int32_t *__restrict__ mloc, DevCounters *ctr) {
  constexpr int TX = SYM ? 4 : NT / 16;
  constexpr int RT = SYM ? 4 : NP / 16;
  constexpr int CT = SYM ? 8 : NP / TX;
  constexpr int NW = NT / 32;
  static_assert(!SYM || (NP == 64 && NT == 96), "symmetric variant: N <= 64, three warps");
  static_assert(CT >= 2 && CT % 2 == 0, "column tile must be pairs");
  static_assert(GRAM_CH == 64, "two evaluating warps");
  extern __shared__ __align__(16) double rowbuf[];  // [2][GRAM_CH][NP]
  __shared__ double s_coef[3][GRAM_CH], s_cd[3][GRAM_CH];
  __shared__ int s_pos[3][GRAM_CH];
  __shared__ int s_cnt[3][2];
  __shared__ int s_rstart[GRAM_MAXR], s_rlen[GRAM_MAXR];
  __shared__ int s_total;
  __shared__ int s_true;   // localise_obs = .false.: observations that pass the relevance predicate

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int zl = blockIdx.x;
  if (zl >= nz) return;
  const int zone = zone0 + zl;
  const ZoneQuery q = oak_zone_query(zg, zone);
  const CellBox box = oak_zone_box(og, q);
  // full product: thread (ty,tx) of a 16 x TX grid; symmetric: warp = block (0: LL, 1: HL, 2: HH), lane = (ty,tx)
  // of an 8 x 4 grid inside the block
  const int ty = SYM ? (lane >> 2) : tid / TX, tx = SYM ? (lane & 3) : tid % TX;
  const int rbase = SYM ? (warp >= 1 ? 32 : 0) : 0, cbase = SYM ? (warp == 2 ? 32 : 0) : 0;

  double acc[RT][CT];
#pragma unroll
  for (int a = 0; a < RT; a++)
#pragma unroll
    for (int b = 0; b < CT; b++) acc[a][b] = 0.;
  double cacc = 0.;
  int nrel_total = 0;
  long long ncand_total = 0;
  if (threadIdx.x == 0) s_true = 0;   // ordered before its first use by the barrier at the head of the cell loop

  // E(c): warps 0 and 1, 32 candidates each, into list buffer lb
  auto eval = [&](int c, int lb, int total) {
    if (warp < 2) {
      int qq = c * GRAM_CH + warp * 32 + lane;
      bool rel = false;
      double w = 0.;
      int p = 0;
      if (qq < total) {
        int r = 0;
        while (qq >= s_rlen[r]) { qq -= s_rlen[r]; r++; }
        p = s_rstart[r] + qq;
        rel = oak_obs_relevant(q, og.sx[p], og.sy[p], w);
        if (q.noloc) {   // localise_obs = .false.: count the relevant ones, take them all
          if (rel) atomicAdd(&s_true, 1);
          rel = true;
        }
      }
      const unsigned bal = __ballot_sync(0xffffffffu, rel);
      if (rel) {
        const int slot = warp * 32 + __popc(bal & ((1u << lane) - 1u));
        const double coef = (w * w) * orows.scoef[p];
        s_pos[lb][slot] = p;
        s_coef[lb][slot] = coef;
        s_cd[lb][slot] = coef * orows.delta[p];
      }
      if (lane == 0) s_cnt[lb][warp] = __popc(bal);
    }
  };
  // L(c): rows of list buffer lb into row buffer rb
  auto load_rows = [&](int lb, int rb) {
    double *dstb = rowbuf + (size_t)rb * GRAM_CH * NP;
#pragma unroll
    for (int seg = 0; seg < 2; seg++) {
      const int cnt = s_cnt[lb][seg];
      for (int r = warp; r < cnt; r += NW) {
        const int slot = seg * 32 + r;
        const double *src = orows.rows + (int64_t)s_pos[lb][slot] * NP;
        for (int cidx = lane * 2; cidx < NP; cidx += 64) cp_async16(dstb + slot * NP + cidx, src + cidx);
      }
    }
  };
  // F(c)
  auto fma_rows = [&](int lb, int rb) {
    const double *srcb = rowbuf + (size_t)rb * GRAM_CH * NP;
#pragma unroll
    for (int seg = 0; seg < 2; seg++) {
      const int cnt = s_cnt[lb][seg];
      nrel_total += cnt;
#pragma unroll 2
      for (int r = 0; r < cnt; r++) {
        const int slot = seg * 32 + r;
        const double coef = s_coef[lb][slot];
        const double *row = srcb + slot * NP;
        double rv[RT], cv[CT];
        lds_vec<RT>(rv, row + rbase + RT * ty);
#pragma unroll
        for (int b = 0; b < CT / 2; b++) lds_vec<2>(cv + 2 * b, row + cbase + 2 * tx + 2 * TX * b);
#pragma unroll
        for (int a = 0; a < RT; a++) {
          const double ra = rv[a] * coef;
#pragma unroll
          for (int b = 0; b < CT; b++) acc[a][b] = fma(ra, cv[b], acc[a][b]);
        }
        if (tid < NP) cacc = fma(s_cd[lb][slot], row[tid], cacc);
      }
    }
  };

  for (int cyg = box.cy0; cyg <= box.cy1; cyg += GRAM_MAXR / 2) {
    __syncthreads();
    if (tid < GRAM_MAXR) {
      const int cy = cyg + (tid >> 1);
      int start = 0, len = 0;
      if (cy <= box.cy1) {
        const int x0 = (tid & 1) ? box.xb0 : box.xa0, x1 = (tid & 1) ? box.xb1 : box.xa1;
        if (x0 <= x1) {
          start = og.cell_start[cy * og.ncx + x0];
          len = og.cell_start[cy * og.ncx + x1 + 1] - start;
        }
      }
      s_rstart[tid] = start;
      s_rlen[tid] = len;
    }
    __syncthreads();
    if (tid == 0) {
      int t = 0;
      for (int r = 0; r < GRAM_MAXR; r++) t += s_rlen[r];
      s_total = t;
    }
    __syncthreads();
    const int total = s_total;
    ncand_total += total;
    const int nchunk = (total + GRAM_CH - 1) / GRAM_CH;
    if (nchunk == 0) continue;

    eval(0, 0, total);
    __syncthreads();
    load_rows(0, 0);
    cp_async_commit();
    if (nchunk > 1) eval(1, 1, total);
    __syncthreads();
    for (int c = 0; c < nchunk; c++) {
      const int lb = c % 3, rb = c & 1;
      if (c + 1 < nchunk) {
        load_rows((c + 1) % 3, rb ^ 1);
        cp_async_commit();
        cp_async_wait<1>();
      } else {
        cp_async_wait<0>();
      }
      __syncthreads();  // rows of chunk c have landed for every thread
      fma_rows(lb, rb);
      if (c + 2 < nchunk) eval(c + 2, (c + 2) % 3, total);
      __syncthreads();  // list c+2 visible; row buffer rb free for chunk c+2
    }
  }

  double *Gz = G + (int64_t)zl * NP * NP;
#pragma unroll
  for (int a = 0; a < RT; a++)
#pragma unroll
    for (int b = 0; b < CT / 2; b++) {
      const int i = rbase + RT * ty + a;
      const int j = cbase + 2 * tx + 2 * TX * b;
      Gz[i + NP * j] = acc[a][2 * b];
      Gz[i + NP * (j + 1)] = acc[a][2 * b + 1];
      if (SYM && warp == 1) {  // mirror HL into LH
        Gz[j + NP * i] = acc[a][2 * b];
        Gz[j + 1 + NP * i] = acc[a][2 * b + 1];
      }
    }
  if (tid < NP) cvec[(int64_t)zl * NP + tid] = cacc;
  if (tid == 0) {
    // localise_obs = .false.: a zone without any relevant observation is still skipped (rrsqrt.F90:371-372)
    const int used = (q.noloc && s_true == 0) ? 0 : nrel_total;
    mloc[zone] = used;
    atomicAdd(&ctr->relevant, (unsigned long long)used);
    atomicAdd(&ctr->candidates, (unsigned long long)ncand_total);
    if (used == 0) atomicAdd(&ctr->skipped, 1ull);
  }
}

template <int NP, int NT, bool SYM = false>
int launch(cudaStream_t st, const ZoneGeom &zg, const ObsGrid &og, const ObsRows &orows, int zone0, int nz,
           double *G, double *c, int32_t *mloc, DevCounters *ctr) {
  const size_t smem = sizeof(double) * 2 * GRAM_CH * NP;
  { int rc_ = oak_func_smem(k_gram<NP, NT, SYM>, (size_t)((int)smem)); if (rc_) return rc_; }
  k_gram<NP, NT, SYM><<<nz, NT, smem, st>>>(zg, og, orows, zone0, nz, G, c, mloc, ctr);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

}  // namespace

int oak_launch_gram(cudaStream_t st, int NP, const ZoneGeom &zg, const ObsGrid &og, const ObsRows &orows,
                    int zone0, int nz, double *G, double *c, int32_t *mloc, DevCounters *ctr) {
  if (nz <= 0) return 0;
  switch (NP) {
    case 16: return launch<16, 128>(st, zg, og, orows, zone0, nz, G, c, mloc, ctr);
    case 32: return launch<32, 128>(st, zg, og, orows, zone0, nz, G, c, mloc, ctr);
#if GRAM_SYM
    case 64: return launch<64, 96, true>(st, zg, og, orows, zone0, nz, G, c, mloc, ctr);
#else
    case 64: return launch<64, 128>(st, zg, og, orows, zone0, nz, G, c, mloc, ctr);
#endif
    case 128: return launch<128, 256>(st, zg, og, orows, zone0, nz, G, c, mloc, ctr);
  }
  oak_set_error("gram: unsupported padded ensemble size %d", NP);
  return OAK_ERR_UNSUPPORTED;
}
