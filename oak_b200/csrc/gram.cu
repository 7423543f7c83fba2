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

namespace {

constexpr int GRAM_CH = 64;    // candidates examined per chunk (warp 0, two per lane)
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

template <int NP, int NT>
__global__ void __launch_bounds__(NT) k_gram(ZoneGeom zg, ObsGrid og, ObsRows orows, int zone0, int nz,
                                             double *__restrict__ G, double *__restrict__ cvec,
                                             int32_t *__restrict__ mloc, DevCounters *ctr) {
  constexpr int TX = NT / 16;
  constexpr int RT = NP / 16;
  constexpr int CT = NP / TX;
  static_assert(CT >= 2 && CT % 2 == 0, "column tile must be pairs");
  extern __shared__ __align__(16) double rowbuf[];  // [GRAM_CH][NP]
  __shared__ double s_coef[GRAM_CH], s_cd[GRAM_CH];
  __shared__ int s_pos[GRAM_CH];
  __shared__ int s_rstart[GRAM_MAXR], s_rlen[GRAM_MAXR];
  __shared__ int s_nrel, s_total;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int zl = blockIdx.x;
  if (zl >= nz) return;
  const int zone = zone0 + zl;
  const ZoneQuery q = oak_zone_query(zg, zone);
  const CellBox box = oak_zone_box(og, q);
  const int ty = tid / TX, tx = tid % TX;

  double acc[RT][CT];
#pragma unroll
  for (int a = 0; a < RT; a++)
#pragma unroll
    for (int b = 0; b < CT; b++) acc[a][b] = 0.;
  double cacc = 0.;
  int nrel_total = 0;
  long long ncand_total = 0;

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

    for (int chunk0 = 0; chunk0 < total; chunk0 += GRAM_CH) {
      if (warp == 0) {
        int nrel = 0;
#pragma unroll
        for (int h = 0; h < GRAM_CH / 32; h++) {
          int qq = chunk0 + h * 32 + lane;
          const bool valid = qq < total;
          bool rel = false;
          double w = 0.;
          int p = 0;
          if (valid) {
            int r = 0;
            while (qq >= s_rlen[r]) { qq -= s_rlen[r]; r++; }
            p = s_rstart[r] + qq;
            rel = oak_obs_relevant(q, og.sx[p], og.sy[p], w);
          }
          const unsigned bal = __ballot_sync(0xffffffffu, rel);
          if (rel) {
            const int slot = nrel + __popc(bal & ((1u << lane) - 1u));
            const double coef = (w * w) * orows.scoef[p];
            s_pos[slot] = p;
            s_coef[slot] = coef;
            s_cd[slot] = coef * orows.delta[p];
          }
          nrel += __popc(bal);
        }
        if (lane == 0) s_nrel = nrel;
      }
      __syncthreads();
      const int nrel = s_nrel;
      nrel_total += nrel;
      for (int r = warp; r < nrel; r += NT / 32) {
        const double *src = orows.rows + (int64_t)s_pos[r] * NP;
        for (int cidx = lane * 2; cidx < NP; cidx += 64)
          *reinterpret_cast<double2 *>(&rowbuf[r * NP + cidx]) = *reinterpret_cast<const double2 *>(src + cidx);
      }
      __syncthreads();
#pragma unroll 2
      for (int r = 0; r < nrel; r++) {
        const double coef = s_coef[r];
        const double *row = rowbuf + r * NP;
        double rv[RT], cv[CT];
        lds_vec<RT>(rv, row + RT * ty);
#pragma unroll
        for (int b = 0; b < CT / 2; b++) lds_vec<2>(cv + 2 * b, row + 2 * tx + 2 * TX * b);
#pragma unroll
        for (int a = 0; a < RT; a++) {
          const double ra = rv[a] * coef;
#pragma unroll
          for (int b = 0; b < CT; b++) acc[a][b] = fma(ra, cv[b], acc[a][b]);
        }
        if (tid < NP) cacc = fma(s_cd[r], row[tid], cacc);
      }
      __syncthreads();
    }
  }

  double *Gz = G + (int64_t)zl * NP * NP;
#pragma unroll
  for (int a = 0; a < RT; a++)
#pragma unroll
    for (int b = 0; b < CT / 2; b++) {
      const int i = RT * ty + a;
      const int j = 2 * tx + 2 * TX * b;
      Gz[i + NP * j] = acc[a][2 * b];
      Gz[i + NP * (j + 1)] = acc[a][2 * b + 1];
    }
  if (tid < NP) cvec[(int64_t)zl * NP + tid] = cacc;
  if (tid == 0) {
    mloc[zone] = nrel_total;
    atomicAdd(&ctr->relevant, (unsigned long long)nrel_total);
    atomicAdd(&ctr->candidates, (unsigned long long)ncand_total);
    if (nrel_total == 0) atomicAdd(&ctr->skipped, 1ull);
  }
}

template <int NP, int NT>
int launch(cudaStream_t st, const ZoneGeom &zg, const ObsGrid &og, const ObsRows &orows, int zone0, int nz,
           double *G, double *c, int32_t *mloc, DevCounters *ctr) {
  const size_t smem = sizeof(double) * GRAM_CH * NP;
  static bool attr_done = false;
  if (!attr_done) {
    CUDA_TRY(cudaFuncSetAttribute(k_gram<NP, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  k_gram<NP, NT><<<nz, NT, smem, st>>>(zg, og, orows, zone0, nz, G, c, mloc, ctr);
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
    case 64: return launch<64, 128>(st, zg, og, orows, zone0, nz, G, c, mloc, ctr);
    case 128: return launch<128, 256>(st, zg, og, orows, zone0, nz, G, c, mloc, ctr);
  }
  oak_set_error("gram: unsupported padded ensemble size %d", NP);
  return OAK_ERR_UNSUPPORTED;
}
