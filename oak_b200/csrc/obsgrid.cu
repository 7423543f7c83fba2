// obsgrid.cu — device cell-grid bucketing of the observations and the observation
// selection kernels.
//
// Replaces the O(m) scan per zone of selectObservations (assimilation.F90:3745-3757) and
// the cellgrid/setupgrid/near neighbour search (ndgrid.F90:1489-1691): observations are
// sorted by cell (stable: increasing observation number inside a cell), a zone visits the
// cells of its conservative box and applies the exact predicate to each candidate.
#include <cub/cub.cuh>

#include "common.cuh"

// ---------------------------------------------------------------------------------------
// cell ids
// ---------------------------------------------------------------------------------------
__global__ void k_cell_ids(int m, const double *bx, const double *by, ObsGrid g, uint32_t *key,
                           int32_t *val, int32_t *hist) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= m) return;
  double x = bx[l];
  if (g.wrap_x) x = oak_fold360(x);
  const int cx = oak_cell_of(x, g.x0, g.csx, g.ncx);
  const int cy = by ? oak_cell_of(by[l], g.y0, g.csy, g.ncy) : 0;
  const uint32_t c = (uint32_t)cy * (uint32_t)g.ncx + (uint32_t)cx;
  key[l] = c;
  val[l] = l;
  atomicAdd(&hist[c], 1);
}

__global__ void k_gather_coords(int m, const int32_t *perm, const double *bx, const double *by,
                                double *sx, double *sy) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= m) return;
  const int l = perm[p];
  sx[p] = bx[l];
  sy[p] = by ? by[l] : 0.;
}

// Builds perm / cell_start / sorted coordinates. All pointers are device pointers; key/val/hist
// and the CUB temp storage are caller-provided scratch.
int oak_build_obsgrid(cudaStream_t st, int m, const double *bx, const double *by, const ObsGrid &g,
                      uint32_t *key_in, uint32_t *key_out, int32_t *val_in, int32_t *perm,
                      int32_t *cell_start, double *sx, double *sy, void *tmp, size_t tmp_bytes) {
  const int ncell = g.ncx * g.ncy;
  CUDA_TRY(cudaMemsetAsync(cell_start, 0, sizeof(int32_t) * (ncell + 1), st));
  if (m == 0) return 0;
  const int nb = (m + 255) / 256;
  k_cell_ids<<<nb, 256, 0, st>>>(m, bx, by, g, key_in, val_in, cell_start);
  CUDA_TRY(cudaGetLastError());
  int bits = 1;
  while ((1ll << bits) < ncell) bits++;
  size_t need = 0;
  CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, need, key_in, key_out, val_in, perm, m, 0, bits, st));
  if (need > tmp_bytes) { oak_set_error("obsgrid: scratch too small (%zu > %zu)", need, tmp_bytes); return OAK_ERR_NOMEM; }
  CUDA_TRY(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, key_in, key_out, val_in, perm, m, 0, bits, st));
  size_t need2 = 0;
  CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, need2, cell_start, cell_start, ncell + 1, st));
  if (need2 > tmp_bytes) { oak_set_error("obsgrid: scan scratch too small"); return OAK_ERR_NOMEM; }
  CUDA_TRY(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, cell_start, cell_start, ncell + 1, st));
  k_gather_coords<<<nb, 256, 0, st>>>(m, perm, bx, by, sx, sy);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

size_t oak_obsgrid_scratch_bytes(int m, int ncell) {
  size_t a = 0, b = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, a, (uint32_t *)nullptr, (uint32_t *)nullptr, (int32_t *)nullptr,
                                  (int32_t *)nullptr, m > 0 ? m : 1);
  cub::DeviceScan::ExclusiveSum(nullptr, b, (int32_t *)nullptr, (int32_t *)nullptr, ncell + 1);
  return (a > b ? a : b) + 256;
}

// ---------------------------------------------------------------------------------------
// k_pack_obs: observation-space arrays into sorted, row-major, padded form.
//   rows[p][k] = HSf(perm[p], k)   (HSf column-major m x N, ld ldH)  k < N ; 0 for N <= k < NP
//   delta[p]   = yo - Hxf          (rrsqrt.F90:142: R%mldivide(yo-Hxf))
//   scoef[p]   = d01^2 / Rdiag     (covariance.F90:431,:618)
// One CTA handles 32 sorted positions; a 32 x 33 shared tile turns the gather (strided by ldH
// in k) into coalesced row writes.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_pack_obs(int m, int N, int NP, const int32_t *perm,
                                                  const double *HSf, int64_t ldH, const double *yo,
                                                  const double *Hxf, const double *Rdiag,
                                                  const double *d01, double *rows, double *delta,
                                                  double *scoef) {
  __shared__ double tile[32][33];
  __shared__ int32_t sl[32];
  const int p0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 8 warps
  if (threadIdx.x < 32) {
    const int p = p0 + threadIdx.x;
    int l = -1;
    if (p < m) {
      l = perm[p];
      delta[p] = yo[l] - Hxf[l];
      const double e = d01 ? d01[l] : 1.;
      scoef[p] = (e * e) / Rdiag[l];
    }
    sl[threadIdx.x] = l;
  }
  __syncthreads();
  for (int k0 = 0; k0 < NP; k0 += 32) {
    // read: lane = position, warp-row = member
    for (int kk = ty; kk < 32; kk += 8) {
      const int k = k0 + kk;
      const int l = sl[tx];
      tile[kk][tx] = (l >= 0 && k < N) ? HSf[l + ldH * (int64_t)k] : 0.;
    }
    __syncthreads();
    // write: lane = member, warp-row = position
    for (int pp = ty; pp < 32; pp += 8) {
      const int p = p0 + pp;
      if (p < m && k0 + tx < NP) rows[(int64_t)p * NP + k0 + tx] = tile[tx][pp];
    }
    __syncthreads();
  }
}

int oak_launch_pack_obs(cudaStream_t st, int m, int N, int NP, const int32_t *perm, const double *HSf,
                        int64_t ldH, const double *yo, const double *Hxf, const double *Rdiag,
                        const double *d01, double *rows, double *delta, double *scoef) {
  if (m == 0) return 0;
  k_pack_obs<<<(m + 31) / 32, 256, 0, st>>>(m, N, NP, perm, HSf, ldH, yo, Hxf, Rdiag, d01, rows, delta, scoef);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------
// k_select: the selectObservations callback materialised (one warp per zone).
//   fill = false: counts[z] = number of relevant observations
//   fill = true : idx/w written at offsets[z] in traversal order (the host entry point sorts each
//                 zone's list by observation number, the order pack() gives in rrsqrt.F90:395-404)
// ---------------------------------------------------------------------------------------
template <bool FILL>
__global__ void __launch_bounds__(128) k_select(ZoneGeom zg, ObsGrid og, int zone0, int nz,
                                                const int64_t *offsets, int32_t *counts, int32_t *idx,
                                                double *wout) {
  const int wz = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (wz >= nz) return;
  const int zone = zone0 + wz;
  const ZoneQuery q = oak_zone_query(zg, zone);
  const CellBox b = oak_zone_box(og, q);
  int count = 0;
  const int64_t base = FILL ? offsets[wz] : 0;
  for (int cy = b.cy0; cy <= b.cy1; cy++) {
    for (int rr = 0; rr < 2; rr++) {
      const int x0 = rr ? b.xb0 : b.xa0, x1 = rr ? b.xb1 : b.xa1;
      if (x0 > x1) continue;
      const int s = og.cell_start[cy * og.ncx + x0], e = og.cell_start[cy * og.ncx + x1 + 1];
      for (int p0 = s; p0 < e; p0 += 32) {
        const int p = p0 + lane;
        bool rel = false;
        double w = 0.;
        if (p < e) rel = oak_obs_relevant(q, og.sx[p], og.sy[p], w);
        const unsigned bal = __ballot_sync(0xffffffffu, rel);
        if (FILL && rel) {
          const int64_t o = base + count + __popc(bal & ((1u << lane) - 1u));
          idx[o] = og.perm[p] + 1;
          wout[o] = w;
        }
        count += __popc(bal);
      }
    }
  }
  if (!FILL && lane == 0) counts[wz] = count;
}

int oak_launch_select(cudaStream_t st, const ZoneGeom &zg, const ObsGrid &og, int zone0, int nz,
                      const int64_t *offsets, int32_t *counts, int32_t *idx, double *w, bool fill) {
  if (nz == 0) return 0;
  const int nb = (nz + 3) / 4;
  if (fill)
    k_select<true><<<nb, 128, 0, st>>>(zg, og, zone0, nz, offsets, counts, idx, w);
  else
    k_select<false><<<nb, 128, 0, st>>>(zg, og, zone0, nz, offsets, counts, idx, w);
  CUDA_TRY(cudaGetLastError());
  return 0;
}
