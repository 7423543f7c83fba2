// ensemble.cu — streaming kernels of the ensemble branch of Assim around the local analysis.
//   prologue  assimilation.F90:3106-3134 : HE = H E + Hshift (COO SpMV, matoper_inc.F90:220-242),
//             forward anamorphosis, mean and scaled anomalies (scaling = sqrt(N-1), ppdef.h:52)
//   epilogue  assimilation.F90:3301-3326,:3558-3562 : inflation, saturation of the correction,
//             Ea = xa + scaling*Sa, inverse anamorphosis.
// All of them are HBM-bound passes over member-major (column-major n x N) arrays: lanes run over
// rows so every warp access is one contiguous run of a member column.
#include <cub/cub.cuh>

#include "common.cuh"

namespace {

// oak_interp1 / oak_anam / oak_anam_row: common.cuh (shared with the fused apply kernel, apply.cu)

// ---- COO -> row-sorted (stable: the entries of a row keep the caller's order, so the sum is
// accumulated in the same order as the sequential loop of matoper_inc.F90:238-240) ----
__global__ void k_coo_keys(int64_t nnz, int m, const int32_t *Hi, const int32_t *Hj, uint32_t *key,
                           int32_t *val, int32_t *rowcount) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= nnz) return;
  int i = Hi[e] - 1;
  const bool ok = i >= 0 && i < m && Hj[e] > 0;  // model index <= 0: out-of-grid observation (assimilation.F90:2597-2611)
  if (!ok) i = m;                                // parked behind the last row
  key[e] = (uint32_t)i;
  val[e] = (int32_t)e;
  if (ok) atomicAdd(&rowcount[i], 1);
}

__global__ void k_obsoper(int m, const int32_t *rowstart, const int32_t *order, const int32_t *Hj,
                          const double *Hs, const double *Hshift, const double *E, int64_t ldE, double *HE) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int k = blockIdx.y;
  if (i >= m) return;
  double acc = 0.;
  const int s = rowstart[i], e = rowstart[i + 1];
  for (int q = s; q < e; q++) {
    const int en = order[q];
    acc = __dadd_rn(acc, __dmul_rn(Hs[en], E[(int64_t)(Hj[en] - 1) + ldE * k]));
  }
  if (Hshift) acc = __dadd_rn(acc, Hshift[i]);
  HE[i + (int64_t)m * k] = acc;
}

// ---- mean + scaled anomalies, one pass: a tile of 64 rows x N members staged in shared memory ----
__global__ void __launch_bounds__(256) k_mean_anom(int64_t rows, int N, int anamtype, AnamTab at, const double *E,
                                                   int64_t ldE, double *mean, double *S, int64_t ldS,
                                                   double scaling) {
  extern __shared__ double tile[];  // [N][64]
  __shared__ double s_mean[64];
  const int64_t r0 = (int64_t)blockIdx.x * 64;
  const int tid = threadIdx.x, lr = tid & 63, kq = tid >> 6;  // 4 members in flight
  const int64_t row = r0 + lr;
  for (int k = kq; k < N; k += 4) tile[k * 64 + lr] = row < rows ? oak_anam_row(anamtype, true, at, row, E[row + ldE * k]) : 0.;
  __syncthreads();
  if (tid < 64) {
    double s = 0.;
    for (int k = 0; k < N; k++) s = __dadd_rn(s, tile[k * 64 + tid]);  // sum(Sf,2)/N  assimilation.F90:3127
    s = s / (double)N;
    s_mean[tid] = s;
    if (r0 + tid < rows) mean[r0 + tid] = s;
  }
  __syncthreads();
  if (row < rows) {
    const double mu = s_mean[lr];
    for (int k = kq; k < N; k += 4) S[row + ldS * k] = __ddiv_rn(__dsub_rn(tile[k * 64 + lr], mu), scaling);
  }
}

__global__ void k_epilogue(int64_t rows, int N, int anamtype, AnamTab at, double inflation, double scaling,
                           const double *maxCorr, const double *xf, double *xa, const double *Sa,
                           int64_t ldSa, double *Ea, int64_t ldEa) {
  const int64_t row = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (row >= rows) return;
  double x = xa[row];
  if (maxCorr) {  // the two `where` statements of assimilation.F90:3311-3312, in their order, xa untouched otherwise
    const double mc = maxCorr[row], f = xf[row];
    if (__dsub_rn(x, mc) > f) x = __dadd_rn(f, mc);
    if (x < __dsub_rn(f, mc)) x = __dsub_rn(f, mc);
  }
  double sum = 0.;
  for (int k = 0; k < N; k++) {
    double s = Sa[row + ldSa * k];
    if (inflation != 1.) s = __dmul_rn(s, inflation);               // :3301-3304
    const double ea = oak_anam_row(anamtype, false, at, row, __dadd_rn(x, __dmul_rn(s, scaling)));  // :3318-3326
    Ea[row + ldEa * k] = ea;
    sum = __dadd_rn(sum, ea);
  }
  xa[row] = sum / (double)N;   // xa = sum(Sa,2)/size(Sa,2) of the back-transformed ensemble (:3343-3349)
}

}  // namespace

// scratch requirements for the COO sort
size_t oak_coo_scratch_bytes(int64_t nnz, int m) {
  size_t a = 0, b = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, a, (uint32_t *)nullptr, (uint32_t *)nullptr, (int32_t *)nullptr,
                                  (int32_t *)nullptr, nnz > 0 ? nnz : 1);
  cub::DeviceScan::ExclusiveSum(nullptr, b, (int32_t *)nullptr, (int32_t *)nullptr, m + 2);
  return (a > b ? a : b) + 256;
}

// rowstart[m+2], order[nnz] out; key_in/key_out/val_in scratch of nnz entries
int oak_coo_to_rows(cudaStream_t st, int64_t nnz, int m, const int32_t *Hi, const int32_t *Hj, uint32_t *key_in,
                    uint32_t *key_out, int32_t *val_in, int32_t *order, int32_t *rowstart, void *tmp,
                    size_t tmp_bytes) {
  CUDA_TRY(cudaMemsetAsync(rowstart, 0, sizeof(int32_t) * (m + 2), st));
  if (nnz == 0) return 0;
  if (nnz > 0x7fffffffll) { oak_set_error("obsoper: nnz too large"); return OAK_ERR_UNSUPPORTED; }
  k_coo_keys<<<(unsigned)((nnz + 255) / 256), 256, 0, st>>>(nnz, m, Hi, Hj, key_in, val_in, rowstart);
  CUDA_TRY(cudaGetLastError());
  int bits = 1;
  while ((1ll << bits) < (long long)m + 1) bits++;
  CUDA_TRY(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, key_in, key_out, val_in, order, (int)nnz, 0, bits, st));
  CUDA_TRY(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, rowstart, rowstart, m + 1, st));
  return 0;
}

int oak_launch_obsoper_rows(cudaStream_t st, int m, int N, const int32_t *rowstart, const int32_t *order,
                            const int32_t *Hj, const double *Hs, const double *Hshift, const double *E,
                            int64_t ldE, double *HE) {
  if (m == 0) return 0;
  dim3 grid((m + 127) / 128, N);
  k_obsoper<<<grid, 128, 0, st>>>(m, rowstart, order, Hj, Hs, Hshift, E, ldE, HE);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int oak_launch_mean_anom(cudaStream_t st, int64_t rows, int N, int anamtype, AnamTab at, const double *E, int64_t ldE,
                         double *mean, double *S, int64_t ldS) {
  if (rows == 0) return 0;
  const size_t smem = sizeof(double) * 64 * N;
  { int rc_ = oak_func_smem(k_mean_anom, (size_t)(64 * 128 * 8)); if (rc_) return rc_; }
  const double scaling = sqrt((double)N - 1.);
  k_mean_anom<<<(unsigned)((rows + 63) / 64), 256, smem, st>>>(rows, N, anamtype, at, E, ldE, mean, S, ldS, scaling);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int oak_launch_epilogue(cudaStream_t st, int64_t rows, int N, int anamtype, AnamTab at, double inflation,
                        const double *maxCorr, const double *xf, double *xa, const double *Sa,
                        int64_t ldSa, double *Ea, int64_t ldEa) {
  if (rows == 0) return 0;
  const double scaling = sqrt((double)N - 1.);
  dim3 grid((unsigned)((rows + 255) / 256), 1);
  k_epilogue<<<grid, 256, 0, st>>>(rows, N, anamtype, at, inflation, scaling, maxCorr, xf, xa, Sa, ldSa, Ea, ldEa);
  CUDA_TRY(cudaGetLastError());
  return 0;
}
