// hgen.cu — observation-operator generation on the GPU: the batched form of OAK's `cinterp`
// (ndgrid.F90:1183-1257), the arithmetic inside genObservationOper (assimilation.F90:2471-2656): for every observation
// locate the grid cell that contains it, then find its interpolation weights on the 2^n corners of that cell.
//
// What the reference does per observation, one after the other on the master thread:
//   g%locate        a lazily built bounding-box tree over the grid (databox, ndgrid_inc.F90:928-1001) whose leaves are
//                   tested with InCube (:304-367)
//   interp_cube     ndgrid.F90:636-665: the cell is cut into n! 2^(n-1) simplices that share the cell centre (split,
//                   :357-435); the first simplex that contains the point wins
//   interp_tetrahedron  :464-629: barycentric coordinates from the inverse of the (n+1) x (n+1) vertex matrix, relative
//                   to the vertex average, |det| > 1e-8, admissible when -1e-8 <= c <= 1 + 1e-8
//   c_cube = tetrahedron(:,:,l) c_simplex, corner indices 1-based, nbp = 2^n (0: outside the grid or a masked corner)
//
// Here: one thread per observation (the work per observation is a few hundred to a few thousand flops, the kernel is
// bound by the scattered reads of the axes and the stores of the result), grids whose coordinate k depends on
// subscript k only (regular and rectilinear grids: `dependence` diagonal).  For those the tree search has a closed
// form: a point on a shared face is inside several cells, the tree visits the upper half of a box first
// (sub-box m = 1 has no lower half, :880-893), so the cell with the HIGHEST subscript in every dimension is the one the
// reference finds; the root box test (xi < xmin or xi > xmax -> out) is the axis range.  The simplex table comes from
// the host (the recursion of split, rewritten over bit masks) and is read through the read-only path.
// Degenerate simplices (|det| <= 1e-8: singleton dimensions, cells thinner than the tolerance) take the SVD branch
// in the reference (:527-627); that branch is not on the device: such observations are reported (nbp = -1) and the
// call fails loudly instead of guessing.
#include <vector>

#include "common.cuh"

namespace {

constexpr int HG_NDMAX = 4;

// ---- the simplices of the unit cell: tet[(l*(n+1) + j)*2^n + q] = weight of cube corner q in vertex j of simplex l
void split_host(int n, int subn, unsigned fixedmask, unsigned selmask, std::vector<double> &tet, int l0) {
  const int twon = 1 << n;
  auto at = [&](int l, int j, int q) -> double & { return tet[((size_t)l * (n + 1) + j) * twon + q]; };
  if (subn == 1) {  // an edge: its two end points
    int j = 0;
    for (int q = 0; q < twon; q++) {
      at(l0, 0, q) = 0.; at(l0, 1, q) = 0.;
    }
    for (int q = 0; q < twon; q++)
      if (selmask >> q & 1u) at(l0, j++, q) = 1.;
    return;
  }
  int cnt = 0;
  for (int q = 0; q < twon; q++) cnt += selmask >> q & 1u;
  int fact = 1, factm = 1;
  for (int i = 2; i <= subn; i++) fact *= i;
  for (int i = 2; i <= subn - 1; i++) factm *= i;
  const int nbth = fact << (subn - 1), subnbth = factm << (subn - 2);
  for (int l = 0; l < nbth; l++)   // the centre of the selected face is the last vertex of all its simplices
    for (int q = 0; q < twon; q++) at(l0 + l, subn, q) = (selmask >> q & 1u) ? 1. / cnt : 0.;
  int m = 0;
  for (int i = 0; i < n; i++) {
    if (fixedmask >> i & 1u) continue;
    for (int side = 0; side < 2; side++) {
      unsigned sub = 0;
      for (int q = 0; q < twon; q++)
        if ((selmask >> q & 1u) && ((q >> i & 1) == side)) sub |= 1u << q;
      split_host(n, subn - 1, fixedmask | (1u << i), sub, tet, l0 + m);
      m += subnbth;
    }
  }
}

int nsimplex(int n) {
  int f = 1;
  for (int i = 2; i <= n; i++) f *= i;
  return f << (n - 1);
}

struct HgenGrid {
  int32_t gshape[HG_NDMAX];
  int32_t axoff[HG_NDMAX];    // start of axis k in the concatenated axes array
  int64_t ioffset[HG_NDMAX];  // linear offset of subscript k in the mask
};

// largest cell subscript i in [0, g-2] whose interval contains v (ascending or descending axis); -1 outside
__device__ __forceinline__ int locate_axis(const double *__restrict__ x, int g, double v) {
  if (g == 1) return (v == x[0]) ? 0 : -1;   // box test of a singleton dimension: xmin = xmax = x(1)
  const bool asc = x[g - 1] >= x[0];
  const double lo = asc ? x[0] : x[g - 1], hi = asc ? x[g - 1] : x[0];
  if (v < lo || v > hi) return -1;
  // ascending: largest i with x[i] <= v ; descending: largest i with x[i] >= v ; both capped at g-2
  int a = 0, b = g - 1;   // invariant: node a satisfies the predicate, node b is the first known not to (or g-1)
  while (b - a > 1) {
    const int mid = (a + b) >> 1;
    const bool ok = asc ? (x[mid] <= v) : (x[mid] >= v);
    if (ok) a = mid; else b = mid;
  }
  return a < g - 1 ? a : g - 2;
}

// c = M^-1 d by Gaussian elimination with the pivoting rule of dgetrf (largest modulus, first on ties); returns det
template <int K>
__device__ __forceinline__ double lu_solve(double (&M)[K][K], double (&c)[K]) {
  double det = 1.;
#pragma unroll
  for (int j = 0; j < K; j++) {
    int p = j;
#pragma unroll
    for (int i = j + 1; i < K; i++)
      if (fabs(M[i][j]) > fabs(M[p][j])) p = i;
    if (p != j) {
#pragma unroll
      for (int q = 0; q < K; q++) {
        // row exchange written with selects so that the matrix stays in registers (no dynamically indexed array)
        double rj = M[j][q], rp = rj;
#pragma unroll
        for (int i = j + 1; i < K; i++)
          if (i == p) rp = M[i][q];
        M[j][q] = rp;
#pragma unroll
        for (int i = j + 1; i < K; i++)
          if (i == p) M[i][q] = rj;
      }
      double cj = c[j], cp = cj;
#pragma unroll
      for (int i = j + 1; i < K; i++)
        if (i == p) cp = c[i];
      c[j] = cp;
#pragma unroll
      for (int i = j + 1; i < K; i++)
        if (i == p) c[i] = cj;
      det = -det;
    }
    const double piv = M[j][j];
    det *= piv;
    if (piv != 0.) {
#pragma unroll
      for (int i = j + 1; i < K; i++) {
        const double f = M[i][j] / piv;
#pragma unroll
        for (int q = j + 1; q < K; q++) M[i][q] -= f * M[j][q];
        c[i] -= f * c[j];
      }
    }
  }
#pragma unroll
  for (int j = K - 1; j >= 0; j--) {
    c[j] /= M[j][j];
#pragma unroll
    for (int i = 0; i < j; i++) c[i] -= M[i][j] * c[j];
  }
  return det;
}

template <int N>
__global__ void __launch_bounds__(128) k_cinterp(HgenGrid g, const double *__restrict__ axes,
                                                 const uint8_t *__restrict__ masked, const double *__restrict__ tet,
                                                 int nbth, int m, const double *__restrict__ xi,
                                                 int32_t *__restrict__ indexes, double *__restrict__ coeff,
                                                 int32_t *__restrict__ nbp, int *__restrict__ ndegenerate) {
  constexpr int TWON = 1 << N, K = N + 1;
  constexpr double tol = 1e-8;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= m) return;
  double x[N];
  int ind[N];
  bool out = false;
#pragma unroll
  for (int k = 0; k < N; k++) {
    x[k] = xi[(size_t)p * N + k];
    ind[k] = locate_axis(axes + g.axoff[k], g.gshape[k], x[k]);
    out = out || ind[k] < 0;
  }
  int32_t *ix = indexes + (size_t)p * TWON * N;
  double *cf = coeff + (size_t)p * TWON;
  if (out) {
#pragma unroll
    for (int q = 0; q < TWON * N; q++) ix[q] = 0;
#pragma unroll
    for (int q = 0; q < TWON; q++) cf[q] = 0.;
    nbp[p] = 0;
    return;
  }
  // corners of the cell: lower / upper node per dimension (a singleton dimension has one node), mask
  double lo[N], hi[N];
  bool anymasked = false;
#pragma unroll
  for (int k = 0; k < N; k++) {
    const double *ax = axes + g.axoff[k];
    lo[k] = ax[ind[k]];
    hi[k] = g.gshape[k] > 1 ? ax[ind[k] + 1] : lo[k];
  }
#pragma unroll
  for (int j = 0; j < TWON; j++) {
    int64_t lin = 0;
#pragma unroll
    for (int k = 0; k < N; k++) {
      const int c = ((j >> k & 1) && g.gshape[k] > 1) ? ind[k] + 1 : ind[k];
      ix[j * N + k] = c + 1;
      lin += c * g.ioffset[k];
    }
    if (masked && masked[lin]) anymasked = true;
  }
#pragma unroll
  for (int q = 0; q < TWON; q++) cf[q] = 0.;
  if (anymasked) { nbp[p] = 0; return; }

  // interp_cube: first simplex that contains the point
  bool degenerate = false;
  for (int l = 0; l < nbth; l++) {
    const double *T = tet + (size_t)l * K * TWON;
    // vertices X(:,j) = sum_q px(:,q) T(q,j); px(k,q) = hi_k or lo_k by bit k of q
    double X[N][K];
#pragma unroll
    for (int j = 0; j < K; j++) {
#pragma unroll
      for (int k = 0; k < N; k++) X[k][j] = 0.;
#pragma unroll
      for (int q = 0; q < TWON; q++) {
        const double t = __ldg(T + j * TWON + q);
#pragma unroll
        for (int k = 0; k < N; k++) X[k][j] = fma((q >> k & 1) ? hi[k] : lo[k], t, X[k][j]);
      }
    }
    double M[K][K], c[K];
#pragma unroll
    for (int k = 0; k < N; k++) {
      double s = 0.;
#pragma unroll
      for (int j = 0; j < K; j++) s += X[k][j];
      const double xc = s / K;
#pragma unroll
      for (int j = 0; j < K; j++) M[1 + k][j] = X[k][j] - xc;
      c[1 + k] = x[k] - xc;
    }
#pragma unroll
    for (int j = 0; j < K; j++) M[0][j] = 1.;
    c[0] = 1.;
    const double det = lu_solve<K>(M, c);
    if (!(fabs(det) > tol)) { degenerate = true; continue; }   // SVD branch of the reference: not on the device
    bool in = true;
#pragma unroll
    for (int j = 0; j < K; j++) in = in && (0. - tol <= c[j] && c[j] <= 1. + tol);
    if (in) {
#pragma unroll
      for (int q = 0; q < TWON; q++) {
        double s = 0.;
#pragma unroll
        for (int j = 0; j < K; j++) s = fma(__ldg(T + j * TWON + q), c[j], s);
        cf[q] = s;
      }
      nbp[p] = TWON;
      return;
    }
  }
  // no simplex took the point.  With a degenerate simplex on the way the reference would have gone through its SVD
  // branch: report; otherwise cinterp still returns nbp = 2^n with the coefficients it was given (zeros here)
  if (degenerate) { nbp[p] = -1; atomicAdd(ndegenerate, 1); }
  else nbp[p] = TWON;
}

}  // namespace

// indexes[m][2^n][n] (1-based), coeff[m][2^n], nbp[m] (2^n, 0 = outside / masked corner, -1 = degenerate cell);
// everything in device memory except the grid descriptor; *ndeg_host = number of degenerate observations
int oak_launch_cinterp(cudaStream_t st, int n, const int32_t *gshape, const double *d_axes, const uint8_t *d_masked,
                       double *d_tet /* >= oak_cinterp_tet_doubles(n) */, int m, const double *d_xi, int32_t *d_indexes,
                       double *d_coeff, int32_t *d_nbp, int *d_ndeg) {
  if (n < 1 || n > HG_NDMAX) { oak_set_error("cinterp: %d dimensions (1 .. %d supported)", n, HG_NDMAX); return OAK_ERR_UNSUPPORTED; }
  HgenGrid g{};
  int32_t off = 0;
  int64_t io = 1;
  for (int k = 0; k < n; k++) {
    if (gshape[k] < 1) { oak_set_error("cinterp: empty dimension"); return OAK_ERR_ARG; }
    g.gshape[k] = gshape[k]; g.axoff[k] = off; g.ioffset[k] = io;
    off += gshape[k]; io *= gshape[k];
  }
  const int nb = nsimplex(n);
  std::vector<double> tet((size_t)nb * (n + 1) * (1 << n), 0.);
  split_host(n, n, 0u, (1u << (1 << n)) - 1u, tet, 0);   // n <= 4: at most 16 corners
  CUDA_TRY(cudaMemcpyAsync(d_tet, tet.data(), sizeof(double) * tet.size(), cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaStreamSynchronize(st));   // `tet` goes out of scope
  CUDA_TRY(cudaMemsetAsync(d_ndeg, 0, sizeof(int), st));
  if (m <= 0) return 0;
  const int grid = (m + 127) / 128;
  switch (n) {
    case 1: k_cinterp<1><<<grid, 128, 0, st>>>(g, d_axes, d_masked, d_tet, nb, m, d_xi, d_indexes, d_coeff, d_nbp, d_ndeg); break;
    case 2: k_cinterp<2><<<grid, 128, 0, st>>>(g, d_axes, d_masked, d_tet, nb, m, d_xi, d_indexes, d_coeff, d_nbp, d_ndeg); break;
    case 3: k_cinterp<3><<<grid, 128, 0, st>>>(g, d_axes, d_masked, d_tet, nb, m, d_xi, d_indexes, d_coeff, d_nbp, d_ndeg); break;
    case 4: k_cinterp<4><<<grid, 128, 0, st>>>(g, d_axes, d_masked, d_tet, nb, m, d_xi, d_indexes, d_coeff, d_nbp, d_ndeg); break;
  }
  CUDA_TRY(cudaGetLastError());
  return 0;
}

size_t oak_cinterp_tet_doubles(int n) { return (size_t)nsimplex(n) * (n + 1) * (1 << n); }
