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
// Here: one thread per observation, grids whose coordinate k depends on
// subscript k only (regular and rectilinear grids: `dependence` diagonal).  For those the tree search has a closed
// form: a point on a shared face is inside several cells, the tree visits the upper half of a box first
// (sub-box m = 1 has no lower half, :880-893), so the cell with the HIGHEST subscript in every dimension is the one the
// reference finds; the root box test (xi < xmin or xi > xmax -> out) is the axis range.  The simplex table comes from
// the host (the recursion of split, rewritten over bit masks) and is read through the read-only path.
// Degenerate simplices (|det| <= 1e-8: singleton dimensions, cells thinner than the tolerance) take the SVD branch
// in the reference (:527-627); that branch is not on the device: such observations are reported (nbp = -1) and the
// call fails loudly instead of guessing.
#include <algorithm>
#include <cmath>
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

// Per simplex l, in the coordinates u in [0,1]^n of the cell (device table, built on the host):
//   B_l ((n+1) x (n+1)) with c = B_l [1; u] the barycentric coordinates of u, det_l the determinant of the vertex matrix
//   [1; W - mean(W)] of the UNIT cell.
// A rectilinear cell is the unit cell stretched by h_k along dimension k: barycentric coordinates do not change under
// an affine map, and the determinant interp_tetrahedron tests (ndgrid.F90:514-516, vertices relative to their
// mean) is det_l * prod_k h_k.  The reference's inverse of the (n+1) x (n+1) matrix per simplex and observation
// (matoper_inc.F90:569-602: ~1400 instructions here, the kernel then sits on the fp64 pipe: 533 M warp instructions
// per 1e6 observations, profiles/r2_ncu_full_k_cinterp_lu.txt) becomes n (n+1) FMAs; the two forms agree to rounding
// (1e-15), far inside the 1e-8 acceptance band that decides which simplex is taken.
void simplex_tables(int n, const std::vector<double> &tet, std::vector<double> &B, std::vector<double> &det) {
  const int twon = 1 << n, K = n + 1, nb = nsimplex(n);
  B.assign((size_t)nb * K * K, 0.);
  det.assign(nb, 0.);
  for (int l = 0; l < nb; l++) {
    double W[HG_NDMAX][HG_NDMAX + 1], wbar[HG_NDMAX], M[HG_NDMAX + 1][2 * (HG_NDMAX + 1)];
    for (int k = 0; k < n; k++) {
      wbar[k] = 0.;
      for (int j = 0; j < K; j++) {
        double sacc = 0.;
        for (int q = 0; q < twon; q++)
          if (q >> k & 1) sacc += tet[((size_t)l * K + j) * twon + q];
        W[k][j] = sacc;
        wbar[k] += sacc;
      }
      wbar[k] /= K;
    }
    for (int i = 0; i < K; i++)
      for (int j = 0; j < K; j++) {
        M[i][j] = i == 0 ? 1. : W[i - 1][j] - wbar[i - 1];
        M[i][K + j] = i == j ? 1. : 0.;
      }
    double d = 1.;
    for (int j = 0; j < K; j++) {  // Gauss-Jordan with partial pivoting
      int p = j;
      for (int i = j + 1; i < K; i++)
        if (fabs(M[i][j]) > fabs(M[p][j])) p = i;
      if (p != j) { for (int q = 0; q < 2 * K; q++) std::swap(M[j][q], M[p][q]); d = -d; }
      const double piv = M[j][j];
      d *= piv;
      if (piv == 0.) continue;   // cannot happen: the simplices of the unit cell have volume 1 / (n! 2^(n-1))
      for (int q = 0; q < 2 * K; q++) M[j][q] /= piv;
      for (int i = 0; i < K; i++)
        if (i != j) {
          const double f = M[i][j];
          for (int q = 0; q < 2 * K; q++) M[i][q] -= f * M[j][q];
        }
    }
    det[l] = d;
    // c = Ainv [1; u - wbar]  ->  c_j = (Ainv[j][0] - sum_k Ainv[j][1+k] wbar_k) + sum_k Ainv[j][1+k] u_k
    for (int j = 0; j < K; j++) {
      double a0 = M[j][K + 0];
      for (int k = 0; k < n; k++) a0 -= M[j][K + 1 + k] * wbar[k];
      B[((size_t)l * K + j) * K + 0] = a0;
      for (int k = 0; k < n; k++) B[((size_t)l * K + j) * K + 1 + k] = M[j][K + 1 + k];
    }
  }
}

// largest cell subscript i in [0, g-2] whose interval contains v (ascending or descending axis); -1 outside
__device__ __forceinline__ int locate_axis(const double *__restrict__ x, int g, double v) {
  if (g == 1) return (v == x[0]) ? 0 : -1;   // box test of a singleton dimension: xmin = xmax = x(1)
  const bool asc = x[g - 1] >= x[0];
  const double lo = asc ? x[0] : x[g - 1], hi = asc ? x[g - 1] : x[0];
  if (v < lo || v > hi) return -1;
  // ascending: largest i with x[i] <= v ; descending: largest i with x[i] >= v ; both capped at g-2
  int a = 0, b = g - 1;   // node a satisfies the predicate, node b is the first known not to (or g-1)
  while (b - a > 1) {
    const int mid = (a + b) >> 1;
    const bool ok = asc ? (x[mid] <= v) : (x[mid] >= v);
    if (ok) a = mid; else b = mid;
  }
  return a;
}

// One thread per observation; the block's results are staged in shared memory and leave in contiguous runs (the
// per-observation records are 96 + 64 + 4 bytes in 3-D: written by their own threads every store instruction would
// touch 32 sectors).
template <int N, int NT>
__global__ void __launch_bounds__(NT) k_cinterp(HgenGrid g, const double *__restrict__ axes,
                                                const uint8_t *__restrict__ masked, const double *tet,
                                                const double *Btab, const double *dettab,
                                                int nbth, int m, const double *__restrict__ xi,
                                                int32_t *__restrict__ indexes, double *__restrict__ coeff,
                                                int32_t *__restrict__ nbp, int *__restrict__ ndegenerate) {
  constexpr int TWON = 1 << N, K = N + 1;
  constexpr int LI = TWON * N + 1, LC = TWON + 1;   // odd strides: conflict-free staging
  constexpr double tol = 1e-8;
  __shared__ int32_t s_ix[NT * LI];
  __shared__ double s_cf[NT * LC];
  __shared__ int32_t s_nbp[NT];
  // the simplex tables (9.4 KB in 3-D) in shared memory: the lanes of a warp read them at different simplices (a gather:
  // through the read-only path this was 49 % long-scoreboard stalls); 4-D (163 KB) keeps them in global memory
  constexpr bool SMTAB = N <= 3;
  constexpr int NBTH = N == 1 ? 1 : (N == 2 ? 4 : 24);
  // odd strides per simplex (the lanes of a warp sit at different simplices: with the dense strides of 32 and 16 doubles
  // they would all hit the same banks: 28 M conflicts per 1e6 observations)
  constexpr int ST = SMTAB ? K * TWON + 1 : K * TWON, SB = SMTAB ? K * K + 1 : K * K;
  constexpr int TABD = SMTAB ? NBTH * (ST + SB + 1) : 1;
  __shared__ double s_tab[TABD];
  const int tid = threadIdx.x;
  if (SMTAB) {
    for (int i = tid; i < NBTH * K * TWON; i += NT) s_tab[(i / (K * TWON)) * ST + i % (K * TWON)] = tet[i];
    for (int i = tid; i < NBTH * K * K; i += NT) s_tab[NBTH * ST + (i / (K * K)) * SB + i % (K * K)] = Btab[i];
    for (int i = tid; i < NBTH; i += NT) s_tab[NBTH * (ST + SB) + i] = dettab[i];
    __syncthreads();
    tet = s_tab; Btab = s_tab + NBTH * ST; dettab = s_tab + NBTH * (ST + SB);
  }
  const int p0 = blockIdx.x * NT;
  const int p = p0 + tid;
  int32_t *ix = s_ix + tid * LI;
  double *cf = s_cf + tid * LC;
  int mynbp = 0;
#pragma unroll
  for (int q = 0; q < TWON * N; q++) ix[q] = 0;
#pragma unroll
  for (int q = 0; q < TWON; q++) cf[q] = 0.;
  if (p < m) {
    double x[N];
    int ind[N];
    bool out = false;
#pragma unroll
    for (int k = 0; k < N; k++) {
      x[k] = xi[(size_t)p * N + k];
      ind[k] = locate_axis(axes + g.axoff[k], g.gshape[k], x[k]);
      out = out || ind[k] < 0;
    }
    if (!out) {
      // the cell: lower / upper node per dimension (a singleton dimension has one node), position inside it, mask
      double u[N], vol = 1.;
      bool anymasked = false;
#pragma unroll
      for (int k = 0; k < N; k++) {
        const double *ax = axes + g.axoff[k];
        const double lo = ax[ind[k]];
        const double h = g.gshape[k] > 1 ? ax[ind[k] + 1] - lo : 0.;
        u[k] = h != 0. ? (x[k] - lo) / h : 0.;
        vol *= h;
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
      if (!anymasked) {
        // interp_cube: first simplex (in the order of split) that contains the point.
        // Fast path: the simplex that contains u follows from the order of |u_k - 1/2| (each level of split cuts the
        // current face into pyramids over its sub-faces, apex = face centre; the pyramid that holds the point is the one
        // over the sub-face in the direction of the largest |u_k - 1/2| among the free dimensions).  If the point is inside
        // that simplex by a margin of 1e-5 in every barycentric coordinate it is outside every other simplex of the
        // tiling by far more than the acceptance band of 1e-8, so "the first simplex that accepts" is this one and the
        // ordered search is not needed.  Points closer to a simplex boundary (and degenerate cells) take the ordered loop.
        bool degenerate = false, found = false;
        auto test = [&](int l, double margin) -> bool {
          const double *B = Btab + (size_t)l * SB;
          double c[K];
          bool in = true;
#pragma unroll
          for (int j = 0; j < K; j++) {
            double sacc = B[j * K];
#pragma unroll
            for (int k = 0; k < N; k++) sacc = fma(B[j * K + 1 + k], u[k], sacc);
            c[j] = sacc;
            in = in && (margin <= sacc && sacc <= 1. - margin);
          }
          if (in) {
            const double *T = tet + (size_t)l * ST;
#pragma unroll
            for (int q = 0; q < TWON; q++) {
              double sacc = 0.;
#pragma unroll
              for (int j = 0; j < K; j++) sacc = fma(T[j * TWON + q], c[j], sacc);
              cf[q] = sacc;
            }
          }
          return in;
        };
        {
          int lstar = 0, sub = nbth;
          unsigned fixedm = 0;
#pragma unroll
          for (int level = 0; level < N - 1; level++) {
            int best = 0;
            double bv = -1.;
#pragma unroll
            for (int k = 0; k < N; k++) {
              const double a = fabs(u[k] - 0.5);
              if (!(fixedm >> k & 1u) && a > bv) { bv = a; best = k; }
            }
            const int pos = __popc(~fixedm & ((1u << best) - 1u));
            sub /= 2 * (N - level);
            lstar += (2 * pos + (u[best] > 0.5 ? 1 : 0)) * sub;
            fixedm |= 1u << best;
          }
          if (fabs(dettab[lstar] * vol) > tol) found = test(lstar, 1e-5);
        }
        for (int l = 0; l < nbth && !found; l++) {
          if (!(fabs(dettab[l] * vol) > tol)) { degenerate = true; continue; }   // SVD branch of the reference
          found = test(l, 0. - tol);
        }
        // no simplex took the point: with a degenerate simplex on the way the reference would have gone through its SVD
        // branch (reported); otherwise cinterp still returns nbp = 2^n
        if (!found && degenerate) { mynbp = -1; atomicAdd(ndegenerate, 1); }
        else mynbp = TWON;
      }
    }
  }
  s_nbp[tid] = mynbp;
  __syncthreads();
  const int nv = min(NT, m - p0);   // observations of this block
  int32_t *gi = indexes + (size_t)p0 * TWON * N;
  for (int i = tid; i < nv * TWON * N; i += NT) gi[i] = s_ix[(i / (TWON * N)) * LI + i % (TWON * N)];
  double *gc = coeff + (size_t)p0 * TWON;
  for (int i = tid; i < nv * TWON; i += NT) gc[i] = s_cf[(i / TWON) * LC + i % TWON];
  if (tid < nv) nbp[p0 + tid] = s_nbp[tid];
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
  const int nb = nsimplex(n), K = n + 1, twon = 1 << n;
  std::vector<double> tab((size_t)nb * K * twon, 0.), B, det;
  split_host(n, n, 0u, (1u << twon) - 1u, tab, 0);   // n <= 4: at most 16 corners
  simplex_tables(n, tab, B, det);
  const size_t ntet = tab.size();
  tab.insert(tab.end(), B.begin(), B.end());
  tab.insert(tab.end(), det.begin(), det.end());
  CUDA_TRY(cudaMemcpyAsync(d_tet, tab.data(), sizeof(double) * tab.size(), cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaStreamSynchronize(st));   // `tab` goes out of scope
  CUDA_TRY(cudaMemsetAsync(d_ndeg, 0, sizeof(int), st));
  if (m <= 0) return 0;
  const double *dB = d_tet + ntet, *dD = dB + B.size();
  switch (n) {
    case 1: k_cinterp<1, 128><<<(m + 127) / 128, 128, 0, st>>>(g, d_axes, d_masked, d_tet, dB, dD, nb, m, d_xi, d_indexes, d_coeff, d_nbp, d_ndeg); break;
    case 2: k_cinterp<2, 128><<<(m + 127) / 128, 128, 0, st>>>(g, d_axes, d_masked, d_tet, dB, dD, nb, m, d_xi, d_indexes, d_coeff, d_nbp, d_ndeg); break;
    case 3: k_cinterp<3, 128><<<(m + 127) / 128, 128, 0, st>>>(g, d_axes, d_masked, d_tet, dB, dD, nb, m, d_xi, d_indexes, d_coeff, d_nbp, d_ndeg); break;
    case 4: k_cinterp<4, 64><<<(m + 63) / 64, 64, 0, st>>>(g, d_axes, d_masked, d_tet, dB, dD, nb, m, d_xi, d_indexes, d_coeff, d_nbp, d_ndeg); break;
  }
  CUDA_TRY(cudaGetLastError());
  return 0;
}

// simplex table + barycentric matrices + determinants
size_t oak_cinterp_tet_doubles(int n) { return (size_t)nsimplex(n) * ((n + 1) * (1 << n) + (n + 1) * (n + 1) + 1); }
