// eig_simple.cu — straightforward shared-memory version of the per-zone transform
// (cross-check kernel, option "eig_kernel" = 1).  Same mathematics as eig_fast.cu, written for
// clarity: everything in shared memory, one CTA of 128 threads per zone, any NP <= 128.
//
// What it computes (rrsqrt.F90:136-185) from G = HSf^T R_loc^-1 HSf
// and c = HSf^T R_loc^-1 (yo-Hxf):
//     ampl = U (1+Lambda)^-1 U^T c                      rrsqrt.F90:137-142
//     T    = U (1+Lambda)^-1/2 U^T Omega                rrsqrt.F90:162-185
// Both are matrix functions of A = I + G, so instead of dsyev (matoper_inc.F90:991-995) we factor
// A = L L^T (Cholesky) and orthogonalise the columns of L by one-sided Jacobi rotations:
// Z = L V = U diag(sigma), sigma_j^2 = 1 + lambda_j.  Then with s2_j = |z_j|^2:
//     (I+G)^-1/2 = sum_j z_j z_j^T / (s2_j sqrt(max(s2_j,1)))        (lambda <- max(lambda,0), :137)
//     ampl       = sum_j z_j (z_j . c) / (s2_j max(s2_j,1))
//     v ~ (I+G)^1/2 1 = sum_j z_j (z_j . 1) sqrt(max(s2_j,1)) / s2_j  (rrsqrt.F90:176-178)
// Omega = RotateVector(w,v) (rrsqrt.F90:737-744, perpSpace matoper.F90:509-535) equals
// H_v diag(1,..,1, sign(v_N) sign(w_N)) H_w with H_x = I - u_x u_x^T/(1+|x_N|), u_x = x + sign(x_N) e_N,
// which is applied as two rank-one updates.
#include "common.cuh"
#include "eig_common.cuh"

namespace {

template <int NP>
__global__ void __launch_bounds__(128) k_eig_simple(int N, const int32_t *__restrict__ mloc,
                                                    const double *__restrict__ G,
                                                    const double *__restrict__ cin,
                                                    double *__restrict__ Tout,
                                                    double *__restrict__ ampl_out, double tol,
                                                    int max_sweeps, DevCounters *ctr) {
  constexpr int LD = NP;
  extern __shared__ __align__(16) double sm[];
  double *W = sm;                 // [NP][LD] column-major
  double *s_vec = sm + NP * LD;   // 8 vectors of NP
  double *s_c = s_vec, *s_a = s_vec + NP, *s_b = s_vec + 2 * NP, *s_d = s_vec + 3 * NP;
  double *s_t1 = s_vec + 4 * NP, *s_t2 = s_vec + 5 * NP, *s_g1 = s_vec + 6 * NP, *s_g2 = s_vec + 7 * NP;
  __shared__ int s_maxcos, s_maxt;
  __shared__ double s_red[4];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int zl = blockIdx.x;
  if (mloc[zl] == 0) return;
  const double *Gz = G + (int64_t)zl * NP * NP;

  for (int idx = tid; idx < NP * NP; idx += 128) {
    const int i = idx % NP, j = idx / NP;
    W[i + LD * j] = Gz[idx] + (i == j ? 1. : 0.);
  }
  if (tid < NP) s_c[tid] = cin[(int64_t)zl * NP + tid];
  __syncthreads();

  // ---- Cholesky A = L L^T, thread i owns row i (right-looking) ----
  for (int j = 0; j < NP; j++) {
    const double d = sqrt(W[j + LD * j]);
    __syncthreads();
    if (tid == j) W[j + LD * j] = d;
    if (tid > j && tid < NP) W[tid + LD * j] /= d;
    __syncthreads();
    if (tid > j && tid < NP) {
      const double lij = W[tid + LD * j];
      for (int k = j + 1; k <= tid; k++) W[tid + LD * k] = fma(-lij, W[k + LD * j], W[tid + LD * k]);
    }
    __syncthreads();
  }
  for (int idx = tid; idx < NP * NP; idx += 128) {
    const int i = idx % NP, j = idx / NP;
    if (j > i) W[i + LD * j] = 0.;
  }
  __syncthreads();

  // ---- one-sided Jacobi, round-robin ordering ----
  int sweeps = 0;
  for (int sweep = 0; sweep < max_sweeps; sweep++) {
    if (tid == 0) { s_maxcos = 0; s_maxt = 0; }
    __syncthreads();
    float mymax = 0.f, mymaxt = 0.f;
    for (int r = 0; r < NP - 1; r++) {
      for (int pi = warp; pi < NP / 2; pi += 4) {
        int p, qcol;
        if (pi == 0) { p = NP - 1; qcol = r; }
        else { p = (r + pi) % (NP - 1); qcol = (r - pi + NP - 1) % (NP - 1); }
        double a = 0., b = 0., g = 0.;
        for (int i = lane; i < NP; i += 32) {
          const double x = W[i + LD * p], y = W[i + LD * qcol];
          a = fma(x, x, a); b = fma(y, y, b); g = fma(x, y, g);
        }
        for (int o = 16; o > 0; o >>= 1) {
          a += __shfl_xor_sync(0xffffffffu, a, o);
          b += __shfl_xor_sync(0xffffffffu, b, o);
          g += __shfl_xor_sync(0xffffffffu, g, o);
        }
        double cs, sn, tt;
        const float cosang = jacobi_params(a, b, g, cs, sn, tt);
        if (cosang > JACOBI_SKIP) {
          mymax = fmaxf(mymax, cosang);
          mymaxt = fmaxf(mymaxt, fabsf((float)tt));
          for (int i = lane; i < NP; i += 32) {
            const double x = W[i + LD * p], y = W[i + LD * qcol];
            W[i + LD * p] = cs * x - sn * y;
            W[i + LD * qcol] = sn * x + cs * y;
          }
        }
      }
      __syncthreads();
    }
    atomicMax(&s_maxcos, __float_as_int(mymax));
    atomicMax(&s_maxt, __float_as_int(mymaxt));
    __syncthreads();
    sweeps = sweep + 1;
    const float mc = __int_as_float(s_maxcos), mt = __int_as_float(s_maxt);
    __syncthreads();
    if (jacobi_converged(mc, mt, (float)tol)) break;
    if (sweep == max_sweeps - 1 && tid == 0 && tol >= 0) atomicAdd(&ctr->not_converged, 1);
  }

  // ---- column statistics: s2, z.c, z.1 ----
  for (int j = warp; j < NP; j += 4) {
    double s2 = 0., zc = 0., z1 = 0.;
    for (int i = lane; i < NP; i += 32) {
      const double z = W[i + LD * j];
      s2 = fma(z, z, s2);
      zc = fma(z, s_c[i], zc);
      if (i < N) z1 += z;
    }
    for (int o = 16; o > 0; o >>= 1) {
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
      zc += __shfl_xor_sync(0xffffffffu, zc, o);
      z1 += __shfl_xor_sync(0xffffffffu, z1, o);
    }
    if (lane == 0) {
      const double s2c = fmax(s2, 1.);
      s_d[j] = 1. / (s2 * sqrt(s2c));
      s_a[j] = zc / (s2 * s2c);
      s_b[j] = z1 * sqrt(s2c) / s2;
    }
  }
  __syncthreads();
  // ampl, v
  double vi = 0.;
  if (tid < NP) {
    double am = 0.;
    for (int j = 0; j < NP; j++) {
      const double z = W[tid + LD * j];
      am = fma(z, s_a[j], am);
      vi = fma(z, s_b[j], vi);
    }
    if (tid >= N) { am = 0.; vi = 0.; }
    if (am != am) atomicExch(&ctr->nan_flag, 1);
    ampl_out[(int64_t)zl * NP + tid] = am;
  }
  // |v|
  double part = (tid < N) ? vi * vi : 0.;
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  if (lane == 0) s_red[warp] = part;
  __syncthreads();
  const double vnorm = sqrt(s_red[0] + s_red[1] + s_red[2] + s_red[3]);
  // Householder vectors u_v, u_w (w = 1/sqrt(N)) : s_t1 = u_v, s_t2 = D u_w
  const double wN = 1. / sqrt((double)N);
  if (tid < NP) {
    const double v = (tid < N) ? vi / vnorm : 0.;
    s_g1[tid] = v;  // temporarily v
  }
  __syncthreads();
  const double vN = s_g1[N - 1];
  const double sv = copysign(1., vN), sw = 1.;
  const double dNN = sv * sw;
  const double hv = 1. / (1. + fabs(vN)), hw = 1. / (1. + fabs(wN));
  __syncthreads();
  if (tid < NP) {
    double uv = (tid < N) ? s_g1[tid] : 0.;
    double uw = (tid < N) ? wN : 0.;
    if (tid == N - 1) { uv += sv; uw += sw; }
    s_t1[tid] = uv;                               // u_v
    s_t2[tid] = (tid == N - 1) ? dNN * uw : uw;   // D u_w
    s_b[tid] = uw;                                // u_w (s_b no longer needed)
  }
  __syncthreads();
  // g1 = M u_v, gm2 = M (D u_w) with M = Z diag(d) Z^T: first t = d * (Z^T x)
  for (int j = warp; j < NP; j += 4) {
    double p1 = 0., p2 = 0.;
    for (int i = lane; i < NP; i += 32) {
      const double z = W[i + LD * j];
      p1 = fma(z, s_t1[i], p1);
      p2 = fma(z, s_t2[i], p2);
    }
    for (int o = 16; o > 0; o >>= 1) {
      p1 += __shfl_xor_sync(0xffffffffu, p1, o);
      p2 += __shfl_xor_sync(0xffffffffu, p2, o);
    }
    if (lane == 0) { s_a[j] = s_d[j] * p1; s_c[j] = s_d[j] * p2; }
  }
  __syncthreads();
  if (tid < NP) {
    double g1 = 0., gm2 = 0.;
    for (int j = 0; j < NP; j++) {
      const double z = W[tid + LD * j];
      g1 = fma(z, s_a[j], g1);
      gm2 = fma(z, s_c[j], gm2);
    }
    s_g1[tid] = g1;
    s_g2[tid] = gm2;
  }
  // kappa = (u_v/(1+|vN|)) . (D u_w)
  double kp = (tid < NP) ? s_t1[tid] * hv * s_t2[tid] : 0.;
  for (int o = 16; o > 0; o >>= 1) kp += __shfl_xor_sync(0xffffffffu, kp, o);
  __syncthreads();
  if (lane == 0) s_red[warp] = kp;
  __syncthreads();
  const double kappa = s_red[0] + s_red[1] + s_red[2] + s_red[3];
  if (tid < NP) s_g2[tid] = s_g2[tid] - kappa * s_g1[tid];
  __syncthreads();
  // T[i][k] = (M[i][k] - g1[i] hv u_v[k]) * D_k - g2[i] hw u_w[k]   (row-major output T[i*NP+k])
  double *Tz = Tout + (int64_t)zl * NP * NP;
  for (int idx = tid; idx < NP * NP; idx += 128) {
    const int i = idx % NP, k = idx / NP;
    double mval = 0.;
    for (int j = 0; j < NP; j++) mval = fma(s_d[j] * W[i + LD * j], W[k + LD * j], mval);
    double t = mval - s_g1[i] * hv * s_t1[k];
    if (k == N - 1) t *= dNN;
    t -= s_g2[i] * hw * s_b[k];
    Tz[(int64_t)i * NP + k] = t;
  }
  if (tid == 0) atomicAdd(&ctr->sweeps, (unsigned long long)sweeps);
}

template <int NP>
int launch(cudaStream_t st, int N, int nz, const int32_t *mloc, const double *G, const double *c, double *T,
           double *ampl, double tol, int max_sweeps, DevCounters *ctr) {
  const size_t smem = sizeof(double) * (NP * NP + 8 * NP);
  { int rc_ = oak_func_smem(k_eig_simple<NP>, (size_t)((int)smem)); if (rc_) return rc_; }
  k_eig_simple<NP><<<nz, 128, smem, st>>>(N, mloc, G, c, T, ampl, tol, max_sweeps, ctr);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

}  // namespace

int oak_launch_eig_simple(cudaStream_t st, int N, int NP, int nz, const int32_t *mloc, const double *G,
                          const double *c, double *T, double *ampl, double tol, int max_sweeps,
                          DevCounters *ctr) {
  switch (NP) {
    case 16: return launch<16>(st, N, nz, mloc, G, c, T, ampl, tol, max_sweeps, ctr);
    case 32: return launch<32>(st, N, nz, mloc, G, c, T, ampl, tol, max_sweeps, ctr);
    case 64: return launch<64>(st, N, nz, mloc, G, c, T, ampl, tol, max_sweeps, ctr);
    case 128: return launch<128>(st, N, nz, mloc, G, c, T, ampl, tol, max_sweeps, ctr);
  }
  oak_set_error("eig: unsupported padded ensemble size %d", NP);
  return OAK_ERR_UNSUPPORTED;
}
