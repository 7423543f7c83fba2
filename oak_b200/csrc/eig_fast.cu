// eig_fast.cu — register-resident batched block-Jacobi kernel: the per-zone transform
// (ampl, T) from (G, c).  Mathematics as in eig_simple.cu (Cholesky of I+G, one-sided Jacobi on
// the columns of L, matrix functions from Z = L V); this file is the production layout.
//
// One CTA per zone.  The NP x NP matrix W lives in registers: the CTA is NP/4 groups of TL lanes,
// a group owns two blocks of two columns (P = block position 2g, Q = position 2g+1), a lane owns
// R = NP/TL = 16 rows of each of its 4 columns (64 doubles).  Row pairs are interleaved across the
// lanes of a group so that the 16-byte shared-memory transfers of neighbouring lanes are contiguous.
//
// Ordering: odd-even transposition on the NB = NP/2 block positions; after every block rotation the
// two blocks swap positions, which is what makes every pair of blocks meet exactly once in NB steps.
// The swap is never a register move: it is realised by WHICH register block travels.  Even step: a
// group rotates its own (P,Q) in place; logically position 2g is now in Q, 2g+1 in P.  Odd step, pairs
// (2g+1, 2g+2): the group lends Q (position 2g) through shared memory to its left neighbour, borrows
// the right neighbour's Q (position 2g+2) into its Q registers, rotates (P,Q) in place, returns its P
// registers (the new position 2g+2) and recovers the new position 2g into P.  After the pair of steps
// P = position 2g and Q = 2g+1 again, so the loop body is register-invariant (no moves, no swapped
// assignments).  Inside a block pair the 4 cross column pairs are rotated in two sub-rounds of two
// independent rotations; the two columns of a block are rotated against each other once per sweep.
// Dot products are reduced over the TL lanes of a group with warp shuffles.
#include "common.cuh"
#include "eig_common.cuh"

#ifndef EIG_CTAS_PER_SM
#define EIG_CTAS_PER_SM 4
#endif

namespace {

constexpr unsigned FULL = 0xffffffffu;

template <int NP, int TL, int KB>
struct Cfg {
  static constexpr int R = NP / TL;          // rows of a column per lane
  static constexpr int NCG = 2 * KB;         // columns per group: two blocks of KB columns
  static constexpr int NG = NP / NCG;        // groups
  static constexpr int NTH = NG * TL;
  static constexpr int NB = NP / KB;         // block positions
  static_assert(KB == 1 || KB == 2, "one or two columns per block");
  static constexpr int LDW = NP;      // W / Y leading dimension in shared memory
  static constexpr int LDX = NP + 4;  // exchange-buffer column stride
  static_assert(R == 16 || R == 8, "8 or 16 rows per lane");
  static_assert(NTH % 32 == 0, "whole warps");
};

template <int R>
__device__ __forceinline__ double dotR(const double (&x)[R], const double (&y)[R]) {
  // two accumulators: the callers interleave two to four independent dot products
  double s0 = 0., s1 = 0.;
#pragma unroll
  for (int i = 0; i < R; i += 2) {
    s0 = fma(x[i], y[i], s0);
    s1 = fma(x[i + 1], y[i + 1], s1);
  }
  return s0 + s1;
}

template <int TL>
__device__ __forceinline__ double group_sum(double v) {
#pragma unroll
  for (int o = 1; o < TL; o <<= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}

// ---------------------------------------------------------------------------------------------
// Scaled ("fast") rotations, in place.  A column is stored as X with a deferred scale: true column =
// sg * X.  The rotation x' = c (x - t y), y' = c (y + t x) is applied to the stored columns as two
// sequential shears that need no temporaries:
//     X <- X - t1 Y            t1 = t sg_y / sg_x                  sg_x <- c sg_x
//     Y <- Y + t2 X(new)       t2 = t c^2 sg_x / sg_y              sg_y <- sg_y / c
// (y + t x = (1+t^2) (y + t c^2 x'), the factor 1+t^2 = 1/c^2 goes into the scale).  Two FMAs per
// row pair instead of four multiply-adds, and every result overwrites its own operand, so the loop
// body keeps every column in the same registers.  ig = 1/sg is carried along so that no division is
// needed; nn = |true column|^2 follows the Jacobi identities |x'|^2 = |x|^2 - t g, |y'|^2 = |y|^2 + t g.
// Scales are folded back into the columns at the start of every sweep.
// ---------------------------------------------------------------------------------------------
struct Col {
  double sg, ig;  // deferred scale of the stored column and its inverse
  double nn;      // |true column|^2.  fp64 on purpose: the difference of two nearly equal norms decides the
                  // angle inside clusters; tracking it in fp32 costs up to 8 extra sweeps there (measured)
  float sf, nf;   // fp32 shadows of sg and nn for the quantities that only steer the angle / the tests
};

#define JACOBI_SKIP2 1e-26f  // (JACOBI_SKIP)^2, on cos^2

struct Rot {
  double t1, t2, c, ic, tg;
  float cf, sf, tgf, k2, ta;  // fp32 cosine / sine for the Gram algebra, t*g for the fp32 norm shadow
  bool on;
};

// Parameters of the rotation of columns (p,q) from the reduced stored dot product Gam.
// The angle only steers convergence, so t = tan(theta) is computed in fp32; c = (1+t^2)^-1/2 must make
// the transformation orthogonal to fp64 accuracy: fp32 rsqrt + one fp32 and one fp64 Newton step
// (error 1.5 eps32^2 ~ 5e-15 per rotation, unbiased, far below the 1e-9 budget after a few 1e3 rotations).
// fp32 helpers as single instructions: plain cvt (one F2F; the compiler's -ftz conversion adds a fp64
// compare and a fix-up multiply per cast), approximate reciprocal / rsqrt with flush-to-zero (one MUFU).
#ifdef OAK_CUEMU  // functional CPU emulation for tests (tools/cuemu)
__device__ __forceinline__ float d2f(double x) { return (float)x; }
__device__ __forceinline__ float rcp_approx(float x) { return 1.f / x; }
__device__ __forceinline__ float rsqrt_approx(float x) { return 1.f / sqrtf(x); }
#else
__device__ __forceinline__ float d2f(double x) {
  float y;
  asm("cvt.rn.f32.f64 %0, %1;" : "=f"(y) : "d"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rsqrt_approx(float x) {
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
#endif

// gf is the TRUE dot product of the two columns (stored dot times both scales), in fp32.
__device__ __forceinline__ Rot rot_params(const Col &p, const Col &q, float gf, bool active) {
  Rot r;
  r.k2 = (gf * gf) * rcp_approx(p.nf * q.nf);
  r.on = active && (r.k2 > JACOBI_SKIP2);
  const float df = d2f(q.nn - p.nn), g2f = gf + gf;
  const float h2 = fmaf(df, df, g2f * g2f);
  const float h = h2 * rsqrt_approx(h2);
  float tf = g2f * rcp_approx(df + copysignf(h, df));
  r.ta = fabsf(tf);
  tf = r.on ? tf : 0.f;
  const float yf = fmaf(tf, tf, 1.f);
  float cf = rsqrt_approx(yf);
  cf = cf * fmaf(-0.5f * yf, cf * cf, 1.5f);
  r.cf = cf;
  r.sf = tf * cf;
  r.tgf = tf * gf;
  const double tt = (double)tf;
  const double y = fma(tt, tt, 1.);
  double c = (double)cf;
  c = c * fma(-0.5 * y, c * c, 1.5);
  r.c = c;
  r.ic = y * c;
  r.t1 = tt * (q.sg * p.ig);
  r.t2 = ((tt * c) * c) * (p.sg * q.ig);
  r.tg = (double)r.tgf;
  return r;
}

template <int R>
__device__ __forceinline__ void rot_apply(double (&x)[R], double (&y)[R], const Rot &r) {
  const double m1 = -r.t1;
#pragma unroll
  for (int i = 0; i < R; i++) x[i] = fma(m1, y[i], x[i]);
#pragma unroll
  for (int i = 0; i < R; i++) y[i] = fma(r.t2, x[i], y[i]);
}
__device__ __forceinline__ void col_update(Col &p, Col &q, const Rot &r) {
  p.sg *= r.c; p.ig *= r.ic; p.nn -= r.tg;
  q.sg *= r.ic; q.ig *= r.c; q.nn += r.tg;
  const float icf = fmaf(r.sf, r.sf * rcp_approx(r.cf), r.cf);  // 1/c = c + s^2/c
  p.sf *= r.cf; p.nf -= r.tgf;
  q.sf *= icf; q.nf += r.tgf;
}

struct SweepStat {  // maxima over the rotated pairs of a sweep (jacobi_converged)
  float mx2, mt;
  __device__ __forceinline__ void add(const Rot &r) {
    if (r.on) { mx2 = fmaxf(mx2, r.k2); mt = fmaxf(mt, r.ta); }
  }
};

// Rotates the 4 cross pairs of blocks X={X0,X1}, Y={Y0,Y1} in place.
template <int R, int TL>
__device__ __forceinline__ void rotate_block_pair(double (&X0)[R], double (&X1)[R], double (&Y0)[R],
                                                  double (&Y1)[R], Col &cX0, Col &cX1, Col &cY0, Col &cY1,
                                                  bool active, SweepStat &ss) {
  {  // sub-round 1: (X0,Y0) (X1,Y1)
    const float g1 = (cX0.sf * cY0.sf) * d2f(group_sum<TL>(dotR<R>(X0, Y0)));
    const float g2 = (cX1.sf * cY1.sf) * d2f(group_sum<TL>(dotR<R>(X1, Y1)));
    const Rot r1 = rot_params(cX0, cY0, g1, active), r2 = rot_params(cX1, cY1, g2, active);
    ss.add(r1); ss.add(r2);
    if (__any_sync(FULL, r1.on || r2.on)) {
      rot_apply<R>(X0, Y0, r1);
      rot_apply<R>(X1, Y1, r2);
      col_update(cX0, cY0, r1);
      col_update(cX1, cY1, r2);
    }
  }
  {  // sub-round 2: (X0,Y1) (X1,Y0)
    const float g1 = (cX0.sf * cY1.sf) * d2f(group_sum<TL>(dotR<R>(X0, Y1)));
    const float g2 = (cX1.sf * cY0.sf) * d2f(group_sum<TL>(dotR<R>(X1, Y0)));
    const Rot r1 = rot_params(cX0, cY1, g1, active), r2 = rot_params(cX1, cY0, g2, active);
    ss.add(r1); ss.add(r2);
    if (__any_sync(FULL, r1.on || r2.on)) {
      rot_apply<R>(X0, Y1, r1);
      rot_apply<R>(X1, Y0, r2);
      col_update(cX0, cY1, r1);
      col_update(cX1, cY0, r2);
    }
  }
}

// Same four rotations with ONE reduction phase: the four cross dot products of the block pair are taken
// up front; the dot products the second sub-round needs, (X0',Y1') and (X1',Y0'), follow from them and
// from the intra-block dots gX = x0.x1, gY = y0.y1 (carried with the blocks) by the 4x4 Gram algebra of
// the two first rotations (a'=c1 a - s1 c, c'=s1 a + c1 c, b'=c2 b - s2 d, d'=s2 b + c2 d):
//   a'.d' = c1 s2 ab + c1 c2 ad - s1 s2 cb - s1 c2 cd      c'.b' = s1 c2 ab - s1 s2 ad + c1 c2 cb - c1 s2 cd
//   a'.b' = c1 c2 ab - c1 s2 ad - s1 c2 cb + s1 s2 cd      c'.d' = s1 s2 ab + s1 c2 ad + c1 s2 cb + c1 c2 cd
// and after the second sub-round (a''=c3 a' - s3 d', d''=.., b''=c4 b' - s4 c', c''=..):
//   gX'' = a''.b'' = c3 c4 a'.b' + s3 s4 c'.d'              gY'' = c''.d'' = s3 s4 a'.b' + c3 c4 c'.d'
// All scalar work is fp64 on true (unscaled) quantities; the columns are then updated in one pass of
// 8 in-place FMAs per row.  Halves the number of latency-exposed dot -> shuffle -> parameter chains.
template <int R, int TL>
__device__ __forceinline__ void rotate_block_pair_gram(double (&X0)[R], double (&X1)[R], double (&Y0)[R],
                                                       double (&Y1)[R], Col &cX0, Col &cX1, Col &cY0, Col &cY1,
                                                       float &gX, float &gY, bool active, SweepStat &ss) {
  const double d_ac = group_sum<TL>(dotR<R>(X0, Y0)), d_bd = group_sum<TL>(dotR<R>(X1, Y1));
  const double d_ad = group_sum<TL>(dotR<R>(X0, Y1)), d_cb = group_sum<TL>(dotR<R>(X1, Y0));
  // true dot products; they only steer angles and tests: fp32 from here on
  const float ac = (cX0.sf * cY0.sf) * d2f(d_ac), bd = (cX1.sf * cY1.sf) * d2f(d_bd);
  const float ad = (cX0.sf * cY1.sf) * d2f(d_ad), cb = (cX1.sf * cY0.sf) * d2f(d_cb);
  const float ab = gX, cd = gY;
  const Rot r1 = rot_params(cX0, cY0, ac, active), r2 = rot_params(cX1, cY1, bd, active);
  col_update(cX0, cY0, r1);
  col_update(cX1, cY1, r2);
  const float c1c2 = r1.cf * r2.cf, c1s2 = r1.cf * r2.sf, s1c2 = r1.sf * r2.cf, s1s2 = r1.sf * r2.sf;
  const float a1d1 = fmaf(c1s2, ab, fmaf(c1c2, ad, -fmaf(s1s2, cb, s1c2 * cd)));
  const float c1b1 = fmaf(s1c2, ab, fmaf(c1c2, cb, -fmaf(s1s2, ad, c1s2 * cd)));
  const float a1b1 = fmaf(c1c2, ab, fmaf(s1s2, cd, -fmaf(c1s2, ad, s1c2 * cb)));
  const float c1d1 = fmaf(s1s2, ab, fmaf(s1c2, ad, fmaf(c1s2, cb, c1c2 * cd)));
  const Rot r3 = rot_params(cX0, cY1, a1d1, active), r4 = rot_params(cX1, cY0, c1b1, active);
  col_update(cX0, cY1, r3);
  col_update(cX1, cY0, r4);
  const float c3c4 = r3.cf * r4.cf, s3s4 = r3.sf * r4.sf;
  gX = fmaf(c3c4, a1b1, s3s4 * c1d1);
  gY = fmaf(s3s4, a1b1, c3c4 * c1d1);
  ss.add(r1); ss.add(r2); ss.add(r3); ss.add(r4);
  if (__any_sync(FULL, r1.on || r2.on || r3.on || r4.on)) {
    const double m1 = -r1.t1, m2 = -r2.t1, m3 = -r3.t1, m4 = -r4.t1;
#pragma unroll
    for (int i = 0; i < R; i++) {
      double x0 = X0[i], x1 = X1[i], y0 = Y0[i], y1 = Y1[i];
      x0 = fma(m1, y0, x0); x1 = fma(m2, y1, x1);
      y0 = fma(r1.t2, x0, y0); y1 = fma(r2.t2, x1, y1);
      x0 = fma(m3, y1, x0); x1 = fma(m4, y0, x1);
      y1 = fma(r3.t2, x0, y1); y0 = fma(r4.t2, x1, y0);
      X0[i] = x0; X1[i] = x1; Y0[i] = y0; Y1[i] = y1;
    }
  }
}

#ifndef EIG_GRAM_ALGEBRA
#define EIG_GRAM_ALGEBRA 1
#endif

template <int NP, int TL, int KB>
__global__ void __launch_bounds__(Cfg<NP, TL, KB>::NTH, (NP == 64 ? (TL == 4 && KB == 2 ? EIG_CTAS_PER_SM : 4) : 1))
    k_eig_fast(int N, const int32_t *__restrict__ mloc, const double *__restrict__ G,
               const double *__restrict__ cin, double *__restrict__ Tout, double *__restrict__ ampl_out,
               float tol, int max_sweeps, DevCounters *ctr) {
  using C = Cfg<NP, TL, KB>;
  constexpr int R = C::R, NG = C::NG, NTH = C::NTH, NB = C::NB, LDW = C::LDW, LDX = C::LDX, NCG = C::NCG;
  constexpr int NW = NTH / 32;
  constexpr int TG = NP / 8;  // Cholesky thread grid TG x TG, 8 x 8 elements per thread (cyclic)
  static_assert(TG * TG <= NTH, "Cholesky thread grid");
  const bool chol = threadIdx.x < TG * TG;  // threads beyond the TG x TG grid only take part in the barriers
  extern __shared__ __align__(16) double sm[];
  double *sW = sm;                   // NP x LDW : L, then exchange buffer, partial sums, Y
  double *s_vec = sm + NP * LDW;     // vectors of NP
  double *s_c = s_vec, *s_uv = s_vec + NP, *s_duw = s_vec + 2 * NP, *s_uw = s_vec + 3 * NP;
  double *s_g1 = s_vec + 4 * NP, *s_g2 = s_vec + 5 * NP, *s_v = s_vec + 6 * NP;
  double *s_xs = s_vec + 7 * NP;     // 6 scalars per lent block, NG+1 regions (< 2*NP doubles)
  __shared__ int s_maxi, s_maxt;
  __shared__ double s_red[NW];
  __shared__ double s_piv;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = tid / TL, r = tid % TL;
  const int zl = blockIdx.x;
  if (mloc[zl] == 0) return;

  // ---- Cholesky A = I + G = L L^T, register tiled: thread (ti,tk) owns A(ti+TG*a, tk+TG*b) ----
  {
    const int ti = tid % TG, tk = chol ? tid / TG : 0;
    double A[8][8];
    const double *Gz = G + (int64_t)zl * NP * NP;
#pragma unroll
    for (int b = 0; b < 8; b++)
#pragma unroll
      for (int a = 0; a < 8; a++) {
        const int i = ti + TG * a, k = tk + TG * b;
        A[a][b] = chol ? Gz[i + NP * k] + (i == k ? 1. : 0.) : 0.;
      }
    if (tid < NP) s_c[tid] = cin[(int64_t)zl * NP + tid];
    double *s_col = s_uv;  // column j of L below the diagonal, zeros above (free until the epilogue)
#pragma unroll
    for (int ja = 0; ja < 8; ja++) {
      for (int jt = 0; jt < TG; jt++) {
        const int j = jt + TG * ja;
        if (chol && ti == jt && tk == jt) s_piv = A[ja][ja];
        __syncthreads();
        if (chol && tk == jt) {  // owners of column j
          const double piv = s_piv;
          const double rinv = rsqrt(piv);
#pragma unroll
          for (int a = 0; a < 8; a++) {
            const int i = ti + TG * a;
            const double l = A[a][ja] * rinv;
            s_col[i] = (i > j) ? l : 0.;
            A[a][ja] = (i > j) ? l : (i == j ? piv * rinv : 0.);
          }
        }
        __syncthreads();
        if (chol) {
          double lc[8];
#pragma unroll
          for (int b = 0; b < 8; b++) lc[b] = s_col[tk + TG * b];
#pragma unroll
          for (int a = 0; a < 8; a++) {
            const double la = -s_col[ti + TG * a];
#pragma unroll
            for (int b = 0; b < 8; b++) A[a][b] = fma(la, lc[b], A[a][b]);
          }
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int b = 0; b < 8; b++)
#pragma unroll
      for (int a = 0; a < 8; a++) {
        const int i = ti + TG * a, k = tk + TG * b;
        if (chol) sW[i + LDW * k] = (k <= i) ? A[a][b] : 0.;
      }
  }
  __syncthreads();

  // ---- registers: group g owns columns 4g..4g+3 ; lane r rows 2*(pi*TL + r) + e ----
  double P0[R], P1[R], Q0[R], Q1[R];
  auto ld_col = [&](double(&X)[R], const double *base) {
#pragma unroll
    for (int pi = 0; pi < R / 2; pi++) {
      const double2 v = *reinterpret_cast<const double2 *>(base + 2 * (pi * TL + r));
      X[2 * pi] = v.x;
      X[2 * pi + 1] = v.y;
    }
  };
  auto st_col = [&](const double(&X)[R], double *base) {
#pragma unroll
    for (int pi = 0; pi < R / 2; pi++)
      *reinterpret_cast<double2 *>(base + 2 * (pi * TL + r)) = make_double2(X[2 * pi], X[2 * pi + 1]);
  };
  if constexpr (KB == 2) {
    ld_col(P0, sW + LDW * (4 * g + 0));
    ld_col(P1, sW + LDW * (4 * g + 1));
    ld_col(Q0, sW + LDW * (4 * g + 2));
    ld_col(Q1, sW + LDW * (4 * g + 3));
  } else {  // one column per block: P1, Q1 unused
    ld_col(P0, sW + LDW * (2 * g + 0));
    ld_col(Q0, sW + LDW * (2 * g + 1));
#pragma unroll
    for (int i = 0; i < R; i++) { P1[i] = 0.; Q1[i] = 0.; }
  }
  __syncthreads();  // sW is now free: exchange buffer
  double *xbuf = sW;
  Col cP0, cP1, cQ0, cQ1;

  constexpr int SX = (KB == 2) ? 7 : 3;  // scalars travelling with a block: (sg, ig, nn) per column + intra dot
  float gP = 0.f, gQ = 0.f;              // true dot product of the two columns of block P / Q (fp32)
  auto lend = [&](int region, const double(&B0)[R], const double(&B1)[R], const Col &c0, const Col &c1, float gi) {
    st_col(B0, xbuf + LDX * (KB * region));
    if constexpr (KB == 2) st_col(B1, xbuf + LDX * (KB * region + 1));
    if (r == 0) {
      double *q = s_xs + SX * region;
      q[0] = c0.sg; q[1] = c0.ig; q[2] = c0.nn;
      if constexpr (KB == 2) { q[3] = c1.sg; q[4] = c1.ig; q[5] = c1.nn; q[6] = (double)gi; }
    }
  };
  auto take = [&](int region, double(&B0)[R], double(&B1)[R], Col &c0, Col &c1, float &gi) {
    ld_col(B0, xbuf + LDX * (KB * region));
    if constexpr (KB == 2) ld_col(B1, xbuf + LDX * (KB * region + 1));
    const double *q = s_xs + SX * region;
    c0.sg = q[0]; c0.ig = q[1]; c0.nn = q[2];
    c0.sf = d2f(c0.sg); c0.nf = d2f(c0.nn);
    if constexpr (KB == 2) {
      c1.sg = q[3]; c1.ig = q[4]; c1.nn = q[5]; gi = d2f(q[6]);
      c1.sf = d2f(c1.sg); c1.nf = d2f(c1.nn);
    }
  };
  // folds the deferred scale into the stored column and refreshes its norm
  auto renorm = [&](double(&X)[R], Col &cx, bool first) {
    if (!first) {
#pragma unroll
      for (int i = 0; i < R; i++) X[i] *= cx.sg;
    }
    cx.sg = 1.; cx.ig = 1.; cx.sf = 1.f;
    cx.nn = group_sum<TL>(dotR<R>(X, X));
    cx.nf = d2f(cx.nn);
  };

  int sweeps = 0;
  for (int sweep = 0; sweep < max_sweeps; sweep++) {
    if (tid == 0) { s_maxi = 0; s_maxt = 0; }
    renorm(P0, cP0, sweep == 0);
    renorm(Q0, cQ0, sweep == 0);
    if constexpr (KB == 2) { renorm(P1, cP1, sweep == 0); renorm(Q1, cQ1, sweep == 0); }
    SweepStat ss{0.f, 0.f};
    // one block pair = KB x KB column pairs, rotated in place
    auto rotate_pair = [&](bool active) {
      if constexpr (KB == 2 && EIG_GRAM_ALGEBRA) {
        rotate_block_pair_gram<R, TL>(P0, P1, Q0, Q1, cP0, cP1, cQ0, cQ1, gP, gQ, active, ss);
      } else if constexpr (KB == 2) {
        rotate_block_pair<R, TL>(P0, P1, Q0, Q1, cP0, cP1, cQ0, cQ1, active, ss);
      } else {
        const float g1 = (cP0.sf * cQ0.sf) * d2f(group_sum<TL>(dotR<R>(P0, Q0)));
        const Rot r1 = rot_params(cP0, cQ0, g1, active);
        ss.add(r1);
        if (__any_sync(FULL, r1.on)) {
          rot_apply<R>(P0, Q0, r1);
          col_update(cP0, cQ0, r1);
        }
      }
    };
    if constexpr (KB == 2) {  // the two columns of each block against each other
      // scales are 1 here (just folded): stored dots are true dots
      const float g1 = d2f(group_sum<TL>(dotR<R>(P0, P1))), g2 = d2f(group_sum<TL>(dotR<R>(Q0, Q1)));
      const Rot r1 = rot_params(cP0, cP1, g1, true), r2 = rot_params(cQ0, cQ1, g2, true);
      ss.add(r1); ss.add(r2);
      if (__any_sync(FULL, r1.on || r2.on)) {
        rot_apply<R>(P0, P1, r1);
        rot_apply<R>(Q0, Q1, r2);
        col_update(cP0, cP1, r1);
        col_update(cQ0, cQ1, r2);
      }
      gP = r1.on ? 0.f : g1;  // rotated pairs are orthogonal; skipped ones keep their (negligible) dot
      gQ = r2.on ? 0.f : g2;
    }
    for (int step = 0; step < NB; step += 2) {
      // even step: positions (2g, 2g+1); afterwards position 2g lives in Q, 2g+1 in P
      rotate_pair(true);
      // odd step: positions (2g+1, 2g+2)
      const bool act = g < NG - 1;
      lend(g, Q0, Q1, cQ0, cQ1, gQ);
      if (!act) lend(NG, P0, P1, cP0, cP1, gP);  // the last group parks its idle block (position NB-1)
      __syncthreads();
      if (act) take(g + 1, Q0, Q1, cQ0, cQ1, gQ);
      rotate_pair(act);
      if (act) lend(g + 1, P0, P1, cP0, cP1, gP);  // the new position 2g+2 goes home
      __syncthreads();
      take(g, P0, P1, cP0, cP1, gP);               // the new position 2g
      if (!act) take(NG, Q0, Q1, cQ0, cQ1, gQ);
    }
    atomicMax(&s_maxi, __float_as_int(ss.mx2));
    atomicMax(&s_maxt, __float_as_int(ss.mt));
    __syncthreads();
    const float mc = sqrtf(__int_as_float(s_maxi)), mt = __int_as_float(s_maxt);
    sweeps = sweep + 1;
    __syncthreads();
    if (jacobi_converged(mc, mt, tol)) break;
    if (sweep == max_sweeps - 1 && tid == 0 && tol >= 0) atomicAdd(&ctr->not_converged, 1);
  }
  // fold the scales: from here on the stored columns are the true z_j
#pragma unroll
  for (int i = 0; i < R; i++) {
    P0[i] *= cP0.sg; Q0[i] *= cQ0.sg;
    if constexpr (KB == 2) { P1[i] *= cP1.sg; Q1[i] *= cQ1.sg; }
  }

  // ---- epilogue: matrix functions from the orthogonal columns z_j = sigma_j u_j ----
  // lane-local views of the vectors c and 1_{i<N}
  double dcol[4], acol[4], bcol[4];
  {
    double cr[R], one[R];
#pragma unroll
    for (int pi = 0; pi < R / 2; pi++)
#pragma unroll
      for (int e = 0; e < 2; e++) {
        const int row = 2 * (pi * TL + r) + e;
        cr[2 * pi + e] = s_c[row];
        one[2 * pi + e] = row < N ? 1. : 0.;
      }
    auto stats = [&](const double(&Z)[R], int slot) {
      const double s2 = group_sum<TL>(dotR<R>(Z, Z));
      const double zc = group_sum<TL>(dotR<R>(Z, cr));
      const double z1 = group_sum<TL>(dotR<R>(Z, one));
      const double s2c = fmax(s2, 1.);  // lambda <- max(lambda,0)   rrsqrt.F90:137
      const double rs = sqrt(s2c);
      dcol[slot] = 1. / (s2 * rs);      // (1+lambda)^-1/2 / |z|^2
      acol[slot] = zc / (s2 * s2c);     // (1+lambda)^-1 (u.c) / |z|
      bcol[slot] = z1 * rs / s2;        // (1+lambda)^+1/2 (u.1) / |z|
    };
    if constexpr (KB == 2) {
      stats(P0, 0); stats(P1, 1); stats(Q0, 2); stats(Q1, 3);
    } else {  // slots 0,1 = P0,Q0 ; the unused P1,Q1 registers are zero and get zero weights
      stats(P0, 0); stats(Q0, 2);
      dcol[1] = acol[1] = bcol[1] = 0.; dcol[3] = acol[3] = bcol[3] = 0.;
    }
  }
  // partial sums over the 4 columns of a group, then over groups through shared memory
  double *s_part = sW;  // [2][NG][NP]
  auto scatter2 = [&](const double(&w1)[4], const double(&w2)[4]) {
#pragma unroll
    for (int pi = 0; pi < R / 2; pi++) {
      double2 o1, o2;
      {
        const int i = 2 * pi;
        o1.x = P0[i] * w1[0] + P1[i] * w1[1] + Q0[i] * w1[2] + Q1[i] * w1[3];
        o2.x = P0[i] * w2[0] + P1[i] * w2[1] + Q0[i] * w2[2] + Q1[i] * w2[3];
        o1.y = P0[i + 1] * w1[0] + P1[i + 1] * w1[1] + Q0[i + 1] * w1[2] + Q1[i + 1] * w1[3];
        o2.y = P0[i + 1] * w2[0] + P1[i + 1] * w2[1] + Q0[i + 1] * w2[2] + Q1[i + 1] * w2[3];
      }
      const int row = 2 * (pi * TL + r);
      *reinterpret_cast<double2 *>(s_part + (size_t)g * NP + row) = o1;
      *reinterpret_cast<double2 *>(s_part + (size_t)(NG + g) * NP + row) = o2;
    }
  };
  __syncthreads();  // everybody is past the last take()
  scatter2(acol, bcol);
  __syncthreads();
  double vi = 0.;
  if (tid < NP) {
    double am = 0.;
    for (int gg = 0; gg < NG; gg++) {
      am += s_part[(size_t)gg * NP + tid];
      vi += s_part[(size_t)(NG + gg) * NP + tid];
    }
    if (tid >= N) { am = 0.; vi = 0.; }
    if (am != am) atomicExch(&ctr->nan_flag, 1);  // rrsqrt.F90:145-149
    ampl_out[(int64_t)zl * NP + tid] = am;
  }
  // v = normate(v)  rrsqrt.F90:178
  {
    double part = (tid < N) ? vi * vi : 0.;
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(FULL, part, o);
    if (lane == 0) s_red[warp] = part;
  }
  __syncthreads();
  double vnorm2 = 0.;
#pragma unroll
  for (int w = 0; w < NW; w++) vnorm2 += s_red[w];
  const double vnorm = sqrt(vnorm2);
  if (tid < NP) s_v[tid] = (tid < N) ? vi / vnorm : 0.;
  __syncthreads();
  // Omega = H_v diag(1,..,1,sign(v_N) sign(w_N)) H_w ; w = 1/sqrt(N)   (rrsqrt.F90:176-185,:737-744)
  const double wN = 1. / sqrt((double)N);
  const double vN = s_v[N - 1];
  const double sv = copysign(1., vN);
  const double dNN = sv;  // sign(w_N) = +1
  const double hv = 1. / (1. + fabs(vN)), hw = 1. / (1. + fabs(wN));
  if (tid < NP) {
    double uv = (tid < N) ? s_v[tid] : 0.;
    double uw = (tid < N) ? wN : 0.;
    if (tid == N - 1) { uv += sv; uw += 1.; }
    s_uv[tid] = uv;
    s_uw[tid] = uw;
    s_duw[tid] = (tid == N - 1) ? dNN * uw : uw;
  }
  __syncthreads();
  // g1 = M u_v, gm2 = M (D u_w):  t_j = d_j (z_j . x) per column, then sum_j z_j t_j
  {
    double xr1[R], xr2[R];
#pragma unroll
    for (int pi = 0; pi < R / 2; pi++)
#pragma unroll
      for (int e = 0; e < 2; e++) {
        const int row = 2 * (pi * TL + r) + e;
        xr1[2 * pi + e] = s_uv[row];
        xr2[2 * pi + e] = s_duw[row];
      }
    double t1[4], t2[4];
    t1[0] = dcol[0] * group_sum<TL>(dotR<R>(P0, xr1)); t2[0] = dcol[0] * group_sum<TL>(dotR<R>(P0, xr2));
    t1[2] = dcol[2] * group_sum<TL>(dotR<R>(Q0, xr1)); t2[2] = dcol[2] * group_sum<TL>(dotR<R>(Q0, xr2));
    if constexpr (KB == 2) {
      t1[1] = dcol[1] * group_sum<TL>(dotR<R>(P1, xr1)); t2[1] = dcol[1] * group_sum<TL>(dotR<R>(P1, xr2));
      t1[3] = dcol[3] * group_sum<TL>(dotR<R>(Q1, xr1)); t2[3] = dcol[3] * group_sum<TL>(dotR<R>(Q1, xr2));
    } else {
      t1[1] = t2[1] = t1[3] = t2[3] = 0.;
    }
    scatter2(t1, t2);
  }
  // kappa = hv u_v . (D u_w)
  {
    double part = (tid < NP) ? s_uv[tid] * hv * s_duw[tid] : 0.;
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(FULL, part, o);
    if (lane == 0) s_red[warp] = part;  // previous readers of s_red are behind two barriers
  }
  __syncthreads();
  double kappa = 0.;
#pragma unroll
  for (int w = 0; w < NW; w++) kappa += s_red[w];
  if (tid < NP) {
    double g1 = 0., gm2 = 0.;
    for (int gg = 0; gg < NG; gg++) {
      g1 += s_part[(size_t)gg * NP + tid];
      gm2 += s_part[(size_t)(NG + gg) * NP + tid];
    }
    s_g1[tid] = g1;
    s_g2[tid] = gm2 - kappa * g1;
  }
  __syncthreads();
  // Y = Z diag(sqrt(d)) to shared memory (column j = 4g+slot)
  double *Ys = sW;
  {
    auto put = [&](const double(&Z)[R], int slot, int colg) {
      const double sc = sqrt(dcol[slot]);
      double *base = Ys + LDW * (NCG * g + colg);
#pragma unroll
      for (int pi = 0; pi < R / 2; pi++)
        *reinterpret_cast<double2 *>(base + 2 * (pi * TL + r)) = make_double2(Z[2 * pi] * sc, Z[2 * pi + 1] * sc);
    };
    if constexpr (KB == 2) { put(P0, 0, 0); put(P1, 1, 1); put(Q0, 2, 2); put(Q1, 3, 3); }
    else { put(P0, 0, 0); put(Q0, 2, 1); }
  }
  __syncthreads();
  // M = Y Y^T in 8x8 register tiles, then T = (M - g1 (hv u_v)^T) D - g2 (hw u_w)^T, row-major
  {
    constexpr int TPR = NP / 8;  // tiles per row
    constexpr int HPT = 2 * TPR * TPR / NTH;  // 8 x 4 half tiles per thread (2, or 1 with twice the threads)
    static_assert(HPT == 1 || HPT == 2, "tile mapping");
    const int tt = tid % (TPR * TPR), h0 = (tid / (TPR * TPR)) * HPT;
    const int ti = tt / TPR, tj = tt % TPR;
    double *Tz = Tout + (int64_t)zl * NP * NP;
#pragma unroll 1
    for (int half = h0; half < h0 + HPT; half++) {  // 8 x 4 register tiles: columns 2tj + (NP/4)(2 half + b)
      double acc[8][4];
#pragma unroll
      for (int a = 0; a < 8; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) acc[a][b] = 0.;
#pragma unroll 2
      for (int j = 0; j < NP; j++) {
        const double *col = Ys + LDW * j;
        double rv[8], cv[4];
#pragma unroll
        for (int a = 0; a < 4; a++) {
          const double2 v = *reinterpret_cast<const double2 *>(col + 8 * ti + 2 * a);
          rv[2 * a] = v.x; rv[2 * a + 1] = v.y;
        }
#pragma unroll
        for (int b = 0; b < 2; b++) {
          const double2 u = *reinterpret_cast<const double2 *>(col + 2 * tj + (NP / 4) * (2 * half + b));
          cv[2 * b] = u.x; cv[2 * b + 1] = u.y;
        }
#pragma unroll
        for (int a = 0; a < 8; a++)
#pragma unroll
          for (int b = 0; b < 4; b++) acc[a][b] = fma(rv[a], cv[b], acc[a][b]);
      }
#pragma unroll
      for (int a = 0; a < 8; a++) {
        const int i = 8 * ti + a;
        const double g1i = s_g1[i] * hv, g2i = s_g2[i] * hw;
#pragma unroll
        for (int b = 0; b < 2; b++) {
          const int k = 2 * tj + (NP / 4) * (2 * half + b);
          double t0 = acc[a][2 * b] - g1i * s_uv[k];
          double t1 = acc[a][2 * b + 1] - g1i * s_uv[k + 1];
          if (k == N - 1) t0 *= dNN;
          if (k + 1 == N - 1) t1 *= dNN;
          t0 -= g2i * s_uw[k];
          t1 -= g2i * s_uw[k + 1];
          *reinterpret_cast<double2 *>(Tz + (int64_t)i * NP + k) = make_double2(t0, t1);
        }
      }
    }
  }
  if (tid == 0) atomicAdd(&ctr->sweeps, (unsigned long long)sweeps);
}

template <int NP, int TL, int KB>
int launch(cudaStream_t st, int N, int nz, const int32_t *mloc, const double *G, const double *c, double *T,
           double *ampl, double tol, int max_sweeps, DevCounters *ctr) {
  using C = Cfg<NP, TL, KB>;
  const size_t smem = sizeof(double) * (NP * C::LDW + 9 * NP);
  static_assert((C::NG + 1) * KB * C::LDX <= NP * C::LDW && (KB == 2 ? 7 : 3) * (C::NG + 1) <= 2 * NP, "exchange buffer fits");
  static_assert(2 * C::NG * NP <= NP * C::LDW, "partial-sum buffer fits");
  { int rc_ = oak_func_smem(k_eig_fast<NP, TL, KB>, (size_t)((int)smem)); if (rc_) return rc_; }
  k_eig_fast<NP, TL, KB><<<nz, C::NTH, smem, st>>>(N, mloc, G, c, T, ampl, (float)tol, max_sweeps, ctr);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

}  // namespace

int oak_launch_eig_simple(cudaStream_t st, int N, int NP, int nz, const int32_t *mloc, const double *G,
                          const double *c, double *T, double *ampl, double tol, int max_sweeps,
                          DevCounters *ctr);

int oak_launch_eig(cudaStream_t st, int kernel, int N, int NP, int zone0, int nz, const int32_t *mloc,
                   const double *G, const double *c, double *T, double *ampl, double tol, int max_sweeps,
                   DevCounters *ctr) {
  if (nz <= 0) return 0;
  const int32_t *ml = mloc + zone0;
  // 0: two columns per block, 4 lanes per group at NP = 64 / 8 at NP = 128.  The variants measured and rejected in round 1
  // (8 lanes per group: -15 %, one column per block: -64 %; profiles/r1_notes.md) are no longer instantiated.
  if (kernel == 0 && NP == 64) return launch<64, 4, 2>(st, N, nz, ml, G, c, T, ampl, tol, max_sweeps, ctr);
  if (kernel == 0 && NP == 128) return launch<128, 8, 2>(st, N, nz, ml, G, c, T, ampl, tol, max_sweeps, ctr);
  if (kernel != 0 && kernel != 1) { oak_set_error("eig_kernel = %d (4 tridiagonal route, 0 block Jacobi, 1 shared-memory cross-check)", kernel); return OAK_ERR_ARG; }
  return oak_launch_eig_simple(st, N, NP, nz, ml, G, c, T, ampl, tol, max_sweeps, ctr);
}
