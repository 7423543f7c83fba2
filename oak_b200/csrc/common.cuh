// common.cuh — shared declarations of the oak_b200 CUDA library (sm_100a only).
//
// Device-side restatement of the observation-selection arithmetic of OAK
// (assimilation.F90:3635-3672 `distance`, :3683-3771 `selectObservations`,
// covariance.F90:645-667 `locfun`) with every operation that feeds the relevance
// predicate written as an explicit round-to-nearest intrinsic (never contracted to
// FMA), so that the index sets are bit-identical to a strict-IEEE evaluation.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#include "../../include/oak_b200.h"
#include "../../include/oak_b200_math.h"

#define OAK_ERR_CUDA (-1)
#define OAK_ERR_ARG (-2)
#define OAK_ERR_STATE (-3)
#define OAK_ERR_NOMEM (-4)
#define OAK_ERR_CAPACITY (-5)
#define OAK_ERR_UNSUPPORTED (-6)
#define OAK_ERR_NAN (-7)

void oak_set_error(const char *fmt, ...);

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device attribute: handles on several GPUs in one process each
// need it, so the "already set" record is keyed by (kernel, device) and guarded by a mutex.
int oak_func_smem_impl(const void *func, int bytes);
template <class F>
inline int oak_func_smem(F *func, size_t bytes) { return oak_func_smem_impl(reinterpret_cast<const void *>(func), (int)bytes); }

#define CUDA_TRY(expr)                                                                     \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      oak_set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return OAK_ERR_CUDA;                                                                 \
    }                                                                                      \
  } while (0)

// ---------------------------------------------------------------------------------------
// fp64 tensor-core instruction: D(8x8) += A(8x4) B(4x8), lane = 4 g + t holds
//   a = A[g][t] ,  b = B[t][g] ,  c0, c1 = C[g][2t], C[g][2t+1]      (PTX ISA, mma.m8n8k4 .f64 fragments)
// OAK_CUEMU is the functional CPU emulation of the kernels used by tests (tools/cuemu); the product build never
// defines it.
// ---------------------------------------------------------------------------------------
#ifdef OAK_CUEMU
__device__ __forceinline__ void oak_dmma_m8n8k4(double &c0, double &c1, double a, double b) {
  cuemu_dmma_m8n8k4(c0, c1, a, b, c0, c1);
}
#else
__device__ __forceinline__ void oak_dmma_m8n8k4(double &c0, double &c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}
#endif

// ---------------------------------------------------------------------------------------
// 16-byte asynchronous global -> shared copies (LDGSTS); the emulation copies at once
// ---------------------------------------------------------------------------------------
#ifdef OAK_CUEMU
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) { memcpy(smem_dst, gmem_src, 16); }
__device__ __forceinline__ void cp_async_commit() {}
template <int N>
__device__ __forceinline__ void cp_async_wait() {}
#else
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
#endif

// ---------------------------------------------------------------------------------------
// Zone / observation-grid descriptors passed by value to kernels
// ---------------------------------------------------------------------------------------
struct ZoneGeom {
  const double *zx, *zy;      // metric coordinates of each zone's first element (zy may alias zeros)
  const double *corrLen, *maxLen;
  const int64_t *zstart;      // [nzones+1] prefix sums of zoneSize
  int32_t loctype, metrictype, weightfun;
  int32_t noloc;              // 1: localise_obs = .false. (rrsqrt.F90:374-385): every observation enters with its weight;
                              //    the relevance predicate only decides whether the zone is analysed at all
};

struct ObsGrid {
  int32_t m;
  int32_t ncx, ncy;           // cells; sorted order = cell-major (cy*ncx+cx), ascending obs index inside
  double x0, y0, csx, csy;    // cell origin / size in bucketing coordinates
  int32_t wrap_x;             // 1: x is a longitude folded to [0,360) (spherical metric)
  const int32_t *cell_start;  // [ncx*ncy+1]
  const int32_t *perm;        // sorted position -> original 0-based observation index
  const double *sx, *sy;      // metric coordinates in sorted order (bit copies of the caller's)
};

// Packed observation-space rows in sorted order (built per analysis by k_pack_obs)
struct ObsRows {
  const double *rows;   // [m][NP] : HSf(l, 0..N-1) zero padded to NP
  const double *delta;  // [m] yo - Hxf
  const double *scoef;  // [m] d01^2 / Rdiag   (R_loc^-1 = w * (e*((e*(w x))/r)), covariance.F90:612-619,:425-431)
};

// ---------------------------------------------------------------------------------------
// exact distance (assimilation.F90:3635-3672); p0 = observation, p1 = zone
// ---------------------------------------------------------------------------------------
#define OAK_PI 3.141592653589793238462643383279502884197
#define OAK_EARTH_RADIUS 6378137.

__device__ __forceinline__ double oak_distance(int metrictype, double x0, double y0, double x1,
                                               double y1) {
  if (metrictype == OAKB200_METRIC_CARTESIAN) {
    const double dx = __dsub_rn(x1, x0), dy = __dsub_rn(y1, y0);
    return __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
  } else if (metrictype == OAKB200_METRIC_SPHERICAL_APPROX) {
    const double pi = OAK_PI;
    const double coeff = __ddiv_rn(__dmul_rn(pi, OAK_EARTH_RADIUS), 180.);
    const double cc = oakm_cos(__dmul_rn(__dadd_rn(y0, y1), __ddiv_rn(pi, 360.)));
    const double u = __dmul_rn(__dmul_rn(coeff, cc), __dsub_rn(x1, x0));
    const double v = __dmul_rn(coeff, __dsub_rn(y1, y0));
    return __dsqrt_rn(__dadd_rn(__dmul_rn(u, u), __dmul_rn(v, v)));
  } else {
    const double pi = OAK_PI;
    const double d2r = __ddiv_rn(pi, 180.);
    const double a = __dmul_rn(y0, d2r), b = __dmul_rn(y1, d2r), C = __dmul_rn(__dsub_rn(x1, x0), d2r);
    double coeff = __dadd_rn(__dmul_rn(oakm_sin(b), oakm_sin(a)),
                             __dmul_rn(__dmul_rn(oakm_cos(b), oakm_cos(a)), oakm_cos(C)));
    coeff = fmax(fmin(coeff, 1.), -1.);
    return __dmul_rn(OAK_EARTH_RADIUS, oakm_acos(coeff));
  }
}

// Gaspari-Cohn, Horner form exactly as covariance.F90:652-663
__device__ __forceinline__ double oak_locfun(double r) {
  if (r <= 1.) {
    double p = __dadd_rn(__ddiv_rn(-r, 4.), __ddiv_rn(1., 2.));
    p = __dadd_rn(__dmul_rn(p, r), __ddiv_rn(5., 8.));
    p = __dsub_rn(__dmul_rn(p, r), __ddiv_rn(5., 3.));
    return __dadd_rn(__dmul_rn(p, __dmul_rn(r, r)), 1.);
  } else if (r <= 2.) {
    double p = __dsub_rn(__ddiv_rn(r, 12.), __ddiv_rn(1., 2.));
    p = __dadd_rn(__dmul_rn(p, r), __ddiv_rn(5., 8.));
    p = __dadd_rn(__dmul_rn(p, r), __ddiv_rn(5., 3.));
    p = __dsub_rn(__dmul_rn(p, r), 5.);
    p = __dadd_rn(__dmul_rn(p, r), 4.);
    return __dsub_rn(p, __ddiv_rn(2., __dmul_rn(3., r)));
  }
  return 0.;
}

struct ZoneQuery {
  double x, y, corr, maxl;
  int32_t loctype, metrictype, weightfun, noloc;
};

// the callback body for one (zone, observation) pair: relevance flag + weight
__device__ __forceinline__ bool oak_obs_relevant(const ZoneQuery &q, double ox, double oy, double &w) {
  double d;
  if (q.loctype == OAKB200_LOC_HORIZONTAL)
    d = oak_distance(q.metrictype, ox, oy, q.x, q.y);
  else
    d = fabs(__dsub_rn(ox, q.x));  // |obsZ - z| or |obsT - t| (assimilation.F90:3750-3753)
  if (q.weightfun == OAKB200_WEIGHT_GAUSSIAN) {
    const double t = __ddiv_rn(d, q.corr);
    w = exp(-__dmul_rn(t, t));
    return d <= q.maxl;
  } else if (q.weightfun == OAKB200_WEIGHT_GASPARI_COHN) {
    w = oak_locfun(__ddiv_rn(d, q.corr));
    return w != 0.;
  }
  w = 1.;
  return true;
}

// ---------------------------------------------------------------------------------------
// Conservative cell box of a zone: every observation that can pass the predicate lies in a
// cell of rows [cy0,cy1] and of the x-ranges [xa0,xa1] (and [xb0,xb1] when the longitude
// interval wraps).  Same contract as ndgrid.F90:1589 `near` ("or more"): a superset that the
// exact predicate then filters.
// ---------------------------------------------------------------------------------------
struct CellBox {
  int32_t cy0, cy1, xa0, xa1, xb0, xb1;  // xb0 > xb1 : no second range
};

__host__ __device__ __forceinline__ int32_t oak_cell_of(double v, double origin, double cs, int32_t nc) {
  double f = floor((v - origin) / cs);
  if (!(f > 0.)) return 0;  // also NaN
  if (f >= (double)nc) return nc - 1;
  return (int32_t)f;
}

__host__ __device__ __forceinline__ double oak_fold360(double lon) {
  double r = lon - 360. * floor(lon / 360.);
  if (!(r >= 0.)) r = 0.;
  if (r >= 360.) r = 0.;
  return r;
}

__device__ __forceinline__ CellBox oak_zone_box(const ObsGrid &g, const ZoneQuery &q) {
  CellBox b;
  b.xb0 = 1; b.xb1 = 0;
  double R = (q.weightfun == OAKB200_WEIGHT_GAUSSIAN) ? q.maxl
             : (q.weightfun == OAKB200_WEIGHT_GASPARI_COHN ? 2. * q.corr : INFINITY);
  const bool all = !(R < 1e300) || !(R == R) || q.noloc;
  if (all || (g.ncx == 1 && g.ncy == 1)) {
    b.cy0 = 0; b.cy1 = g.ncy - 1; b.xa0 = 0; b.xa1 = g.ncx - 1;
    return b;
  }
  if (R < 0.) R = 0.;
  if (q.loctype != OAKB200_LOC_HORIZONTAL || q.metrictype == OAKB200_METRIC_CARTESIAN) {
    const double mx = R * 1e-9 + 1e-12 * (fabs(q.x) + fabs(g.x0)) + R;
    const double my = R * 1e-9 + 1e-12 * (fabs(q.y) + fabs(g.y0)) + R;
    b.xa0 = oak_cell_of(q.x - mx, g.x0, g.csx, g.ncx);
    b.xa1 = oak_cell_of(q.x + mx, g.x0, g.csx, g.ncx);
    if (q.loctype != OAKB200_LOC_HORIZONTAL) { b.cy0 = 0; b.cy1 = g.ncy - 1; return b; }
    b.cy0 = oak_cell_of(q.y - my, g.y0, g.csy, g.ncy);
    b.cy1 = oak_cell_of(q.y + my, g.y0, g.csy, g.ncy);
    return b;
  }
  // spherical metrics: x = longitude, y = latitude in degrees
  const double deg = 180. / OAK_PI;
  // acos() near 1 resolves angles only to ~sqrt(eps) = 2e-8 rad: absolute slack of 2e-7 rad
  const double dlat = ((R / OAK_EARTH_RADIUS) * (1. + 1e-6) + 2e-7) * deg;
  b.cy0 = oak_cell_of(q.y - dlat, g.y0, g.csy, g.ncy);
  b.cy1 = oak_cell_of(q.y + dlat, g.y0, g.csy, g.ncy);
  const double latmax = fmin(fabs(q.y) + dlat, 90.);
  const double cl = cos(latmax / deg);
  double dlon;
  if (q.metrictype == OAKB200_METRIC_SPHERICAL) {
    // small circle of angular radius delta around latitude phi: |dlon| <= asin(sin(delta)/cos(phi))
    const double delta = (R / OAK_EARTH_RADIUS) * (1. + 1e-6) + 2e-7;
    const double cphi = cos(fabs(q.y) / deg);
    const double sd = sin(fmin(delta, OAK_PI / 2));
    if (delta >= OAK_PI / 2 || !(cphi > 0.) || sd >= cphi * (1. - 1e-6) || fabs(q.y) + dlat >= 90.)
      dlon = 1e30;
    else
      dlon = asin(sd / cphi) * deg * (1. + 1e-6) + 1e-9;
  } else {
    // approx metric: |c*cos(mean lat)*dlon| <= R, mean lat within [phi-dlat/2, phi+dlat/2]
    const double latm = fmin(fabs(q.y) + 0.5 * dlat, 90.);
    const double cm = cos(latm / deg);
    (void)cl;
    if (!(cm > 1e-9)) dlon = 1e30;
    else dlon = ((R / (OAK_EARTH_RADIUS * cm)) * (1. + 1e-6) + 2e-7) * deg;
  }
  if (!g.wrap_x) {
    b.xa0 = oak_cell_of(q.x - dlon, g.x0, g.csx, g.ncx);
    b.xa1 = oak_cell_of(q.x + dlon, g.x0, g.csx, g.ncx);
    return b;
  }
  if (dlon >= 180.) { b.xa0 = 0; b.xa1 = g.ncx - 1; return b; }
  const double lo = oak_fold360(q.x - dlon), hi = oak_fold360(q.x + dlon);
  if (lo <= hi) {
    b.xa0 = oak_cell_of(lo, g.x0, g.csx, g.ncx);
    b.xa1 = oak_cell_of(hi, g.x0, g.csx, g.ncx);
  } else {  // wraps through 360 -> 0
    b.xa0 = oak_cell_of(lo, g.x0, g.csx, g.ncx);
    b.xa1 = g.ncx - 1;
    b.xb0 = 0;
    b.xb1 = oak_cell_of(hi, g.x0, g.csx, g.ncx);
    if (b.xb1 >= b.xa0) { b.xa0 = 0; b.xb0 = 1; b.xb1 = 0; }  // overlap: take everything
  }
  return b;
}

__device__ __forceinline__ ZoneQuery oak_zone_query(const ZoneGeom &zg, int32_t zone) {
  ZoneQuery q;
  q.x = zg.zx[zone];
  q.y = zg.zy ? zg.zy[zone] : 0.;
  q.corr = zg.corrLen[zone];
  q.maxl = zg.maxLen[zone];
  q.loctype = zg.loctype; q.metrictype = zg.metrictype; q.weightfun = zg.weightfun; q.noloc = zg.noloc;
  return q;
}

// ---------------------------------------------------------------------------------------
// kernel launchers (one per translation unit)
// ---------------------------------------------------------------------------------------
struct BatchWs {          // per-stream workspace for a batch of zones
  double *G;              // [zb][NP*NP]  Gram matrix, then reused for T (row-major T[k][k'])
  double *T;              // [zb][NP*NP]
  double *c;              // [zb][NP]     HSf^T R_loc^-1 (yo-Hxf), then ampl
  double *ampl;           // [zb][NP]
};

// tabulated anamorphosis (AnamTrans%anam(v)%transform, K x 2 column-major on the device; oakb200_set_anamorphosis_table)
struct AnamTab {
  const double *tab;   // [0..K) physical values, [K..2K) transformed values
  int32_t K;
  int32_t monotone;    // both columns strictly increasing: the bracket search may bisect
  // per-variable transforms (anamtype 0; assimilation.F90:4531-4535 looks AnamTrans%anam(v) up for every element):
  const int32_t *rowvar;   // [n] 0-based variable of each (zone-permuted) row
  const int32_t *vdesc;    // [nvar][4] : type, K, offset of the K x 2 table in vtab (doubles), monotone
  const double *vtab;
};

// interp1 (anamorphosis.F90:304-339): first bracket x_k <= xi < x_k+1, (1-alpha) y_k + alpha y_k+1; without a
// bracket y_1 (xi < x_1) or y_K, and `out`.  A strictly increasing x has at most one bracket, found by bisection;
// otherwise the reference's linear scan.  Explicitly rounded operations (no FMA contraction): same bits as the
// Fortran expression evaluated in IEEE arithmetic.
__device__ __forceinline__ double oak_interp1(int K, const double *x, const double *y, bool monotone, double xi, bool &out) {
  int k = -1;
  if (monotone) {
    if (xi >= x[0] && xi < x[K - 1]) {
      int lo = 0, hi = K - 1;  // x[lo] <= xi < x[hi]
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (x[mid] <= xi) lo = mid; else hi = mid;
      }
      k = lo;
    }
  } else {
    for (int kp = 0; kp < K - 1; kp++)
      if (x[kp] <= xi && xi < x[kp + 1]) { k = kp; break; }
  }
  out = (k == -1);
  if (k != -1) {
    const double alpha = __ddiv_rn(__dsub_rn(xi, x[k]), __dsub_rn(x[k + 1], x[k]));
    return __dadd_rn(__dmul_rn(__dsub_rn(1., alpha), y[k]), __dmul_rn(alpha, y[k + 1]));
  }
  return (xi < x[0]) ? y[0] : y[K - 1];
}

// anamtransform for one element (assimilation.F90:4516-4576): 1 identity, 2 log/exp, 3 tabulated.  An
// extrapolated tabulated value is replaced by the first / last entry of the table's INPUT-side column,
// chosen by comparing the already interpolated value with transform(1,ti) (:4560-4567, reproduced as is).
__device__ __forceinline__ double oak_anam(int type, bool forward, const AnamTab &at, double x) {
  if (type == 2) return forward ? log(x) : exp(x);
  if (type == 3) {
    const double *ti = forward ? at.tab : at.tab + at.K, *tj = forward ? at.tab + at.K : at.tab;
    bool out;
    double v = oak_interp1(at.K, ti, tj, at.monotone != 0, x, out);
    if (out) v = (v < ti[0]) ? ti[0] : ti[at.K - 1];
    return v;
  }
  return x;
}

// anamtype 0: the transform of the row's own variable
__device__ __forceinline__ double oak_anam_row(int type, bool forward, const AnamTab &at, int64_t row, double x) {
  if (type != 0) return oak_anam(type, forward, at, x);
  const int32_t *d = at.vdesc + 4 * at.rowvar[row];
  AnamTab sub = at;
  sub.tab = at.vtab + d[2]; sub.K = d[1]; sub.monotone = d[3];
  return oak_anam(d[0], forward, sub, x);
}

// Ensemble branch of Assim folded into the apply kernel (oakb200_assim_ensemble[_dev], option "ens_fuse"): the array
// the kernel reads holds the raw ensemble E; per staged chunk of rows it does the prologue (forward anamorphosis,
// xf = mean, Sf = (E - xf)/scaling; assimilation.F90:3123-3131) before the product and the epilogue (inflation,
// saturation of the correction, Ea = xa + scaling Sa, inverse anamorphosis, xa = mean(Ea); :3301-3349) after it,
// with the operations and summation order of k_mean_anom / k_epilogue, so E is read once and Ea written once.
struct EnsFuse {
  int32_t on;              // 0: plain apply (the array holds anomalies)
  int32_t anamtype;
  AnamTab at;
  double inflation, scaling;
  const double *maxCorr;   // [n] or NULL
  double *xf_out;          // [n] mean of the (transformed) forecast ensemble
};

// destinations of the fused all-gather (oakb200_set_peer_outputs), passed by value to k_apply
struct PeerOut {
  double *Sa[OAKB200_MAX_PEERS];
  double *xa[OAKB200_MAX_PEERS];
  int64_t ld, row0;
  int32_t n;
};

// state arrays of the batch for the apply fused into the transform kernel (option "fuse_apply")
struct FusedApplyArgs {
  const int64_t *zstart;   // prefix sums of the zone sizes, offset to the first zone of the batch
  int64_t rowbase;         // global row of the first row held in xf / Sf / xa / Sa
  const double *xf, *Sf;
  double *xa, *Sa;
  int64_t ldS, ldSa;
};

struct DevCounters {      // device-side statistics / status
  unsigned long long relevant, candidates, sweeps, skipped;
  int nan_flag;
  int not_converged;
  unsigned long long fallback;  // zones the tridiagonal route handed to the Jacobi kernel
  unsigned long long fb_reason[4];  // ... because of: QL iterations, residual test, group size, parallel vectors
  unsigned long long gs_pairs;      // Gram-Schmidt projections done inside close groups
  unsigned long long tw_fallback;   // eigenvectors of T computed by the pivot form because the Sturm products underflowed
};

int oak_launch_pack_obs(cudaStream_t st, int m, int N, int NP, const int32_t *perm, const double *HSf,
                        int64_t ldH, const double *yo, const double *Hxf, const double *Rdiag,
                        const double *d01, double *rows, double *delta, double *scoef);
int oak_launch_select(cudaStream_t st, const ZoneGeom &zg, const ObsGrid &og, int zone0, int nz,
                      const int64_t *offsets, int32_t *counts, int32_t *idx, double *w, bool fill);
int oak_launch_gram(cudaStream_t st, int NP, const ZoneGeom &zg, const ObsGrid &og, const ObsRows &orows,
                    int zone0, int nz, double *G, double *c, int32_t *mloc, DevCounters *ctr);
int oak_launch_gram_mma(cudaStream_t st, int variant, int NP, const ZoneGeom &zg, const ObsGrid &og,
                        const ObsRows &orows, int zone0, int nz, double *G, double *c, int32_t *mloc,
                        DevCounters *ctr);
int oak_launch_eig(cudaStream_t st, int kernel, int N, int NP, int zone0, int nz, const int32_t *mloc,
                   const double *G, const double *c, double *T, double *ampl, double tol, int max_sweeps,
                   DevCounters *ctr);
size_t oak_eig_tridiag_ws_bytes(int NP, int nz);
// k_tql on a stream of its own (experiment / option "tql_side"): it is latency bound (one thread per zone), so it can
// be given a high-priority stream whose few CTAs are placed as soon as an SM has room
struct TqlSide { cudaStream_t qst; cudaEvent_t e0, e1, e2, e3; };
int oak_launch_eig_tridiag(cudaStream_t st, int N, int NP, int nz, const int32_t *mloc, const double *G,
                           const double *c, double *T, double *ampl, void *ws, int32_t **flags_out,
                           DevCounters *ctr, cudaEvent_t *ev /* optional: [0] after k_tridiag, [1] after k_tql */,
                           double orthtol /* <= 0: default */, int maxgroup /* < 0: default */,
                           const FusedApplyArgs *fuse /* NULL: k_tvec writes T */,
                           double *Wg /* NULL, or [nz][NP*NP] workspace: k_tvec runs as two kernels (option tvec_split) */,
                           const TqlSide *side = nullptr);
int oak_launch_apply(cudaStream_t st, int N, int NP, const ZoneGeom &zg, int zone0, int nz,
                     int64_t rowbase, const int32_t *mloc, const double *T, const double *ampl,
                     const double *xf, const double *Sf, int64_t ldS, double *xa, double *Sa,
                     int64_t ldSa, const PeerOut &peers,
                     const int32_t *only_flagged = nullptr /* batch-local flags: analysed zones with flag 0 are skipped */,
                     bool shared_transform = false /* global scheme: every block of rows uses T[0], ampl[0]; mloc may be NULL */,
                     int uniform_rows = 0 /* > 0: every zone has this many rows (enables the TMA-staged kernel) */,
                     int64_t rows_in_buffers = 0 /* rows held in Sf / Sa from their first element (tensor-map extent) */,
                     const EnsFuse *ens = nullptr /* ensemble prologue / epilogue inside the kernel (Sf = raw ensemble) */);
int oak_launch_apply_mma(cudaStream_t st, int N, int NP, const ZoneGeom &zg, int zone0, int nz, int64_t rowbase,
                         const int32_t *mloc, const double *T, const double *ampl, const double *xf, const double *Sf,
                         int64_t ldS, double *xa, double *Sa, int64_t ldSa, const int32_t *only_flagged,
                         bool shared_transform);
// global scheme (global.cu)
size_t oak_global_gram_ws_bytes(int NP, int nparts);
int oak_global_gram_parts(int m);
int oak_launch_global_gram(cudaStream_t st, int m, int N, int NP, const double *HSf, int64_t ldH, const double *yo,
                           const double *Hxf, const double *Rdiag, const double *d01, void *ws, int nparts, double *G,
                           double *c, int32_t *mloc);
int oak_launch_block_starts(cudaStream_t st, int64_t n, int rows_per_block, int nblocks, int64_t *zstart);
int oak_fp64_peak(int mode, double *tflops);

// observation-operator generation: batched cinterp (hgen.cu)
size_t oak_cinterp_tet_doubles(int n);
int oak_launch_cinterp(cudaStream_t st, int n, const int32_t *gshape, const double *d_axes, const uint8_t *d_masked,
                       double *d_tet, int m, const double *d_xi, int32_t *d_indexes, double *d_coeff, int32_t *d_nbp,
                       int *d_ndeg);

// ensemble prologue / epilogue (assimilation.F90:3106-3134, :3301-3357)
int oak_launch_obsoper(cudaStream_t st, int m, int N, int64_t nnz, const int32_t *Hi, const int32_t *Hj,
                       const double *Hs, const double *Hshift, const double *E, int64_t ldE, double *HE);
int oak_launch_mean_anom(cudaStream_t st, int64_t rows, int N, int anamtype, AnamTab at, const double *E, int64_t ldE,
                         double *mean, double *S, int64_t ldS);
int oak_launch_epilogue(cudaStream_t st, int64_t rows, int N, int anamtype, AnamTab at, double inflation,
                        const double *maxCorr, const double *xf, double *xa, const double *Sa,
                        int64_t ldSa, double *Ea, int64_t ldEa);
