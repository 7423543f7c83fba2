// api.cu — the C ABI (include/oak_b200.h): handle, zone/observation setup, batching of the zone loop
// over CUDA streams, host-buffer streaming, statistics.
//
// The zone loop of locAnalysisIncrement (rrsqrt.F90:357-418) becomes, per batch of zones,
//     k_gram (selection + Gram)  ->  k_eig (transform)  ->  k_apply (update of the zone rows)
// on one of NSLOT streams; consecutive batches go to different streams so that the tail of one
// batch overlaps the head of the next.  With host buffers the state is cut in chunks of whole
// zones; chunk c uses slot c % NSLOT: H2D copy, batches, D2H copy, all on the slot's stream, so the
// copies of one chunk overlap the kernels of the others.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "common.cuh"

// from the other translation units
int oak_build_obsgrid(cudaStream_t st, int m, const double *bx, const double *by, const ObsGrid &g,
                      uint32_t *key_in, uint32_t *key_out, int32_t *val_in, int32_t *perm,
                      int32_t *cell_start, double *sx, double *sy, void *tmp, size_t tmp_bytes);
size_t oak_obsgrid_scratch_bytes(int m, int ncell);
size_t oak_coo_scratch_bytes(int64_t nnz, int m);
int oak_coo_to_rows(cudaStream_t st, int64_t nnz, int m, const int32_t *Hi, const int32_t *Hj, uint32_t *key_in,
                    uint32_t *key_out, int32_t *val_in, int32_t *order, int32_t *rowstart, void *tmp,
                    size_t tmp_bytes);
int oak_launch_obsoper_rows(cudaStream_t st, int m, int N, const int32_t *rowstart, const int32_t *order,
                            const int32_t *Hj, const double *Hs, const double *Hshift, const double *E,
                            int64_t ldE, double *HE);

static thread_local std::string g_err;
void oak_set_error(const char *fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
}

#include <map>
#include <mutex>
int oak_func_smem_impl(const void *func, int bytes) {
  static std::mutex mu;
  static std::map<std::pair<const void *, int>, int> done;  // (kernel, device) -> bytes granted
  int dev = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(mu);
  auto it = done.find({func, dev});
  if (it != done.end() && it->second >= bytes) return 0;
  CUDA_TRY(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  done[{func, dev}] = bytes;
  return 0;
}
extern "C" OAKB200_API const char *oakb200_last_error(void) { return g_err.c_str(); }
extern "C" OAKB200_API int oakb200_version(void) { return 100; }
#ifdef OAK_CUEMU
// only in the CPU emulation build used by tests (tools/cuemu): lets the Python layer refuse it as a product library
extern "C" OAKB200_API int oakb200_emulated(void) { return 1; }
#endif

namespace {

#ifndef OAK_NSLOT
#define OAK_NSLOT 4
#endif
constexpr int NSLOT = OAK_NSLOT;

struct DevBuf {
  void *p = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes) {
    if (bytes <= cap) return 0;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) {
      oak_set_error("cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
      p = nullptr;
      return OAK_ERR_NOMEM;
    }
    cap = bytes;
    return 0;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
  template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

// Pinned staging of a chunk of the caller's PAGEABLE state arrays (option host_stage): copies from pageable memory are
// staged by the driver through one thread (measured 0.37 M columns/s end to end on C3 against 2.6 M from pinned arrays,
// and cudaHostRegister per call costs more than it saves); here the library owns two pinned buffers per stream slot and
// moves the chunk in and out with several host threads, overlapped with the kernels of the other slots.
struct HostStage {
  double *in = nullptr, *out = nullptr;
  size_t cap = 0;
  cudaEvent_t ev_out = nullptr;
  bool pending = false;          // a finished chunk sits in `out` and has not been copied to the caller's arrays yet
  int64_t r0 = 0, rows = 0;
};

struct Slot {
  HostStage hs;
  cudaStream_t st = nullptr;
  cudaStream_t cst = nullptr;  // copy-engine pushes of finished batches to the peers (peer_mode 1)
  bool cst_used = false;
  cudaStream_t qst = nullptr;  // option tql_side: k_tql on a high-priority stream of its own
  cudaEvent_t qev[4] = {};
  DevBuf Wv;                   // option tvec_split: eigenvectors of T between the two halves of k_tvec
  DevBuf G, T, c, ampl, tri;   // batch workspace (tri: d, e, tau, lambda, flags of the tridiagonal route)
  DevBuf S, xf, xa;            // host-streaming chunk buffers
  cudaEvent_t ev[12] = {};
};

}  // namespace

struct oakb200_handle {
  int device = 0;
  // options
  int eig_kernel = 4;
  int apply_kernel = 0;       // 0: k_apply (DFMA register tiles); 1: k_apply_mma (DMMA; NP = 64, not with peer_mode 0)
  int scheme = 1;             // ensemble entry points: 1 = local scheme (default), 0 = global scheme (schemetype, assimilation.F90:292)
  int tvec_split = 0;         // 1: tridiagonal route, k_tvec as two kernels (eigenvectors of T | back-transformation and the rest)
  int fuse_apply = 0;         // 1: tridiagonal route, k_tvec updates the zone rows from the factored transform (no T, no k_apply)
  int push_kernel = 1;        // fused gather: 1 (default: 47.8 -> 46.1 ms per C3 step at 8 GPUs) = k_push on the slot's side stream, 0 = copy-engine copies
  int push_ctas = 24;         // CTAs of the push kernels (8 GPUs, multicast: 8 / 16 / 24 / 32 / 64 / 148 CTAs -> 36.7 / 36.2 / 36.2 / 36.3 / 38.4 / 40.4 ms)
  int ens_fuse = -1;          // oakb200_assim_ensemble[_dev], local scheme: 1 = prologue (anamorphosis, mean, anomalies) and epilogue (inflation,
                              // saturation, Ea, inverse anamorphosis) inside the apply kernel: E read once, Ea written once (4 of 6 passes saved);
                              // 0 = three-pass form (k_mean_anom, analysis in place, k_epilogue); -1 (default) = fused when there is no anamorphosis
                              // (identity): measured on C5 (B200), the 2 x nz x N log / exp per zone cost more inside the fp64-bound apply kernel
                              // (27.9 ms per step) than under the HBM-bound streaming kernels (27.1 ms)
  EnsFuse ens{};              // set for the duration of such a call
  int apply_tma = 1;          // k_apply_tma (zone rows staged by 2-D tensor copies) where its conditions hold, else k_apply
  int host_stage = -1;        // host-buffer entry points, pageable caller arrays: 1 = chunks go through pinned staging buffers filled by
                              // `stage_threads` host threads, 0 = asynchronous copies straight from the caller's arrays, -1 (default) = 1 for
                              // calls of 32 MB and more
  int stage_threads = 0;      // 0: min(16, hardware threads)
  int host_register = 0;      // host-buffer entry points: 1 = page-lock the caller's pageable arrays for the duration of the call
                              // (measured on C3: registering 31 GB per call costs more than the driver's staged copies:
                              // 0.09 vs 0.40 M columns/s; pinned buffers from oakb200_host_alloc: 2.68 M)
  int localise_obs = 1;       // 0: locAnalysis(..., localise_obs=.false.) (rrsqrt.F90:374-385): all observations with their weights, amplitudes filled
  int tql_side = 1;           // 1 (default: 286.7 -> 280.9 ms per C3 step): k_tql on the slot's high-priority side stream
  int gram_kernel = 1;        // 1 (default since round 2: 8.6 -> 5.7 ms per 90 k zones) / 2: k_gram_mma (DMMA, 4 / 2 warps per zone; NP = 64); 3 / 4: same, 32-candidate chunks; 0: k_gram (DFMA register tiles)
  double tri_orthtol = 0.;  // tridiagonal route: accepted loss of orthogonality between neighbouring eigenvectors (0: default)
  int tri_maxgroup = -1;
  DevBuf d_anam;              // tabulated anamorphosis (K x 2), oakb200_set_anamorphosis_table
  DevBuf d_tet;               // simplex table of the batched cinterp (+ its degenerate-cell counter)
  int anam_K = 0, anam_monotone = 0;
  DevBuf d_rowvar, d_vdesc, d_vtab;   // per-variable transforms (oakb200_set_anamorphosis_vars)
  int anam_nvar = 0; int64_t anam_rows = 0;
  PeerOut peers{};            // fused all-gather destinations (oakb200_set_peer_outputs); n = 0: none
  double *mc_Sa = nullptr, *mc_xa = nullptr;   // multicast addresses of the result arrays (oakb200_set_multicast_output)
  int64_t mc_ld = 0, mc_row0 = 0;
  cudaStream_t pstream[OAKB200_MAX_PEERS] = {};  // one copy stream per destination (created on first use)
  cudaEvent_t pev[OAKB200_MAX_PEERS] = {};
  int push_pieces = 1;        // peer_mode 1: the apply of a batch is launched in this many pieces, each pushed as soon as it is done
  int peer_mode = 1;          // 1: copy engines push every finished batch (no SM time); 0: stores of k_apply    // ... largest group of close eigenvalues orthogonalised in place (-1: default)
  int zones_per_batch = 0;
  int taper = 0;              // > 0: batches inside the last (taper + 1) x zones_per_batch zones of a call halve (see local_analysis_dev)
  double tol = 2e-11;  // bound on the remaining non-orthogonality (eig_common.cuh: jacobi_converged)
  int max_sweeps = 30;
  int profile = 0;
  double chunk_mb = 256.;
  int pad_to = 0;
  int async = 0;              // local_analysis_dev returns after enqueueing; oakb200_synchronize collects
  int order_after_caller = 1; // slot streams wait for the caller's stream before starting
  int stream_priority = 0;
  bool pending = false;       // an asynchronous call has not been synchronised yet
  int64_t pending_launches = 0;
  // zones
  bool zones_set = false;
  int nzones = 0;
  int64_t nrows = 0;
  int max_zone_rows = 0, min_zone_rows = 0;
  int loctype = 1, metrictype = 0, weightfun = 0;
  std::vector<int64_t> h_zstart;
  double rmax = 0.;     // largest finite search radius
  bool any_unbounded = false;
  double zlat_absmax = 0.;
  DevBuf d_zx, d_zy, d_corr, d_maxl, d_zstart, d_mloc;
  // observations
  bool obs_set = false;
  int m = 0;
  ObsGrid og{};
  DevBuf d_bx, d_by, d_key_in, d_key_out, d_val_in, d_perm, d_cell_start, d_sx, d_sy, d_tmp;
  // packed observation rows
  DevBuf d_rows, d_delta, d_scoef;
  // host-path staging of observation-space arrays
  DevBuf d_HSf, d_yo, d_Hxf, d_R, d_d01, d_ampzero;
  // ensemble path
  DevBuf d_HE, d_Hi, d_Hj, d_Hs, d_Hshift, d_order, d_rowstart, d_xf, d_xa, d_maxc, d_E;
  DevBuf d_ctr;
  DevBuf d_gws, d_gzstart;    // global scheme: partial Gram matrices, prefix sums of the row blocks
  Slot slot[NSLOT];
  cudaEvent_t ev_a = nullptr, ev_b = nullptr, ev_user = nullptr;

  ZoneGeom geom() const {
    ZoneGeom z;
    z.zx = d_zx.as<double>(); z.zy = d_zy.as<double>();
    z.corrLen = d_corr.as<double>(); z.maxLen = d_maxl.as<double>();
    z.zstart = d_zstart.as<int64_t>();
    z.loctype = loctype; z.metrictype = metrictype; z.weightfun = weightfun;
    z.noloc = localise_obs ? 0 : 1;
    return z;
  }
};

namespace {

int padded(int N) { return N <= 32 ? 32 : (N <= 64 ? 64 : (N <= 128 ? 128 : -1)); }
int padded(const oakb200_handle *h, int N) {
  const int np = padded(N);
  return (np > 0 && h->pad_to > np && h->pad_to <= 128) ? padded(h->pad_to) : np;
}

// Page-locks pageable caller arrays for the duration of a host-buffer call (a Fortran caller's Sf is ordinary
// allocatable memory: asynchronous copies from it would be staged synchronously by the driver) and releases them when
// the call returns, on every path.  Arrays that are already pinned or registered are left alone.
struct HostPins {
  std::vector<void *> regs;
  void pin(const void *p, size_t bytes) {
    if (!p || bytes < (1u << 20)) return;   // small arrays: not worth a system call
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return; }
    if (a.type != cudaMemoryTypeUnregistered) return;
    for (void *q : regs) if (q == p) return;
    if (cudaHostRegister(const_cast<void *>(p), bytes, cudaHostRegisterPortable) == cudaSuccess) regs.push_back(const_cast<void *>(p));
    else cudaGetLastError();                 // not fatal: the copies fall back to the driver's staging
  }
  ~HostPins() { for (void *q : regs) cudaHostUnregister(q); }
};

// dst(rows x ncols, leading dimension dld) = src(rows x ncols, leading dimension sld) with `nt` host threads, each a
// contiguous share of the flattened (column, row) range
void par_copy_cols(double *dst, size_t dld, const double *src, size_t sld, size_t rows, int ncols, int nt) {
  const size_t total = rows * (size_t)ncols;
  if (total == 0) return;
  auto work = [=](size_t a, size_t b) {
    while (a < b) {
      const size_t col = a / rows, r = a - col * rows, len = std::min(rows - r, b - a);
      memcpy(dst + col * dld + r, src + col * sld + r, len * sizeof(double));
      a += len;
    }
  };
  nt = (int)std::max<size_t>(1, std::min<size_t>((size_t)nt, total / (1u << 16) + 1));   // >= 512 KB per thread
  if (nt == 1) { work(0, total); return; }
  std::vector<std::thread> th;
  th.reserve(nt - 1);
  for (int t = 1; t < nt; t++) th.emplace_back(work, total * t / nt, total * (t + 1) / nt);
  work(0, total / nt);
  for (auto &x : th) x.join();
}

int ensure_stage(Slot &s, size_t doubles) {
  if (!s.hs.ev_out) CUDA_TRY(cudaEventCreateWithFlags(&s.hs.ev_out, cudaEventDisableTiming));
  if (doubles <= s.hs.cap) return 0;
  if (s.hs.in) cudaFreeHost(s.hs.in);
  if (s.hs.out) cudaFreeHost(s.hs.out);
  s.hs.in = s.hs.out = nullptr; s.hs.cap = 0;
  void *a = nullptr, *b = nullptr;
  if (cudaHostAlloc(&a, doubles * sizeof(double), cudaHostAllocPortable) != cudaSuccess ||
      cudaHostAlloc(&b, doubles * sizeof(double), cudaHostAllocPortable) != cudaSuccess) {
    cudaGetLastError();
    if (a) cudaFreeHost(a);
    oak_set_error("host_stage: cannot allocate %zu bytes of page-locked staging memory", 2 * doubles * sizeof(double));
    return OAK_ERR_NOMEM;
  }
  s.hs.in = (double *)a; s.hs.out = (double *)b; s.hs.cap = doubles;
  return 0;
}

bool is_pageable(const void *p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeUnregistered;
}

// A pageable host array (rows x ncols, leading dimension ld) to / from a dense device array (leading dimension dld)
// through the pinned staging buffers of slots 0 and 1: pieces of <= 64 MB, filled / drained by host threads while the
// previous piece is on the bus.  Synchronous: complete on return.  Not to be used while chunks of an analysis are in
// flight on those slots (the callers run it before / after the analysis).
int staged_copy(oakb200_handle *h, bool to_device, double *dev, size_t dld, double *host, size_t ld, size_t rows, int ncols,
                int nthreads);

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

int staged_copy(oakb200_handle *h, bool to_device, double *dev, size_t dld, double *host, size_t ld, size_t rows, int ncols,
                int nthreads) {
  if (rows == 0 || ncols == 0) return 0;
  const size_t piece = (size_t)8 << 20;   // doubles per piece (64 MB)
  const size_t rb = std::min(rows, piece);
  const int cb = (int)std::max<size_t>(1, std::min<size_t>((size_t)ncols, piece / rb));
  int rc, k = 0;
  struct Pending { size_t r0, nr; int c0, nc; bool on = false; } pend[2];
  auto finish = [&](int i) -> int {   // device-to-host: the piece that has landed in the slot's buffer goes to the caller
    Slot &s = h->slot[i];
    CUDA_TRY(cudaEventSynchronize(s.hs.ev_out));
    if (!to_device && pend[i].on)
      par_copy_cols(host + (size_t)pend[i].c0 * ld + pend[i].r0, ld, s.hs.out, pend[i].nr, pend[i].nr, pend[i].nc, nthreads);
    pend[i].on = false;
    return 0;
  };
  for (size_t r0 = 0; r0 < rows; r0 += rb)
    for (int c0 = 0; c0 < ncols; c0 += cb, k++) {
      const size_t nr = std::min(rb, rows - r0);
      const int nc = std::min(cb, ncols - c0);
      const int i = k & 1;
      Slot &s = h->slot[i];
      if (k >= 2 && (rc = finish(i))) return rc;
      if ((rc = ensure_stage(s, std::max(s.hs.cap, nr * (size_t)nc)))) return rc;
      if (to_device) {
        par_copy_cols(s.hs.in, nr, host + (size_t)c0 * ld + r0, ld, nr, nc, nthreads);
        CUDA_TRY(cudaMemcpy2DAsync(dev + (size_t)c0 * dld + r0, 8 * dld, s.hs.in, 8 * nr, 8 * nr, nc, cudaMemcpyHostToDevice, s.st));
      } else {
        CUDA_TRY(cudaMemcpy2DAsync(s.hs.out, 8 * nr, dev + (size_t)c0 * dld + r0, 8 * dld, 8 * nr, nc, cudaMemcpyDeviceToHost, s.st));
      }
      CUDA_TRY(cudaEventRecord(s.hs.ev_out, s.st));
      pend[i] = Pending{r0, nr, c0, nc, true};
    }
  for (int i = 0; i < 2; i++)
    if (pend[i].on && (rc = finish(i))) return rc;
  return 0;
}

int ensure_ws(oakb200_handle *h, Slot &s, int NP, int zb) {
  int rc;
  if ((rc = s.G.ensure(sizeof(double) * (size_t)zb * NP * NP))) return rc;
  if ((rc = s.T.ensure(sizeof(double) * (size_t)zb * NP * NP))) return rc;
  if ((rc = s.c.ensure(sizeof(double) * (size_t)zb * NP))) return rc;
  if ((rc = s.ampl.ensure(sizeof(double) * (size_t)zb * NP))) return rc;
  if (h->eig_kernel == 4 && NP <= 64 && (rc = s.tri.ensure(oak_eig_tridiag_ws_bytes(NP, zb)))) return rc;
  if (h->eig_kernel == 4 && NP <= 64 && h->tvec_split && (rc = s.Wv.ensure(sizeof(double) * (size_t)zb * NP * NP))) return rc;
  return 0;
}

// Zones per batch: a whole number of waves of the transform kernel (4 CTAs x 148 SMs at NP = 64) so that a
// batch does not end on a nearly empty wave, small enough that a call has >= ~8 batches per stream slot to
// overlap the tails, capped by the workspace budget (2 x 128 MB per slot at NP = 64).
int batch_size(const oakb200_handle *h, int NP, int nzones_call) {
  if (h->zones_per_batch > 0) return h->zones_per_batch;
  const size_t per_zone = sizeof(double) * 2 * (size_t)NP * NP;
  // the tridiagonal route has a latency-bound kernel (k_tql: one thread per zone, ~1.6 ms whatever the
  // batch size up to ~28 k zones), so its batches are larger: 1 GB of workspace instead of 256 MB
  const bool tri = h->eig_kernel == 4 && NP <= 64;
  // (round 2, with k_tql at ~0.7 ms per launch: 16.5 k -> 24.3 k zones per batch, C3 234.0 -> 230.1 ms per step)
  const int cap = (int)((size_t)(tri ? 1536 : 256) * 1024 * 1024 / per_zone);
  const int wave = NP <= 64 ? 592 : 148;
  int zb = nzones_call / (NSLOT * (tri ? 4 : 8));
  // ... and never small when the call is (multi-GPU phases of ~30 k zones): 12 waves per batch, or one batch
  // per stream slot, measured 56.3 -> 47.4 ms per step on 125 k zones per rank (profiles/r1_notes.md)
  if (tri) zb = std::max(zb, std::min(7104, (nzones_call + NSLOT - 1) / NSLOT));
  zb = std::max(wave, (zb / wave) * wave);
  zb = std::min(zb, std::max(wave, (cap / wave) * wave));
  return zb;
}

// Packs the observation-space arrays (device pointers) into sorted rows. Enqueued on `st`.
int pack_obs(oakb200_handle *h, cudaStream_t st, int N, int NP, const double *HSf, int64_t ldH, const double *yo,
             const double *Hxf, const double *R, const double *d01) {
  int rc;
  const int m = h->m;
  if ((rc = h->d_rows.ensure(sizeof(double) * (size_t)std::max(m, 1) * NP))) return rc;
  if ((rc = h->d_delta.ensure(sizeof(double) * (size_t)std::max(m, 1)))) return rc;
  if ((rc = h->d_scoef.ensure(sizeof(double) * (size_t)std::max(m, 1)))) return rc;
  return oak_launch_pack_obs(st, m, N, NP, h->og.perm, HSf, ldH, yo, Hxf, R, d01, h->d_rows.as<double>(),
                             h->d_delta.as<double>(), h->d_scoef.as<double>());
}

// amplitudes(:, zone) of the zones of a batch (localise_obs = .false.; zero for the zones that were skipped)
__global__ void k_copy_ampl(int N, int NP, int nz, const int32_t *__restrict__ mloc, const double *__restrict__ ampl,
                            double *__restrict__ out) {
  const int z = blockIdx.x, j = threadIdx.x;
  if (z < nz && j < N) out[(int64_t)z * N + j] = mloc[z] != 0 ? ampl[(int64_t)z * NP + j] : 0.;
}

// Fused gather, kernel flavour (option "push_kernel"): a few small CTAs on a high-priority side stream read the rows a
// batch has just finished once and store them into the result arrays of all ranks (posted 16-byte stores over
// NVLink).  The copy engines move the same 13.4 GB per rank and step at ~335 GB/s (8 GPUs: 47.8 ms against 34.1 ms
// without any push); stores issued by SMs are not limited by the engines, and a dedicated low-footprint kernel does not
// hold the resources of the apply kernel while they drain (the problem of peer_mode 0).
__global__ void __launch_bounds__(256) k_push(PeerOut P, const double *__restrict__ Sa, int64_t ldSa,
                                              const double *__restrict__ xa, int64_t r0, int64_t g0, int64_t L, int N,
                                              int vec) {
  const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
  if (vec) {
    const int64_t LV = L >> 1, total = LV * (N + 1);
    for (int64_t i = tid; i < total; i += nth) {
      const int64_t k = i / LV, v = i - k * LV;
      if (k < N) {
        const double2 x = *reinterpret_cast<const double2 *>(Sa + r0 + 2 * v + ldSa * k);
#pragma unroll 4
        for (int d = 0; d < P.n; d++) *reinterpret_cast<double2 *>(P.Sa[d] + g0 + 2 * v + P.ld * k) = x;
      } else {
        const double2 x = *reinterpret_cast<const double2 *>(xa + r0 + 2 * v);
        for (int d = 0; d < P.n; d++) *reinterpret_cast<double2 *>(P.xa[d] + g0 + 2 * v) = x;
      }
    }
  } else {
    const int64_t total = L * (N + 1);
    for (int64_t i = tid; i < total; i += nth) {
      const int64_t k = i / L, v = i - k * L;
      if (k < N) {
        const double x = Sa[r0 + v + ldSa * k];
        for (int d = 0; d < P.n; d++) P.Sa[d][g0 + v + P.ld * k] = x;
      } else {
        const double x = xa[r0 + v];
        for (int d = 0; d < P.n; d++) P.xa[d][g0 + v] = x;
      }
    }
  }
}

// Fused gather, NVSwitch-multicast flavour (oakb200_set_multicast_output): the result arrays of all ranks are bound to one
// multicast object (cuMulticastCreate / cuMulticastBindMem, set up by the caller: oak_b200.dist.MulticastResult); a
// 16-byte multimem.st to the multicast address is replicated BY THE SWITCH into every rank's array, so a rank sends its
// slab once (1.9 GB per C3 step at 8 GPUs) instead of once per peer (13.4 GB).
#ifndef OAK_CUEMU
__device__ __forceinline__ void mc_store16(double *p, double2 v) {
  const float4 f = *reinterpret_cast<const float4 *>(&v);
  asm volatile("multimem.st.weak.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(f.x), "f"(f.y), "f"(f.z), "f"(f.w) : "memory");
}
__device__ __forceinline__ void mc_store8(double *p, double v) {
  asm volatile("multimem.st.weak.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
#else
__device__ __forceinline__ void mc_store16(double *p, double2 v) { *reinterpret_cast<double2 *>(p) = v; }
__device__ __forceinline__ void mc_store8(double *p, double v) { *p = v; }
#endif
__global__ void __launch_bounds__(256) k_push_mc(double *__restrict__ Smc, double *__restrict__ xmc, int64_t ld,
                                                 const double *__restrict__ Sa, int64_t ldSa,
                                                 const double *__restrict__ xa, int64_t r0, int64_t g0, int64_t L, int N,
                                                 int vec) {
  const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
  if (vec) {
    const int64_t LV = L >> 1, total = LV * (N + 1);
    for (int64_t i = tid; i < total; i += nth) {
      const int64_t k = i / LV, v = i - k * LV;
      if (k < N) mc_store16(Smc + g0 + 2 * v + ld * k, *reinterpret_cast<const double2 *>(Sa + r0 + 2 * v + ldSa * k));
      else mc_store16(xmc + g0 + 2 * v, *reinterpret_cast<const double2 *>(xa + r0 + 2 * v));
    }
  } else {
    const int64_t total = L * (N + 1);
    for (int64_t i = tid; i < total; i += nth) {
      const int64_t k = i / L, v = i - k * L;
      if (k < N) mc_store8(Smc + g0 + v + ld * k, Sa[r0 + v + ldSa * k]);
      else mc_store8(xmc + g0 + v, xa[r0 + v]);
    }
  }
}

struct ProfAcc { double gram = 0, eig = 0, apply = 0, tridiag = 0, tql = 0, tvec = 0; };

// Runs zones [z0, z1) on slot s. The state buffers hold rows starting at global row `rowbase`.
int run_zones(oakb200_handle *h, Slot &s, int N, int NP, int z0, int z1, int64_t rowbase, const double *xf,
              const double *Sf, int64_t ldS, double *xa, double *Sa, int64_t ldSa, int64_t *launches,
              ProfAcc *prof, bool use_peers = false, double *ampl_out = nullptr /* device, [N x nzones] */,
              int64_t rows_in_buffers = 0 /* rows held in Sf / Sa (extent of the TMA tensor maps; 0: no TMA) */) {
  int zb = batch_size(h, NP, h->nzones);
  if (h->zones_per_batch <= 0 && z1 - z0 > zb) {
    // a chunk of the host path holds a little more than one batch: split it evenly instead of one full batch
    // plus a tiny one (the tiny one would still pay the fixed latency of k_tql)
    const int nb = (z1 - z0 + zb - 1) / zb;
    zb = std::min(zb, (z1 - z0 + nb - 1) / nb);
  }
  int rc;
  if ((rc = ensure_ws(h, s, NP, std::min(zb, z1 - z0)))) return rc;
  const ZoneGeom zg = h->geom();
  ObsRows orows{h->d_rows.as<double>(), h->d_delta.as<double>(), h->d_scoef.as<double>()};
  DevCounters *ctr = h->d_ctr.as<DevCounters>();
  int32_t *mloc = h->d_mloc.as<int32_t>();
  for (int b0 = z0; b0 < z1; b0 += zb) {
    const int nz = std::min(zb, z1 - b0);
    if (prof) CUDA_TRY(cudaEventRecord(s.ev[0], s.st));
    if (h->gram_kernel > 0 && NP == 64)
      rc = oak_launch_gram_mma(s.st, h->gram_kernel, NP, zg, h->og, orows, b0, nz, s.G.as<double>(), s.c.as<double>(), mloc, ctr);
    else
      rc = oak_launch_gram(s.st, NP, zg, h->og, orows, b0, nz, s.G.as<double>(), s.c.as<double>(), mloc, ctr);
    if (rc) return rc;
    if (prof) CUDA_TRY(cudaEventRecord(s.ev[1], s.st));
    const int32_t *only_flagged = nullptr;
    if (h->eig_kernel == 4 && NP <= 64) {
      // tridiagonal route; the zones it flags (close eigenvalue groups it could not orthogonalise, ...) are
      // recomputed by the Jacobi kernel, which skips the zones whose flag is 0
      int32_t *flags = nullptr;
      // fused apply: not with the store flavour of the fused all-gather (k_apply is the kernel that stores to the peers)
      // ... and only while the factored form is the cheaper one (4 nr N^2 against 2 N^3 + 2 nr N^2: zones of at most N rows)
      const bool fuse = h->fuse_apply && !(use_peers && h->peer_mode == 0) && h->max_zone_rows <= NP;
      const FusedApplyArgs fa{zg.zstart + b0, rowbase, xf, Sf, xa, Sa, ldS, ldSa};
      const TqlSide tside{s.qst, s.qev[0], s.qev[1], s.qev[2], s.qev[3]};
      if ((rc = oak_launch_eig_tridiag(s.st, N, NP, nz, mloc + b0, s.G.as<double>(), s.c.as<double>(),
                                       s.T.as<double>(), s.ampl.as<double>(), s.tri.p, &flags, ctr,
                                       prof ? &s.ev[8] : nullptr, h->tri_orthtol, h->tri_maxgroup,
                                       fuse ? &fa : nullptr, h->tvec_split ? s.Wv.as<double>() : nullptr,
                                       (h->tql_side && !prof) ? &tside : nullptr))) return rc;
      if (fuse) only_flagged = flags;
      if (prof) CUDA_TRY(cudaEventRecord(s.ev[10], s.st));
      if ((rc = oak_launch_eig(s.st, 0, N, NP, 0, nz, flags, s.G.as<double>(), s.c.as<double>(),
                               s.T.as<double>(), s.ampl.as<double>(), h->tol, h->max_sweeps, ctr))) return rc;
      *launches += 3 + (h->tvec_split ? 1 : 0);
    } else if ((rc = oak_launch_eig(s.st, h->eig_kernel == 4 ? 0 : h->eig_kernel, N, NP, b0, nz, mloc,
                                    s.G.as<double>(), s.c.as<double>(), s.T.as<double>(), s.ampl.as<double>(),
                                    h->tol, h->max_sweeps, ctr))) return rc;
    if (ampl_out && !h->localise_obs) {
      k_copy_ampl<<<nz, NP, 0, s.st>>>(N, NP, nz, mloc + b0, s.ampl.as<double>(), ampl_out + (int64_t)b0 * N);
      CUDA_TRY(cudaGetLastError());
      *launches += 1;
    }
    if (prof) CUDA_TRY(cudaEventRecord(s.ev[2], s.st));
    PeerOut none{};
    // fused all-gather, copy-engine flavour: as soon as (a piece of) the batch is applied its rows go to every
    // peer's array (strided 2-D peer copies over NVLink, no SM involved), overlapping the kernels of the next
    // batches.  What stays exposed at the end of a call is the push of the last piece of every stream slot, so
    // "push_pieces" > 1 launches the apply of a batch in pieces and pushes each one behind it.
    const bool mcast = h->mc_Sa != nullptr && !prof;
    const bool push = (use_peers && h->peer_mode == 1) || mcast;
    const int npiece = push ? std::max(1, std::min(h->push_pieces, nz)) : 1;
    for (int pc = 0; pc < npiece; pc++) {
      const int o0 = (int)((int64_t)nz * pc / npiece), o1 = (int)((int64_t)nz * (pc + 1) / npiece);
      if (h->apply_kernel == 1 && NP == 64 && !(use_peers && h->peer_mode == 0))
        rc = oak_launch_apply_mma(s.st, N, NP, zg, b0 + o0, o1 - o0, rowbase, mloc, s.T.as<double>() + (size_t)o0 * NP * NP,
                                  s.ampl.as<double>() + (size_t)o0 * NP, xf, Sf, ldS, xa, Sa, ldSa,
                                  only_flagged ? only_flagged + o0 : nullptr, false);
      else
        rc = oak_launch_apply(s.st, N, NP, zg, b0 + o0, o1 - o0, rowbase, mloc, s.T.as<double>() + (size_t)o0 * NP * NP,
                              s.ampl.as<double>() + (size_t)o0 * NP, xf, Sf, ldS, xa, Sa, ldSa,
                              (use_peers && h->peer_mode == 0) ? h->peers : none,
                              only_flagged ? only_flagged + o0 : nullptr, false,
                              (h->apply_tma && h->min_zone_rows == h->max_zone_rows) ? h->max_zone_rows : 0, rows_in_buffers,
                              h->ens.on ? &h->ens : nullptr);
      if (rc) return rc;
      if (pc > 0) *launches += 1;
      if (!push) continue;
      const PeerOut &P = h->peers;
      const int64_t r0 = h->h_zstart[b0 + o0], r1 = h->h_zstart[b0 + o1];  // rows of the piece (zone prefix sums of this call)
      if (r1 > r0) {
        // one stream per destination: the copies to different peers run on different copy engines / links
        CUDA_TRY(cudaEventRecord(s.ev[11], s.st));
        if (mcast) {
          CUDA_TRY(cudaStreamWaitEvent(s.cst, s.ev[11], 0));
          const int64_t L = r1 - r0, lr0 = r0 - rowbase, g0 = h->mc_row0 + r0;
          const bool vec = ((L | lr0 | g0 | ldSa | h->mc_ld) & 1) == 0 && ((uintptr_t)Sa & 15) == 0 && ((uintptr_t)xa & 15) == 0 &&
                           ((uintptr_t)h->mc_Sa & 15) == 0 && ((uintptr_t)h->mc_xa & 15) == 0;
          k_push_mc<<<h->push_ctas, 256, 0, s.cst>>>(h->mc_Sa, h->mc_xa, h->mc_ld, Sa, ldSa, xa, lr0, g0, L, N, vec ? 1 : 0);
          CUDA_TRY(cudaGetLastError());
          *launches += 1;
          s.cst_used = true;
        } else if (h->push_kernel) {
          CUDA_TRY(cudaStreamWaitEvent(s.cst, s.ev[11], 0));
          const int64_t L = r1 - r0, lr0 = r0 - rowbase, g0 = P.row0 + r0;
          bool vec = ((L | lr0 | g0 | ldSa | P.ld) & 1) == 0 && ((uintptr_t)Sa & 15) == 0 && ((uintptr_t)xa & 15) == 0;
          for (int d = 0; d < P.n && vec; d++) vec = ((uintptr_t)P.Sa[d] & 15) == 0 && ((uintptr_t)P.xa[d] & 15) == 0;
          k_push<<<h->push_ctas, 256, 0, s.cst>>>(P, Sa, ldSa, xa, lr0, g0, L, N, vec ? 1 : 0);
          CUDA_TRY(cudaGetLastError());
          *launches += 1;
          s.cst_used = true;
        } else
        for (int d = 0; d < P.n; d++) {
          cudaStream_t cs = h->pstream[d];
          CUDA_TRY(cudaStreamWaitEvent(cs, s.ev[11], 0));
          CUDA_TRY(cudaMemcpy2DAsync(P.Sa[d] + P.row0 + r0, sizeof(double) * (size_t)P.ld, Sa + (r0 - rowbase),
                                     sizeof(double) * (size_t)ldSa, sizeof(double) * (size_t)(r1 - r0), (size_t)N,
                                     cudaMemcpyDeviceToDevice, cs));
          CUDA_TRY(cudaMemcpyAsync(P.xa[d] + P.row0 + r0, xa + (r0 - rowbase), sizeof(double) * (size_t)(r1 - r0),
                                   cudaMemcpyDeviceToDevice, cs));
        }
      }
    }
    *launches += 3;
    if (prof) {
      CUDA_TRY(cudaEventRecord(s.ev[3], s.st));
      CUDA_TRY(cudaEventSynchronize(s.ev[3]));
      float a, b, c;
      CUDA_TRY(cudaEventElapsedTime(&a, s.ev[0], s.ev[1]));
      CUDA_TRY(cudaEventElapsedTime(&b, s.ev[1], s.ev[2]));
      CUDA_TRY(cudaEventElapsedTime(&c, s.ev[2], s.ev[3]));
      prof->gram += a; prof->eig += b; prof->apply += c;
      if (h->eig_kernel == 4 && NP <= 64) {
        CUDA_TRY(cudaEventElapsedTime(&a, s.ev[1], s.ev[8]));
        CUDA_TRY(cudaEventElapsedTime(&b, s.ev[8], s.ev[9]));
        CUDA_TRY(cudaEventElapsedTime(&c, s.ev[9], s.ev[10]));
        prof->tridiag += a; prof->tql += b; prof->tvec += c;
      }
    }
  }
  return 0;
}

int check_ready(oakb200_handle *h, int64_t n, int N, int m) {
  if (!h) { oak_set_error("null handle"); return OAK_ERR_ARG; }
  if (!h->zones_set) { oak_set_error("oakb200_set_zones has not been called"); return OAK_ERR_STATE; }
  if (!h->obs_set) { oak_set_error("oakb200_set_observations has not been called"); return OAK_ERR_STATE; }
  if (n != h->nrows) { oak_set_error("n = %lld does not match sum(zoneSize) = %lld", (long long)n, (long long)h->nrows); return OAK_ERR_ARG; }
  if (m != h->m) { oak_set_error("m = %d does not match the %d observations set", m, h->m); return OAK_ERR_ARG; }
  if (N < 2) { oak_set_error("ensemble size N = %d < 2", N); return OAK_ERR_ARG; }
  if (padded(N) < 0) { oak_set_error("ensemble size N = %d > 128 is not supported", N); return OAK_ERR_UNSUPPORTED; }
  return 0;
}

int begin_call(oakb200_handle *h, oakb200_stats *stats) {
  if (stats) memset(stats, 0, sizeof *stats);
  CUDA_TRY(cudaMemsetAsync(h->d_ctr.p, 0, sizeof(DevCounters), h->slot[0].st));
  CUDA_TRY(cudaMemsetAsync(h->d_mloc.p, 0, sizeof(int32_t) * std::max(h->nzones, 1), h->slot[0].st));
  return 0;
}

int end_call(oakb200_handle *h, oakb200_stats *stats, int64_t launches, const ProfAcc &prof, float ms_total,
             float ms_pack) {
  DevCounters ctr;
  CUDA_TRY(cudaMemcpy(&ctr, h->d_ctr.p, sizeof ctr, cudaMemcpyDeviceToHost));
  if (stats) {
    stats->zones_total = h->nzones;
    stats->zones_skipped = (int64_t)ctr.skipped;
    stats->obs_relevant_sum = (int64_t)ctr.relevant;
    stats->obs_candidate_sum = (int64_t)ctr.candidates;
    stats->jacobi_sweeps_sum = (int64_t)ctr.sweeps;
    stats->ms_total = ms_total;
    stats->ms_pack = ms_pack;
    stats->ms_gram = prof.gram; stats->ms_eig = prof.eig; stats->ms_apply = prof.apply;
    stats->launches = launches;
    stats->zones_fallback = (int64_t)ctr.fallback;
    stats->ms_tridiag = prof.tridiag; stats->ms_tql = prof.tql; stats->ms_tvec = prof.tvec;
  }
  if (getenv("OAKB200_DEBUG"))
    fprintf(stderr, "[oak_b200] fallback %llu (ql %llu, residual %llu, group %llu, parallel %llu), gram-schmidt projections %llu, pivot-form vectors %llu, sweeps %llu\n",
            ctr.fallback, ctr.fb_reason[0], ctr.fb_reason[1], ctr.fb_reason[2], ctr.fb_reason[3], ctr.gs_pairs, ctr.tw_fallback, ctr.sweeps);
  if (ctr.nan_flag) { oak_set_error("NaN in the analysis amplitudes (rrsqrt.F90:145-149)"); return OAK_ERR_NAN; }
  if (ctr.not_converged) { oak_set_error("Jacobi eigensolve did not converge in %d sweeps for %d zones", h->max_sweeps, ctr.not_converged); return OAK_ERR_NAN; }
  return 0;
}

}  // namespace

extern "C" OAKB200_API int oakb200_destroy(oakb200_handle *h);

extern "C" OAKB200_API int oakb200_create(int device, oakb200_handle **out) {
  if (!out) { oak_set_error("null output pointer"); return OAK_ERR_ARG; }
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    oak_set_error("no CUDA device available (%s); oak_b200 has no CPU fallback", cudaGetErrorString(e));
    return OAK_ERR_CUDA;
  }
  if (device < 0 || device >= ndev) { oak_set_error("device %d out of range (%d devices)", device, ndev); return OAK_ERR_ARG; }
  DeviceGuard guard(device);
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    oak_set_error("device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major, prop.minor);
    return OAK_ERR_UNSUPPORTED;
  }
  oakb200_handle *h = new oakb200_handle();
  h->device = device;
  // every resource is owned by the handle from the moment it exists: a failure below destroys what was created
  auto build = [&]() -> int {
    for (int i = 0; i < NSLOT; i++) {
      CUDA_TRY(cudaStreamCreateWithFlags(&h->slot[i].st, cudaStreamNonBlocking));
      { int lo = 0, hi = 0; CUDA_TRY(cudaDeviceGetStreamPriorityRange(&lo, &hi)); CUDA_TRY(cudaStreamCreateWithPriority(&h->slot[i].qst, cudaStreamNonBlocking, hi)); CUDA_TRY(cudaStreamCreateWithPriority(&h->slot[i].cst, cudaStreamNonBlocking, hi)); }
      for (auto &ev : h->slot[i].qev) CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
      for (auto &ev : h->slot[i].ev) CUDA_TRY(cudaEventCreate(&ev));
    }
    CUDA_TRY(cudaEventCreate(&h->ev_a));
    CUDA_TRY(cudaEventCreate(&h->ev_b));
    CUDA_TRY(cudaEventCreateWithFlags(&h->ev_user, cudaEventDisableTiming));
    return h->d_ctr.ensure(sizeof(DevCounters));
  };
  const int rc = build();
  if (rc) { oakb200_destroy(h); return rc; }
  *out = h;
  return 0;
}

extern "C" OAKB200_API int oakb200_destroy(oakb200_handle *h) {
  if (!h) return 0;
  DeviceGuard guard(h->device);
  cudaDeviceSynchronize();
  DevBuf *bufs[] = {&h->d_zx, &h->d_zy, &h->d_corr, &h->d_maxl, &h->d_zstart, &h->d_mloc, &h->d_bx, &h->d_by,
                    &h->d_key_in, &h->d_key_out, &h->d_val_in, &h->d_perm, &h->d_cell_start, &h->d_sx, &h->d_sy,
                    &h->d_tmp, &h->d_rows, &h->d_delta, &h->d_scoef, &h->d_HSf, &h->d_yo, &h->d_Hxf, &h->d_R,
                    &h->d_d01, &h->d_ampzero, &h->d_HE, &h->d_Hi, &h->d_Hj, &h->d_Hs, &h->d_Hshift, &h->d_order,
                    &h->d_rowstart, &h->d_xf, &h->d_xa, &h->d_maxc, &h->d_E, &h->d_ctr, &h->d_anam, &h->d_tet, &h->d_gws, &h->d_gzstart, &h->d_rowvar, &h->d_vdesc, &h->d_vtab};
  for (DevBuf *b : bufs) b->release();
  for (int d = 0; d < OAKB200_MAX_PEERS; d++) {
    if (h->pstream[d]) cudaStreamDestroy(h->pstream[d]);
    if (h->pev[d]) cudaEventDestroy(h->pev[d]);
  }
  for (int i = 0; i < NSLOT; i++) {
    Slot &s = h->slot[i];
    s.Wv.release(); s.G.release(); s.T.release(); s.c.release(); s.ampl.release(); s.tri.release(); s.S.release(); s.xf.release(); s.xa.release();
    if (s.hs.in) cudaFreeHost(s.hs.in);
    if (s.hs.out) cudaFreeHost(s.hs.out);
    if (s.hs.ev_out) cudaEventDestroy(s.hs.ev_out);
    for (auto &ev : s.ev) if (ev) cudaEventDestroy(ev);
    if (s.st) cudaStreamDestroy(s.st);
    if (s.cst) cudaStreamDestroy(s.cst);
    if (s.qst) cudaStreamDestroy(s.qst);
    for (auto &ev : s.qev) if (ev) cudaEventDestroy(ev);
  }
  if (h->ev_a) cudaEventDestroy(h->ev_a);
  if (h->ev_b) cudaEventDestroy(h->ev_b);
  if (h->ev_user) cudaEventDestroy(h->ev_user);
  delete h;
  return 0;
}

extern "C" OAKB200_API int oakb200_set_anamorphosis_table(oakb200_handle *h, int32_t K, const double *table) {
  if (!h) { oak_set_error("null handle"); return OAK_ERR_ARG; }
  if (K == 0) { h->anam_K = 0; return 0; }
  if (K < 2 || !table) { oak_set_error("set_anamorphosis_table: K = %d (need >= 2 rows) or null table", K); return OAK_ERR_ARG; }
  for (int i = 0; i < 2 * K; i++)
    if (!(table[i] == table[i])) { oak_set_error("set_anamorphosis_table: NaN in the table"); return OAK_ERR_ARG; }
  int mono = 1;
  for (int c = 0; c < 2; c++)
    for (int i = 0; i + 1 < K; i++)
      if (!(table[c * K + i] < table[c * K + i + 1])) mono = 0;
  DeviceGuard guard(h->device);
  int rc = h->d_anam.ensure(sizeof(double) * 2 * (size_t)K);
  if (rc) return rc;
  CUDA_TRY(cudaMemcpy(h->d_anam.p, table, sizeof(double) * 2 * (size_t)K, cudaMemcpyHostToDevice));
  h->anam_K = K;
  h->anam_monotone = mono;
  return 0;
}

extern "C" OAKB200_API int oakb200_set_anamorphosis_vars(oakb200_handle *h, int32_t nvar, const int32_t *vtype,
                                                         const int32_t *vK, const double *tables, int64_t n,
                                                         const int32_t *rowvar) {
  if (!h) { oak_set_error("null handle"); return OAK_ERR_ARG; }
  if (nvar == 0) { h->anam_nvar = 0; h->anam_rows = 0; return 0; }
  if (nvar < 0 || !vtype || !vK || n < 0 || (n > 0 && !rowvar)) { oak_set_error("set_anamorphosis_vars: invalid argument"); return OAK_ERR_ARG; }
  std::vector<int32_t> desc(4 * (size_t)nvar);
  int64_t off = 0;
  for (int v = 0; v < nvar; v++) {
    if (vtype[v] < 1 || vtype[v] > 3) { oak_set_error("set_anamorphosis_vars: variable %d has unknown type %d", v + 1, vtype[v]); return OAK_ERR_ARG; }
    const int K = vtype[v] == 3 ? vK[v] : 0;
    if (vtype[v] == 3 && (K < 2 || !tables)) { oak_set_error("set_anamorphosis_vars: variable %d is tabulated but has no table", v + 1); return OAK_ERR_ARG; }
    int mono = 1;
    for (int c = 0; c < 2 && K > 0; c++)
      for (int i = 0; i + 1 < K; i++) {
        const double a = tables[off + (int64_t)c * K + i], b = tables[off + (int64_t)c * K + i + 1];
        if (!(a == a) || !(b == b)) { oak_set_error("set_anamorphosis_vars: NaN in the table of variable %d", v + 1); return OAK_ERR_ARG; }
        if (!(a < b)) mono = 0;
      }
    desc[4 * v] = vtype[v]; desc[4 * v + 1] = K; desc[4 * v + 2] = (int32_t)off; desc[4 * v + 3] = mono;
    off += 2 * (int64_t)K;
  }
  std::vector<int32_t> rv((size_t)n);
  for (int64_t i = 0; i < n; i++) {   // 1-based variable numbers, as ind2submv returns them (assimilation.F90:1822-1858)
    if (rowvar[i] < 1 || rowvar[i] > nvar) { oak_set_error("set_anamorphosis_vars: row %lld has variable %d (1..%d)", (long long)i + 1, rowvar[i], nvar); return OAK_ERR_ARG; }
    rv[(size_t)i] = rowvar[i] - 1;
  }
  DeviceGuard guard(h->device);
  int rc;
  if ((rc = h->d_vdesc.ensure(sizeof(int32_t) * desc.size())) || (rc = h->d_vtab.ensure(sizeof(double) * (size_t)std::max<int64_t>(off, 1))) ||
      (rc = h->d_rowvar.ensure(sizeof(int32_t) * (size_t)std::max<int64_t>(n, 1)))) return rc;
  CUDA_TRY(cudaMemcpy(h->d_vdesc.p, desc.data(), sizeof(int32_t) * desc.size(), cudaMemcpyHostToDevice));
  if (off > 0) CUDA_TRY(cudaMemcpy(h->d_vtab.p, tables, sizeof(double) * (size_t)off, cudaMemcpyHostToDevice));
  if (n > 0) CUDA_TRY(cudaMemcpy(h->d_rowvar.p, rv.data(), sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice));
  h->anam_nvar = nvar; h->anam_rows = n;
  return 0;
}

extern "C" OAKB200_API int oakb200_set_peer_outputs(oakb200_handle *h, int32_t npeer, double *const *Sa_peer,
                                                    double *const *xa_peer, int64_t ld_peer, int64_t row0) {
  if (!h) { oak_set_error("null handle"); return OAK_ERR_ARG; }
  if (npeer < 0 || npeer > OAKB200_MAX_PEERS || (npeer > 0 && (!Sa_peer || !xa_peer))) {
    oak_set_error("set_peer_outputs: npeer = %d (0..%d) or null pointer arrays", npeer, OAKB200_MAX_PEERS);
    return OAK_ERR_ARG;
  }
  h->peers = PeerOut{};
  for (int d = 0; d < npeer; d++) {
    if (!Sa_peer[d] || !xa_peer[d]) { oak_set_error("set_peer_outputs: null destination %d", d); h->peers = PeerOut{}; return OAK_ERR_ARG; }
    h->peers.Sa[d] = Sa_peer[d];
    h->peers.xa[d] = xa_peer[d];
  }
  h->peers.n = npeer; h->peers.ld = ld_peer; h->peers.row0 = row0;
  DeviceGuard guard(h->device);
  for (int d = 0; d < npeer; d++) {
    if (!h->pstream[d]) {
      CUDA_TRY(cudaStreamCreateWithFlags(&h->pstream[d], cudaStreamNonBlocking));
      CUDA_TRY(cudaEventCreateWithFlags(&h->pev[d], cudaEventDisableTiming));
    }
  }
  return 0;
}

extern "C" OAKB200_API int oakb200_set_multicast_output(oakb200_handle *h, double *Sa_mc, double *xa_mc, int64_t ld, int64_t row0) {
  if (!h) { oak_set_error("null handle"); return OAK_ERR_ARG; }
  if ((Sa_mc == nullptr) != (xa_mc == nullptr)) { oak_set_error("set_multicast_output: both addresses or none"); return OAK_ERR_ARG; }
  h->mc_Sa = Sa_mc; h->mc_xa = xa_mc; h->mc_ld = ld; h->mc_row0 = row0;
  return 0;
}

extern "C" OAKB200_API int oakb200_ipc_alloc(oakb200_handle *h, int64_t bytes, void **ptr, unsigned char handle[64]) {
  if (!h || !ptr || !handle || bytes <= 0) { oak_set_error("ipc_alloc: bad arguments"); return OAK_ERR_ARG; }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
  DeviceGuard guard(h->device);
  void *p = nullptr;
  if (cudaMalloc(&p, (size_t)bytes) != cudaSuccess) {
    cudaGetLastError();
    oak_set_error("ipc_alloc: cudaMalloc of %lld bytes failed", (long long)bytes);
    return OAK_ERR_NOMEM;
  }
  cudaIpcMemHandle_t hd;
  cudaError_t e = cudaIpcGetMemHandle(&hd, p);
  if (e != cudaSuccess) { cudaFree(p); oak_set_error("cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e)); return OAK_ERR_CUDA; }
  memcpy(handle, &hd, 64);
  *ptr = p;
  return 0;
}

extern "C" OAKB200_API int oakb200_ipc_open(oakb200_handle *h, const unsigned char handle[64], void **ptr) {
  if (!h || !ptr || !handle) { oak_set_error("ipc_open: bad arguments"); return OAK_ERR_ARG; }
  DeviceGuard guard(h->device);
  cudaIpcMemHandle_t hd;
  memcpy(&hd, handle, 64);
  void *p = nullptr;
  cudaError_t e = cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) { cudaGetLastError(); oak_set_error("cudaIpcOpenMemHandle failed: %s", cudaGetErrorString(e)); return OAK_ERR_CUDA; }
  *ptr = p;
  return 0;
}

extern "C" OAKB200_API int oakb200_ipc_close(oakb200_handle *h, void *ptr) {
  if (!h) return 0;
  DeviceGuard guard(h->device);
  if (ptr) CUDA_TRY(cudaIpcCloseMemHandle(ptr));
  return 0;
}

extern "C" OAKB200_API int oakb200_ipc_free(oakb200_handle *h, void *ptr) {
  if (!h) return 0;
  DeviceGuard guard(h->device);
  if (ptr) CUDA_TRY(cudaFree(ptr));
  return 0;
}

// Page-locked host memory for the caller's big arrays (Sf / Sa, HSf): the host-buffer entry points then overlap their
// chunked copies with the kernels (2.68 M columns/s on C3 against 0.40 M from pageable memory).  A Fortran caller maps
// the pointer with c_f_pointer.
extern "C" OAKB200_API int oakb200_host_alloc(int64_t bytes, void **ptr) {
  if (!ptr || bytes <= 0) { oak_set_error("host_alloc: bad arguments"); return OAK_ERR_ARG; }
  void *p = nullptr;
  if (cudaHostAlloc(&p, (size_t)bytes, cudaHostAllocPortable) != cudaSuccess) {
    cudaGetLastError();
    oak_set_error("host_alloc: cudaHostAlloc of %lld bytes failed", (long long)bytes);
    return OAK_ERR_NOMEM;
  }
  *ptr = p;
  return 0;
}

extern "C" OAKB200_API int oakb200_host_free(void *ptr) {
  if (ptr) CUDA_TRY(cudaFreeHost(ptr));
  return 0;
}

extern "C" OAKB200_API int oakb200_set_option(oakb200_handle *h, const char *key, double value) {
  if (!h || !key) { oak_set_error("null argument"); return OAK_ERR_ARG; }
  const std::string k(key);
  if (k == "eig_kernel") {
    if (value != 4. && value != 0. && value != 1.) { oak_set_error("eig_kernel = %g (4 tridiagonal route, 0 block Jacobi, 1 shared-memory cross-check)", value); return OAK_ERR_ARG; }
    h->eig_kernel = (int)value;
  }
  else if (k == "fuse_apply") h->fuse_apply = value != 0.;
  else if (k == "tvec_split") h->tvec_split = value != 0.;
  else if (k == "tql_side") h->tql_side = value != 0.;
  else if (k == "localise_obs") h->localise_obs = value != 0.;
  else if (k == "host_register") h->host_register = value != 0.;
  else if (k == "host_stage") h->host_stage = value < 0. ? -1 : (value != 0.);
  else if (k == "stage_threads") h->stage_threads = std::max(0, (int)value);
  else if (k == "apply_tma") h->apply_tma = value != 0.;
  else if (k == "ens_fuse") h->ens_fuse = value < 0. ? -1 : (value != 0.);
  else if (k == "push_kernel") h->push_kernel = value != 0.;
  else if (k == "push_ctas") h->push_ctas = std::max(1, (int)value);
  else if (k == "apply_kernel") {
    if (value != 0. && value != 1.) { oak_set_error("apply_kernel = %g (0 register tiles, 1 tensor-core tiles)", value); return OAK_ERR_ARG; }
    h->apply_kernel = (int)value;
  }
  else if (k == "scheme") {
    if (value != 0. && value != 1.) { oak_set_error("scheme = %g (0 global, 1 local)", value); return OAK_ERR_ARG; }
    h->scheme = (int)value;
  }
  else if (k == "gram_kernel") {
    if (!(value >= 0. && value <= 4.) || value != (int)value) { oak_set_error("gram_kernel = %g (expected 0 .. 4)", value); return OAK_ERR_ARG; }
    h->gram_kernel = (int)value;
  }
  else if (k == "tri_orthtol") h->tri_orthtol = value;
  else if (k == "tri_maxgroup") h->tri_maxgroup = (int)value;
  else if (k == "peer_mode") h->peer_mode = (int)value;
  else if (k == "push_pieces") h->push_pieces = std::max(1, (int)value);
  else if (k == "zones_per_batch") h->zones_per_batch = (int)value;
  else if (k == "taper") h->taper = std::max(0, (int)value);
  else if (k == "jacobi_tol") h->tol = value;
  else if (k == "max_sweeps") h->max_sweeps = (int)value;
  else if (k == "fixed_sweeps") { h->max_sweeps = (int)value; h->tol = -1.; }  // timing experiments: no convergence test
  else if (k == "profile") h->profile = (int)value;
  else if (k == "chunk_mb") h->chunk_mb = value;
  else if (k == "pad_to") h->pad_to = (int)value;
  else if (k == "async") h->async = (int)value;
  else if (k == "order_after_caller") h->order_after_caller = (int)value;
  else if (k == "stream_priority") {
    // recreate the slot streams with a CUDA stream priority (0 = default, negative = more urgent): lets a
    // caller that enqueues several handles at once tell the block scheduler which one goes first
    DeviceGuard guard(h->device);
    int lo = 0, hi = 0;
    CUDA_TRY(cudaDeviceGetStreamPriorityRange(&lo, &hi));  // lo = least urgent (0), hi = most urgent (negative)
    const int pr = std::max(hi, std::min(lo, (int)value));
    CUDA_TRY(cudaDeviceSynchronize());
    for (int i = 0; i < NSLOT; i++) {
      if (h->slot[i].st) CUDA_TRY(cudaStreamDestroy(h->slot[i].st));
      CUDA_TRY(cudaStreamCreateWithPriority(&h->slot[i].st, cudaStreamNonBlocking, pr));
    }
    h->stream_priority = pr;
  }
  else { oak_set_error("unknown option '%s'", key); return OAK_ERR_ARG; }
  return 0;
}

extern "C" OAKB200_API int oakb200_partition_zones(int32_t nzones, int32_t nranks, int32_t *first) {
  if (nzones < 0 || nranks < 1 || !first) { oak_set_error("bad partition arguments"); return OAK_ERR_ARG; }
  // startZIndex = nzones*cumulSpeed(p)/total + 1 (1-based), unit speeds   parall.F90:176-177
  for (int p = 0; p <= nranks; p++) first[p] = (int32_t)(((int64_t)nzones * p) / nranks);
  return 0;
}

extern "C" OAKB200_API int oakb200_set_zones(oakb200_handle *h, int32_t nzones, const int32_t *zoneSize, const double *zx,
                                 const double *zy, const double *zz, const double *zt, const double *corrLen,
                                 const double *maxLen, int32_t loctype, int32_t metrictype, int32_t weightfun) {
  if (!h || nzones < 0 || !zoneSize || !corrLen || !maxLen) { oak_set_error("set_zones: null/invalid argument"); return OAK_ERR_ARG; }
  if (loctype < 1 || loctype > 3) { oak_set_error("set_zones: loctype %d (expected 1,2,3)", loctype); return OAK_ERR_ARG; }
  if (metrictype < 0 || metrictype > 2) { oak_set_error("Unsupported metric: %d", metrictype); return OAK_ERR_ARG; }
  if (weightfun < 0 || weightfun > 2) { oak_set_error("set_zones: weightfun %d", weightfun); return OAK_ERR_ARG; }
  const double *prim = loctype == 1 ? zx : (loctype == 2 ? zz : zt);
  if (!prim && nzones > 0) { oak_set_error("set_zones: coordinate array needed by loctype %d is NULL", loctype); return OAK_ERR_ARG; }
  DeviceGuard guard(h->device);
  h->zones_set = false;
  h->obs_set = false;  // the cell size depends on the zones' radii
  h->nzones = nzones;
  h->loctype = loctype; h->metrictype = metrictype; h->weightfun = weightfun;
  h->h_zstart.assign((size_t)nzones + 1, 0);
  for (int z = 0; z < nzones; z++) {
    if (zoneSize[z] < 0) { oak_set_error("set_zones: negative zone size"); return OAK_ERR_ARG; }
    h->h_zstart[z + 1] = h->h_zstart[z] + zoneSize[z];
  }
  h->max_zone_rows = 0;
  h->min_zone_rows = nzones > 0 ? zoneSize[0] : 0;
  for (int z = 0; z < nzones; z++) { h->max_zone_rows = std::max(h->max_zone_rows, (int)zoneSize[z]); h->min_zone_rows = std::min(h->min_zone_rows, (int)zoneSize[z]); }
  h->nrows = h->h_zstart[nzones];
  h->rmax = 0.; h->any_unbounded = false; h->zlat_absmax = 0.;
  std::vector<double> zeros;
  for (int z = 0; z < nzones; z++) {
    double R = weightfun == OAKB200_WEIGHT_GAUSSIAN ? maxLen[z] : (weightfun == OAKB200_WEIGHT_GASPARI_COHN ? 2. * corrLen[z] : INFINITY);
    if (!(R < 1e300)) h->any_unbounded = true;
    else h->rmax = std::max(h->rmax, R);
    if (loctype == 1 && zy) h->zlat_absmax = std::max(h->zlat_absmax, std::fabs(zy[z]));
  }
  const size_t nb = sizeof(double) * (size_t)std::max(nzones, 1);
  int rc;
  if ((rc = h->d_zx.ensure(nb)) || (rc = h->d_zy.ensure(nb)) || (rc = h->d_corr.ensure(nb)) ||
      (rc = h->d_maxl.ensure(nb)) || (rc = h->d_zstart.ensure(sizeof(int64_t) * ((size_t)nzones + 1))) ||
      (rc = h->d_mloc.ensure(sizeof(int32_t) * (size_t)std::max(nzones, 1))))
    return rc;
  if (nzones > 0) {
    CUDA_TRY(cudaMemcpy(h->d_zx.p, prim, sizeof(double) * nzones, cudaMemcpyHostToDevice));
    if (loctype == 1 && zy) CUDA_TRY(cudaMemcpy(h->d_zy.p, zy, sizeof(double) * nzones, cudaMemcpyHostToDevice));
    else CUDA_TRY(cudaMemset(h->d_zy.p, 0, sizeof(double) * nzones));
    CUDA_TRY(cudaMemcpy(h->d_corr.p, corrLen, sizeof(double) * nzones, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(h->d_maxl.p, maxLen, sizeof(double) * nzones, cudaMemcpyHostToDevice));
  }
  CUDA_TRY(cudaMemcpy(h->d_zstart.p, h->h_zstart.data(), sizeof(int64_t) * ((size_t)nzones + 1), cudaMemcpyHostToDevice));
  h->zones_set = true;
  return 0;
}

extern "C" OAKB200_API int oakb200_set_observations(oakb200_handle *h, int32_t m, const double *obsx, const double *obsy,
                                        const double *obsz, const double *obst) {
  if (!h || m < 0) { oak_set_error("set_observations: invalid argument"); return OAK_ERR_ARG; }
  if (!h->zones_set) { oak_set_error("set_observations: call oakb200_set_zones first"); return OAK_ERR_STATE; }
  const double *bx = h->loctype == 1 ? obsx : (h->loctype == 2 ? obsz : obst);
  const double *by = h->loctype == 1 ? obsy : nullptr;
  if (m > 0 && !bx) { oak_set_error("set_observations: coordinate array needed by loctype %d is NULL", h->loctype); return OAK_ERR_ARG; }
  DeviceGuard guard(h->device);
  h->obs_set = false;
  h->m = m;
  const bool spherical = h->loctype == 1 && h->metrictype != OAKB200_METRIC_CARTESIAN;
  ObsGrid g{};
  g.m = m;
  g.wrap_x = (h->loctype == 1 && h->metrictype == OAKB200_METRIC_SPHERICAL) ? 1 : 0;
  // extents of the bucketing coordinates
  double xmin = 0, xmax = 0, ymin = 0, ymax = 0;
  bool first = true;
  for (int l = 0; l < m; l++) {
    double x = bx[l], y = by ? by[l] : 0.;
    if (g.wrap_x) x = oak_fold360(x);
    if (!(x == x) || !(y == y) || std::isinf(x) || std::isinf(y)) continue;
    if (first) { xmin = xmax = x; ymin = ymax = y; first = false; }
    else { xmin = std::min(xmin, x); xmax = std::max(xmax, x); ymin = std::min(ymin, y); ymax = std::max(ymax, y); }
  }
  if (g.wrap_x) { xmin = 0.; xmax = 360.; }
  double csx = 0., csy = 0.;
  const double R = h->rmax;
  if (R > 0.) {
    if (spherical) {
      const double deg = 180. / OAK_PI;
      csy = 0.5 * (R / OAK_EARTH_RADIUS) * deg;
      const double cl = std::max(std::cos(std::min(h->zlat_absmax, 89.) / deg), 0.05);
      csx = csy / cl;
    } else {
      csx = csy = 0.5 * R;
    }
  }
  auto ncells = [](double lo, double hi, double cs) -> double {
    if (!(cs > 0.) || !(hi > lo)) return 1.;
    return std::floor((hi - lo) / cs) + 1.;
  };
  double ncx = ncells(xmin, xmax, csx), ncy = by ? ncells(ymin, ymax, csy) : 1.;
  const double limit = std::max(4. * m, 4096.);
  if (ncx * ncy > limit) {  // coarsen: never more than ~4 cells per observation
    const double f = std::sqrt(ncx * ncy / limit);
    if (ncy > 1. && ncx > 1.) { csx *= f; csy *= f; }
    else if (ncx > 1.) csx *= ncx / limit;
    else csy *= ncy / limit;
    ncx = ncells(xmin, xmax, csx); ncy = by ? ncells(ymin, ymax, csy) : 1.;
  }
  g.ncx = (int)std::min(std::max(ncx, 1.), 65536.);
  g.ncy = (int)std::min(std::max(ncy, 1.), 65536.);
  if ((double)g.ncx * g.ncy > 64e6) { g.ncx = 1; g.ncy = 1; }
  g.x0 = xmin; g.y0 = ymin;
  g.csx = csx > 0. ? csx : 1.;
  g.csy = csy > 0. ? csy : 1.;
  const int ncell = g.ncx * g.ncy;
  const size_t mb = (size_t)std::max(m, 1);
  int rc;
  const size_t tmpb = oak_obsgrid_scratch_bytes(m, ncell);
  if ((rc = h->d_bx.ensure(8 * mb)) || (rc = h->d_by.ensure(8 * mb)) || (rc = h->d_key_in.ensure(4 * mb)) ||
      (rc = h->d_key_out.ensure(4 * mb)) || (rc = h->d_val_in.ensure(4 * mb)) || (rc = h->d_perm.ensure(4 * mb)) ||
      (rc = h->d_cell_start.ensure(4 * ((size_t)ncell + 1))) || (rc = h->d_sx.ensure(8 * mb)) ||
      (rc = h->d_sy.ensure(8 * mb)) || (rc = h->d_tmp.ensure(tmpb)))
    return rc;
  if (m > 0) {
    CUDA_TRY(cudaMemcpy(h->d_bx.p, bx, 8 * (size_t)m, cudaMemcpyHostToDevice));
    if (by) CUDA_TRY(cudaMemcpy(h->d_by.p, by, 8 * (size_t)m, cudaMemcpyHostToDevice));
  }
  g.cell_start = h->d_cell_start.as<int32_t>();
  g.perm = h->d_perm.as<int32_t>();
  g.sx = h->d_sx.as<double>();
  g.sy = h->d_sy.as<double>();
  cudaStream_t st = h->slot[0].st;
  rc = oak_build_obsgrid(st, m, h->d_bx.as<double>(), by ? h->d_by.as<double>() : nullptr, g, h->d_key_in.as<uint32_t>(),
                         h->d_key_out.as<uint32_t>(), h->d_val_in.as<int32_t>(), h->d_perm.as<int32_t>(),
                         h->d_cell_start.as<int32_t>(), h->d_sx.as<double>(), h->d_sy.as<double>(), h->d_tmp.p,
                         h->d_tmp.cap);
  if (rc) return rc;
  CUDA_TRY(cudaStreamSynchronize(st));
  h->og = g;
  h->obs_set = true;
  return 0;
}

extern "C" OAKB200_API int oakb200_select_observations(oakb200_handle *h, int32_t zone_first, int32_t zone_count,
                                           int64_t capacity, int64_t *offsets, int32_t *idx, double *weight) {
  if (!h || !offsets) { oak_set_error("select_observations: null argument"); return OAK_ERR_ARG; }
  if (!h->zones_set || !h->obs_set) { oak_set_error("select_observations: zones/observations not set"); return OAK_ERR_STATE; }
  if (zone_first < 0 || zone_count < 0 || zone_first + zone_count > h->nzones) { oak_set_error("select_observations: zone range out of bounds"); return OAK_ERR_ARG; }
  DeviceGuard guard(h->device);
  cudaStream_t st = h->slot[0].st;
  const ZoneGeom zg = h->geom();
  DevBuf d_counts, d_off, d_idx, d_w;
  int rc = 0;
  std::vector<int32_t> counts((size_t)zone_count);
  do {
    if ((rc = d_counts.ensure(4 * (size_t)std::max(zone_count, 1)))) break;
    if ((rc = oak_launch_select(st, zg, h->og, zone_first, zone_count, nullptr, d_counts.as<int32_t>(), nullptr, nullptr, false))) break;
    if (cudaMemcpyAsync(counts.data(), d_counts.p, 4 * (size_t)zone_count, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaStreamSynchronize(st) != cudaSuccess) { oak_set_error("select_observations: copy failed: %s", cudaGetErrorString(cudaGetLastError())); rc = OAK_ERR_CUDA; break; }
    offsets[0] = 0;
    for (int z = 0; z < zone_count; z++) offsets[z + 1] = offsets[z] + counts[z];
    const int64_t total = offsets[zone_count];
    if (total > capacity || (total > 0 && (!idx || !weight))) { oak_set_error("select_observations: capacity %lld < required %lld", (long long)capacity, (long long)total); rc = OAK_ERR_CAPACITY; break; }
    if (total == 0) break;
    if ((rc = d_off.ensure(8 * ((size_t)zone_count + 1))) || (rc = d_idx.ensure(4 * (size_t)total)) || (rc = d_w.ensure(8 * (size_t)total))) break;
    if (cudaMemcpyAsync(d_off.p, offsets, 8 * ((size_t)zone_count + 1), cudaMemcpyHostToDevice, st) != cudaSuccess) { oak_set_error("select_observations: copy failed"); rc = OAK_ERR_CUDA; break; }
    if ((rc = oak_launch_select(st, zg, h->og, zone_first, zone_count, d_off.as<int64_t>(), nullptr, d_idx.as<int32_t>(), d_w.as<double>(), true))) break;
    if (cudaMemcpyAsync(idx, d_idx.p, 4 * (size_t)total, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaMemcpyAsync(weight, d_w.p, 8 * (size_t)total, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaStreamSynchronize(st) != cudaSuccess) { oak_set_error("select_observations: copy failed: %s", cudaGetErrorString(cudaGetLastError())); rc = OAK_ERR_CUDA; break; }
    // increasing observation number inside each zone, the order pack() produces (rrsqrt.F90:395-404)
    std::vector<std::pair<int32_t, double>> tmp;
    for (int z = 0; z < zone_count; z++) {
      const int64_t a = offsets[z], b = offsets[z + 1];
      tmp.resize((size_t)(b - a));
      for (int64_t q = a; q < b; q++) tmp[(size_t)(q - a)] = {idx[q], weight[q]};
      std::sort(tmp.begin(), tmp.end(), [](const auto &u, const auto &v) { return u.first < v.first; });
      for (int64_t q = a; q < b; q++) { idx[q] = tmp[(size_t)(q - a)].first; weight[q] = tmp[(size_t)(q - a)].second; }
    }
  } while (0);
  d_counts.release(); d_off.release(); d_idx.release(); d_w.release();
  return rc;
}

extern "C" OAKB200_API int oakb200_local_analysis_dev(oakb200_handle *h, int64_t n, int32_t N, int32_t m, const double *xf,
                                          const double *Hxf, const double *yo, const double *Sf, int64_t ldSf,
                                          const double *HSf, int64_t ldHSf, const double *Rdiag, const double *d01,
                                          double *xa, double *Sa, int64_t ldSa, double *amplitudes, void *stream,
                                          oakb200_stats *stats) {
  int rc = check_ready(h, n, N, m);
  if (rc) return rc;
  if ((n > 0 && (!xf || !Sf || !xa || !Sa)) || (m > 0 && (!Hxf || !yo || !HSf || !Rdiag))) { oak_set_error("local_analysis: null array"); return OAK_ERR_ARG; }
  if (ldSf < n || ldSa < n || ldHSf < m) { oak_set_error("local_analysis: leading dimension too small"); return OAK_ERR_ARG; }
  if (n > 0 && xa == xf && !h->ens.on) { oak_set_error("local_analysis: xa must not be the array xf (Sa may be Sf; the mean is read while it is written)"); return OAK_ERR_ARG; }
  DeviceGuard guard(h->device);
  const int NP = padded(h, N);
  cudaStream_t s0 = h->slot[0].st;
  if (h->pending) { oak_set_error("local_analysis_dev: the previous asynchronous call has not been synchronised (oakb200_synchronize)"); return OAK_ERR_STATE; }
  // order after the caller's stream
  if (h->order_after_caller) {
    CUDA_TRY(cudaEventRecord(h->ev_user, (cudaStream_t)stream));
    for (int i = 0; i < NSLOT; i++) CUDA_TRY(cudaStreamWaitEvent(h->slot[i].st, h->ev_user, 0));
  }
  if ((rc = begin_call(h, stats))) return rc;
  CUDA_TRY(cudaEventRecord(h->ev_a, s0));
  if ((rc = pack_obs(h, s0, N, NP, HSf, ldHSf, yo, Hxf, Rdiag, d01))) return rc;
  if (amplitudes) CUDA_TRY(cudaMemsetAsync(amplitudes, 0, sizeof(double) * (size_t)N * h->nzones, s0));
  CUDA_TRY(cudaEventRecord(h->slot[0].ev[4], s0));
  for (int i = 1; i < NSLOT; i++) CUDA_TRY(cudaStreamWaitEvent(h->slot[i].st, h->slot[0].ev[4], 0));
  int64_t launches = 1;
  ProfAcc prof;
  const int zb = batch_size(h, NP, h->nzones);
  int bi = 0;
  // option "taper": the last batches of a call get smaller (half of what is left, in whole waves), so that the chain of
  // kernels that ends the call - which nothing overlaps any more - is short; matters on the small slabs of a multi-GPU run
  const int wave = NP <= 64 ? 592 : 148;
  for (int z0 = 0, zn = 0; z0 < h->nzones; z0 += zn, bi++) {
    Slot &s = h->slot[h->profile ? 0 : bi % NSLOT];
    zn = zb;
    const int left = h->nzones - z0;
    if (h->taper > 0 && !h->profile && left < (h->taper + 1) * zb && left > 2 * wave)
      zn = std::min(zb, std::max(wave, (left / 2 + wave - 1) / wave * wave));
    const int z1 = std::min(h->nzones, z0 + zn);
    if ((rc = run_zones(h, s, N, NP, z0, z1, 0, xf, Sf, ldSf, xa, Sa, ldSa, &launches, h->profile ? &prof : nullptr,
                        h->peers.n > 0, amplitudes, h->nrows))) {
      // nothing of this call may still be writing the caller's (or the peers') arrays when the error is returned
      for (int i = 0; i < NSLOT; i++) { cudaStreamSynchronize(h->slot[i].st); cudaStreamSynchronize(h->slot[i].cst); cudaStreamSynchronize(h->slot[i].qst); }
      for (int d = 0; d < OAKB200_MAX_PEERS; d++) if (h->pstream[d]) cudaStreamSynchronize(h->pstream[d]);
      return rc;
    }
  }
  for (int i = 1; i < NSLOT; i++) {
    CUDA_TRY(cudaEventRecord(h->slot[i].ev[5], h->slot[i].st));
    CUDA_TRY(cudaStreamWaitEvent(s0, h->slot[i].ev[5], 0));
  }
  if ((h->peers.n > 0 && h->peer_mode == 1) || h->mc_Sa) {
    for (int d = 0; d < h->peers.n; d++) {
      CUDA_TRY(cudaEventRecord(h->pev[d], h->pstream[d]));
      CUDA_TRY(cudaStreamWaitEvent(s0, h->pev[d], 0));
    }
    for (int i = 0; i < NSLOT; i++)
      if (h->slot[i].cst_used) {
        CUDA_TRY(cudaEventRecord(h->slot[i].ev[7], h->slot[i].cst));
        CUDA_TRY(cudaStreamWaitEvent(s0, h->slot[i].ev[7], 0));
        h->slot[i].cst_used = false;
      }
  }
  CUDA_TRY(cudaEventRecord(h->ev_b, s0));
  if (h->async && !h->profile) {
    // the caller's stream waits for the result; statistics and status come from oakb200_synchronize
    CUDA_TRY(cudaStreamWaitEvent((cudaStream_t)stream, h->ev_b, 0));
    h->pending = true;
    h->pending_launches = launches;
    if (stats) stats->launches = launches;
    return 0;
  }
  CUDA_TRY(cudaEventSynchronize(h->ev_b));
  float ms = 0.f, msp = 0.f;
  CUDA_TRY(cudaEventElapsedTime(&ms, h->ev_a, h->ev_b));
  CUDA_TRY(cudaEventElapsedTime(&msp, h->ev_a, h->slot[0].ev[4]));
  return end_call(h, stats, launches, prof, ms, msp);
}

// Number of relevant observations of every zone as the PRODUCTION kernel of the last analysis counted them (the
// selection fused into the Gram kernel: k_gram / k_gram_mma), for diagnostics and for the parity tests: the index sets
// themselves come from oakb200_select_observations (k_select), this is the cross-check that both evaluate the same sets.
extern "C" OAKB200_API int oakb200_zone_counts(oakb200_handle *h, int32_t *mloc) {
  if (!h || !mloc) { oak_set_error("zone_counts: null argument"); return OAK_ERR_ARG; }
  if (h->nzones <= 0 || !h->d_mloc.p) { oak_set_error("zone_counts: no zones configured"); return OAK_ERR_STATE; }
  DeviceGuard guard(h->device);
  for (int i = 0; i < NSLOT; i++) CUDA_TRY(cudaStreamSynchronize(h->slot[i].st));
  CUDA_TRY(cudaMemcpy(mloc, h->d_mloc.p, sizeof(int32_t) * (size_t)h->nzones, cudaMemcpyDeviceToHost));
  return 0;
}

extern "C" OAKB200_API int oakb200_synchronize(oakb200_handle *h, oakb200_stats *stats) {
  if (!h) { oak_set_error("null handle"); return OAK_ERR_ARG; }
  DeviceGuard guard(h->device);
  if (!h->pending) {
    for (int i = 0; i < NSLOT; i++) CUDA_TRY(cudaStreamSynchronize(h->slot[i].st));
    if (stats) memset(stats, 0, sizeof *stats);
    return 0;
  }
  CUDA_TRY(cudaEventSynchronize(h->ev_b));
  h->pending = false;
  float ms = 0.f, msp = 0.f;
  CUDA_TRY(cudaEventElapsedTime(&ms, h->ev_a, h->ev_b));
  CUDA_TRY(cudaEventElapsedTime(&msp, h->ev_a, h->slot[0].ev[4]));
  if (stats) memset(stats, 0, sizeof *stats);
  return end_call(h, stats, h->pending_launches, ProfAcc(), ms, msp);
}

extern "C" OAKB200_API int oakb200_local_analysis(oakb200_handle *h, int64_t n, int32_t N, int32_t m, const double *xf,
                                      const double *Hxf, const double *yo, const double *Sf, int64_t ldSf,
                                      const double *HSf, int64_t ldHSf, const double *Rdiag, const double *d01,
                                      double *xa, double *Sa, int64_t ldSa, double *amplitudes,
                                      oakb200_stats *stats) {
  int rc = check_ready(h, n, N, m);
  if (rc) return rc;
  if ((n > 0 && (!xf || !Sf || !xa || !Sa)) || (m > 0 && (!Hxf || !yo || !HSf || !Rdiag))) { oak_set_error("local_analysis: null array"); return OAK_ERR_ARG; }
  if (ldSf < n || ldSa < n || ldHSf < m) { oak_set_error("local_analysis: leading dimension too small"); return OAK_ERR_ARG; }
  DeviceGuard guard(h->device);
  const int NP = padded(h, N);
  cudaStream_t s0 = h->slot[0].st;
  HostPins pins;
  if (h->host_register) {
    pins.pin(Sf, sizeof(double) * ((size_t)ldSf * (N - 1) + (size_t)n));
    if (Sa != Sf) pins.pin(Sa, sizeof(double) * ((size_t)ldSa * (N - 1) + (size_t)n));
    pins.pin(HSf, sizeof(double) * ((size_t)ldHSf * (N - 1) + (size_t)m));
    pins.pin(xf, sizeof(double) * (size_t)n);
    pins.pin(xa, sizeof(double) * (size_t)n);
  }
  if ((rc = begin_call(h, stats))) return rc;
  CUDA_TRY(cudaEventRecord(h->ev_a, s0));
  int64_t h2d = 0, d2h = 0;
  // observation-space arrays: whole, once
  const size_t mb = (size_t)std::max(m, 1);
  if ((rc = h->d_HSf.ensure(8 * mb * N)) || (rc = h->d_yo.ensure(8 * mb)) || (rc = h->d_Hxf.ensure(8 * mb)) ||
      (rc = h->d_R.ensure(8 * mb)) || (rc = h->d_d01.ensure(8 * mb)))
    return rc;
  // pageable caller arrays: through the slot's pinned staging buffers, filled / drained by several host threads
  const bool stage = !h->host_register && n > 0 &&
                     (h->host_stage == 1 || (h->host_stage < 0 && 8. * (double)n * N >= 32. * 1024 * 1024)) &&
                     (is_pageable(Sf) || is_pageable(Sa));
  const int nthreads = h->stage_threads > 0 ? h->stage_threads : (int)std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
  if (m > 0) {
    if (stage && is_pageable(HSf)) {
      if ((rc = staged_copy(h, true, h->d_HSf.as<double>(), (size_t)m, const_cast<double *>(HSf), (size_t)ldHSf, (size_t)m, N, nthreads))) return rc;
    } else
      CUDA_TRY(cudaMemcpy2DAsync(h->d_HSf.p, 8 * (size_t)m, HSf, 8 * (size_t)ldHSf, 8 * (size_t)m, N, cudaMemcpyHostToDevice, s0));
    CUDA_TRY(cudaMemcpyAsync(h->d_yo.p, yo, 8 * (size_t)m, cudaMemcpyHostToDevice, s0));
    CUDA_TRY(cudaMemcpyAsync(h->d_Hxf.p, Hxf, 8 * (size_t)m, cudaMemcpyHostToDevice, s0));
    CUDA_TRY(cudaMemcpyAsync(h->d_R.p, Rdiag, 8 * (size_t)m, cudaMemcpyHostToDevice, s0));
    if (d01) CUDA_TRY(cudaMemcpyAsync(h->d_d01.p, d01, 8 * (size_t)m, cudaMemcpyHostToDevice, s0));
    h2d += 8ll * m * (N + 3 + (d01 ? 1 : 0));
  }
  if ((rc = pack_obs(h, s0, N, NP, h->d_HSf.as<double>(), m, h->d_yo.as<double>(), h->d_Hxf.as<double>(),
                     h->d_R.as<double>(), d01 ? h->d_d01.as<double>() : nullptr)))
    return rc;
  CUDA_TRY(cudaEventRecord(h->slot[0].ev[4], s0));
  for (int i = 1; i < NSLOT; i++) CUDA_TRY(cudaStreamWaitEvent(h->slot[i].st, h->slot[0].ev[4], 0));
  if (amplitudes) memset(amplitudes, 0, sizeof(double) * (size_t)N * h->nzones);  // rrsqrt.F90:324
  double *d_ampl_out = nullptr;
  if (amplitudes && !h->localise_obs) {
    if ((rc = h->d_ampzero.ensure(sizeof(double) * (size_t)N * std::max(h->nzones, 1)))) return rc;
    d_ampl_out = h->d_ampzero.as<double>();
  }
  // chunks of whole zones
  const int64_t rows_target = std::max<int64_t>(1, (int64_t)(h->chunk_mb * 1024. * 1024. / (8. * N)));
  int64_t launches = 1;
  ProfAcc prof;
  for (int i = 0; i < NSLOT; i++) h->slot[i].hs.pending = false;   // nothing of an earlier (failed) call is drained into these arrays
  auto drain = [&](Slot &s) -> int {   // the slot's finished chunk: pinned buffer -> the caller's Sa, xa
    if (!s.hs.pending) return 0;
    CUDA_TRY(cudaEventSynchronize(s.hs.ev_out));
    par_copy_cols(Sa + s.hs.r0, (size_t)ldSa, s.hs.out, (size_t)s.hs.rows, (size_t)s.hs.rows, N, nthreads);
    memcpy(xa + s.hs.r0, s.hs.out + (size_t)s.hs.rows * N, sizeof(double) * (size_t)s.hs.rows);
    s.hs.pending = false;
    return 0;
  };
  int ci = 0;
  for (int z0 = 0; z0 < h->nzones; ci++) {
    int z1 = z0;
    const int64_t r0 = h->h_zstart[z0];
    while (z1 < h->nzones && (z1 == z0 || h->h_zstart[z1 + 1] - r0 <= rows_target)) z1++;
    const int64_t rows = h->h_zstart[z1] - r0;
    Slot &s = h->slot[h->profile ? 0 : ci % NSLOT];
    if (rows > 0) {
      if ((rc = s.S.ensure(8 * (size_t)rows * N)) || (rc = s.xf.ensure(8 * (size_t)rows)) || (rc = s.xa.ensure(8 * (size_t)rows))) return rc;
      if (stage) {
        // the slot's previous chunk leaves its output buffer first (its device-to-host copy is behind its host-to-device
        // copy in the slot's stream, so the input buffer is free as well)
        if ((rc = drain(s)) || (rc = ensure_stage(s, (size_t)rows * (N + 1)))) return rc;
        par_copy_cols(s.hs.in, (size_t)rows, Sf + r0, (size_t)ldSf, (size_t)rows, N, nthreads);
        memcpy(s.hs.in + (size_t)rows * N, xf + r0, sizeof(double) * (size_t)rows);
        CUDA_TRY(cudaMemcpyAsync(s.S.p, s.hs.in, 8 * (size_t)rows * N, cudaMemcpyHostToDevice, s.st));
        CUDA_TRY(cudaMemcpyAsync(s.xf.p, s.hs.in + (size_t)rows * N, 8 * (size_t)rows, cudaMemcpyHostToDevice, s.st));
      } else {
        CUDA_TRY(cudaMemcpy2DAsync(s.S.p, 8 * (size_t)rows, Sf + r0, 8 * (size_t)ldSf, 8 * (size_t)rows, N, cudaMemcpyHostToDevice, s.st));
        CUDA_TRY(cudaMemcpyAsync(s.xf.p, xf + r0, 8 * (size_t)rows, cudaMemcpyHostToDevice, s.st));
      }
      h2d += 8ll * rows * (N + 1);
    }
    if ((rc = run_zones(h, s, N, NP, z0, z1, r0, s.xf.as<double>(), s.S.as<double>(), rows, s.xa.as<double>(),
                        s.S.as<double>(), rows, &launches, h->profile ? &prof : nullptr, false, d_ampl_out, rows)))
      return rc;
    if (rows > 0) {
      if (stage) {
        CUDA_TRY(cudaMemcpyAsync(s.hs.out, s.S.p, 8 * (size_t)rows * N, cudaMemcpyDeviceToHost, s.st));
        CUDA_TRY(cudaMemcpyAsync(s.hs.out + (size_t)rows * N, s.xa.p, 8 * (size_t)rows, cudaMemcpyDeviceToHost, s.st));
        CUDA_TRY(cudaEventRecord(s.hs.ev_out, s.st));
        s.hs.pending = true; s.hs.r0 = r0; s.hs.rows = rows;
      } else {
        CUDA_TRY(cudaMemcpy2DAsync(Sa + r0, 8 * (size_t)ldSa, s.S.p, 8 * (size_t)rows, 8 * (size_t)rows, N, cudaMemcpyDeviceToHost, s.st));
        CUDA_TRY(cudaMemcpyAsync(xa + r0, s.xa.p, 8 * (size_t)rows, cudaMemcpyDeviceToHost, s.st));
      }
      d2h += 8ll * rows * (N + 1);
    }
    z0 = z1;
  }
  if (stage)
    for (int i = 0; i < NSLOT; i++)
      if ((rc = drain(h->slot[i]))) return rc;
  for (int i = 1; i < NSLOT; i++) {
    CUDA_TRY(cudaEventRecord(h->slot[i].ev[5], h->slot[i].st));
    CUDA_TRY(cudaStreamWaitEvent(s0, h->slot[i].ev[5], 0));
  }
  if (d_ampl_out) CUDA_TRY(cudaMemcpyAsync(amplitudes, d_ampl_out, sizeof(double) * (size_t)N * h->nzones, cudaMemcpyDeviceToHost, s0));
  CUDA_TRY(cudaEventRecord(h->ev_b, s0));
  CUDA_TRY(cudaEventSynchronize(h->ev_b));
  float ms = 0.f, msp = 0.f;
  CUDA_TRY(cudaEventElapsedTime(&ms, h->ev_a, h->ev_b));
  CUDA_TRY(cudaEventElapsedTime(&msp, h->ev_a, h->slot[0].ev[4]));
  rc = end_call(h, stats, launches, prof, ms, msp);
  if (stats) { stats->h2d_bytes = h2d; stats->d2h_bytes = d2h; }
  return rc;
}

// ---------------------------------------------------------------------------------------------------------
// global scheme (schemetype = 0): analysis, rrsqrt.F90:196-208
// ---------------------------------------------------------------------------------------------------------
namespace {

constexpr int GLOBAL_BLOCK_ROWS = 512;  // rows of the state one k_apply CTA updates with the shared transform

int global_check(oakb200_handle *h, int64_t n, int N, int m) {
  if (!h) { oak_set_error("null handle"); return OAK_ERR_ARG; }
  if (n < 0 || m < 0) { oak_set_error("global_analysis: negative size"); return OAK_ERR_ARG; }
  if (N < 2) { oak_set_error("ensemble size N = %d < 2", N); return OAK_ERR_ARG; }
  if (padded(N) < 0) { oak_set_error("ensemble size N = %d > 128 is not supported", N); return OAK_ERR_UNSUPPORTED; }
  return 0;
}

// (ampl, T) of the global scheme into slot 0's workspace from device-resident observation-space arrays; enqueued on
// slot 0's stream.  m > 0.
int global_transform(oakb200_handle *h, int N, int NP, int m, const double *HSf, int64_t ldH, const double *yo,
                     const double *Hxf, const double *Rdiag, const double *d01, int64_t *launches) {
  Slot &s = h->slot[0];
  int rc;
  if ((rc = ensure_ws(h, s, NP, 1))) return rc;
  if ((rc = h->d_mloc.ensure(sizeof(int32_t)))) return rc;
  const int nparts = oak_global_gram_parts(m);
  if ((rc = h->d_gws.ensure(oak_global_gram_ws_bytes(NP, nparts)))) return rc;
  DevCounters *ctr = h->d_ctr.as<DevCounters>();
  int32_t *mloc = h->d_mloc.as<int32_t>();
  if ((rc = oak_launch_global_gram(s.st, m, N, NP, HSf, ldH, yo, Hxf, Rdiag, d01, h->d_gws.p, nparts, s.G.as<double>(),
                                   s.c.as<double>(), mloc))) return rc;
  *launches += 2;
  if (h->eig_kernel == 4 && NP <= 64) {
    int32_t *flags = nullptr;
    if ((rc = oak_launch_eig_tridiag(s.st, N, NP, 1, mloc, s.G.as<double>(), s.c.as<double>(), s.T.as<double>(),
                                     s.ampl.as<double>(), s.tri.p, &flags, ctr, nullptr, h->tri_orthtol, h->tri_maxgroup,
                                     nullptr, nullptr))) return rc;
    if ((rc = oak_launch_eig(s.st, 0, N, NP, 0, 1, flags, s.G.as<double>(), s.c.as<double>(), s.T.as<double>(),
                             s.ampl.as<double>(), h->tol, h->max_sweeps, ctr))) return rc;
    *launches += 4;
  } else {
    if ((rc = oak_launch_eig(s.st, h->eig_kernel == 4 ? 0 : h->eig_kernel, N, NP, 0, 1, mloc, s.G.as<double>(),
                             s.c.as<double>(), s.T.as<double>(), s.ampl.as<double>(), h->tol, h->max_sweeps, ctr))) return rc;
    *launches += 1;
  }
  return 0;
}

// Sa = Sf T, xa = xf + Sf ampl for the rows [row0, row0 + rows) held in the given buffers (first row = row0, a
// multiple of GLOBAL_BLOCK_ROWS), on stream slot s; T and ampl are slot 0's.
int global_apply(oakb200_handle *h, Slot &s, int N, int NP, int64_t n, int64_t row0, int64_t rows, const double *xf,
                 const double *Sf, int64_t ldS, double *xa, double *Sa, int64_t ldSa, int64_t *launches) {
  if (rows <= 0) return 0;
  ZoneGeom zg{};
  zg.zstart = h->d_gzstart.as<int64_t>();
  const int b0 = (int)(row0 / GLOBAL_BLOCK_ROWS), b1 = (int)((row0 + rows + GLOBAL_BLOCK_ROWS - 1) / GLOBAL_BLOCK_ROWS);
  (void)n;
  PeerOut none{};
  int rc;
  if (h->apply_kernel == 1 && NP == 64)
    rc = oak_launch_apply_mma(s.st, N, NP, zg, b0, b1 - b0, row0, nullptr, h->slot[0].T.as<double>(),
                              h->slot[0].ampl.as<double>(), xf, Sf, ldS, xa, Sa, ldSa, nullptr, true);
  else
    rc = oak_launch_apply(s.st, N, NP, zg, b0, b1 - b0, row0, nullptr, h->slot[0].T.as<double>(),
                          h->slot[0].ampl.as<double>(), xf, Sf, ldS, xa, Sa, ldSa, none, nullptr, true);
  *launches += 1;
  return rc;
}

int global_block_starts(oakb200_handle *h, cudaStream_t st, int64_t n) {
  const int nblocks = (int)((n + GLOBAL_BLOCK_ROWS - 1) / GLOBAL_BLOCK_ROWS);
  int rc = h->d_gzstart.ensure(sizeof(int64_t) * ((size_t)nblocks + 1));
  if (rc) return rc;
  return oak_launch_block_starts(st, n, GLOBAL_BLOCK_ROWS, nblocks, h->d_gzstart.as<int64_t>());
}

int global_finish(oakb200_handle *h, oakb200_stats *stats, int64_t launches, int m, float ms) {
  const int saved = h->nzones;
  h->nzones = 1;
  int rc = end_call(h, stats, launches, ProfAcc(), ms, 0.f);
  h->nzones = saved;
  if (stats) { stats->obs_relevant_sum = m; stats->zones_skipped = m == 0 ? 1 : 0; }
  return rc;
}

}  // namespace

extern "C" OAKB200_API int oakb200_global_analysis_dev(oakb200_handle *h, int64_t n, int32_t N, int32_t m, const double *xf,
                                           const double *Hxf, const double *yo, const double *Sf, int64_t ldSf,
                                           const double *HSf, int64_t ldHSf, const double *Rdiag, const double *d01,
                                           double *xa, double *Sa, int64_t ldSa, double *amplitudes, void *stream,
                                           oakb200_stats *stats) {
  int rc = global_check(h, n, N, m);
  if (rc) return rc;
  if ((n > 0 && (!xf || !Sf || !xa || !Sa)) || (m > 0 && (!Hxf || !yo || !HSf || !Rdiag))) { oak_set_error("global_analysis: null array"); return OAK_ERR_ARG; }
  if (ldSf < n || ldSa < n || ldHSf < m) { oak_set_error("global_analysis: leading dimension too small"); return OAK_ERR_ARG; }
  DeviceGuard guard(h->device);
  const int NP = padded(h, N);
  cudaStream_t s0 = h->slot[0].st;
  CUDA_TRY(cudaEventRecord(h->ev_user, (cudaStream_t)stream));
  CUDA_TRY(cudaStreamWaitEvent(s0, h->ev_user, 0));
  if (stats) memset(stats, 0, sizeof *stats);
  CUDA_TRY(cudaMemsetAsync(h->d_ctr.p, 0, sizeof(DevCounters), s0));
  CUDA_TRY(cudaEventRecord(h->ev_a, s0));
  int64_t launches = 0;
  if (m == 0) {  // nothing to assimilate: the increment is zero (G = 0, c = 0)
    if (n > 0) {
      if (Sa != Sf) CUDA_TRY(cudaMemcpy2DAsync(Sa, 8 * (size_t)ldSa, Sf, 8 * (size_t)ldSf, 8 * (size_t)n, N, cudaMemcpyDeviceToDevice, s0));
      CUDA_TRY(cudaMemcpyAsync(xa, xf, 8 * (size_t)n, cudaMemcpyDeviceToDevice, s0));
    }
    if (amplitudes) CUDA_TRY(cudaMemsetAsync(amplitudes, 0, sizeof(double) * N, s0));
  } else {
    if ((rc = global_transform(h, N, NP, m, HSf, ldHSf, yo, Hxf, Rdiag, d01, &launches))) return rc;
    if ((rc = global_block_starts(h, s0, n))) return rc;
    if ((rc = global_apply(h, h->slot[0], N, NP, n, 0, n, xf, Sf, ldSf, xa, Sa, ldSa, &launches))) return rc;
    launches += 1;
    if (amplitudes) CUDA_TRY(cudaMemcpyAsync(amplitudes, h->slot[0].ampl.p, sizeof(double) * N, cudaMemcpyDeviceToDevice, s0));
  }
  CUDA_TRY(cudaEventRecord(h->ev_b, s0));
  CUDA_TRY(cudaEventSynchronize(h->ev_b));
  float ms = 0.f;
  CUDA_TRY(cudaEventElapsedTime(&ms, h->ev_a, h->ev_b));
  return global_finish(h, stats, launches, m, ms);
}

extern "C" OAKB200_API int oakb200_global_analysis(oakb200_handle *h, int64_t n, int32_t N, int32_t m, const double *xf,
                                       const double *Hxf, const double *yo, const double *Sf, int64_t ldSf,
                                       const double *HSf, int64_t ldHSf, const double *Rdiag, const double *d01,
                                       double *xa, double *Sa, int64_t ldSa, double *amplitudes, oakb200_stats *stats) {
  int rc = global_check(h, n, N, m);
  if (rc) return rc;
  if ((n > 0 && (!xf || !Sf || !xa || !Sa)) || (m > 0 && (!Hxf || !yo || !HSf || !Rdiag))) { oak_set_error("global_analysis: null array"); return OAK_ERR_ARG; }
  if (ldSf < n || ldSa < n || ldHSf < m) { oak_set_error("global_analysis: leading dimension too small"); return OAK_ERR_ARG; }
  DeviceGuard guard(h->device);
  const int NP = padded(h, N);
  cudaStream_t s0 = h->slot[0].st;
  if (stats) memset(stats, 0, sizeof *stats);
  CUDA_TRY(cudaMemsetAsync(h->d_ctr.p, 0, sizeof(DevCounters), s0));
  CUDA_TRY(cudaEventRecord(h->ev_a, s0));
  int64_t launches = 0, h2d = 0, d2h = 0;
  if (m == 0) {
    for (int k = 0; k < N && n > 0; k++)
      if (Sa != Sf) memmove(Sa + (size_t)ldSa * k, Sf + (size_t)ldSf * k, 8 * (size_t)n);
    if (n > 0) memmove(xa, xf, 8 * (size_t)n);
    if (amplitudes) memset(amplitudes, 0, sizeof(double) * N);
  } else {
    const size_t mb = (size_t)m;
    if ((rc = h->d_HSf.ensure(8 * mb * N)) || (rc = h->d_yo.ensure(8 * mb)) || (rc = h->d_Hxf.ensure(8 * mb)) ||
        (rc = h->d_R.ensure(8 * mb)) || (rc = h->d_d01.ensure(8 * mb)))
      return rc;
    CUDA_TRY(cudaMemcpy2DAsync(h->d_HSf.p, 8 * mb, HSf, 8 * (size_t)ldHSf, 8 * mb, N, cudaMemcpyHostToDevice, s0));
    CUDA_TRY(cudaMemcpyAsync(h->d_yo.p, yo, 8 * mb, cudaMemcpyHostToDevice, s0));
    CUDA_TRY(cudaMemcpyAsync(h->d_Hxf.p, Hxf, 8 * mb, cudaMemcpyHostToDevice, s0));
    CUDA_TRY(cudaMemcpyAsync(h->d_R.p, Rdiag, 8 * mb, cudaMemcpyHostToDevice, s0));
    if (d01) CUDA_TRY(cudaMemcpyAsync(h->d_d01.p, d01, 8 * mb, cudaMemcpyHostToDevice, s0));
    h2d += 8ll * m * (N + 3 + (d01 ? 1 : 0));
    if ((rc = global_transform(h, N, NP, m, h->d_HSf.as<double>(), m, h->d_yo.as<double>(), h->d_Hxf.as<double>(),
                               h->d_R.as<double>(), d01 ? h->d_d01.as<double>() : nullptr, &launches))) return rc;
    if ((rc = global_block_starts(h, s0, n))) return rc;
    launches += 1;
    if (amplitudes) CUDA_TRY(cudaMemcpyAsync(amplitudes, h->slot[0].ampl.p, sizeof(double) * N, cudaMemcpyDeviceToHost, s0));
    CUDA_TRY(cudaEventRecord(h->slot[0].ev[4], s0));
    for (int i = 1; i < NSLOT; i++) CUDA_TRY(cudaStreamWaitEvent(h->slot[i].st, h->slot[0].ev[4], 0));
    // the state streams through the device in chunks of whole row blocks, one stream slot per chunk
    int64_t rows_target = std::max<int64_t>(1, (int64_t)(h->chunk_mb * 1024. * 1024. / (8. * N)));
    rows_target = std::max<int64_t>(GLOBAL_BLOCK_ROWS, rows_target / GLOBAL_BLOCK_ROWS * GLOBAL_BLOCK_ROWS);
    int ci = 0;
    for (int64_t r0 = 0; r0 < n; r0 += rows_target, ci++) {
      const int64_t rows = std::min(rows_target, n - r0);
      Slot &s = h->slot[ci % NSLOT];
      if ((rc = s.S.ensure(8 * (size_t)rows * N)) || (rc = s.xf.ensure(8 * (size_t)rows)) || (rc = s.xa.ensure(8 * (size_t)rows))) return rc;
      CUDA_TRY(cudaMemcpy2DAsync(s.S.p, 8 * (size_t)rows, Sf + r0, 8 * (size_t)ldSf, 8 * (size_t)rows, N, cudaMemcpyHostToDevice, s.st));
      CUDA_TRY(cudaMemcpyAsync(s.xf.p, xf + r0, 8 * (size_t)rows, cudaMemcpyHostToDevice, s.st));
      if ((rc = global_apply(h, s, N, NP, n, r0, rows, s.xf.as<double>(), s.S.as<double>(), rows, s.xa.as<double>(),
                             s.S.as<double>(), rows, &launches))) return rc;
      CUDA_TRY(cudaMemcpy2DAsync(Sa + r0, 8 * (size_t)ldSa, s.S.p, 8 * (size_t)rows, 8 * (size_t)rows, N, cudaMemcpyDeviceToHost, s.st));
      CUDA_TRY(cudaMemcpyAsync(xa + r0, s.xa.p, 8 * (size_t)rows, cudaMemcpyDeviceToHost, s.st));
      h2d += 8ll * rows * (N + 1);
      d2h += 8ll * rows * (N + 1);
    }
    for (int i = 1; i < NSLOT; i++) {
      CUDA_TRY(cudaEventRecord(h->slot[i].ev[5], h->slot[i].st));
      CUDA_TRY(cudaStreamWaitEvent(s0, h->slot[i].ev[5], 0));
    }
  }
  CUDA_TRY(cudaEventRecord(h->ev_b, s0));
  CUDA_TRY(cudaEventSynchronize(h->ev_b));
  float ms = 0.f;
  CUDA_TRY(cudaEventElapsedTime(&ms, h->ev_a, h->ev_b));
  rc = global_finish(h, stats, launches, m, ms);
  if (stats) { stats->h2d_bytes = h2d; stats->d2h_bytes = d2h; }
  return rc;
}

extern "C" OAKB200_API int oakb200_assim_ensemble_dev(oakb200_handle *h, int64_t n, int32_t N, int32_t m, const double *E,
                                          int64_t ldE, int64_t nnz, const int32_t *Hi, const int32_t *Hj,
                                          const double *Hs, const double *Hshift, const double *yo,
                                          const double *Rdiag, const double *d01, int32_t anamtype, double inflation,
                                          const double *maxCorrection, double *Ea, int64_t ldEa, double *xf_out,
                                          double *xa_out, void *stream, oakb200_stats *stats) {
  int rc = (h && h->scheme == 0) ? global_check(h, n, N, m) : check_ready(h, n, N, m);
  if (rc) return rc;
  if (anamtype < 0 || anamtype > 3) { oak_set_error("assim_ensemble: anamorphosis type %d unknown (0 per variable, 1 identity, 2 log, 3 tabulated)", anamtype); return OAK_ERR_ARG; }
  if (anamtype == 3 && h->anam_K < 2) { oak_set_error("assim_ensemble: tabulated anamorphosis without a table (oakb200_set_anamorphosis_table)"); return OAK_ERR_STATE; }
  if (anamtype == 0 && (h->anam_nvar < 1 || h->anam_rows != n)) { oak_set_error("assim_ensemble: per-variable anamorphosis needs oakb200_set_anamorphosis_vars for the %lld rows of this state", (long long)n); return OAK_ERR_STATE; }
  const AnamTab at{h->d_anam.as<double>(), h->anam_K, h->anam_monotone, h->d_rowvar.as<int32_t>(), h->d_vdesc.as<int32_t>(), h->d_vtab.as<double>()};
  if ((n > 0 && (!E || !Ea)) || (nnz > 0 && (!Hi || !Hj || !Hs)) || (m > 0 && (!yo || !Rdiag))) { oak_set_error("assim_ensemble: null array"); return OAK_ERR_ARG; }
  DeviceGuard guard(h->device);
  cudaStream_t s0 = h->slot[0].st;
  CUDA_TRY(cudaEventRecord(h->ev_user, (cudaStream_t)stream));
  CUDA_TRY(cudaStreamWaitEvent(s0, h->ev_user, 0));
  const size_t mb = (size_t)std::max(m, 1), nzb = (size_t)std::max<int64_t>(nnz, 1);
  if ((rc = h->d_HE.ensure(8 * mb * N)) || (rc = h->d_Hxf.ensure(8 * mb)) || (rc = h->d_key_in.ensure(4 * std::max(nzb, mb))) ||
      (rc = h->d_key_out.ensure(4 * std::max(nzb, mb))) || (rc = h->d_val_in.ensure(4 * std::max(nzb, mb))) ||
      (rc = h->d_order.ensure(4 * nzb)) || (rc = h->d_rowstart.ensure(4 * (mb + 2))) ||
      (rc = h->d_tmp.ensure(std::max(oak_coo_scratch_bytes(nnz, m), h->d_tmp.cap))) ||
      (rc = h->d_xf.ensure(8 * (size_t)std::max<int64_t>(n, 1))) || (rc = h->d_xa.ensure(8 * (size_t)std::max<int64_t>(n, 1))))
    return rc;
  // NB: d_key_*/d_val_in/d_tmp are shared with the observation grid build, which is complete (synchronous).
  if ((rc = oak_coo_to_rows(s0, nnz, m, Hi, Hj, h->d_key_in.as<uint32_t>(), h->d_key_out.as<uint32_t>(),
                            h->d_val_in.as<int32_t>(), h->d_order.as<int32_t>(), h->d_rowstart.as<int32_t>(),
                            h->d_tmp.p, h->d_tmp.cap)))
    return rc;
  // HE = H E + Hshift on the untransformed state (assimilation.F90:3112-3114)
  if ((rc = oak_launch_obsoper_rows(s0, m, N, h->d_rowstart.as<int32_t>(), h->d_order.as<int32_t>(), Hj, Hs, Hshift, E, ldE, h->d_HE.as<double>()))) return rc;
  // Hxf, HSf (in place in HE)
  if ((rc = oak_launch_mean_anom(s0, m, N, 1, AnamTab{nullptr, 0, 0, nullptr, nullptr, nullptr}, h->d_HE.as<double>(), m, h->d_Hxf.as<double>(), h->d_HE.as<double>(), m))) return rc;
  // fused form (local scheme, matrix form of the apply): the apply kernel reads E and writes Ea, xf, xa itself
  const bool fuse = (h->ens_fuse == 1 || (h->ens_fuse < 0 && anamtype == 1)) && h->scheme != 0 && !h->fuse_apply && h->apply_kernel == 0 && n > 0;
  if (fuse) {
    CUDA_TRY(cudaStreamSynchronize(s0));
    h->ens = EnsFuse{1, anamtype, at, inflation, sqrt((double)N - 1.), maxCorrection, h->d_xf.as<double>()};
    rc = oakb200_local_analysis_dev(h, n, N, m, h->d_xf.as<double>(), h->d_Hxf.as<double>(), yo, E, ldE,
                                    h->d_HE.as<double>(), m, Rdiag, d01, h->d_xa.as<double>(), Ea, ldEa, nullptr,
                                    (void *)s0, stats);
    h->ens.on = 0;
    if (rc) return rc;
    if (stats) stats->launches += 3;
  } else {
  // xf, Sf (into Ea)
  if ((rc = oak_launch_mean_anom(s0, n, N, anamtype, at, E, ldE, h->d_xf.as<double>(), Ea, ldEa))) return rc;
  CUDA_TRY(cudaStreamSynchronize(s0));
  if (h->scheme == 0)   // Assim's global branch: call analysis(xf,Hxf,yo,Sf,HSf,R,xa,Sa,amplitudes)
    rc = oakb200_global_analysis_dev(h, n, N, m, h->d_xf.as<double>(), h->d_Hxf.as<double>(), yo, Ea, ldEa,
                                     h->d_HE.as<double>(), m, Rdiag, d01, h->d_xa.as<double>(), Ea, ldEa, nullptr,
                                     (void *)s0, stats);
  else
    rc = oakb200_local_analysis_dev(h, n, N, m, h->d_xf.as<double>(), h->d_Hxf.as<double>(), yo, Ea, ldEa,
                                    h->d_HE.as<double>(), m, Rdiag, d01, h->d_xa.as<double>(), Ea, ldEa, nullptr,
                                    (void *)s0, stats);
  if (rc) return rc;
  if ((rc = oak_launch_epilogue(s0, n, N, anamtype, at, inflation, maxCorrection, h->d_xf.as<double>(), h->d_xa.as<double>(), Ea, ldEa, Ea, ldEa))) return rc;
  if (stats) stats->launches += 6;
  }
  if (xf_out) CUDA_TRY(cudaMemcpyAsync(xf_out, h->d_xf.p, 8 * (size_t)n, cudaMemcpyDeviceToDevice, s0));
  if (xa_out) CUDA_TRY(cudaMemcpyAsync(xa_out, h->d_xa.p, 8 * (size_t)n, cudaMemcpyDeviceToDevice, s0));
  CUDA_TRY(cudaStreamSynchronize(s0));
  return 0;
}

extern "C" OAKB200_API int oakb200_assim_ensemble(oakb200_handle *h, int64_t n, int32_t N, int32_t m, const double *E, int64_t ldE,
                                      int64_t nnz, const int32_t *Hi, const int32_t *Hj, const double *Hs,
                                      const double *Hshift, const double *yo, const double *Rdiag, const double *d01,
                                      int32_t anamtype, double inflation, const double *maxCorrection, double *Ea,
                                      int64_t ldEa, double *xf_out, double *xa_out, oakb200_stats *stats) {
  int rc = (h && h->scheme == 0) ? global_check(h, n, N, m) : check_ready(h, n, N, m);
  if (rc) return rc;
  if ((n > 0 && (!E || !Ea)) || (nnz > 0 && (!Hi || !Hj || !Hs)) || (m > 0 && (!yo || !Rdiag))) { oak_set_error("assim_ensemble: null array"); return OAK_ERR_ARG; }
  DeviceGuard guard(h->device);
  // whole ensemble resident (H E needs arbitrary rows of E): n*N*8 bytes must fit on the device
  const size_t nb = (size_t)std::max<int64_t>(n, 1), mb = (size_t)std::max(m, 1), nzb = (size_t)std::max<int64_t>(nnz, 1);
  DevBuf dHi, dHj, dHs, dHshift, dyo, dR, dd01, dmaxc, dxf, dxa;
  auto cleanup = [&]() { dHi.release(); dHj.release(); dHs.release(); dHshift.release(); dyo.release(); dR.release(); dd01.release(); dmaxc.release(); dxf.release(); dxa.release(); };
  if ((rc = h->d_E.ensure(8 * nb * N)) || (rc = dHi.ensure(4 * nzb)) || (rc = dHj.ensure(4 * nzb)) || (rc = dHs.ensure(8 * nzb)) ||
      (rc = dHshift.ensure(8 * mb)) || (rc = dyo.ensure(8 * mb)) || (rc = dR.ensure(8 * mb)) || (rc = dd01.ensure(8 * mb)) ||
      (rc = dmaxc.ensure(8 * nb)) || (rc = dxf.ensure(8 * nb)) || (rc = dxa.ensure(8 * nb))) { cleanup(); return rc; }
  cudaError_t e = cudaSuccess;
  auto up = [&](void *d, const void *s, size_t bytes) { if (e == cudaSuccess && bytes && s) e = cudaMemcpy(d, s, bytes, cudaMemcpyHostToDevice); };
  // pageable E / Ea of 32 MB and more: through the pinned staging buffers with host threads (option host_stage)
  const bool stage = !h->host_register && n > 0 && (h->host_stage == 1 || (h->host_stage < 0 && 8. * (double)n * N >= 32. * 1024 * 1024));
  const int nthreads = h->stage_threads > 0 ? h->stage_threads : (int)std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
  if (n > 0 && stage && is_pageable(E)) {
    if ((rc = staged_copy(h, true, h->d_E.as<double>(), (size_t)n, const_cast<double *>(E), (size_t)ldE, (size_t)n, N, nthreads))) { cleanup(); return rc; }
  } else if (n > 0) e = cudaMemcpy2D(h->d_E.p, 8 * (size_t)n, E, 8 * (size_t)ldE, 8 * (size_t)n, N, cudaMemcpyHostToDevice);
  up(dHi.p, Hi, 4 * (size_t)nnz); up(dHj.p, Hj, 4 * (size_t)nnz); up(dHs.p, Hs, 8 * (size_t)nnz);
  up(dHshift.p, Hshift, 8 * (size_t)m); up(dyo.p, yo, 8 * (size_t)m); up(dR.p, Rdiag, 8 * (size_t)m);
  up(dd01.p, d01, 8 * (size_t)m); up(dmaxc.p, maxCorrection, 8 * (size_t)n);
  if (e != cudaSuccess) { oak_set_error("assim_ensemble: H2D copy failed: %s", cudaGetErrorString(e)); cleanup(); return OAK_ERR_CUDA; }
  rc = oakb200_assim_ensemble_dev(h, n, N, m, h->d_E.as<double>(), n, nnz, dHi.as<int32_t>(), dHj.as<int32_t>(),
                                  dHs.as<double>(), Hshift ? dHshift.as<double>() : nullptr, dyo.as<double>(),
                                  dR.as<double>(), d01 ? dd01.as<double>() : nullptr, anamtype, inflation,
                                  maxCorrection ? dmaxc.as<double>() : nullptr, h->d_E.as<double>(), n,
                                  dxf.as<double>(), dxa.as<double>(), nullptr, stats);
  if (rc == 0) {
    if (n > 0 && stage && is_pageable(Ea)) {
      if ((rc = staged_copy(h, false, h->d_E.as<double>(), (size_t)n, Ea, (size_t)ldEa, (size_t)n, N, nthreads))) { cleanup(); return rc; }
    } else if (n > 0) e = cudaMemcpy2D(Ea, 8 * (size_t)ldEa, h->d_E.p, 8 * (size_t)n, 8 * (size_t)n, N, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && xf_out) e = cudaMemcpy(xf_out, dxf.p, 8 * (size_t)n, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && xa_out) e = cudaMemcpy(xa_out, dxa.p, 8 * (size_t)n, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { oak_set_error("assim_ensemble: D2H copy failed: %s", cudaGetErrorString(e)); rc = OAK_ERR_CUDA; }
    if (stats) { stats->h2d_bytes = 8ll * n * N + 16ll * nnz + 8ll * m * 4; stats->d2h_bytes = 8ll * n * N; }
  }
  cleanup();
  return rc;
}

// ---------------------------------------------------------------------------------------------------------
// Observation-operator generation: batched cinterp (ndgrid.F90:1183-1257) for one model grid
// ---------------------------------------------------------------------------------------------------------
extern "C" OAKB200_API int oakb200_cinterp_dev(oakb200_handle *h, int32_t ndim, const int32_t *gshape, const double *axes,
                                   const uint8_t *masked, int32_t m, const double *xi, int32_t *indexes, double *coeff,
                                   int32_t *nbp, int32_t *ndegenerate, void *stream) {
  if (!h || !gshape || !axes || (m > 0 && (!xi || !indexes || !coeff || !nbp))) { oak_set_error("cinterp: null argument"); return OAK_ERR_ARG; }
  if (ndim < 1 || ndim > 4) { oak_set_error("cinterp: %d dimensions (1 .. 4 supported)", ndim); return OAK_ERR_UNSUPPORTED; }
  DeviceGuard guard(h->device);
  int rc;
  if ((rc = h->d_tet.ensure(sizeof(double) * oak_cinterp_tet_doubles(ndim) + 64))) return rc;
  cudaStream_t st = stream ? (cudaStream_t)stream : h->slot[0].st;
  int *d_ndeg = reinterpret_cast<int *>(h->d_tet.as<double>() + oak_cinterp_tet_doubles(ndim));
  if ((rc = oak_launch_cinterp(st, ndim, gshape, axes, masked, h->d_tet.as<double>(), m, xi, indexes, coeff, nbp, d_ndeg))) return rc;
  int nd = 0;
  CUDA_TRY(cudaMemcpyAsync(&nd, d_ndeg, sizeof(int), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  if (ndegenerate) *ndegenerate = nd;
  if (nd > 0) {
    oak_set_error("cinterp: %d observation(s) fall in a degenerate cell (singleton dimension or |det| <= 1e-8): the SVD branch of "
                  "interp_tetrahedron (ndgrid.F90:527-627) is not implemented on the device; they are marked nbp = -1", nd);
    return OAK_ERR_UNSUPPORTED;
  }
  return 0;
}

extern "C" OAKB200_API int oakb200_cinterp(oakb200_handle *h, int32_t ndim, const int32_t *gshape, const double *axes,
                               const uint8_t *masked, int32_t m, const double *xi, int32_t *indexes, double *coeff,
                               int32_t *nbp, int32_t *ndegenerate) {
  if (!h || !gshape || !axes || (m > 0 && (!xi || !indexes || !coeff || !nbp))) { oak_set_error("cinterp: null argument"); return OAK_ERR_ARG; }
  if (ndim < 1 || ndim > 4) { oak_set_error("cinterp: %d dimensions (1 .. 4 supported)", ndim); return OAK_ERR_UNSUPPORTED; }
  DeviceGuard guard(h->device);
  size_t nax = 0, total = 1;
  for (int k = 0; k < ndim; k++) {
    if (gshape[k] < 1) { oak_set_error("cinterp: empty dimension"); return OAK_ERR_ARG; }
    nax += (size_t)gshape[k]; total *= (size_t)gshape[k];
  }
  const size_t twon = (size_t)1 << ndim, mb = (size_t)std::max(m, 1);
  DevBuf dax, dmask, dxi, didx, dco, dnbp;
  auto cleanup = [&]() { dax.release(); dmask.release(); dxi.release(); didx.release(); dco.release(); dnbp.release(); };
  int rc;
  if ((rc = dax.ensure(8 * nax)) || (masked && (rc = dmask.ensure(total))) || (rc = dxi.ensure(8 * mb * ndim)) ||
      (rc = didx.ensure(4 * mb * twon * ndim)) || (rc = dco.ensure(8 * mb * twon)) || (rc = dnbp.ensure(4 * mb))) { cleanup(); return rc; }
  cudaStream_t st = h->slot[0].st;
  cudaError_t e = cudaMemcpyAsync(dax.p, axes, 8 * nax, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess && masked) e = cudaMemcpyAsync(dmask.p, masked, total, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess && m > 0) e = cudaMemcpyAsync(dxi.p, xi, 8 * (size_t)m * ndim, cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) { oak_set_error("cinterp: H2D copy failed: %s", cudaGetErrorString(e)); cleanup(); return OAK_ERR_CUDA; }
  rc = oakb200_cinterp_dev(h, ndim, gshape, dax.as<double>(), masked ? dmask.as<uint8_t>() : nullptr, m, dxi.as<double>(),
                           didx.as<int32_t>(), dco.as<double>(), dnbp.as<int32_t>(), ndegenerate, (void *)st);
  if ((rc == 0 || rc == OAK_ERR_UNSUPPORTED) && m > 0) {
    e = cudaMemcpy(indexes, didx.p, 4 * (size_t)m * twon * ndim, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(coeff, dco.p, 8 * (size_t)m * twon, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(nbp, dnbp.p, 4 * (size_t)m, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { oak_set_error("cinterp: D2H copy failed: %s", cudaGetErrorString(e)); rc = OAK_ERR_CUDA; }
  }
  cleanup();
  return rc;
}

extern "C" OAKB200_API int oakb200_fp64_peak(oakb200_handle *h, int32_t mode, double *tflops) {
  if (!h || !tflops) { oak_set_error("fp64_peak: null argument"); return OAK_ERR_ARG; }
  DeviceGuard guard(h->device);
  return oak_fp64_peak(mode, tflops);
}
