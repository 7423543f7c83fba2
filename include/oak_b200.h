/*
 * oak_b200.h — C ABI of the B200-native local ensemble analysis for OAK.
 *
 * This is the drop-in boundary.  The Fortran driver keeps `Assim`
 * (assimilation.F90:2977) and replaces the body of its LocalScheme branch
 *     call locanalysis(zoneSize,selectObservations,xf,Hxf,yo,Sf,HSf,R,xa,Sa,locAmplitudes)
 * (assimilation.F90:3235-3236; signature rrsqrt.F90:433-457) by calls into this library through
 * ISO_C_BINDING (fortran/oak_b200_shim.F90).
 *
 * Conventions
 *   - every entry point returns 0 on success, <0 on error (oakb200_last_error() has the text);
 *     the reference prints to unit 0 and `call exit(1)` (ppdef.h:22) — the shim maps !=0 to ERROR_STOP.
 *   - all arrays are contiguous, column-major, fp64 / int32, owned by the caller; nothing is retained
 *     after return except what oakb200_set_zones / oakb200_set_observations copy to the device.
 *   - observation and zone indices in results are 1-based, as the Fortran caller stores them.
 *   - there is NO CPU fallback: every entry point needs a CUDA device (sm_100a).
 *   - a handle is not re-entrant; call from one thread (the `!$omp master` thread, between the
 *     barriers at assimilation.F90:3215 and :3295).
 */
#ifndef OAK_B200_H
#define OAK_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define OAKB200_API __attribute__((visibility("default")))
#else
#define OAKB200_API
#endif

typedef struct oakb200_handle oakb200_handle;

/* loctype (assimilation.F90:204-208), metrictype (:108-112) */
enum { OAKB200_LOC_HORIZONTAL = 1, OAKB200_LOC_DEPTH = 2, OAKB200_LOC_TIME = 3 };
enum { OAKB200_METRIC_CARTESIAN = 0, OAKB200_METRIC_SPHERICAL = 1, OAKB200_METRIC_SPHERICAL_APPROX = 2 };
/* weight function of the selectObservations callback:
 *   GAUSSIAN      relevant = d <= maxLen ; w = exp(-(d/corrLen)^2)       assimilation.F90:3756,:3767
 *   GASPARI_COHN  w = locfun(d/corrLen) ; relevant = w /= 0               test/test_rrsqrt.F90:254-271, covariance.F90:645-667
 *   UNIFORM       all observations, w = 1 (selectAllObservations)         test/test_rrsqrt.F90:239-247 */
enum { OAKB200_WEIGHT_GAUSSIAN = 0, OAKB200_WEIGHT_GASPARI_COHN = 1, OAKB200_WEIGHT_UNIFORM = 2 };

typedef struct {
  int64_t zones_total;      /* zones visited                                                     */
  int64_t zones_skipped;    /* zones without relevant observation (rrsqrt.F90:370-371)            */
  int64_t obs_relevant_sum; /* sum over zones of m_loc                                            */
  int64_t obs_candidate_sum;/* sum over zones of cell-grid candidates examined                    */
  int64_t jacobi_sweeps_sum;/* sum over analysed zones of Jacobi sweeps                           */
  int64_t h2d_bytes, d2h_bytes; /* bytes copied by the host-buffer entry points                   */
  double ms_total;          /* CUDA-event time of the whole call on the library's streams         */
  double ms_pack, ms_gram, ms_eig, ms_apply; /* per-kernel-family CUDA-event sums (only filled when
                                                option "profile" = 1, which serialises the batches) */
  int64_t launches;         /* kernels launched by this call                                      */
  int64_t zones_fallback;   /* zones the tridiagonal transform route re-did with the Jacobi kernel */
  double ms_tridiag, ms_tql, ms_tvec; /* "profile": the three kernels of the tridiagonal route (ms_eig = their
                                         sum + the fallback launch)                                        */
} oakb200_stats;

OAKB200_API const char *oakb200_last_error(void);
OAKB200_API int oakb200_version(void);

/* Create / destroy a context on CUDA device `device` (one handle per GPU; one process per GPU
 * replaces the MPI ranks of parall.F90:53-186). */
OAKB200_API int oakb200_create(int device, oakb200_handle **h);
OAKB200_API int oakb200_destroy(oakb200_handle *h);

/* Options (all optional; a key the library does not know is an error):
 *   "scheme"          oakb200_assim_ensemble[_dev]: 1 = local scheme (default), 0 = global scheme (schemetype of the init
 *                     file, assimilation.F90:292,:3227; zones and observation positions are then not needed)
 *   "eig_kernel"      4 = Householder tridiagonalisation + QL + twisted factorisation (default for N <= 64;
 *                     flagged zones fall back to 0), 0 = register-resident block Jacobi (N > 64), 1 = simple
 *                     shared-memory Jacobi (cross-check)
 *   "gram_kernel"     1 = fp64 tensor-core tiles (mma.m8n8k4, 4 warps per zone; default since the round-2 timing:
 *                     8.6 -> 5.7 ms per 90 k zones), 2 = 2 warps per zone, 3 / 4 = the same with chunks of 32 instead of
 *                     64 candidates (half the shared memory), 0 = DFMA register tiles; padded ensemble size 64 only,
 *                     other sizes use 0
 *   "apply_kernel"    0 = DFMA register tiles (default), 1 = fp64 tensor-core tiles (padded ensemble size 64 only;
 *                     zones the fused transform kernel leaves over, zones with many rows, the global scheme)
 *   "fuse_apply"      route 4: 1 = the transform kernel updates the zone rows itself from the factored transform
 *                     (no T written, k_apply only for the zones it did not finish); used while every zone has at most
 *                     N (padded) rows, where the factored form is the cheaper one. Default 0
 *   "tvec_split"      route 4: 1 = the eigenvector kernel runs as two kernels (vectors of T with few registers and
 *                     high occupancy | back-transformation and the rest), 32 KB more workspace per zone. Default 0
 *   "tql_side"        route 4: 1 (default) = the QL eigenvalue kernel runs on a high-priority side stream of its slot
 *   "localise_obs"    1 (default) = locAnalysis' default branch; 0 = localise_obs=.false. (rrsqrt.F90:374-385): a zone with
 *                     at least one relevant observation is analysed with ALL observations (their weights, no cut-off)
 *                     and amplitudes(:,zone) is returned
 *   "host_stage"      host-buffer entry points, PAGEABLE caller arrays (ordinary Fortran allocatables): 1 = every chunk of the
 *                     state goes through two page-locked staging buffers of its stream slot, copied in and out by
 *                     "stage_threads" host threads (default min(16, hardware threads)) while the other slots compute;
 *                     0 = asynchronous copies straight from / to the caller's arrays (staged by the driver, one thread);
 *                     -1 (default) = 1 for calls of 32 MB and more.  Measured on C3 end to end: 1.18 M columns/s against
 *                     0.40 M (pinned caller arrays: 2.6 M).  Results are identical.  Host memory: two page-locked buffers of one
 *                     chunk ("chunk_mb", 256 MB) per stream slot, 2 GB in total, allocated on first use and kept by the handle
 *   "host_register"   host-buffer entry points: 1 = pageable caller arrays are page-locked (cudaHostRegister) for the
 *                     duration of the call; 0 (default) = left as they are (the driver stages the copies).  Measured on
 *                     C3: pinned arrays (oakb200_host_alloc) 2.68 M columns/s, pageable 0.40 M, registered per call 0.09 M
 *   "apply_tma"       1 (default) = where every zone has the same even number of rows <= 32 (water columns) and the
 *                     leading dimensions are even, the apply kernel stages the zone's rows with 2-D tensor copies (TMA:
 *                     cp.async.bulk.tensor.2d in and out); 0 = always the plain kernel
 *   "ens_fuse"        oakb200_assim_ensemble[_dev], local scheme: 1 = the prologue (forward anamorphosis, mean, anomalies) and the
 *                     epilogue (inflation, saturation, Ea, inverse anamorphosis, mean) run inside the apply kernel on the
 *                     staged rows: E is read once and Ea written once; 0 = three passes (k_mean_anom, analysis in place,
 *                     k_epilogue); -1 (default) = fused when there is no anamorphosis (measured on C5: log / exp inside the
 *                     fp64-bound apply kernel cost more than the saved passes).  Bit-identical results either way
 *   "push_kernel", "push_ctas"   fused gather: rows pushed by a small kernel (SM stores / multimem stores) on a side stream
 *                     instead of copy-engine copies (default 1), its number of CTAs (24)
 *   "tri_orthtol"     route 4: accepted loss of orthogonality between neighbouring eigenvectors (default 1e-11)
 *   "tri_maxgroup"    route 4: largest group of close eigenvalues orthogonalised in place (default 6; 0 sends
 *                     every zone with a close pair to the Jacobi kernel)
 *   "jacobi_tol", "max_sweeps", "fixed_sweeps"   stopping rule of the Jacobi kernels
 *   "zones_per_batch" zones per kernel batch (default: from the workspace budget), "pad_to" padded ensemble size
 *   "chunk_mb"        host-buffer entry points: size of the state chunks streamed through the device (256)
 *   "profile"         1: batches serialised, CUDA-event time per kernel family in the statistics
 *   "async", "order_after_caller", "stream_priority"   see oakb200_synchronize
 *   "peer_mode"       see oakb200_set_peer_outputs: 1 copy engines (default), 0 stores of the apply kernel
 *   "push_pieces"     peer_mode 1: the apply of a batch runs in this many launches, each pushed to the peers as soon
 *                     as it is done (shorter exposed tail at the end of a call); default 1 */
OAKB200_API int oakb200_set_option(oakb200_handle *h, const char *key, double value);

/* Page-locked host memory for the caller's large arrays (Sf / Sa, HSf of assimilation.F90:2987-3000): with it the
 * host-buffer entry points overlap the chunked copies with the kernels.  Fortran: c_f_pointer on the returned pointer. */
OAKB200_API int oakb200_host_alloc(int64_t bytes, void **ptr);
OAKB200_API int oakb200_host_free(void *ptr);

/* Zones = the partition of the (zone-permuted) state vector (assimilation.F90:578-641).
 * zoneSize[nzones] as passed to locanalysis; zone z owns rows sum(zoneSize[0..z-1]) ... of the state.
 * zx,zy,zz,zt: coordinate of each zone's FIRST element (rrsqrt.F90:368, assimilation.F90:3713-3740);
 * zy/zz/zt may be NULL when unused.  corrLen/maxLen: hCorrLengthToObs / hMaxCorrLengthToObs of that
 * element (assimilation.F90:3756,:3767).  Replaces the module globals the callback reads. */
OAKB200_API int oakb200_set_zones(oakb200_handle *h, int32_t nzones, const int32_t *zoneSize, const double *zx,
                      const double *zy, const double *zz, const double *zt, const double *corrLen,
                      const double *maxLen, int32_t loctype, int32_t metrictype, int32_t weightfun);

/* Observation positions obsGridX/Y/Z/T(m) (assimilation.F90:3155-3160); builds the device cell grid
 * that replaces the O(m) scan of assimilation.F90:3745-3757 and the cellgrid/near search of
 * ndgrid.F90:1489-1691.  Arrays not needed by `loctype` may be NULL.  Call after oakb200_set_zones. */
OAKB200_API int oakb200_set_observations(oakb200_handle *h, int32_t m, const double *obsx, const double *obsy,
                             const double *obsz, const double *obst);

/* The selectObservations callback for zones [zone_first, zone_first+zone_count) (0-based zone
 * numbers): CSR output offsets[zone_count+1], idx[] = 1-based observation numbers in increasing
 * order (the order pack() uses, rrsqrt.F90:395-404), weight[] their weights.  If capacity is too
 * small returns -5 with offsets[] filled (offsets[zone_count] = required capacity). */
OAKB200_API int oakb200_select_observations(oakb200_handle *h, int32_t zone_first, int32_t zone_count,
                                int64_t capacity, int64_t *offsets, int32_t *idx, double *weight);

/* locAnalysis (rrsqrt.F90:433-466) with R = DiagCovar(Rdiag) optionally wrapped by
 * DCDCovar(d01,.) (assimilation.F90:3086-3092; d01 may be NULL), HOST buffers:
 *   xf[n], Hxf[m], yo[m], Sf[n x N, ld ldSf], HSf[m x N, ld ldHSf]  ->  xa[n], Sa[n x N, ld ldSa]
 *   amplitudes[N x nzones] (may be NULL) is zero-filled, as the reference leaves it on the default
 *   local_obs branch (rrsqrt.F90:324,:386-412).
 * Sa may alias Sf (in-place update).  The state is streamed through the device in zone chunks, so
 * n*N may exceed device memory.  Returns -7 if an analysis produced NaN (rrsqrt.F90:145-149). */
OAKB200_API int oakb200_local_analysis(oakb200_handle *h, int64_t n, int32_t N, int32_t m, const double *xf,
                           const double *Hxf, const double *yo, const double *Sf, int64_t ldSf,
                           const double *HSf, int64_t ldHSf, const double *Rdiag, const double *d01,
                           double *xa, double *Sa, int64_t ldSa, double *amplitudes,
                           oakb200_stats *stats);

/* Same, all array arguments are DEVICE pointers on the handle's device (state resident in HBM).
 * The work is ordered after everything already enqueued on `stream` (a cudaStream_t passed as
 * void*, NULL = the legacy default stream), runs on the library's own streams, and the call returns
 * only when the results are complete (it synchronises), so they are visible to any stream. */
OAKB200_API int oakb200_local_analysis_dev(oakb200_handle *h, int64_t n, int32_t N, int32_t m, const double *xf,
                               const double *Hxf, const double *yo, const double *Sf, int64_t ldSf,
                               const double *HSf, int64_t ldHSf, const double *Rdiag,
                               const double *d01, double *xa, double *Sa, int64_t ldSa,
                               double *amplitudes, void *stream, oakb200_stats *stats);

/* The global scheme (schemetype = 0, assimilation.F90:292): `analysis` (rrsqrt.F90:196-208), i.e. analysisIncrement
 * (rrsqrt.F90:100-190) with every observation and weight 1 — one N x N Gram matrix over the m observations, one
 * transform, applied to all n rows.  Needs neither zones nor observation positions.  Same arguments as the local
 * entry points; amplitudes[N] (may be NULL) receives ampl (rrsqrt.F90:142,:153), which `analysis` returns.
 * HOST buffers (the state is streamed through the device in chunks) / DEVICE pointers. */
OAKB200_API int oakb200_global_analysis(oakb200_handle *h, int64_t n, int32_t N, int32_t m, const double *xf,
                            const double *Hxf, const double *yo, const double *Sf, int64_t ldSf,
                            const double *HSf, int64_t ldHSf, const double *Rdiag, const double *d01,
                            double *xa, double *Sa, int64_t ldSa, double *amplitudes, oakb200_stats *stats);
OAKB200_API int oakb200_global_analysis_dev(oakb200_handle *h, int64_t n, int32_t N, int32_t m, const double *xf,
                                const double *Hxf, const double *yo, const double *Sf, int64_t ldSf,
                                const double *HSf, int64_t ldHSf, const double *Rdiag, const double *d01,
                                double *xa, double *Sa, int64_t ldSa, double *amplitudes, void *stream,
                                oakb200_stats *stats);

/* Asynchronous use of oakb200_local_analysis_dev (option "async" = 1): the call returns after enqueueing,
 * `stream` waits for the result (stream-ordered consumers, e.g. an NCCL all-gather, need no host
 * synchronisation); status (NaN, convergence) and statistics are collected here.  One outstanding call per
 * handle.  Related options: "order_after_caller" = 0 (inputs are already complete: do not wait for `stream`
 * before starting), "stream_priority" (CUDA priority of the library's streams, 0 .. negative = urgent). */
OAKB200_API int oakb200_synchronize(oakb200_handle *h, oakb200_stats *stats);

/* Fused all-gather (multi-GPU, one process per GPU; replaces parallGather, parall.F90:507-566): in addition to
 * Sa / xa, the kernel that applies the transform stores every row of this rank's slab into up to
 * OAKB200_MAX_PEERS result arrays that live on other GPUs (device pointers valid on this device: CUDA-IPC
 * mappings of the other ranks' arrays, written over NVLink), so that no separate collective is needed:
 *     Sa_peer[d][(row0 + i) + ld_peer * k] ,  xa_peer[d][row0 + i]      local row i, member k
 * npeer = 0 switches it off.  Only the device-pointer entry point uses it.  The caller orders the readers
 * after the writers of all ranks (e.g. a one-element all-reduce enqueued behind the analysis). */
#define OAKB200_MAX_PEERS 8
OAKB200_API int oakb200_set_peer_outputs(oakb200_handle *h, int32_t npeer, double *const *Sa_peer,
                             double *const *xa_peer, int64_t ld_peer, int64_t row0);
/* Device memory that other processes of the same node can map (cudaIpc): allocate + export the 64-byte handle,
 * open a peer's handle (lazy peer access over NVLink), close, free. */
OAKB200_API int oakb200_ipc_alloc(oakb200_handle *h, int64_t bytes, void **ptr, unsigned char handle[64]);
OAKB200_API int oakb200_ipc_open(oakb200_handle *h, const unsigned char handle[64], void **ptr);
OAKB200_API int oakb200_ipc_close(oakb200_handle *h, void *ptr);
OAKB200_API int oakb200_ipc_free(oakb200_handle *h, void *ptr);

/* The same gather through NVSwitch multicast: Sa_mc / xa_mc are the MULTICAST addresses of the result arrays (every rank
 * has bound its own array at the same offsets of one multicast object: cuMulticastCreate / cuMulticastAddDevice /
 * cuMulticastBindMem, done by the caller, see oak_b200.dist.MulticastResult).  Every finished batch is then stored once
 * with multimem.st and replicated by the switch into all ranks' arrays (1/world of the NVLink egress of the peer
 * flavour).  NULL, NULL switches it off.  Takes precedence over oakb200_set_peer_outputs. */
OAKB200_API int oakb200_set_multicast_output(oakb200_handle *h, double *Sa_mc, double *xa_mc, int64_t ld, int64_t row0);

/* Table of the tabulated anamorphosis (type 3): AnamTrans%anam(v)%transform, K x 2 column-major HOST array,
 * column 1 = physical values, column 2 = transformed values (assimilation.F90:4539-4567, interp1
 * anamorphosis.F90:304-339 incl. its clamping rule).  One table for the whole state vector.  K = 0 clears it. */
OAKB200_API int oakb200_set_anamorphosis_table(oakb200_handle *h, int32_t K, const double *table);

/* Per-variable anamorphosis, as anamtransform applies it (assimilation.F90:4531-4567: the variable v of every element
 * comes from ind2submv, the transform is AnamTrans%anam(v)): vtype[nvar] (1 identity, 2 log, 3 tabulated), vK[nvar]
 * rows of each tabulated variable's table (ignored otherwise), tables = those K_v x 2 column-major tables one after
 * the other (HOST), rowvar[n] = 1-based variable number of every row of the zone-permuted state.  Selected by
 * anamtype = 0 in oakb200_assim_ensemble[_dev].  nvar = 0 clears it. */
OAKB200_API int oakb200_set_anamorphosis_vars(oakb200_handle *h, int32_t nvar, const int32_t *vtype, const int32_t *vK,
                                  const double *tables, int64_t n, const int32_t *rowvar);

/* Ensemble branch of Assim around the local scheme (assimilation.F90:3083,:3106-3134 prologue,
 * :3235 analysis, :3301-3357,:3558-3562 epilogue), HOST buffers:
 *   E[n x N] ensemble (zone-permuted), H as COO (Hi,Hj 1-based int32, Hs, nnz; matoper.F90:30-39),
 *   Hshift[m] (may be NULL), yo, Rdiag, d01, anamtype 0 per variable (oakb200_set_anamorphosis_vars) / 1 identity / 2 log / 3 tabulated (anamorphosis.F90:78-120,
 *   :304-339; the table comes from oakb200_set_anamorphosis_table),
 *   inflation (inflation.mult), maxCorrection[n] (may be NULL)  ->  Ea[n x N]; xf_out/xa_out[n] optional:
 *   xf_out = forecast mean in transformed space (:3127), xa_out = mean of the back-transformed analysis ensemble
 *   (xa = sum(Sa,2)/size(Sa,2), :3343-3349). */
OAKB200_API int oakb200_assim_ensemble(oakb200_handle *h, int64_t n, int32_t N, int32_t m, const double *E,
                           int64_t ldE, int64_t nnz, const int32_t *Hi, const int32_t *Hj,
                           const double *Hs, const double *Hshift, const double *yo,
                           const double *Rdiag, const double *d01, int32_t anamtype,
                           double inflation, const double *maxCorrection, double *Ea, int64_t ldEa,
                           double *xf_out, double *xa_out, oakb200_stats *stats);
/* Same with DEVICE pointers. */
OAKB200_API int oakb200_assim_ensemble_dev(oakb200_handle *h, int64_t n, int32_t N, int32_t m, const double *E,
                               int64_t ldE, int64_t nnz, const int32_t *Hi, const int32_t *Hj,
                               const double *Hs, const double *Hshift, const double *yo,
                               const double *Rdiag, const double *d01, int32_t anamtype,
                               double inflation, const double *maxCorrection, double *Ea,
                               int64_t ldEa, double *xf_out, double *xa_out, void *stream,
                               oakb200_stats *stats);

/* Contiguous zone ranges per rank, the formula of parallPartion (parall.F90:166-186) with unit
 * speeds: rank p of P owns zones [first[p], first[p+1]) ; first has P+1 entries (0-based). */
OAKB200_API int oakb200_partition_zones(int32_t nzones, int32_t nranks, int32_t *first);

/* mloc[nzones]: the number of relevant observations per zone as counted by the production kernel of the LAST analysis
 * (the selection fused into the Gram kernel); cross-check of oakb200_select_observations, which evaluates the same
 * predicate in a kernel of its own. */
OAKB200_API int oakb200_zone_counts(oakb200_handle *h, int32_t *mloc);

/* Observation-operator generation: the batched form of `cinterp`
 * (ndgrid.F90:1183-1257), the arithmetic inside genObservationOper (assimilation.F90:2471-2656, the calls at
 * :2569-2585): for each of the m observation positions xi[m][ndim] (row-major, one row per observation) locate the cell
 * of the model grid that contains it and return the interpolation weights on the 2^ndim corners of that cell, from the
 * first of the n! 2^(n-1) simplices of the cell (split, ndgrid.F90:357-435) that contains the point
 * (interp_cube / interp_tetrahedron, :464-665).
 *   gshape[ndim]            shape of the variable's grid (ndim 1 .. 4)
 *   axes                    the coordinate axes one after the other (gshape[0] + ... + gshape[ndim-1] values): grids whose
 *                           coordinate k depends on subscript k only (regular / rectilinear; ascending or descending)
 *   masked[prod(gshape)]    1 = land / invalid point (first subscript fastest), or NULL
 *   indexes[m][2^ndim][ndim] 1-based subscripts of the corners (corner j has the upper node in dimension k when bit k of
 *                           j is set), coeff[m][2^ndim], nbp[m] = 2^ndim, or 0 when the point is outside the grid or a
 *                           corner is masked (genObservationOper then writes its zero row with index -1, :2597-2611)
 * Cells with a degenerate simplex (singleton dimension, |det| <= 1e-8) take an SVD branch in the reference
 * (:527-627) that is not implemented on the device: they get nbp = -1, *ndegenerate counts them and the call returns
 * OAKB200 error -6 (the other observations are valid).  The name matching of variables and the choice of the finest
 * grid (hres, :2545-2594) stay in the Fortran driver, which calls this once per model variable. */
OAKB200_API int oakb200_cinterp(oakb200_handle *h, int32_t ndim, const int32_t *gshape, const double *axes,
                    const uint8_t *masked, int32_t m, const double *xi, int32_t *indexes, double *coeff,
                    int32_t *nbp, int32_t *ndegenerate);
/* Same with DEVICE pointers for axes, masked, xi, indexes, coeff, nbp (gshape and ndegenerate on the host). */
OAKB200_API int oakb200_cinterp_dev(oakb200_handle *h, int32_t ndim, const int32_t *gshape, const double *axes,
                        const uint8_t *masked, int32_t m, const double *xi, int32_t *indexes, double *coeff,
                        int32_t *nbp, int32_t *ndegenerate, void *stream);

/* Measured fp64 pipe peak on this device, used as the roofline denominator:
 * mode 0 = DFMA (register-resident FMA chains), 1 = DMMA (mma.sync.m8n8k4.f64). TFLOP/s. */
OAKB200_API int oakb200_fp64_peak(oakb200_handle *h, int32_t mode, double *tflops);

#ifdef __cplusplus
}
#endif
#endif /* OAK_B200_H */
