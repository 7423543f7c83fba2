// gram_mma.cu — selection + weighted Gram matrix of a zone on the fp64 tensor-core path (option
// "gram_kernel" = 1 or 2; NP = 64).  Same contract and the same selection pipeline as k_gram (gram.cu):
//     G = sum_{l in L} coef_l a_l a_l^T ,  c = sum_{l in L} coef_l delta_l a_l       rrsqrt.F90:135,:142
// Why: ncu shows k_gram limited by the shared-memory return path (55-66 %) with the fp64 pipe at 43-51 %
// (profiles/r1_ncu_full_gram.txt): a register tile gets 2.67 FMA out of every double it reads.  One
// mma.sync.m8n8k4.f64 does 256 FMA from two doubles per lane, and the DMMA rate of a B200 equals its DFMA rate
// (37.1 vs 36.9 TFLOP/s, microbench.cu), so the same pipe is fed with 1/6 of the shared-memory traffic and 1/8 of
// the issue slots.
//
// Tiles: G is cut into 8 x 8 blocks of 8 x 8; only the 36 blocks on or below the diagonal are accumulated
// (56 % of the full product; k_gram's three 32 x 32 blocks are 75 %) and mirrored on store.  For a k-step of four
// staged rows k0..k0+3, lane 4g+t reads v_b = row[k0+t][8b+g] for the block indices b its warp needs: that one
// value is the B fragment of block column b and, times coef[k0+t], the A fragment of block row b.
//   NW = 4 warps: LL triangle (10 tiles) | HH triangle (10) | HL rows 4,5 (8) | HL rows 6,7 (8)
//   NW = 2 warps: LL + HL rows 4,5 (18 tiles) | HH + HL rows 6,7 (18)
// Staged rows have a stride of NP + 4 doubles: the 16 lanes of a half-warp then read 16 different 8-byte banks.
// The rows the two evaluating warps found are merged into one dense list when they are staged; the rows that pad
// the last k-step of a chunk enter with coef = 0, and the row buffers are zeroed once so that 0 x stale is 0.
#include "common.cuh"

namespace {

constexpr int GRAM_CHMAX = 64; // candidates examined per chunk: 64 (warps 0 and 1, one per lane) or 32 (warp 0)
constexpr int GRAM_MAXR = 64;  // cell ranges per row group

// ---- static tile map (NB = 8 blocks per side, HB = 4) ----
template <int NW, int W>
struct TileMap {
  static constexpr int HB = 4;
  // which of the 36 lower tiles warp W owns
  static __host__ __device__ constexpr bool owns(int bi, int bj) {
    if (bj > bi) return false;
    const bool LL = bi < HB, HH = bj >= HB, HL = !LL && !HH;
    if (NW == 4) return W == 0 ? LL : (W == 1 ? HH : (W == 2 ? (HL && bi < HB + 2) : (HL && bi >= HB + 2)));
    return W == 0 ? (LL || (HL && bi < HB + 2)) : (HH || (HL && bi >= HB + 2));
  }
  // dense accumulator slot of an owned tile
  static __host__ __device__ constexpr int slot(int bi, int bj) {
    const bool LL = bi < HB, HH = bj >= HB;
    if (LL) return bi * (bi + 1) / 2 + bj;
    if (HH) return (bi - HB) * (bi - HB + 1) / 2 + (bj - HB);
    const int hl = ((bi - HB) & 1) * HB + bj;   // rows (4,5) or (6,7) x columns 0..3
    return NW == 4 ? hl : 10 + hl;
  }
  static constexpr int NACC = NW == 4 ? 10 : 18;
  static __host__ __device__ constexpr bool needs_row(int b) {
    for (int j = 0; j <= b; j++) if (owns(b, j)) return true;
    return false;
  }
  static __host__ __device__ constexpr bool needs_col(int b) {
    for (int i = b; i < 2 * HB; i++) if (owns(i, b)) return true;
    return false;
  }
};

#ifndef GRAM_C4
#define GRAM_C4 0   // 1: c = sum coef delta a in four partial sums: measured SLOWER (k_gram_mma 59.0 -> 63.0 ms per C3 step)
#endif
#ifndef GRAM_EVAL_PREFETCH
#define GRAM_EVAL_PREFETCH 1
#endif

template <int NP, int NW, int W>
struct GramMma {
  using TM = TileMap<NW, W>;
  static constexpr int NB = NP / 8, LDR = NP + 4;

  // the `total` staged rows of a chunk; the coefficient of merged row q sits in list slot q (q < cnt0, found by
  // warp 0) or 32 + q - cnt0 (found by warp 1); the rows that pad the last k-step enter with coefficient 0
  static __device__ __forceinline__ void chunk(double (&acc)[TM::NACC][2], const double *rows, const double *coef,
                                               int cnt0, int total, int g, int t) {
#pragma unroll 2
    for (int k0 = 0; k0 < total; k0 += 4) {
      const int qr = k0 + t;
      const double *r = rows + qr * LDR + g;
      const double cf = qr < total ? coef[qr < cnt0 ? qr : 32 + qr - cnt0] : 0.;
      double v[NB];
#pragma unroll
      for (int b = 0; b < NB; b++)
        if (TM::needs_row(b) || TM::needs_col(b)) v[b] = r[8 * b];
#pragma unroll
      for (int bi = 0; bi < NB; bi++) {
        if (!TM::needs_row(bi)) continue;
        const double a = v[bi] * cf;
#pragma unroll
        for (int bj = 0; bj <= bi; bj++)
          if (TM::owns(bi, bj)) oak_dmma_m8n8k4(acc[TM::slot(bi, bj)][0], acc[TM::slot(bi, bj)][1], a, v[bj]);
      }
    }
  }

  static __device__ __forceinline__ void store(const double (&acc)[TM::NACC][2], double *Gz, int g, int t) {
#pragma unroll
    for (int bi = 0; bi < NB; bi++)
#pragma unroll
      for (int bj = 0; bj <= bi; bj++)
        if (TM::owns(bi, bj)) {
          const int i = 8 * bi + g, j = 8 * bj + 2 * t;
          const double c0 = acc[TM::slot(bi, bj)][0], c1 = acc[TM::slot(bi, bj)][1];
          Gz[i + NP * j] = c0;
          Gz[i + NP * (j + 1)] = c1;
          if (bi != bj) {  // mirror the strictly lower tiles
            Gz[j + NP * i] = c0;
            Gz[j + 1 + NP * i] = c1;
          }
        }
  }
};

// CH = 64: two evaluating warps, 2 x 64 staged rows (70 KB: 3 CTAs per SM); CH = 32: one evaluating warp, 2 x 32
// rows (35 KB: 6 CTAs per SM, twice the barriers per zone)
template <int NP, int NW, int CH>
__global__ void __launch_bounds__(32 * NW) k_gram_mma(ZoneGeom zg, ObsGrid og, ObsRows orows, int zone0, int nz,
                                                      double *__restrict__ G, double *__restrict__ cvec,
                                                      int32_t *__restrict__ mloc, DevCounters *ctr) {
  static_assert(NP == 64 && (NW == 2 || NW == 4), "tile map is written for 8 x 8 blocks");
  static_assert(CH == 32 || CH == 64, "one or two evaluating warps");
  constexpr int GRAM_CH = CH;
  constexpr int NT = 32 * NW, LDR = NP + 4;
  constexpr int NACC = NW == 4 ? 10 : 18;
  extern __shared__ __align__(16) double rowbuf[];  // [2][GRAM_CH][LDR]
  __shared__ double s_coef[3][GRAM_CHMAX], s_cd[3][GRAM_CHMAX];
  __shared__ int s_pos[3][GRAM_CHMAX];
  __shared__ int s_cnt[3][2];
  __shared__ int s_rstart[GRAM_MAXR], s_rlen[GRAM_MAXR];
  __shared__ int s_total;
  __shared__ int s_true;   // localise_obs = .false.: observations that pass the relevance predicate

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  // c = sum coef delta a: element ci per thread; with four warps it is the job of warps 2 and 3, which own 8 tiles
  // each against the 10 of warps 0 and 1 (which also evaluate the predicate)
  const int ci = NW == 4 ? tid - 64 : tid;
  const int zl = blockIdx.x;
  if (zl >= nz) return;
  const int zone = zone0 + zl;
  const ZoneQuery q = oak_zone_query(zg, zone);
  const CellBox box = oak_zone_box(og, q);

  double acc[NACC][2];
#pragma unroll
  for (int a = 0; a < NACC; a++) acc[a][0] = acc[a][1] = 0.;
  double cacc = 0., cacc1 = 0., cacc2 = 0., cacc3 = 0.;
  int nrel_total = 0;
  long long ncand_total = 0;
  if (threadIdx.x == 0) s_true = 0;   // ordered before its first use by the barrier at the head of the cell loop
  for (int i = tid; i < 2 * GRAM_CH * LDR; i += NT) rowbuf[i] = 0.;  // 0 x (never written) must be 0, not NaN

  // E(c): warps 0 and 1, 32 candidates each, into list buffer lb (as k_gram), in two halves: the loads of the candidate's
  // position, coefficient and innovation (global memory, scattered) and the predicate with the list update.  Inside the
  // chunk loop the loads of chunk c+2 are issued BEFORE the tensor-core burst of chunk c and consumed after it
  // (GRAM_EVAL_PREFETCH), so their latency is not waited for between the burst and the barrier.
  int pf_p = 0;
  bool pf_valid = false;
  double pf_sx = 0., pf_sy = 0., pf_sc = 0., pf_dl = 0.;
  auto eval_load = [&](int c, int total) {
    if (warp < GRAM_CH / 32) {
      int qq = c * GRAM_CH + warp * 32 + lane;
      pf_valid = qq < total;
      if (pf_valid) {
        int r = 0;
        while (qq >= s_rlen[r]) { qq -= s_rlen[r]; r++; }
        pf_p = s_rstart[r] + qq;
        pf_sx = og.sx[pf_p]; pf_sy = og.sy[pf_p];
        pf_sc = orows.scoef[pf_p]; pf_dl = orows.delta[pf_p];
      }
    }
  };
  auto eval_finish = [&](int lb) {
    if (warp < GRAM_CH / 32) {
      bool rel = false;
      double w = 0.;
      if (pf_valid) {
        rel = oak_obs_relevant(q, pf_sx, pf_sy, w);
        if (q.noloc) {   // localise_obs = .false.: count the relevant ones, take them all
          if (rel) atomicAdd(&s_true, 1);
          rel = true;
        }
      }
      const unsigned bal = __ballot_sync(0xffffffffu, rel);
      const int cnt = __popc(bal);
      if (rel) {
        const int slot = warp * 32 + __popc(bal & ((1u << lane) - 1u));
        const double coef = (w * w) * pf_sc;
        s_pos[lb][slot] = pf_p;
        s_coef[lb][slot] = coef;
        s_cd[lb][slot] = coef * pf_dl;
      }
      if (lane == 0) {
        s_cnt[lb][warp] = cnt;
        if (GRAM_CH == 32) s_cnt[lb][1] = 0;
      }
    }
  };
  auto eval = [&](int c, int lb, int total) { eval_load(c, total); eval_finish(lb); };
  // L(c): rows of list buffer lb into row buffer rb (one 16-byte cp.async per lane and row); the two warps'
  // finds are merged here: row buffer slot = r (warp 0's) or cnt0 + r (warp 1's), so the k-steps of four rows
  // run over one dense list
  auto load_rows = [&](int lb, int rb) {
    double *dstb = rowbuf + (size_t)rb * GRAM_CH * LDR;
    const int cnt0 = s_cnt[lb][0];
#pragma unroll
    for (int seg = 0; seg < 2; seg++) {
      const int cnt = s_cnt[lb][seg];
      for (int r = warp; r < cnt; r += NW) {
        const double *src = orows.rows + (int64_t)s_pos[lb][seg * 32 + r] * NP;
        double *dst = dstb + (seg * cnt0 + r) * LDR;
        for (int cidx = lane * 2; cidx < NP; cidx += 64) cp_async16(dst + cidx, src + cidx);
      }
    }
  };
  // F(c)
  auto fma_rows = [&](int lb, int rb) {
    const double *rows = rowbuf + (size_t)rb * GRAM_CH * LDR;
    const double *coef = s_coef[lb];
    const int cnt0 = s_cnt[lb][0], total = cnt0 + s_cnt[lb][1];
    nrel_total += total;
    if constexpr (NW == 4) {
      switch (warp) {
        case 0: GramMma<NP, NW, 0>::chunk(acc, rows, coef, cnt0, total, g, t); break;
        case 1: GramMma<NP, NW, 1>::chunk(acc, rows, coef, cnt0, total, g, t); break;
        case 2: GramMma<NP, NW, 2>::chunk(acc, rows, coef, cnt0, total, g, t); break;
        default: GramMma<NP, NW, 3>::chunk(acc, rows, coef, cnt0, total, g, t); break;
      }
    } else {
      if (warp == 0) GramMma<NP, NW, 0>::chunk(acc, rows, coef, cnt0, total, g, t);
      else GramMma<NP, NW, 1>::chunk(acc, rows, coef, cnt0, total, g, t);
    }
    if (ci >= 0) {
      const double *cd = s_cd[lb];
      // the two evaluating warps' finds one after the other (same order as the merged list, no index select per row)
      const double *rp = rows + ci;
      const int cnt1 = total - cnt0;
#if GRAM_C4
      // four partial sums instead of one chain of up to 64 dependent FMAs per chunk (experiment, off: see GRAM_C4)
      int r = 0;
      for (; r + 3 < cnt0; r += 4, rp += 4 * LDR) {
        cacc = fma(cd[r], rp[0], cacc); cacc1 = fma(cd[r + 1], rp[LDR], cacc1);
        cacc2 = fma(cd[r + 2], rp[2 * LDR], cacc2); cacc3 = fma(cd[r + 3], rp[3 * LDR], cacc3);
      }
      for (; r < cnt0; r++, rp += LDR) cacc = fma(cd[r], *rp, cacc);
      r = 0;
      for (; r + 3 < cnt1; r += 4, rp += 4 * LDR) {
        cacc = fma(cd[32 + r], rp[0], cacc); cacc1 = fma(cd[33 + r], rp[LDR], cacc1);
        cacc2 = fma(cd[34 + r], rp[2 * LDR], cacc2); cacc3 = fma(cd[35 + r], rp[3 * LDR], cacc3);
      }
      for (; r < cnt1; r++, rp += LDR) cacc = fma(cd[32 + r], *rp, cacc);
#else
#pragma unroll 4
      for (int r = 0; r < cnt0; r++, rp += LDR) cacc = fma(cd[r], *rp, cacc);
#pragma unroll 4
      for (int r = 0; r < cnt1; r++, rp += LDR) cacc = fma(cd[32 + r], *rp, cacc);
#endif
    }
  };

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
      int tt = 0;
      for (int r = 0; r < GRAM_MAXR; r++) tt += s_rlen[r];
      s_total = tt;
    }
    __syncthreads();
    const int total = s_total;
    ncand_total += total;
    const int nchunk = (total + GRAM_CH - 1) / GRAM_CH;
    if (nchunk == 0) continue;

    eval(0, 0, total);
    __syncthreads();
    load_rows(0, 0);
    cp_async_commit();
    if (nchunk > 1) eval(1, 1, total);
    __syncthreads();
    for (int c = 0; c < nchunk; c++) {
      const int lb = c % 3, rb = c & 1;
      if (c + 1 < nchunk) {
        load_rows((c + 1) % 3, rb ^ 1);
        cp_async_commit();
        cp_async_wait<1>();
      } else {
        cp_async_wait<0>();
      }
      __syncthreads();  // rows of chunk c have landed for every thread
#if GRAM_EVAL_PREFETCH
      if (c + 2 < nchunk) eval_load(c + 2, total);
      fma_rows(lb, rb);
      if (c + 2 < nchunk) eval_finish((c + 2) % 3);
#else
      fma_rows(lb, rb);
      if (c + 2 < nchunk) eval(c + 2, (c + 2) % 3, total);
#endif
      __syncthreads();  // list c+2 visible; row buffer rb free for chunk c+2
    }
  }

  double *Gz = G + (int64_t)zl * NP * NP;
  if constexpr (NW == 4) {
    switch (warp) {
      case 0: GramMma<NP, NW, 0>::store(acc, Gz, g, t); break;
      case 1: GramMma<NP, NW, 1>::store(acc, Gz, g, t); break;
      case 2: GramMma<NP, NW, 2>::store(acc, Gz, g, t); break;
      default: GramMma<NP, NW, 3>::store(acc, Gz, g, t); break;
    }
  } else {
    if (warp == 0) GramMma<NP, NW, 0>::store(acc, Gz, g, t);
    else GramMma<NP, NW, 1>::store(acc, Gz, g, t);
  }
  if (ci >= 0) cvec[(int64_t)zl * NP + ci] = (cacc + cacc1) + (cacc2 + cacc3);
  if (tid == 0) {
    // localise_obs = .false.: a zone without any relevant observation is still skipped (rrsqrt.F90:371-372)
    const int used = (q.noloc && s_true == 0) ? 0 : nrel_total;
    mloc[zone] = used;
    atomicAdd(&ctr->relevant, (unsigned long long)used);
    atomicAdd(&ctr->candidates, (unsigned long long)ncand_total);
    if (used == 0) atomicAdd(&ctr->skipped, 1ull);
  }
}

template <int NP, int NW, int CH>
int launch(cudaStream_t st, const ZoneGeom &zg, const ObsGrid &og, const ObsRows &orows, int zone0, int nz,
           double *G, double *c, int32_t *mloc, DevCounters *ctr) {
  const size_t smem = sizeof(double) * 2 * CH * (NP + 4);
  { int rc_ = oak_func_smem(k_gram_mma<NP, NW, CH>, (size_t)((int)smem)); if (rc_) return rc_; }
  k_gram_mma<NP, NW, CH><<<nz, 32 * NW, smem, st>>>(zg, og, orows, zone0, nz, G, c, mloc, ctr);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

}  // namespace

// variant 1: four warps per zone, 2: two warps per zone (chunks of 64 candidates); 3, 4: the same with chunks of 32
// candidates (half the shared memory, twice the CTAs per SM).  Only NP = 64 (the caller keeps k_gram for the rest).
int oak_launch_gram_mma(cudaStream_t st, int variant, int NP, const ZoneGeom &zg, const ObsGrid &og,
                        const ObsRows &orows, int zone0, int nz, double *G, double *c, int32_t *mloc,
                        DevCounters *ctr) {
  if (nz <= 0) return 0;
  if (NP != 64) { oak_set_error("gram_mma: padded ensemble size %d (only 64)", NP); return OAK_ERR_UNSUPPORTED; }
  switch (variant) {
    case 2: return launch<64, 2, 64>(st, zg, og, orows, zone0, nz, G, c, mloc, ctr);
    case 3: return launch<64, 4, 32>(st, zg, og, orows, zone0, nz, G, c, mloc, ctr);
    case 4: return launch<64, 2, 32>(st, zg, og, orows, zone0, nz, G, c, mloc, ctr);
  }
  return launch<64, 4, 64>(st, zg, og, orows, zone0, nz, G, c, mloc, ctr);
}
