// Long curves (N >= 224: the matrix does not fit shared memory): the same bordered factorisation as gp_block.cuh, blocked one level
// up.  The packed 8x8 tiles live in a per-CTA slab in global memory (L2 / HBM); every O(N^3) product streams operand tiles through a
// ring of shared-memory stages filled by bulk-asynchronous copies (cp.async.bulk + mbarrier: one elected thread issues, the copy
// engine delivers, SASS UBLKCP) and runs on the FP64 tensor-core instruction from there.  A warp owns RW tile-rows x 8 tile-columns of
// the output block in registers (the 8 columns are one PANEL = 64 matrix columns), so one staged operand tile feeds 8 or 2 RW DMMA
// pairs instead of one: RW = 2 with 16 warps is a 256 x 64 block, 12.8 flop per byte moved, above the HBM ridge of the FP64 pipe
// (37 TFLOP/s / 6.5 TB/s = 5.7); the global-slab version this replaces ran at about 1 flop per byte of L1/L2 traffic.
//
// Three streamed products (reference: LAPACK dpotrf / dtrtri / dlauum behind sklearn/gaussian_process/_gpr.py:589-591,631-633 as driven
// by fulu/gp_aug.py:75-77; same results to rounding):
//   Cholesky, left-looking by panels:  C(r, p) = K(r, p) - sum_{k < p} L(r, k) L(p, k)^T        "NT": both operands are row panels
//       then the 64 x 64 diagonal block is factored AND inverted in shared memory by the 8x8 code of gp_block.cuh, and the rows
//       below are multiplied by the inverted block (register-local: a warp holds complete panel rows);
//       K(r, p) is assembled on the fly where the accumulation ends - the matrix is never stored as K;
//   inverse, in place, panels from the last to the first:  W = -L(:, p) X_pp;  X(r, p) = sum_{p < k <= r} X(r, k) W(k)      "NN"
//   K^-1 = X^T X met by the recomputed dK/dtheta in the epilogue (never stored):  G(i, j) = sum_{k >= i} X(k, i)^T X(k, j)   "TN"
// Tile-row I of the packed layout holds its tiles 0..I contiguously, so every operand piece is a contiguous run of tiles inside one
// tile-row: one bulk copy each, no tensor map needed.
//
// Written against the same block context as gp_block.cuh (plus the copy/mbarrier primitives), so tests/_host runs it on host threads.
#pragma once
#include "gp_block.cuh"

#define GP_PW 8   // panel width in tiles

#ifndef GP_TILED_RW
#define GP_TILED_RW 2
#endif
#ifndef GP_TILED_KC
#define GP_TILED_KC 4
#endif
#ifndef GP_TILED_NS16
#define GP_TILED_NS16 2   // stages of a 16-warp CTA (the evaluation kernels)
#endif
#ifndef GP_TILED_SLACK
#define GP_TILED_SLACK 0  // stages kept out of the look-ahead
#endif
#ifndef GP_TILED_NS8
#define GP_TILED_NS8 2    // stages of an 8-warp CTA (the prediction kernel)
#endif
#if defined(__CUDACC__)
#define GP_MBAR_WORDS 1   // an mbarrier is 8 bytes of shared memory
#else
#define GP_MBAR_WORDS 2   // (the host emulation of tests/_host keeps 16 bytes per barrier)
#endif

enum { GP_MODE_NT = 0, GP_MODE_NN = 1, GP_MODE_TN = 2 };

struct GpTiledCfgDefault {
  static constexpr int RW = GP_TILED_RW;   // tile-rows of the output block per warp
  static constexpr int KC = GP_TILED_KC;   // k-tiles per stage
  static constexpr int NS_MAX = GP_TILED_NS16 > GP_TILED_NS8 ? GP_TILED_NS16 : GP_TILED_NS8;
  GP_HD static int stages(int nwarp) { return nwarp >= 16 ? GP_TILED_NS16 : GP_TILED_NS8; }
  // chunks in flight ahead of the one being multiplied.  Measured on B200, N = 1024 / 2048, one evaluation per CTA, algorithmic
  // TFLOP/s: KC = 2 with three stages and two chunks ahead 14.8 / 18.8; four stages, one of them slack (so that the issuing warp never
  // waits for the slowest warp to leave the stage it refills) 13.6 / 17.4; four stages, three ahead 13.4 / 17.1; KC = 4 with two
  // stages (2 KB copies, half the barrier rounds per k-tile) 15.7 / 20.4
  GP_HD static int lookahead(int ns) { return ns - 1 - GP_TILED_SLACK > 0 ? ns - 1 - GP_TILED_SLACK : 1; }
};

// ---------------------------------------------------------------------------------------------- sizes
GP_HD int gp_tiled_panels(int T) { return (T + GP_PW - 1) / GP_PW; }
// per-CTA slab in global memory (doubles): packed lower tiles + the scratch column (gp_load_curve stages alpha / y there), the
// inverse's W panel (T x 8 tiles) and the inverted diagonal blocks (36 tiles per panel)
GP_HD size_t gp_tiled_slab_doubles(int np) {
  const size_t T = (size_t)np / GP_NB;
  return 64 * (T * (T + 1) / 2 + T + GP_PW * T + 36 * (size_t)gp_tiled_panels((int)T));
}
GP_HD size_t gp_tiled_dblock_doubles() { return 64 * (36 + GP_PW); }
template <class Cfg>
GP_HD size_t gp_tiled_stage_doubles(int nwarp) { return (size_t)64 * Cfg::KC * (nwarp * Cfg::RW + GP_PW); }
// shared memory of a tiled CTA (doubles): the barriers, the diagonal block (+ its scratch column), the stages, the vectors
template <class Cfg>
GP_HD size_t gp_tiled_smem_doubles(int n, int nwarp) {
  return gp_vec_doubles(gp_pad(n)) + gp_tiled_dblock_doubles() + Cfg::stages(nwarp) * gp_tiled_stage_doubles<Cfg>(nwarp) +
         2 * Cfg::NS_MAX * GP_MBAR_WORDS;
}

struct GpTiled {
  GpSmem s;            // A = the slab's packed tiles (global); u, v, red, tab, flag in shared memory
  double* W;           // global: T x 8 tiles
  double* XD;          // global: 36 tiles per panel
  double* D;           // shared: the diagonal block being factored / its inverse (36 + 8 tiles)
  double* stage;       // shared
  uint64_t* bars;      // shared: full[ns], empty[ns]
  uint32_t it;         // chunks streamed so far by this CTA (same value in every thread): stage = it % ns, use = it / ns
  int ns;              // stages
};
GP_DEV uint64_t* gp_bar_full(const GpTiled& g, int st) { return g.bars + (size_t)GP_MBAR_WORDS * st; }
GP_DEV uint64_t* gp_bar_empty(const GpTiled& g, int st) { return g.bars + (size_t)GP_MBAR_WORDS * (g.ns + st); }

// Shared memory: barriers, diagonal block and stages at fixed offsets (they outlive a curve: the barriers keep their phase across
// the evaluations of a kernel), the curve's vectors last.  `it` is carried by the kernel from curve to curve.
template <class Cfg>
GP_HD GpTiled gp_tiled_carve(double* slab, double* smem, int n, int nwarp, uint32_t it) {
  GpTiled g;
  g.ns = Cfg::stages(nwarp);
  double* p = smem;
  g.bars = reinterpret_cast<uint64_t*>(p); p += 2 * Cfg::NS_MAX * GP_MBAR_WORDS;
  g.D = p; p += gp_tiled_dblock_doubles();
  g.stage = p; p += g.ns * gp_tiled_stage_doubles<Cfg>(nwarp);
  g.s = gp_carve(slab, p, n);
  const size_t T = (size_t)g.s.T;
  g.W = slab + 64 * (T * (T + 1) / 2 + T);
  g.XD = g.W + 64 * GP_PW * T;
  g.it = it;
  return g;
}

// once per kernel, before the first streamed product (the barriers keep their phase across evaluations: g.it carries on)
template <class Cfg, class Ctx>
GP_DEV void gp_tiled_init(Ctx& c, GpTiled& g) {
  if (c.tid == 0) {
    for (int s = 0; s < g.ns; ++s) {
      c.mbar_init(gp_bar_full(g, s), 1);         // full: the producer's arrive.expect_tx (+ the copies' bytes)
      c.mbar_init(gp_bar_empty(g, s), c.nwarp);  // empty: one arrival per warp
    }
    c.mbar_fence_init();
  }
  // stale stage contents are only ever multiplied by zero operands of masked tiles, but they must be finite for that
  const size_t ns = g.ns * gp_tiled_stage_doubles<Cfg>(c.nwarp);
  for (size_t i = c.tid; i < ns; i += c.nthr) g.stage[i] = 0.0;
  g.it = 0;
  c.fence_async();   // (the stages are written by the copy engine from now on)
  c.sync();
}

// ---------------------------------------------------------------------------------------------- the streamed product
struct GpStreamJob {
  const double* A;   // packed tiles
  const double* Wp;  // NN: the W panel (row k at 64 * 8 * k)
  int r0, nrows;     // output rows: tile-rows r0 .. r0 + nrows - 1 (NT, NN) / tile-columns of X (TN); nrows <= nwarp * RW
  int k0, k1;        // k-tiles [k0, k1)
  int b0, nb;        // output columns: tile-rows (NT) / tile-columns (TN) b0 .. b0 + nb - 1 of the panel, nb <= 8 (NN: nb only)
  int kborder;       // TN: the k-tile that holds the border row (T - 1), whose row `rb` is left out of the sum
  int rb;
  int tri;           // NT, TN: output tile (row r, column tile b0 + j) is wanted only for b0 + j <= r (blocks that meet the diagonal)
};

// slot of (warp w, i-th row of the warp): a block with few rows spreads them over all warps, and a warp's rows pair up from the two
// ends of the block (w and 2 nwarp - 1 - w), so that where a row's k-range depends on the row (NN: k <= r, TN: k >= r) every warp
// gets the same number of tile products
#define GP_SLOT(w, i, nwarp) (((i) & 1) ? ((i) + 1) * (nwarp) - 1 - (w) : (i) * (nwarp) + (w))

// Chunk `chunk` of the job into its stage: called by every lane of warp 0; lane e issues the copies e, e + 32, ...
template <int MODE, class Cfg, class Ctx>
GP_DEV void gp_stream_issue(Ctx& c, GpTiled& g, const GpStreamJob& job, int chunk, uint32_t it) {
  constexpr int KC = Cfg::KC;
  const int BM = c.nwarp * Cfg::RW, NS = g.ns;
  const uint32_t st = it % NS, use = it / NS;
  uint64_t* full = gp_bar_full(g, st);
  if (use > 0) c.mbar_wait_reuse(gp_bar_empty(g, st), (use - 1) & 1);   // every warp has released the previous contents of this stage
  double* dst = g.stage + (size_t)st * gp_tiled_stage_doubles<Cfg>(c.nwarp);
  const int kc = job.k0 + chunk * KC;
  const int kn = (job.k1 - kc < KC) ? job.k1 - kc : KC;
  // the bytes of the whole chunk first (every lane computes the same number; lane 0 announces it)
  int tiles = 0;
  if (MODE == GP_MODE_NT) {
    tiles = (job.nrows + job.nb) * kn;
  } else if (MODE == GP_MODE_NN) {
    // row r holds k-tiles <= r: rows r >= kc + kn - 1 take the whole chunk, rows kc + d (d < kn - 1) take d + 1 tiles
    const int rlast = job.r0 + job.nrows - 1;
    int full_rows = rlast - (kc + kn - 1) + 1;
    if (job.r0 > kc + kn - 1) full_rows = job.nrows;
    if (full_rows < 0) full_rows = 0;
    tiles = full_rows * kn;
    for (int d = 0; d < kn - 1; ++d) {
      const int r = kc + d;
      if (r >= job.r0 && r <= rlast) tiles += d + 1;
    }
    tiles += GP_PW * kn;
  } else {
    for (int kk = 0; kk < kn; ++kk) {
      const int k = kc + kk;
      int ca = k + 1 - job.r0, cb = k + 1 - job.b0;
      if (ca > job.nrows) ca = job.nrows;
      if (cb > job.nb) cb = job.nb;
      tiles += (ca > 0 ? ca : 0) + (cb > 0 ? cb : 0);
    }
  }
#ifdef GP_TILED_NOCOPY   // timing experiment only (results are garbage): the products without their operand traffic
  if (c.lane == 0) c.mbar_expect_tx(full, 0u);
  c.warp_sync();
  return;
#endif
  if (c.lane == 0) c.mbar_expect_tx(full, (uint32_t)tiles * 512u);
  c.warp_sync();
  if (MODE == GP_MODE_NT) {
    for (int e = c.lane; e < job.nrows + job.nb; e += 32) {
      if (e < job.nrows) c.bulk_g2s(dst + 64 * e * KC, job.A + gp_tile(job.r0 + e, kc), 512u * kn, full);
      else c.bulk_g2s(dst + 64 * (BM + e - job.nrows) * KC, job.A + gp_tile(job.b0 + e - job.nrows, kc), 512u * kn, full);
    }
  } else if (MODE == GP_MODE_NN) {
    for (int e = c.lane; e <= job.nrows; e += 32) {
      if (e < job.nrows) {
        const int r = job.r0 + e;
        int cnt = r + 1 - kc;
        if (cnt > kn) cnt = kn;
        if (cnt > 0) c.bulk_g2s(dst + 64 * e * KC, job.A + gp_tile(r, kc), 512u * cnt, full);
      } else {
        c.bulk_g2s(dst + 64 * BM * KC, job.Wp + (size_t)64 * GP_PW * kc, 512u * GP_PW * kn, full);
      }
    }
  } else {
    for (int e = c.lane; e < 2 * kn; e += 32) {
      const int kk = e >> 1, k = kc + kk;
      if ((e & 1) == 0) {
        int ca = k + 1 - job.r0;
        if (ca > job.nrows) ca = job.nrows;
        if (ca > 0) c.bulk_g2s(dst + 64 * kk * BM, job.A + gp_tile(k, job.r0), 512u * ca, full);
      } else {
        int cb = k + 1 - job.b0;
        if (cb > job.nb) cb = job.nb;
        if (cb > 0) c.bulk_g2s(dst + 64 * (KC * BM + kk * GP_PW), job.A + gp_tile(k, job.b0), 512u * cb, full);
      }
    }
  }
}

// B fragment of column tile j, k-tile kk of the chunk in stage S
template <int MODE, int KC>
GP_DEV void gp_stream_bfrag(const double* S, int BM, int kk, int j, const GpFrag& f, double& b0, double& b1) {
  if (MODE == GP_MODE_NT) {
    const double* Bp = S + 64 * ((BM + j) * KC + kk);
    b0 = Bp[f.b0];
    b1 = Bp[f.b1];
  } else if (MODE == GP_MODE_NN) {
    const double* Bp = S + 64 * (BM * KC + kk * GP_PW + j);
    b0 = Bp[f.p0];
    b1 = Bp[f.p1];
  } else {
    const double* Bp = S + 64 * (KC * BM + kk * GP_PW + j);
    b0 = Bp[f.t0];
    b1 = Bp[f.t1];
  }
}

// acc[i][j][0..1] += sum over the job's k-tiles of A(row i of this warp, k) B(k, column tile j); accumulators in the PERMUTED layout
// (column t, column 4 + t) for NT and NN, in the natural one (columns 2t, 2t + 1) for TN - as in gp_block.cuh.  Tile products that
// are not wanted (above the diagonal, k outside a row's range) are skipped, warp by warp.
// Every thread of the block must call this with the same job (block-wide barrier inside).
template <int MODE, class Cfg, class Ctx>
GP_DEV void gp_stream_mma(Ctx& c, GpTiled& g, const GpStreamJob& job, const GpFrag& f, double (&acc)[Cfg::RW][GP_PW][2]) {
  constexpr int KC = Cfg::KC, RW = Cfg::RW;
  const int BM = c.nwarp * RW, NS = g.ns, LA = Cfg::lookahead(g.ns);
#pragma unroll
  for (int i = 0; i < RW; ++i)
#pragma unroll
    for (int j = 0; j < GP_PW; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
  const int nk = job.k1 - job.k0;
  if (nk <= 0) return;
  const int nchunk = (nk + KC - 1) / KC;
  const uint32_t it0 = g.it;
  c.sync();   // whatever this block wrote to global memory before (and fenced) is visible to the copies issued from here on
  if (c.warp == 0) {
    const int pre = nchunk < LA ? nchunk : LA;
    for (int q = 0; q < pre; ++q) gp_stream_issue<MODE, Cfg>(c, g, job, q, it0 + q);
  }
  const int t4 = c.lane & 3;
  int rowid[RW], jlim[RW];   // this warp's rows and how many of the panel's column tiles each wants (0: no such row)
  int jmax = 0;
  bool jfull = true;
#pragma unroll
  for (int i = 0; i < RW; ++i) {
    const int sl = GP_SLOT(c.warp, i, c.nwarp);
    rowid[i] = job.r0 + sl;
    int jl = job.nb;
    if (job.tri && rowid[i] - job.b0 + 1 < jl) jl = rowid[i] - job.b0 + 1;
    if (sl >= job.nrows || jl < 0) jl = 0;
    jlim[i] = jl;
    if (jl > jmax) jmax = jl;
    jfull = jfull && jl == GP_PW;
  }
  for (int q = 0; q < nchunk; ++q) {
    if (c.warp == 0 && q + LA < nchunk) gp_stream_issue<MODE, Cfg>(c, g, job, q + LA, it0 + q + LA);
    const uint32_t it = it0 + q, st = it % NS, use = it / NS;
    c.mbar_wait(gp_bar_full(g, st), use & 1);
    const double* S = g.stage + (size_t)st * gp_tiled_stage_doubles<Cfg>(c.nwarp);
    const int kc = job.k0 + q * KC;
    const int kn = (job.k1 - kc < KC) ? job.k1 - kc : KC;
    if (jmax > 0) {
#pragma unroll
      for (int kk = 0; kk < KC; ++kk) {
        if (kk < kn) {
          const int k = kc + kk;
          double a0[RW], a1[RW];
          bool act[RW];
          bool anyact = false;
#pragma unroll
          for (int i = 0; i < RW; ++i) {
            const int sl = GP_SLOT(c.warp, i, c.nwarp);
            act[i] = jlim[i] > 0;
            a0[i] = 0.0; a1[i] = 0.0;
            if (MODE == GP_MODE_TN) {
              act[i] = act[i] && rowid[i] <= k;        // tile (k, i) of X exists
              if (act[i]) {
                const double* Ap = S + 64 * (kk * BM + sl);
                a0[i] = Ap[f.t0];
                a1[i] = Ap[f.t1];
                if (k == job.kborder) {   // the border row of M^-1 (-alpha) is not a row of X
                  if (t4 == job.rb) a0[i] = 0.0;
                  if (4 + t4 == job.rb) a1[i] = 0.0;
                }
              }
            } else {
              if (MODE == GP_MODE_NN) act[i] = act[i] && k <= rowid[i];   // tile (r, k) of X exists
              if (act[i]) {
                const double* Ap = S + 64 * (sl * KC + kk);
                a0[i] = Ap[f.a0];
                a1[i] = Ap[f.a1];
              }
            }
            anyact = anyact || act[i];
          }
          bool allact = jfull;
#pragma unroll
          for (int i = 0; i < RW; ++i) allact = allact && act[i];
          if (allact) {
            // the common case - every row of the warp active, all eight column tiles wanted - without a predicate in sight; the B
            // fragment of the next column tile is loaded before the products of this one are issued
            double b0n, b1n;
            gp_stream_bfrag<MODE, KC>(S, BM, kk, 0, f, b0n, b1n);
#pragma unroll
            for (int j = 0; j < GP_PW; ++j) {
              const double b0 = b0n, b1 = b1n;
              if (j + 1 < GP_PW) gp_stream_bfrag<MODE, KC>(S, BM, kk, j + 1, f, b0n, b1n);
#pragma unroll
              for (int i = 0; i < RW; ++i) {
                c.mma(acc[i][j][0], acc[i][j][1], a0[i], b0);
                c.mma(acc[i][j][0], acc[i][j][1], a1[i], b1);
              }
            }
          } else if (anyact) {
            // one row of the warp active with all eight column tiles wanted (the triangular ends of the k-range, where the warp's
            // two rows - one from each end of the block - are never both outside): again free of predicates, because a DMMA that
            // is predicated off is still issued (ncu: 1.26x / 1.47x the needed DMMA issued at N = 2048 / 1024 before this)
            bool done = false;
#pragma unroll
            for (int i = 0; i < RW; ++i) {
              bool only = act[i] && jlim[i] == GP_PW;
#pragma unroll
              for (int i2 = 0; i2 < RW; ++i2)
                if (i2 != i) only = only && !act[i2];
              if (only) {
                done = true;
                double b0n, b1n;
                gp_stream_bfrag<MODE, KC>(S, BM, kk, 0, f, b0n, b1n);
#pragma unroll
                for (int j = 0; j < GP_PW; ++j) {
                  const double b0 = b0n, b1 = b1n;
                  if (j + 1 < GP_PW) gp_stream_bfrag<MODE, KC>(S, BM, kk, j + 1, f, b0n, b1n);
                  c.mma(acc[i][j][0], acc[i][j][1], a0[i], b0);
                  c.mma(acc[i][j][0], acc[i][j][1], a1[i], b1);
                }
              }
            }
            if (!done) {
              // what is left (blocks that meet the diagonal: fewer than eight column tiles wanted): row by row, leaving the column
              // loop by a branch at the row's limit - not a single predicated-off DMMA is issued
#pragma unroll
              for (int i = 0; i < RW; ++i) {
                if (act[i]) {
                  const int jl = jlim[i];
#pragma unroll
                  for (int j = 0; j < GP_PW; ++j) {
                    if (j >= jl) break;
                    double b0, b1;
                    gp_stream_bfrag<MODE, KC>(S, BM, kk, j, f, b0, b1);   // (TN: tile (k, b0 + j) exists: b0 + j <= row <= k)
                    c.mma(acc[i][j][0], acc[i][j][1], a0[i], b0);
                    c.mma(acc[i][j][0], acc[i][j][1], a1[i], b1);
                  }
                }
              }
            }
          }
        }
      }
    }
    c.warp_sync();
    if (c.lane == 0) c.mbar_arrive(gp_bar_empty(g, st));
  }
  g.it = it0 + nchunk;
}

// ---------------------------------------------------------------------------------------------- K tile on the fly
// The values gp_assemble writes for tile (I, J), J <= I, in accumulator layout: lane (g, t) gets elements (g, t) and (g, 4 + t).
// al / yv: the noise diagonal and y as gp_load_curve staged them.
template <class Kern>
GP_DEV void gp_k_tile(int lane, int n, const GpSmem& s, const Kern& kn, const double* al, const double* yv, int I, int J, double& v0,
                      double& v1) {
  const int g = lane >> 2, t = lane & 3;
  const int i = 8 * I + g, j0 = 8 * J + t, j1 = j0 + 4;
  const double ui = s.u[i], vi = s.v[i];
  const double du0 = ui - s.u[j0], dv0 = vi - s.v[j0], du1 = ui - s.u[j1], dv1 = vi - s.v[j1];
  v0 = kn.value(du0 * du0, dv0 * dv0);
  v1 = kn.value(du1 * du1, dv1 * dv1);
  if (i >= n || j0 >= n) v0 = 0.0;
  if (i >= n || j1 >= n) v1 = 0.0;
  if (i == j0) v0 = (i < n) ? kn.diag() + al[i] : (i == n ? 0.0 : 1.0);
  if (i == j1) v1 = (i < n) ? kn.diag() + al[i] : (i == n ? 0.0 : 1.0);
  if (i == n) {
    if (j0 < n) v0 = yv[j0];
    if (j1 < n) v1 = yv[j1];
  } else if (i < n) {
    if (j0 == n) v0 = yv[i];
    if (j1 == n) v1 = yv[i];
  }
}

// view of the shared-memory diagonal block for the 8x8 routines of gp_block.cuh (packed lower tiles of a pw x pw tile matrix)
GP_DEV GpSmem gp_tiled_dview(const GpTiled& g, int pw) {
  GpSmem d = g.s;
  d.A = g.D;
  d.T = pw;
  d.np = 8 * pw;
  d.kslab = nullptr;
  d.W = d.A + gp_tile(pw, 0);   // the block's own scratch column (gp_trtri)
  d.stage = d.W;
  return d;
}

// ---------------------------------------------------------------------------------------------- Cholesky
// On exit the slab holds what gp_cholesky leaves in shared memory for short curves: strictly-lower tiles = L, diagonal 8x8 tiles = the
// inverses of L's diagonal 8x8 blocks, row n = z^T; XD holds the inverse of every 64 x 64 diagonal block (lower tiles).
// logdet / quad valid in thread 0.  Sets s.flag on breakdown (and returns early; every thread sees the same flag).
template <class Kern, class Cfg, class Ctx>
GP_DEV void gp_tiled_cholesky(Ctx& c, GpTiled& g, int n, const Kern& kn, double& logdet, double& quad) {
  constexpr int RW = Cfg::RW;
  const GpSmem& s = g.s;
  const int T = s.T, S = gp_tiled_panels(T), BM = c.nwarp * RW;
  const GpFrag f = gp_frag(c.lane);
  double* A = s.A;
  const double* al = A + gp_tile(T, 0);   // staged by gp_load_curve in the scratch column
  const double* yv = al + s.np;
  double ld = 0.0, qd = 0.0;
  for (int p = 0; p < S; ++p) {
    const int c0 = GP_PW * p;                              // first tile-column of the panel
    const int pw = (T - c0 < GP_PW) ? T - c0 : GP_PW;      // its width in tiles
    // (1) every row from the diagonal block down: C = K - sum_{k < c0} L(r, k) L(panel, k)^T, raw into the slab
    const int R = T - c0, nblk = (R + BM - 1) / BM, per = (R + nblk - 1) / nblk;
    for (int blk = 0; blk < nblk; ++blk) {
      GpStreamJob job;
      job.A = A; job.Wp = nullptr;
      job.r0 = c0 + blk * per;
      job.nrows = (T - job.r0 < per) ? T - job.r0 : per;
      job.k0 = 0; job.k1 = c0;
      job.b0 = c0; job.nb = pw;
      job.kborder = -1; job.rb = 0; job.tri = 1;
      double acc[RW][GP_PW][2];
      gp_stream_mma<GP_MODE_NT, Cfg>(c, g, job, f, acc);
#pragma unroll
      for (int i = 0; i < RW; ++i) {
        const int sl = GP_SLOT(c.warp, i, c.nwarp);
        if (sl < job.nrows) {
          const int r = job.r0 + sl;
#pragma unroll
          for (int j = 0; j < GP_PW; ++j) {
            if (j < pw && c0 + j <= r) {
              double v0, v1;
              gp_k_tile(c.lane, n, s, kn, al, yv, r, c0 + j, v0, v1);
              double* tp = A + gp_tile(r, c0 + j);
              tp[f.a0] = v0 - acc[i][j][0];
              tp[f.a1] = v1 - acc[i][j][1];
            }
          }
        }
      }
    }
    c.sync();
    // (2) the diagonal block into shared memory, factored and inverted there
    GpSmem d = gp_tiled_dview(g, pw);
    const int ntd = pw * (pw + 1) / 2;
    {
      int I = 0, J = c.warp;
      while (J > I) { J -= I + 1; ++I; }
      for (int q = c.warp; q < ntd; q += c.nwarp) {
        const double* tp = A + gp_tile(c0 + I, c0 + J);
        double* dp = g.D + 64 * q;
        dp[f.a0] = tp[f.a0];
        dp[f.a1] = tp[f.a1];
        J += c.nwarp;
        while (J > I) { J -= I + 1; ++I; }
      }
    }
    c.sync();
    double l1, q1;
    gp_cholesky(c, d, n, l1, q1, p == S - 1);
    ld += l1;
    qd += q1;
#ifdef GP_TILED_NOCOPY
    const bool bad = false;
#else
    const bool bad = (s.flag[0] != 0);
#endif
    if (bad) { logdet = 0.0; quad = 0.0; c.sync(); return; }
    {   // L format back into the slab
      int I = 0, J = c.warp;
      while (J > I) { J -= I + 1; ++I; }
      for (int q = c.warp; q < ntd; q += c.nwarp) {
        double* tp = A + gp_tile(c0 + I, c0 + J);
        const double* dp = g.D + 64 * q;
        tp[f.a0] = dp[f.a0];
        tp[f.a1] = dp[f.a1];
        J += c.nwarp;
        while (J > I) { J -= I + 1; ++I; }
      }
    }
    c.sync();
    gp_trtri(c, d);   // D <- inverse of the block's factor (ends with a block barrier)
    {   // kept for the inverse phase
      double* xd = g.XD + (size_t)64 * 36 * p;
      for (int q = c.tid; q < 64 * ntd; q += c.nthr) xd[q] = g.D[q];
    }
    // (3) rows below the block: P <- P X_D^T, a warp per tile-row, in place in registers (tile j needs the tiles k <= j: descending j)
    for (int r = c0 + pw + c.warp; r < T; r += c.nwarp) {
      double a[GP_PW][2];
      double* rp = A + gp_tile(r, c0);
#pragma unroll
      for (int k = 0; k < GP_PW; ++k)
        if (k < pw) { a[k][0] = rp[64 * k + f.a0]; a[k][1] = rp[64 * k + f.a1]; }
#pragma unroll
      for (int j = GP_PW - 1; j >= 0; --j) {
        if (j < pw) {
          double w0 = 0.0, w1 = 0.0;
#pragma unroll
          for (int k = 0; k < GP_PW; ++k) {
            if (k <= j) {
              const double* X = g.D + gp_tile(j, k);   // B[kk][nn] = X_D(j, k)[nn][kk]
              c.mma(w0, w1, a[k][0], X[f.b0]);
              c.mma(w0, w1, a[k][1], X[f.b1]);
            }
          }
          a[j][0] = w0; a[j][1] = w1;
        }
      }
#pragma unroll
      for (int k = 0; k < GP_PW; ++k)
        if (k < pw) { rp[64 * k + f.a0] = a[k][0]; rp[64 * k + f.a1] = a[k][1]; }
    }
    c.fence_async();   // the rows written here are read by the copy engine from the next panel on
    c.sync();
  }
  logdet = ld;
  quad = qd;
}

// ---------------------------------------------------------------------------------------------- inverse
// M -> M^-1 in place (strictly-lower tiles and diagonal tiles = X; row n = -alpha^T), panels from the last to the first.
template <class Cfg, class Ctx>
GP_DEV void gp_tiled_trtri(Ctx& c, GpTiled& g) {
  constexpr int RW = Cfg::RW;
  const GpSmem& s = g.s;
  const int T = s.T, S = gp_tiled_panels(T), BM = c.nwarp * RW;
  const GpFrag f = gp_frag(c.lane);
  double* A = s.A;
  for (int p = S - 1; p >= 0; --p) {
    const int c0 = GP_PW * p;
    const int pw = (T - c0 < GP_PW) ? T - c0 : GP_PW;
    const int ntd = pw * (pw + 1) / 2;
    // (1) X_pp from the Cholesky phase: into shared memory and into the slab's diagonal block
    {
      const double* xd = g.XD + (size_t)64 * 36 * p;
      for (int q = c.tid; q < 64 * ntd; q += c.nthr) g.D[q] = xd[q];
      int I = 0, J = c.warp;
      while (J > I) { J -= I + 1; ++I; }
      for (int q = c.warp; q < ntd; q += c.nwarp) {
        const double* xp = xd + 64 * q;
        double* tp = A + gp_tile(c0 + I, c0 + J);
        tp[f.a0] = xp[f.a0];
        tp[f.a1] = xp[f.a1];
        J += c.nwarp;
        while (J > I) { J -= I + 1; ++I; }
      }
    }
    c.sync();
    const int r1 = c0 + pw;   // first tile-row below the block
    if (r1 >= T) { c.fence_async(); continue; }
    // (2) W(r) = -L(r, panel) X_pp for the rows below (tile j needs the tiles k >= j: ascending j, in place in registers)
    for (int r = r1 + c.warp; r < T; r += c.nwarp) {
      double a[GP_PW][2];
      const double* rp = A + gp_tile(r, c0);
#pragma unroll
      for (int k = 0; k < GP_PW; ++k)
        if (k < pw) { a[k][0] = rp[64 * k + f.a0]; a[k][1] = rp[64 * k + f.a1]; }
      double* wp = g.W + (size_t)64 * GP_PW * r;
#pragma unroll
      for (int j = 0; j < GP_PW; ++j) {
        double w0 = 0.0, w1 = 0.0;
        if (j < pw) {
#pragma unroll
          for (int k = 0; k < GP_PW; ++k) {
            if (k >= j && k < pw) {
              const double* X = g.D + gp_tile(k, j);   // B[kk][nn] = X_pp(k, j)[kk][nn]
              c.mma(w0, w1, a[k][0], X[f.p0]);
              c.mma(w0, w1, a[k][1], X[f.p1]);
            }
          }
        }
        wp[64 * j + f.a0] = -w0;   // (tiles j >= pw of a narrow last panel are zero: the copy engine moves whole rows of 8)
        wp[64 * j + f.a1] = -w1;
      }
    }
    c.fence_async();
    c.sync();
    // (3) X(r, panel) = sum_{r1 <= k <= r} X(r, k) W(k)
    const int R = T - r1, nblk = (R + BM - 1) / BM, per = (R + nblk - 1) / nblk;
    for (int blk = 0; blk < nblk; ++blk) {
      GpStreamJob job;
      job.A = A; job.Wp = g.W;
      job.r0 = r1 + blk * per;
      job.nrows = (T - job.r0 < per) ? T - job.r0 : per;
      job.k0 = r1; job.k1 = job.r0 + job.nrows;   // k <= r: masked per row inside
      job.b0 = c0; job.nb = pw;
      job.kborder = -1; job.rb = 0; job.tri = 0;
      double acc[RW][GP_PW][2];
      gp_stream_mma<GP_MODE_NN, Cfg>(c, g, job, f, acc);
#pragma unroll
      for (int i = 0; i < RW; ++i) {
        const int sl = GP_SLOT(c.warp, i, c.nwarp);
        if (sl < job.nrows) {
          const int r = job.r0 + sl;
#pragma unroll
          for (int j = 0; j < GP_PW; ++j) {
            if (j < pw) {
              double* tp = A + gp_tile(r, c0 + j);
              tp[f.a0] = acc[i][j][0];
              tp[f.a1] = acc[i][j][1];
            }
          }
        }
      }
    }
    c.fence_async();
    c.sync();
  }
}

// ---------------------------------------------------------------------------------------------- K^-1 and the gradient traces
// tr[k] = sum_{i > j} W_ij dK_ij / dtheta_k, trd = sum_i W_ii, W = alpha alpha^T - K^-1 (per-thread partial sums, as gp_inverse_traces)
template <class Kern, class Cfg, class Ctx>
GP_DEV void gp_tiled_traces(Ctx& c, GpTiled& g, int n, const Kern& kn, double (&tr)[Kern::NTR], double& trd) {
  constexpr int RW = Cfg::RW;
  const GpSmem& s = g.s;
  const int T = s.T, BM = c.nwarp * RW, gq = c.lane >> 2, t = c.lane & 3, rb = n & 7;
  const GpFrag f = gp_frag(c.lane);
  const double* A = s.A;
  const double* bord = A + gp_tile(T - 1, 0);   // row n of M^-1 holds -alpha
  const int ob_g = gp_el(rb, gq), ob_0 = gp_el(rb, 2 * t), ob_1 = gp_el(rb, 2 * t + 1);
#pragma unroll
  for (int k = 0; k < Kern::NTR; ++k) tr[k] = 0.0;
  trd = 0.0;
  const int nblk = (T + BM - 1) / BM, per = (T + nblk - 1) / nblk;
  for (int blk = 0; blk < nblk; ++blk) {
    const int i0 = blk * per;
    const int ni = (T - i0 < per) ? T - i0 : per;
    const int jmax = i0 + ni - 1;   // last tile-column any row of the block needs
    for (int b0 = 0; b0 <= jmax; b0 += GP_PW) {
      GpStreamJob job;
      job.A = A; job.Wp = nullptr;
      job.r0 = i0; job.nrows = ni;
      job.k0 = i0; job.k1 = T;
      job.b0 = b0; job.nb = (jmax + 1 - b0 < GP_PW) ? jmax + 1 - b0 : GP_PW;
      job.kborder = T - 1; job.rb = rb; job.tri = 1;
      double acc[RW][GP_PW][2];
      gp_stream_mma<GP_MODE_TN, Cfg>(c, g, job, f, acc);
#pragma unroll
      for (int ii = 0; ii < RW; ++ii) {
        const int sl = GP_SLOT(c.warp, ii, c.nwarp);
        if (sl >= ni) continue;
        const int I = i0 + sl;
        const int i = 8 * I + gq;
        const double ai = bord[64 * I + ob_g], ui = s.u[i], vi = s.v[i];
        const bool in = i < n;
#pragma unroll
        for (int jj = 0; jj < GP_PW; ++jj) {
          const int J = b0 + jj;
          if (jj >= job.nb || J > I) continue;
          const int j = 8 * J + 2 * t;
          const double w0 = ai * bord[64 * J + ob_0] - acc[ii][jj][0], w1 = ai * bord[64 * J + ob_1] - acc[ii][jj][1];
          const GpPair uj = *reinterpret_cast<const GpPair*>(s.u + j), vj = *reinterpret_cast<const GpPair*>(s.v + j);
          const double du0 = ui - uj.x, dv0 = vi - vj.x, du1 = ui - uj.y, dv1 = vi - vj.y;
          kn.accumulate((in && i > j) ? w0 : 0.0, du0 * du0, dv0 * dv0, tr);
          kn.accumulate((in && i > j + 1) ? w1 : 0.0, du1 * du1, dv1 * dv1, tr);
          if (I == J) trd += (in && i == j) ? w0 : ((in && i == j + 1) ? w1 : 0.0);
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------- one LML + gradient evaluation
template <class Kern, class Cfg, class Ctx>
GP_DEV void gp_tiled_eval_lml_grad(Ctx& c, const GpCurveView& cv, const double* theta, GpTiled& g, double& lml, double* grad, bool& ok) {
  const GpSmem& s = g.s;
  Kern kn;
  kn.init(theta);
  c.mark(0);
  gp_load_curve(c, cv, s, kn);   // u, v, tables into shared memory; alpha and y into the slab's scratch column
  c.sync();
  c.mark(1);
  c.mark(2);
  constexpr int NT = Kern::NTR;
  double v[NT + 3];   // traces, trd, quadratic form, log-determinant
  gp_tiled_cholesky<Kern, Cfg>(c, g, cv.n, kn, v[NT + 2], v[NT + 1]);
  c.mark(3);
  ok = (s.flag[0] == 0);
#ifdef GP_TILED_NOCOPY
  ok = true;
#endif
  if (!ok) {
    lml = -INFINITY;
    for (int k = 0; k < Kern::P; ++k) grad[k] = 0.0;
    c.sync();
    return;
  }
  c.mark(4);
  gp_tiled_trtri<Cfg>(c, g);
  c.mark(5);
  c.mark(6);
  {
    double tr[NT];
    gp_tiled_traces<Kern, Cfg>(c, g, cv.n, kn, tr, v[NT]);
#pragma unroll
    for (int k = 0; k < NT; ++k) v[k] = tr[k];
  }
  c.mark(7);
  gp_block_sum(c, v, s.red);
  c.mark(8);
  lml = -0.5 * v[NT + 1] - v[NT + 2] - 0.5 * (double)cv.n * GP_LOG_2PI;
  kn.gradient(v, v[NT], grad);
}

// L_ = chol(K(theta)), z = L^-1 y for prediction (gp_factor's long-curve form); LML as a by-product
template <class Kern, class Cfg, class Ctx>
GP_DEV void gp_tiled_factor(Ctx& c, const GpCurveView& cv, Kern& kn, GpTiled& g, double& lml, bool& ok) {
  const GpSmem& s = g.s;
  gp_load_curve(c, cv, s, kn);
  c.sync();
  double v[2];
  gp_tiled_cholesky<Kern, Cfg>(c, g, cv.n, kn, v[1], v[0]);
  ok = (s.flag[0] == 0);
  if (!ok) { lml = -INFINITY; c.sync(); return; }
  gp_block_sum(c, v, s.red);
  lml = -0.5 * v[0] - v[1] - 0.5 * (double)cv.n * GP_LOG_2PI;
}

// whole-curve prediction for a long curve: tiled factorisation, then the query tiles against the slab (gp_predict_tile)
template <class Kern, class Cfg, class Ctx>
GP_DEV int gp_tiled_predict_curve(Ctx& c, const GpCurveView& cv, const double* theta, const double* stats, const GpQuery& qd, GpTiled& g,
                                  double* Vw, double* mean, double* stdv, double& lml) {
  bool ok;
  Kern kn;
  kn.init(theta);
  gp_tiled_factor<Kern, Cfg>(c, cv, kn, g, lml, ok);
  if (!ok) {
    for (int q = c.tid; q < qd.m; q += c.nthr) { mean[q] = NAN; stdv[q] = NAN; }
    return 2;
  }
  c.sync();
  double neg[1] = {0.0};
  double* mine = Vw + (size_t)c.warp * g.s.T * 64;
  for (int q0 = 8 * c.warp; q0 < qd.m; q0 += 8 * c.nwarp) neg[0] += gp_predict_tile(c, cv, kn, stats, qd, g.s, mine, q0, mean, stdv);
  gp_block_sum(c, neg, g.s.red);
  return neg[0] > 0.0 ? 16 : 0;
}
