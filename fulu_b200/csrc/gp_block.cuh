// Block-level FP64 GP algebra: one CTA owns one light curve's N x N kernel matrix in shared memory.
//
// Functions here are written against a small "block context" (tid/lane/warp, sync(), warp_sync(), mma(), shfl_xor()) so the
// very same code is executed by a CUDA thread block (DevCtx in gp_kernels.cu: __syncthreads, mma.sync.m8n8k4.f64, shuffles)
// and, for tests without a GPU, by host threads with barriers standing in for the warp/block primitives
// (tests/_host/host_gp_block.cpp).  Nothing in here is a CPU fallback of the product: the host build exists only under tests/.
//
// What is computed (reference lines: sklearn/gaussian_process/_gpr.py:583-651, kernels.py:1558-1585):
//   K   = c * exp(-1/2 |x_i/l - x_j/l|^2) + (s2 + alpha_i) I            (a4)
//   L   = chol(K)                       blocked right-looking, NB = 8    (a5; LAPACK dpotrf in the reference)
//   X   = L^-1                          blocked in-place triangular inverse (replaces the identity-RHS dpotrs)
//   a   = X^T z, z = L^-1 y;  LML = -1/2 z.z - sum log L_ii - N/2 log 2pi
//   K^-1 = X^T X is never stored: each 8x8 tile is reduced straight into the gradient traces
//   g_k = 1/2 sum_ij (a_i a_j - K^-1_ij) dK_ij/dtheta_k                  (Eq. 5.9)
//
// Every O(N^3) piece runs on 8x8 tiles through the FP64 tensor-core instruction (DMMA m8n8k4; on B200 it sustains the full
// FP64 rate with an eighth of the issue slots of a DFMA loop, measured 37.2 vs 33.9 TFLOP/s).  The inherently serial part of
// the Cholesky - the chain of N reciprocal square roots down the diagonal - is done by one thread on an 8x8 block *while* the
// other warps apply the previous panel to the rest of the trailing matrix (lookahead), so a step costs two barriers.
//
// Storage: A is column-major, np = N padded to a multiple of 8 (identity padding), np columns of ld rows with
// ld = np + 12 (= 4 mod 8): lane (g, t) of a fragment load touches row g/column t (or row t/column g), i.e. 64-bit bank
// (g + 4t) mod 16 - conflict free for both orientations.  The strict upper triangle keeps K for the gradient pass while the lower
// triangle is overwritten by L and then L^-1.  Rows np..np+7 are a border tile-row: row np carries y^T, so the panel solves and
// trailing updates of the Cholesky turn it into z^T = (L^-1 y)^T for free (bordered factorisation); after that the border
// rows are reused as the 8-wide panel scratch of the triangular inverse.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define GP_DEV __device__ __forceinline__
#define GP_HD __host__ __device__ __forceinline__
#else
#define GP_DEV inline
#define GP_HD inline
#endif

#define GP_NB 8
#define GP_LOG_2PI 1.8378770664093453

struct GpCurveView {
  int n;               // observations
  const double* xs;    // [n] standardised time              (StandardScaler column 0)
  const int32_t* band; // [n] band index
  const double* y;     // [n] normalised flux
  const double* al;    // [n] noise diagonal (regressor alpha)
  const double* lam;   // [n_bands] standardised log10 wavelength (StandardScaler column 1 per band)
};

struct GpSmem {
  double* A;      // np columns of ld rows
  double* u;      // np   x_i0 / l_t
  double* v;      // np   x_i1 / l_lambda
  double* y;      // np
  double* al;     // np
  double* alpha;  // np   K^-1 y
  double* z;      // np   L^-1 y
  double* rdiag;  // np   1 / L_ii
  double* red;    // reduction scratch, 64 doubles
  int* flag;      // [0] = Cholesky breakdown
  int np, ld;
};

GP_HD int gp_pad(int n) { return (n + GP_NB - 1) / GP_NB * GP_NB; }
GP_HD int gp_ld(int np) { return np + GP_NB + 4; }
GP_HD size_t gp_vec_doubles(int np) { return 7 * (size_t)np + 64 + 2; }
// doubles of shared memory needed for a curve of n observations (matrix + vectors)
GP_HD size_t gp_smem_doubles(int n) {
  const size_t np = (size_t)gp_pad(n);
  return np * (size_t)gp_ld((int)np) + gp_vec_doubles((int)np);
}

// vectors at `vec`, matrix at `mat` (shared memory, or a per-CTA global slab for long curves)
GP_HD GpSmem gp_carve(double* mat, double* vec, int n) {
  GpSmem s;
  s.np = gp_pad(n);
  s.ld = gp_ld(s.np);
  s.A = mat;
  double* p = vec;
  s.u = p; p += s.np;
  s.v = p; p += s.np;
  s.y = p; p += s.np;
  s.al = p; p += s.np;
  s.alpha = p; p += s.np;
  s.z = p; p += s.np;
  s.rdiag = p; p += s.np;
  s.red = p; p += 64;
  s.flag = (int*)p;
  return s;
}

// column permutation of the accumulator fragment: register c0 of lane (g, t) holds tile column t, c1 holds column 4 + t
// (instead of 2t, 2t+1), so that accumulator loads/stores hit banks (g + 4t) as well.  B-fragment lane g must then supply
// tile column gp_perm(g).
GP_HD int gp_perm(int g) { return (g & 1) ? 4 + (g >> 1) : (g >> 1); }

// ---------------------------------------------------------------------------------------------- reductions
// Sum of v[k] over the block (result in every thread).  Warp shuffles for the column reduction, one smem hop across warps.
template <class Ctx, int K>
GP_DEV void gp_block_sum(Ctx& c, double (&v)[K], double* red) {
#pragma unroll
  for (int k = 0; k < K; ++k) {
    double x = v[k];
    for (int o = 16; o > 0; o >>= 1) x += c.shfl_xor(x, o);
    v[k] = x;
  }
  c.sync();
  if (c.lane == 0)
    for (int k = 0; k < K; ++k) red[c.warp * K + k] = v[k];
  c.sync();
#pragma unroll
  for (int k = 0; k < K; ++k) {
    double x = 0.0;
    for (int w = 0; w < c.nwarp; ++w) x += red[w * K + k];
    v[k] = x;
  }
  c.sync();
}

// ---------------------------------------------------------------------------------------------- load + assemble
template <class Ctx>
GP_DEV void gp_load_curve(Ctx& c, const GpCurveView& cv, const GpSmem& s, double lt, double ll) {
  for (int i = c.tid; i < s.np; i += c.nthr) {
    if (i < cv.n) {
      s.u[i] = cv.xs[i] / lt;              // X / length_scale  (kernels.py:1561)
      s.v[i] = cv.lam[cv.band[i]] / ll;
      s.y[i] = cv.y[i];
      s.al[i] = cv.al[i];
    } else {
      s.u[i] = 0.0; s.v[i] = 0.0; s.y[i] = 0.0; s.al[i] = 0.0;
    }
  }
  if (c.tid == 0) s.flag[0] = 0;
}

// K into the lower triangle (to be factored), a copy into the strict upper triangle if keep_upper, and the border rows.
// One warp per column, lanes down the rows: conflict-free stores for the lower triangle.
template <class Ctx>
GP_DEV void gp_assemble(Ctx& c, const GpSmem& s, int n, double cst, double s2, bool keep_upper) {
  const int np = s.np, ld = s.ld;
  for (int j = c.warp; j < np; j += c.nwarp) {
    const double uj = s.u[j], vj = s.v[j];
    double* col = s.A + j * ld;
    for (int i = j + c.lane; i < np + GP_NB; i += 32) {
      double val;
      if (i == j) {
        val = (j < n) ? (cst + s2) + s.al[j] : 1.0;    // White on the diagonal, then K[diag] += alpha (_gpr.py:589)
      } else if (i < np) {
        val = 0.0;
        if (i < n) {   // j < i < n
          const double du = s.u[i] - uj, dv = s.v[i] - vj;
          val = cst * exp(-0.5 * (du * du + dv * dv));
        }
        if (keep_upper) s.A[j + i * ld] = val;
      } else {
        val = (i == np) ? s.y[j] : 0.0;                // border: y^T then seven zero rows
      }
      col[i] = val;
    }
  }
}

// ---------------------------------------------------------------------------------------------- Cholesky
// 8x8 diagonal block at A[j0.., j0..] (lower) -> L in place, rdiag[j0..] = 1/L_jj.  Division-free: r = rsqrt(d), L_jj = d r.
// Executed by ONE thread; everything is unrolled into registers so the scheduler starts column j+1's rsqrt as early as possible.
GP_DEV bool gp_chol8(double* A, int ld, int j0, double* rdiag) {
  double a[8][8];
#pragma unroll
  for (int cc = 0; cc < 8; ++cc)
#pragma unroll
    for (int r = cc; r < 8; ++r) a[r][cc] = A[(j0 + r) + (j0 + cc) * ld];
  bool ok = true;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    double d = a[j][j];
    if (!(d > 0.0) || !(d < 1.0e300)) { ok = false; d = 1.0; }
#if defined(__CUDA_ARCH__)
    const double r = rsqrt(d);
#else
    const double r = 1.0 / sqrt(d);
#endif
    a[j][j] = d * r;
    rdiag[j0 + j] = r;
#pragma unroll
    for (int i = j + 1; i < 8; ++i) a[i][j] *= r;
#pragma unroll
    for (int cc = j + 1; cc < 8; ++cc)
#pragma unroll
      for (int i = cc; i < 8; ++i) a[i][cc] -= a[i][j] * a[cc][j];
  }
#pragma unroll
  for (int cc = 0; cc < 8; ++cc)
#pragma unroll
    for (int r = cc; r < 8; ++r) A[(j0 + r) + (j0 + cc) * ld] = a[r][cc];
  return ok;
}

// rows below the diagonal block j0 (border included): solve x L_D^T = p by forward substitution, one row per thread
template <class Ctx>
GP_DEV void gp_panel_solve(Ctx& c, const GpSmem& s, int j0) {
  const int ld = s.ld, rem = s.np - j0;   // rows j0+8 .. np+7
  if (c.tid >= rem) return;
  double* A = s.A;
  double l[8][8], r[8];
#pragma unroll
  for (int cc = 0; cc < 8; ++cc) {
    r[cc] = s.rdiag[j0 + cc];
#pragma unroll
    for (int k = 0; k < cc; ++k) l[cc][k] = A[(j0 + cc) + (j0 + k) * ld];
  }
  for (int i = c.tid; i < rem; i += c.nthr) {
    double* row = A + (j0 + GP_NB + i) + j0 * ld;
    double x[8];
#pragma unroll
    for (int cc = 0; cc < 8; ++cc) {
      double acc = row[cc * ld];
#pragma unroll
      for (int k = 0; k < cc; ++k) acc -= x[k] * l[cc][k];
      x[cc] = acc * r[cc];
    }
#pragma unroll
    for (int cc = 0; cc < 8; ++cc) row[cc * ld] = x[cc];
  }
}

// Trailing update of step j0 for the tiles q0, q0 + stride, ... of the linear lower-triangular enumeration
// (q = ta (ta + 1) / 2 + tb):  C(ta, tb) -= P_ta P_tb^T  with P = the solved panel (8 columns at j0).  Warp-wide.
// Lane (g, t): A-fragment = P[row g][col t], B-fragment = P[row perm(g)][col t], accumulator c0/c1 = C[g][t], C[g][4 + t].
template <class Ctx>
GP_DEV void gp_chol_update_tiles(Ctx& c, const GpSmem& s, int j0, int q0, int stride, int ntile) {
  const int ld = s.ld, g = c.lane >> 2, t = c.lane & 3;
  double* A = s.A;
  const int base = j0 + GP_NB;
  const double* Pa0 = A + base + g + (j0 + t) * ld;            // + 8 ta
  const double* Pb0 = A + base + gp_perm(g) + (j0 + t) * ld;   // + 8 tb
  double* C00 = A + base + g + (base + t) * ld;                // + 8 ta + 8 tb ld
  const int ld4 = 4 * ld, ld8 = 8 * ld;
  // locate the first tile: ta (ta + 1) / 2 <= q0
  int ta = 0, tb = q0;
  while (tb > ta) { tb -= ta + 1; ++ta; }
  for (int q = q0; q < ntile; q += stride) {
    const double* Pa = Pa0 + 8 * ta;
    const double* Pb = Pb0 + 8 * tb;
    double* C = C00 + 8 * ta + tb * ld8;
    double c0v = C[0], c1v = C[ld4];
    c.mma(c0v, c1v, -Pa[0], Pb[0]);
    c.mma(c0v, c1v, -Pa[ld4], Pb[ld4]);
    // the strict upper triangle of a diagonal tile belongs to K: store the lower part only
    const bool diag = (ta == tb);
    if (!diag || g >= t) C[0] = c0v;
    if (!diag || g >= 4 + t) C[ld4] = c1v;
    tb += stride;
    while (tb > ta) { tb -= ta + 1; ++ta; }
  }
}

// Blocked right-looking Cholesky with one-step lookahead.  Sets flag on breakdown.  On exit the lower triangle holds L,
// rdiag holds 1/L_ii and border row np holds z^T = (L^-1 y)^T.
template <class Ctx>
GP_DEV void gp_cholesky(Ctx& c, const GpSmem& s) {
  const int np = s.np, nblk = np / GP_NB;
  if (c.tid == 0) {
    if (!gp_chol8(s.A, s.ld, 0, s.rdiag)) s.flag[0] = 1;
  }
  c.sync();
  gp_panel_solve(c, s, 0);
  c.sync();
  for (int kb = 0; kb + 1 < nblk; ++kb) {
    const int j0 = kb * GP_NB;
    // trailing matrix: T x T lower tiles plus the border tile-row (ta = T) over the T tile-columns
    const int T = nblk - kb - 1, ntile = T * (T + 1) / 2 + T;
    if (c.warp == 0) {
      // critical path first: the next diagonal tile, then its factorisation by one thread
      gp_chol_update_tiles(c, s, j0, 0, ntile, ntile);   // tile (0, 0) only
      c.warp_sync();
      if (c.lane == 0) {
        if (!gp_chol8(s.A, s.ld, j0 + GP_NB, s.rdiag)) s.flag[0] = 1;
      }
      if (c.nwarp == 1) {
        c.warp_sync();
        gp_chol_update_tiles(c, s, j0, 1, 1, ntile);
      }
    } else {
      gp_chol_update_tiles(c, s, j0, c.warp, c.nwarp - 1, ntile);
    }
    c.sync();
    gp_panel_solve(c, s, j0 + GP_NB);
    c.sync();
  }
}

// After gp_cholesky: copies z (border row np) to s.z and returns this thread's share of z.z (= y^T K^-1 y) and of sum log L_ii.
template <class Ctx>
GP_DEV void gp_collect_z(Ctx& c, const GpSmem& s, double& quad, double& logdet) {
  const int np = s.np, ld = s.ld;
  quad = 0.0;
  logdet = 0.0;
  for (int i = c.tid; i < np; i += c.nthr) {
    const double zi = s.A[np + i * ld];
    s.z[i] = zi;
    quad += zi * zi;
    logdet += log(s.A[i + i * ld]);
  }
  c.sync();
}

// ---------------------------------------------------------------------------------------------- triangular inverse
// inverse of every 8x8 diagonal block in place (one thread per block, division-free thanks to rdiag)
template <class Ctx>
GP_DEV void gp_invert_diag_blocks(Ctx& c, const GpSmem& s) {
  const int ld = s.ld, nblk = s.np / GP_NB;
  for (int kb = c.tid; kb < nblk; kb += c.nthr) {
    const int j0 = kb * GP_NB;
    double l[8][8], r[8];
#pragma unroll
    for (int cc = 0; cc < 8; ++cc) {
      r[cc] = s.rdiag[j0 + cc];
#pragma unroll
      for (int i = cc + 1; i < 8; ++i) l[i][cc] = s.A[(j0 + i) + (j0 + cc) * ld];
    }
    // column by column; L stays in registers, so X can overwrite it in memory as soon as a column is done
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      double x[8];
      x[j] = r[j];
#pragma unroll
      for (int i = j + 1; i < 8; ++i) {
        double acc = 0.0;
#pragma unroll
        for (int k = j; k < i; ++k) acc -= l[i][k] * x[k];
        x[i] = acc * r[i];
      }
#pragma unroll
      for (int i = j; i < 8; ++i) s.A[(j0 + i) + (j0 + j) * ld] = x[i];
    }
  }
}

// In place: lower triangle L -> X = L^-1, block columns from the last to the first:  X21 = -X22 (L21 X11).
template <class Ctx>
GP_DEV void gp_trtri(Ctx& c, const GpSmem& s) {
  const int np = s.np, ld = s.ld, nblk = np / GP_NB, g = c.lane >> 2, t = c.lane & 3;
  double* A = s.A;
  gp_invert_diag_blocks(c, s);
  c.sync();
  for (int kb = nblk - 2; kb >= 0; --kb) {
    const int j0 = kb * GP_NB, T = nblk - kb - 1;   // T row-tiles below the diagonal block
    // (1) W = -(L21 X11), stored transposed in the border rows: W[i][cc] at A[(np + cc) + (j0 + 8 + i) * ld]
    for (int I = c.warp; I < T; I += c.nwarp) {
      const int r0 = j0 + GP_NB + 8 * I;
      double w0 = 0.0, w1 = 0.0;
#pragma unroll
      for (int kc = 0; kc < 2; ++kc) {
        const double a = A[(r0 + g) + (j0 + 4 * kc + t) * ld];
        const int n = gp_perm(g);   // tile column supplied by this lane
        const double b = (4 * kc + t >= n) ? A[(j0 + 4 * kc + t) + (j0 + n) * ld] : 0.0;   // X11 is lower triangular
        c.mma(w0, w1, -a, b);
      }
      A[(np + t) + (r0 + g) * ld] = w0;
      A[(np + 4 + t) + (r0 + g) * ld] = w1;
    }
    c.sync();
    // (2) X21[I] = sum_{K <= I} X22[I][K] W[K]   (two accumulator chains to hide the DMMA latency)
    {
      const int base = j0 + GP_NB, ld4 = 4 * ld, ld8 = 8 * ld;
      const double* Xa0 = A + base + g + (base + t) * ld;              // + 8 I + K ld8
      const double* Wb0 = A + np + gp_perm(g) + (base + t) * ld;       // + K ld8
      for (int I = T - 1 - c.warp; I >= 0; I -= c.nwarp) {
        double p0 = 0.0, p1 = 0.0, q0 = 0.0, q1 = 0.0;
        const double* Xa = Xa0 + 8 * I;
        const double* Wb = Wb0;
        for (int K = 0; K < I; ++K) {
          c.mma(p0, p1, Xa[0], Wb[0]);
          c.mma(q0, q1, Xa[ld4], Wb[ld4]);
          Xa += ld8;
          Wb += ld8;
        }
        // K == I: diagonal block of X22, lower triangular
        c.mma(p0, p1, (g < t) ? 0.0 : Xa[0], Wb[0]);
        c.mma(q0, q1, (g < 4 + t) ? 0.0 : Xa[ld4], Wb[ld4]);
        double* O = A + base + 8 * I + g + (j0 + t) * ld;
        O[0] = p0 + q0;
        O[ld4] = p1 + q1;
      }
    }
    c.sync();
  }
}

// alpha = X^T z  (X = L^-1 in the lower triangle): one warp per column, lanes down the rows, shuffle reduction
template <class Ctx>
GP_DEV void gp_alpha_from_inverse(Ctx& c, const GpSmem& s) {
  const int np = s.np, ld = s.ld;
  for (int j = c.warp; j < np; j += c.nwarp) {
    const double* col = s.A + j * ld;
    double acc = 0.0;
    for (int i = j + c.lane; i < np; i += 32) acc += col[i] * s.z[i];
    for (int o = 16; o > 0; o >>= 1) acc += c.shfl_xor(acc, o);
    if (c.lane == 0) s.alpha[j] = acc;
  }
  c.sync();
}

// Fused K^-1 = X^T X (8x8 DMMA tiles) and gradient traces.  tr[0] = sum_{i>j} W_ij K_ij, tr[1]/tr[2] the same weighted by the
// squared scaled time / wavelength differences, tr[3] = sum_i W_ii  (W = alpha alpha^T - K^-1).
template <class Ctx>
GP_DEV void gp_inverse_traces(Ctx& c, const GpSmem& s, int n, double tr[4]) {
  const int np = s.np, ld = s.ld, T = np / GP_NB, ntile = T * (T + 1) / 2, g = c.lane >> 2, t = c.lane & 3;
  const double* A = s.A;
  tr[0] = tr[1] = tr[2] = tr[3] = 0.0;
  int I = 0, J = c.warp;
  while (J > I) { J -= I + 1; ++I; }
  for (int q = c.warp; q < ntile; q += c.nwarp) {
    const int i0 = 8 * I, j0 = 8 * J;
    double p0 = 0.0, p1 = 0.0, q0 = 0.0, q1 = 0.0;
    // K^-1[i][j] = sum_{k >= i} X[k][i] X[k][j]:  A-fragment (m = g, k = t) = X[k0 + t][i0 + g], B (k = t, n = g) = X[k0 + t][j0 + g]
    const double* Xi = A + t + (i0 + g) * ld;
    const double* Xj = A + t + (j0 + g) * ld;
    {
      // k-block K = I: X[k][i] is lower triangular there (k >= i), and so is X[k][j] when J == I
      const int k0 = i0;
      double a0 = Xi[k0], a1 = Xi[k0 + 4], b0 = Xj[k0], b1 = Xj[k0 + 4];
      if (t < g) a0 = 0.0;
      if (4 + t < g) a1 = 0.0;
      if (I == J) { b0 = a0; b1 = a1; }
      c.mma(p0, p1, a0, b0);
      c.mma(q0, q1, a1, b1);
    }
    for (int k0 = i0 + 8; k0 < np; k0 += 8) {
      c.mma(p0, p1, Xi[k0], Xj[k0]);
      c.mma(q0, q1, Xi[k0 + 4], Xj[k0 + 4]);
    }
    // epilogue: lane (g, t) owns (i, j) = (i0 + g, j0 + 2t) and (i0 + g, j0 + 2t + 1)
    const int i = i0 + g;
    const double ai = s.alpha[i], ui = s.u[i], vi = s.v[i];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int j = j0 + 2 * t + e;
      const double kinv = e ? (p1 + q1) : (p0 + q0);
      const double w = ai * s.alpha[j] - kinv;
      if (i > j) {
        const double kij = A[j + i * ld];   // K kept in the strict upper triangle
        const double du = ui - s.u[j], dv = vi - s.v[j];
        const double wk = w * kij;
        tr[0] += wk;
        tr[1] += wk * (du * du);
        tr[2] += wk * (dv * dv);
      } else if (i == j && i < n) {
        tr[3] += w;
      }
    }
    J += c.nwarp;
    while (J > I) { J -= I + 1; ++I; }
  }
}

// ---------------------------------------------------------------------------------------------- one LML + gradient evaluation
// theta = log[c, l_t, l_lambda, s2].  Result valid in every thread.  ok=false <=> K not positive definite.
template <class Ctx>
GP_DEV void gp_eval_lml_grad(Ctx& c, const GpCurveView& cv, const double* theta, const GpSmem& s, double& lml, double grad[4], bool& ok) {
  const double cst = exp(theta[0]), lt = exp(theta[1]), ll = exp(theta[2]), s2 = exp(theta[3]);
  c.mark(0);
  gp_load_curve(c, cv, s, lt, ll);
  c.sync();
  c.mark(1);
  gp_assemble(c, s, cv.n, cst, s2, true);
  c.sync();
  c.mark(2);
  gp_cholesky(c, s);
  c.mark(3);
  ok = (s.flag[0] == 0);
  if (!ok) {
    lml = -INFINITY;
    grad[0] = grad[1] = grad[2] = grad[3] = 0.0;
    c.sync();
    return;
  }
  double v[6];
  gp_collect_z(c, s, v[4], v[5]);
  c.mark(4);
  gp_trtri(c, s);
  c.mark(5);
  gp_alpha_from_inverse(c, s);
  c.mark(6);
  gp_inverse_traces(c, s, cv.n, v);
  c.mark(7);
  gp_block_sum(c, v, s.red);
  c.mark(8);
  lml = -0.5 * v[4] - v[5] - 0.5 * (double)cv.n * GP_LOG_2PI;
  grad[0] = v[0] + 0.5 * cst * v[3];
  grad[1] = v[1];
  grad[2] = v[2];
  grad[3] = 0.5 * s2 * v[3];
}

// ---------------------------------------------------------------------------------------------- factor for prediction
// L_ = chol(K(theta)) and z = L^-1 y (_gpr.py:349-367; alpha_ = L^-T z is never needed: mean* = v.z with v = L^-1 k*).
// Returns LML as a by-product.
template <class Ctx>
GP_DEV void gp_factor(Ctx& c, const GpCurveView& cv, const double* theta, const GpSmem& s, double& lml, bool& ok) {
  const double cst = exp(theta[0]), lt = exp(theta[1]), ll = exp(theta[2]), s2 = exp(theta[3]);
  gp_load_curve(c, cv, s, lt, ll);
  c.sync();
  gp_assemble(c, s, cv.n, cst, s2, false);
  c.sync();
  gp_cholesky(c, s);
  ok = (s.flag[0] == 0);
  if (!ok) { lml = -INFINITY; c.sync(); return; }
  double v[2];
  gp_collect_z(c, s, v[0], v[1]);
  gp_block_sum(c, v, s.red);
  lml = -0.5 * v[0] - v[1] - 0.5 * (double)cv.n * GP_LOG_2PI;
}

// ---------------------------------------------------------------------------------------------- prediction
// Query points against a factored curve (s.A lower = L, s.z = L^-1 y), one query point per thread:
//   k*_i = c exp(-1/2 |x*/l - x_i/l|^2)          (kernels.py:1569-1570; White contributes 0 off the training set)
//   v    = L^-1 k*  by forward substitution       (_gpr.py:460-462: solve_triangular, NOT an explicit inverse - at noise
//                                                  levels of 1e-5 the explicit-inverse product loses ~40x in the variance)
//   mean* = k*.alpha = v.z                        (_gpr.py:446-447)
//   |v|^2 -> var = (c + s2) - |v|^2               (_gpr.py:480-481)
// Each thread owns column q of V (V[i * ldv + q]); rows are processed four at a time so each v_j load feeds four FMAs.
#define GP_QC 128   // query columns per pass (= threads that own a column)

GP_DEV void gp_predict_column(const GpSmem& s, int n, double cst, double uq, double vq, double* V, int ldv, double& mean, double& v2) {
  const int ld = s.ld;
  const double* A = s.A;
  double m = 0.0, q2 = 0.0;
  for (int i0 = 0; i0 < n; i0 += 4) {
    double acc[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int i = i0 + r;
      double k = 0.0;
      if (i < n) {
        const double du = uq - s.u[i], dv = vq - s.v[i];
        k = cst * exp(-0.5 * (du * du + dv * dv));
      }
      acc[r] = k;
    }
    for (int j = 0; j < i0; ++j) {
      const double vj = V[j * ldv];
      const double* col = A + i0 + j * ld;
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[r] -= col[r] * vj;
    }
    // 4x4 diagonal block by substitution (rows >= n are identity padding)
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int i = i0 + r;
      double a = acc[r];
#pragma unroll
      for (int k = 0; k < r; ++k) a -= A[i + (i0 + k) * ld] * acc[k];
      a = a * s.rdiag[i];
      acc[r] = a;
      V[i * ldv] = a;
      m += a * s.z[i];
      q2 += a * a;
    }
  }
  mean = m;
  v2 = q2;
}

// ---------------------------------------------------------------------------------------------- per-curve preparation (a1-a3)
// StandardScaler statistics with the corrected two-pass variance (sklearn/utils/extmath.py:1203-1262), near-constant
// columns -> scale 1 (sklearn/preprocessing/_data.py:83-98,1073-1078); y normalisation (_gpr.py:275-280);
// noise diagonal (fulu/gp_aug.py:64,72-73).
GP_DEV void gp_scaler_stats(double sum, double csum, double csq, int n, double& mean, double& scale) {
  const double eps = 2.220446049250313e-16;
  mean = sum / n;
  double unnorm = csq - csum * csum / n;
  double var = unnorm / n;
  double ub = n * eps * var + (n * mean * eps) * (n * mean * eps);
  scale = sqrt(var);
  if (var <= ub || scale < 10.0 * eps) scale = 1.0;
}

// Per-curve preparation by one block.  Outputs: xs/y/al [n] (workspace), lam [n_bands], stats[6] =
// (t mean, t scale, loglam mean, loglam scale, flux mean, flux std).  Returns a FULU_GP_STATUS_* value (valid in all threads).
template <class Ctx>
GP_DEV int gp_prepare_curve(Ctx& c, int n, const double* t, const double* flux, const double* ferr, const int32_t* band,
                            const double* loglam, int n_bands, int use_err, double alpha0, double* xs, double* y, double* al,
                            double* lam, double* stats, double* red) {
  double v[4] = {0.0, 0.0, 0.0, 0.0};   // sum t, sum loglam, sum flux, bad-count
  for (int i = c.tid; i < n; i += c.nthr) {
    const int b = band[i];
    const bool badb = (b < 0 || b >= n_bands);
    const double ti = t[i], fi = flux[i], ei = ferr[i];
    const double li = badb ? 0.0 : loglam[b];
    const bool bad = badb || !(fabs(ti) <= 1.7e308) || !(fabs(fi) <= 1.7e308) || (use_err && !(fabs(ei) <= 1.7e308));
    v[0] += ti; v[1] += li; v[2] += fi; v[3] += bad ? 1.0 : 0.0;
  }
  gp_block_sum(c, v, red);
  if (v[3] > 0.0 || n <= 0) return 1;  // FULU_GP_STATUS_BAD_INPUT
  const double mt0 = v[0] / n, ml0 = v[1] / n, mf = v[2] / n;
  double w[5] = {0.0, 0.0, 0.0, 0.0, 0.0};   // centred sums: t, t^2, l, l^2, f^2
  for (int i = c.tid; i < n; i += c.nthr) {
    const double dt = t[i] - mt0, dl = loglam[band[i]] - ml0, df = flux[i] - mf;
    w[0] += dt; w[1] += dt * dt; w[2] += dl; w[3] += dl * dl; w[4] += df * df;
  }
  gp_block_sum(c, w, red);
  double mt, st, ml, sl;
  gp_scaler_stats(v[0], w[0], w[1], n, mt, st);
  gp_scaler_stats(v[1], w[2], w[3], n, ml, sl);
  const double fstd_raw = sqrt(w[4] / n);             // np.std(flux), ddof = 0
  const double fstd = (fstd_raw == 0.0) ? 1.0 : fstd_raw;   // _handle_zeros_in_scale (_data.py:116-119)
  int status = 0;
  for (int i = c.tid; i < n; i += c.nthr) {
    xs[i] = (t[i] - mt) / st;
    y[i] = (flux[i] - mf) / fstd;
    double a = alpha0;
    if (use_err) { const double r = ferr[i] / fstd_raw; a = r * r; }   // (flux_err / np.std(flux))**2, gp_aug.py:64,73
    if (!(fabs(a) <= 1.7e308)) status = 1;
    al[i] = a;
  }
  for (int b = c.tid; b < n_bands; b += c.nthr) lam[b] = (loglam[b] - ml) / sl;
  if (c.tid == 0) { stats[0] = mt; stats[1] = st; stats[2] = ml; stats[3] = sl; stats[4] = mf; stats[5] = fstd; }
  double bad[1] = {(double)status};
  gp_block_sum(c, bad, red);
  return bad[0] > 0.0 ? 1 : 0;
}

// ---------------------------------------------------------------------------------------------- whole-curve prediction
struct GpQuery {
  int m;                 // number of query points of this curve
  int n_obs;             // grid mode: points per band (m = n_bands * n_obs); 0 = explicit points
  double t_min, t_max;   // grid mode
  const double* qt;      // explicit mode [m]
  const int32_t* qband;  // explicit mode [m]
};

// np.linspace(t_min, t_max, n_obs)[k]: arange * step + start with the end point forced (numpy/_core/function_base.py)
GP_DEV double gp_linspace(double t_min, double t_max, int n_obs, int k) {
  if (n_obs <= 1) return t_min;
  if (k == n_obs - 1) return t_max;
  const double step = (t_max - t_min) / (double)(n_obs - 1);
#if defined(__CUDA_ARCH__)
  return __dadd_rn(__dmul_rn((double)k, step), t_min);   // no FMA contraction: numpy rounds the product first
#else
  volatile double prod = (double)k * step;
  return prod + t_min;
#endif
}

// mean/std [m] of one curve at hyperparameters theta.  stats = per-curve (t mean, t scale, ., ., flux mean, flux std).
// V: scratch of gp_pad(n) * GP_QC doubles (column q of thread q).  Threads >= GP_QC only help with the factorisation.
// Returns status bits (valid in all threads): 2 = not PD, 16 = a negative variance was clamped.
template <class Ctx>
GP_DEV int gp_predict_curve(Ctx& c, const GpCurveView& cv, const double* theta, const double* stats, const GpQuery& qd,
                            const GpSmem& s, double* V, double* mean, double* stdv, double& lml) {
  bool ok;
  gp_factor(c, cv, theta, s, lml, ok);
  if (!ok) {
    for (int q = c.tid; q < qd.m; q += c.nthr) { mean[q] = NAN; stdv[q] = NAN; }
    return 2;
  }
  const double cst = exp(theta[0]), lt = exp(theta[1]), ll = exp(theta[2]), s2 = exp(theta[3]);
  const double kss = cst + s2;                       // kernel_.diag(X*) (kernels.py:490,1315-1319,1438-1440)
  const double ymean = stats[4], ystd = stats[5];
  double neg[1] = {0.0};
  const int ncol = c.nthr < GP_QC ? c.nthr : GP_QC;
  if (c.tid < ncol) {
    for (int q = c.tid; q < qd.m; q += ncol) {
      double tq; int bq;
      if (qd.n_obs > 0) { bq = q / qd.n_obs; tq = gp_linspace(qd.t_min, qd.t_max, qd.n_obs, q - bq * qd.n_obs); }
      else { tq = qd.qt[q]; bq = qd.qband[q]; }
      const double uq = ((tq - stats[0]) / stats[1]) / lt;   // ss.transform then X / length_scale
      const double vq = cv.lam[bq] / ll;
      double m, v2;
      gp_predict_column(s, cv.n, cst, uq, vq, V + c.tid, GP_QC, m, v2);
      double var = kss - v2;
      if (var < 0.0) { var = 0.0; neg[0] = 1.0; }      // _gpr.py:485-491
      mean[q] = ystd * m + ymean;                      // _gpr.py:450
      stdv[q] = sqrt(var * (ystd * ystd));             // _gpr.py:494,500
    }
  }
  gp_block_sum(c, neg, s.red);
  return neg[0] > 0.0 ? 16 : 0;
}
