// Block-level FP64 GP algebra: one CTA owns one light curve's N x N kernel matrix in shared memory.
//
// Functions here are written against a tiny "block context" (tid, nthr, sync()) so the very same code
// can be executed by a CUDA thread block (DevCtx, gp_kernels.cu) and, for tests without a GPU, by a
// pool of host threads with a barrier (tests/_host/host_gp_block.cpp).  Nothing in here is a CPU
// fallback of the product: the host build exists only under tests/.
//
// What is computed (reference lines: sklearn/gaussian_process/_gpr.py:583-651, kernels.py:1558-1585):
//   K   = c * exp(-1/2 |x_i/l - x_j/l|^2) + (s2 + alpha_i) I            (a4)
//   L   = chol(K)                       blocked right-looking, NB = 4    (a5, LAPACK dpotrf in the reference)
//   X   = L^-1                          blocked in-place triangular inverse (replaces the identity-RHS dpotrs)
//   a   = X^T (X y), LML = -1/2 y.a - sum log L_ii - N/2 log 2pi
//   K^-1 = X^T X is never stored: each 4x4 tile is reduced straight into the gradient traces
//   g_k = 1/2 sum_ij (a_i a_j - K^-1_ij) dK_ij/dtheta_k                  (Eq. 5.9)
//
// Storage: A is column-major with leading dimension ld (odd, so rows and columns are both conflict-free
// for 64-bit accesses).  The strict upper triangle keeps K for the gradient pass while the lower
// triangle is overwritten by L and then L^-1.  N is padded to a multiple of 4 (np) with an identity block,
// and four extra rows np..np+3 are appended below the square: row np carries y^T, so the panel solves and
// trailing updates of the Cholesky turn it into z^T = (L^-1 y)^T for free (bordered factorisation).
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

#define GP_NB 4
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
  double* A;      // np columns of ld >= np + 4 rows
  double* u;      // np   x_i0 / l_t
  double* v;      // np   x_i1 / l_lambda
  double* y;      // np
  double* al;     // np
  double* alpha;  // np   K^-1 y
  double* z;      // np   L^-1 y
  double* dinv;   // (np/4) * 16   inverses of the 4x4 diagonal blocks of L
  double* panel;  // np * 4
  double* red;    // reduction scratch, >= max(nthr, 64)
  int* flag;      // [0] = Cholesky breakdown
  int np, ld;
};

GP_HD int gp_pad(int n) { return (n + GP_NB - 1) / GP_NB * GP_NB; }
GP_HD int gp_ld(int np) { return (np + GP_NB) | 1; }

// number of doubles of shared memory needed for a curve of n observations
GP_HD size_t gp_smem_doubles(int n, int nthr) {
  size_t np = (size_t)gp_pad(n), ld = (size_t)gp_ld((int)np);
  size_t red = nthr > 64 ? nthr : 64;
  return np * ld + 6 * np + (np / GP_NB) * 16 + np * GP_NB + red + 2;
}

GP_DEV GpSmem gp_carve(double* base, int n, int nthr) {
  GpSmem s;
  s.np = gp_pad(n);
  s.ld = gp_ld(s.np);
  double* p = base;
  s.A = p; p += (size_t)s.np * s.ld;
  s.u = p; p += s.np;
  s.v = p; p += s.np;
  s.y = p; p += s.np;
  s.al = p; p += s.np;
  s.alpha = p; p += s.np;
  s.z = p; p += s.np;
  s.dinv = p; p += (s.np / GP_NB) * 16;
  s.panel = p; p += s.np * GP_NB;
  s.red = p; p += (nthr > 64 ? nthr : 64);
  s.flag = (int*)p;
  return s;
}

// ---------------------------------------------------------------------------------------------- reductions
#if defined(__CUDA_ARCH__)
template <class Ctx, int K>
GP_DEV void gp_block_sum(Ctx& c, double (&v)[K], double* red) {
  // column reductions by warp shuffles, then one shared-memory hop across warps
#pragma unroll
  for (int k = 0; k < K; ++k) {
    double x = v[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    v[k] = x;
  }
  const int lane = c.tid & 31, warp = c.tid >> 5, nwarp = (c.nthr + 31) >> 5;
  c.sync();
  if (lane == 0)
    for (int k = 0; k < K; ++k) red[warp * K + k] = v[k];
  c.sync();
#pragma unroll
  for (int k = 0; k < K; ++k) {
    double x = 0.0;
    for (int w = 0; w < nwarp; ++w) x += red[w * K + k];
    v[k] = x;
  }
  c.sync();
}
#else
template <class Ctx, int K>
GP_DEV void gp_block_sum(Ctx& c, double (&v)[K], double* red) {
  for (int k = 0; k < K; ++k) {
    c.sync();
    red[c.tid] = v[k];
    c.sync();
    double x = 0.0;
    for (int t = 0; t < c.nthr; ++t) x += red[t];
    v[k] = x;
  }
  c.sync();
}
#endif

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

// K into the lower triangle (to be factored) and, if keep_upper, a copy into the strict upper triangle.
template <class Ctx>
GP_DEV void gp_assemble(Ctx& c, const GpSmem& s, int n, double cst, double s2, bool keep_upper) {
  const int np = s.np, ld = s.ld;
  for (int idx = c.tid; idx < np * np; idx += c.nthr) {
    const int j = idx / np, i = idx - j * np;
    if (i > j) {
      double val = 0.0;
      if (i < n) {
        const double du = s.u[i] - s.u[j], dv = s.v[i] - s.v[j];
        val = cst * exp(-0.5 * (du * du + dv * dv));
      }
      s.A[i + j * ld] = val;
      if (keep_upper) s.A[j + i * ld] = val;
    } else if (i == j) {
      s.A[i + i * ld] = (i < n) ? (cst + s2) + s.al[i] : 1.0;   // White on the diagonal, then K[diag] += alpha (_gpr.py:589)
    }
  }
  // border rows: y^T then three zero rows
  for (int idx = c.tid; idx < np * GP_NB; idx += c.nthr) {
    const int j = idx / GP_NB, r = idx - j * GP_NB;
    s.A[(np + r) + j * ld] = (r == 0) ? s.y[j] : 0.0;
  }
}

// ---------------------------------------------------------------------------------------------- Cholesky
// 4x4 diagonal block: a (lower) -> L in place, li = L^-1 (lower, row-major [r*4+c]).  false on a non-positive pivot.
GP_DEV bool gp_chol4(double a[4][4], double* li) {
  bool ok = true;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    double d = a[j][j];
#pragma unroll
    for (int k = 0; k < j; ++k) d -= a[j][k] * a[j][k];
    if (!(d > 0.0)) { ok = false; d = 1.0; }
    d = sqrt(d);
    a[j][j] = d;
    const double r = 1.0 / d;
#pragma unroll
    for (int i = j + 1; i < 4; ++i) {
      double v = a[i][j];
#pragma unroll
      for (int k = 0; k < j; ++k) v -= a[i][k] * a[j][k];
      a[i][j] = v * r;
    }
  }
  // inverse of the lower-triangular 4x4
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    li[j * 4 + j] = 1.0 / a[j][j];
#pragma unroll
    for (int i = j + 1; i < 4; ++i) {
      double v = 0.0;
#pragma unroll
      for (int k = j; k < i; ++k) v -= a[i][k] * li[k * 4 + j];
      li[i * 4 + j] = v / a[i][i];
    }
#pragma unroll
    for (int i = 0; i < j; ++i) li[i * 4 + j] = 0.0;
  }
  return ok;
}

GP_DEV void gp_tile_index(int q, int& a, int& b) {
  a = (int)((sqrtf(8.0f * (float)q + 1.0f) - 1.0f) * 0.5f);
  while (a * (a + 1) / 2 > q) --a;
  while ((a + 1) * (a + 2) / 2 <= q) ++a;
  b = q - a * (a + 1) / 2;
}

// Blocked right-looking Cholesky of the lower triangle of A (np x np, np % 4 == 0).  Sets flag on breakdown.
template <class Ctx>
GP_DEV void gp_cholesky(Ctx& c, const GpSmem& s) {
  const int np = s.np, ld = s.ld, nblk = np / GP_NB;
  double* A = s.A;
  for (int kb = 0; kb < nblk; ++kb) {
    const int j0 = kb * GP_NB;
    // (A) diagonal block: factor + invert (one thread; the dependent sqrt/div chain is the critical path anyway)
    if (c.tid == 0) {
      double a[4][4];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int cc = 0; cc <= r; ++cc) a[r][cc] = A[(j0 + r) + (j0 + cc) * ld];
      if (!gp_chol4(a, s.dinv + kb * 16)) s.flag[0] = 1;
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int cc = 0; cc <= r; ++cc) A[(j0 + r) + (j0 + cc) * ld] = a[r][cc];
    }
    c.sync();
    const int rem = np - j0;   // rows below the diagonal block, border rows included
    // (B) panel: rows below solve x L_D^T = p by forward substitution
    {
      const double l00 = A[(j0) + (j0) * ld], l10 = A[(j0 + 1) + (j0) * ld], l11 = A[(j0 + 1) + (j0 + 1) * ld];
      const double l20 = A[(j0 + 2) + (j0) * ld], l21 = A[(j0 + 2) + (j0 + 1) * ld], l22 = A[(j0 + 2) + (j0 + 2) * ld];
      const double l30 = A[(j0 + 3) + (j0) * ld], l31 = A[(j0 + 3) + (j0 + 1) * ld], l32 = A[(j0 + 3) + (j0 + 2) * ld];
      const double l33 = A[(j0 + 3) + (j0 + 3) * ld];
      const double r0 = 1.0 / l00, r1 = 1.0 / l11, r2 = 1.0 / l22, r3 = 1.0 / l33;
      for (int i = c.tid; i < rem; i += c.nthr) {
        double* row = A + (j0 + GP_NB + i) + j0 * ld;
        const double x0 = row[0] * r0;
        const double x1 = (row[ld] - x0 * l10) * r1;
        const double x2 = (row[2 * ld] - x0 * l20 - x1 * l21) * r2;
        const double x3 = (row[3 * ld] - x0 * l30 - x1 * l31 - x2 * l32) * r3;
        row[0] = x0; row[ld] = x1; row[2 * ld] = x2; row[3 * ld] = x3;
      }
    }
    c.sync();
    // (C) trailing update A22 -= P P^T on 4x4 register tiles (lower tiles only; the upper triangle holds K)
    {
      // square part: T x T lower tiles; border: one more tile-row (ta = T) over the T remaining tile-columns
      const int T = rem / GP_NB - 1, ntile = T * (T + 1) / 2 + T;
      const double* P = A + (j0 + GP_NB) + j0 * ld;   // P[i + k*ld], i in [0,rem), k in [0,4)
      for (int q = c.tid; q < ntile; q += c.nthr) {
        int ta, tb;
        gp_tile_index(q, ta, tb);
        double pa[4][4], pb[4][4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
          for (int r = 0; r < 4; ++r) { pa[r][k] = P[(4 * ta + r) + k * ld]; pb[r][k] = P[(4 * tb + r) + k * ld]; }
        double* C = A + (j0 + GP_NB + 4 * ta) + (j0 + GP_NB + 4 * tb) * ld;
#pragma unroll
        for (int cc = 0; cc < 4; ++cc)
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            if (ta != tb || r >= cc) {
              double acc = C[r + cc * ld];
#pragma unroll
              for (int k = 0; k < 4; ++k) acc -= pa[r][k] * pb[cc][k];
              C[r + cc * ld] = acc;
            }
          }
      }
    }
    c.sync();
  }
}

// ---------------------------------------------------------------------------------------------- triangular inverse
// In place: lower triangle L -> X = L^-1 (blocked, last block column first).  Uses s.dinv from gp_cholesky.
template <class Ctx>
GP_DEV void gp_trtri(Ctx& c, const GpSmem& s) {
  const int np = s.np, ld = s.ld, nblk = np / GP_NB;
  double* A = s.A;
  for (int kb = nblk - 1; kb >= 0; --kb) {
    const int j0 = kb * GP_NB;
    const int rem = np - j0 - GP_NB;
    const double* D = s.dinv + kb * 16;
    // (1) panel <- -(A21 * D)   (right-multiply by the inverse of the diagonal block), staged in scratch
    for (int i = c.tid; i < rem; i += c.nthr) {
      const double* row = A + (j0 + GP_NB + i) + j0 * ld;
      const double p0 = row[0], p1 = row[ld], p2 = row[2 * ld], p3 = row[3 * ld];
      s.panel[i * 4 + 0] = -(p0 * D[0] + p1 * D[4] + p2 * D[8] + p3 * D[12]);
      s.panel[i * 4 + 1] = -(p1 * D[5] + p2 * D[9] + p3 * D[13]);
      s.panel[i * 4 + 2] = -(p2 * D[10] + p3 * D[14]);
      s.panel[i * 4 + 3] = -(p3 * D[15]);
    }
    c.sync();
    // (2) A21 <- X22 * panel  (X22 = already inverted trailing block), diagonal block <- D
    for (int w = c.tid; w < rem * 4; w += c.nthr) {
      const int cc = w / rem, i = w - cc * rem;
      const double* xrow = A + (j0 + GP_NB + i) + (j0 + GP_NB) * ld;   // X22[i][k] at xrow[k*ld]
      double acc = 0.0;
      for (int k = 0; k <= i; ++k) acc += xrow[k * ld] * s.panel[k * 4 + cc];
      A[(j0 + GP_NB + i) + (j0 + cc) * ld] = acc;
    }
    if (c.tid < 16) {
      const int r = c.tid >> 2, cc = c.tid & 3;
      if (r >= cc) A[(j0 + r) + (j0 + cc) * ld] = D[r * 4 + cc];
    }
    c.sync();
  }
}

// After gp_cholesky: z^T sits in border row np.  Copies it to s.z and returns this thread's share of z.z (= y^T K^-1 y) and of
// sum log L_ii.  Must run before gp_trtri overwrites the diagonal.
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

// alpha = X^T z  (X = L^-1 in the lower triangle)
template <class Ctx>
GP_DEV void gp_alpha_from_inverse(Ctx& c, const GpSmem& s) {
  const int np = s.np, ld = s.ld;
  for (int j = c.tid; j < np; j += c.nthr) {
    double acc = 0.0;
    for (int i = j; i < np; ++i) acc += s.A[i + j * ld] * s.z[i];
    s.alpha[j] = acc;
  }
  c.sync();
}

// Fused K^-1 = X^T X (4x4 tiles) and gradient traces.  tr[0] = sum_{i>j} W_ij K_ij, tr[1]/tr[2] the same weighted by the
// squared scaled time / wavelength differences, tr[3] = sum_i W_ii  (W = alpha alpha^T - K^-1).
template <class Ctx>
GP_DEV void gp_inverse_traces(Ctx& c, const GpSmem& s, int n, double tr[4]) {
  const int np = s.np, ld = s.ld, T = np / GP_NB, ntile = T * (T + 1) / 2;
  const double* A = s.A;
  tr[0] = tr[1] = tr[2] = tr[3] = 0.0;
  for (int q = c.tid; q < ntile; q += c.nthr) {
    int ta, tb;
    gp_tile_index(q, ta, tb);
    const int i0 = 4 * ta, j0 = 4 * tb;
    double acc[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) acc[r][cc] = 0.0;
    // head: rows k inside the diagonal block of tile-row ta need the triangular mask
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const int k = i0 + kk;
      double xi[4], xj[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) xi[r] = (kk >= r) ? A[k + (i0 + r) * ld] : 0.0;
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) xj[cc] = (ta != tb || kk >= cc) ? A[k + (j0 + cc) * ld] : 0.0;
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) acc[r][cc] += xi[r] * xj[cc];
    }
    for (int k = i0 + 4; k < np; ++k) {
      double xi[4], xj[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) xi[r] = A[k + (i0 + r) * ld];
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) xj[cc] = A[k + (j0 + cc) * ld];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) acc[r][cc] += xi[r] * xj[cc];
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int i = i0 + r;
      const double ai = s.alpha[i], ui = s.u[i], vi = s.v[i];
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        const int j = j0 + cc;
        const double w = ai * s.alpha[j] - acc[r][cc];
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
    }
  }
}

// ---------------------------------------------------------------------------------------------- one LML + gradient evaluation
// theta = log[c, l_t, l_lambda, s2].  Result valid in every thread.  ok=false <=> K not positive definite.
template <class Ctx>
GP_DEV void gp_eval_lml_grad(Ctx& c, const GpCurveView& cv, const double* theta, const GpSmem& s, double& lml, double grad[4], bool& ok) {
  const double cst = exp(theta[0]), lt = exp(theta[1]), ll = exp(theta[2]), s2 = exp(theta[3]);
  gp_load_curve(c, cv, s, lt, ll);
  c.sync();
  gp_assemble(c, s, cv.n, cst, s2, true);
  c.sync();
  gp_cholesky(c, s);
  ok = (s.flag[0] == 0);
  if (!ok) {
    lml = -INFINITY;
    grad[0] = grad[1] = grad[2] = grad[3] = 0.0;
    c.sync();
    return;
  }
  double v[6];
  gp_collect_z(c, s, v[4], v[5]);
  gp_trtri(c, s);
  gp_alpha_from_inverse(c, s);
  gp_inverse_traces(c, s, cv.n, v);
  gp_block_sum(c, v, s.red);
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
      a = a / A[i + i * ld];
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
// noise diagonal (fulu/gp_aug.py:64,72-73).  Serial helper used by one thread per statistic is enough: N is small and this
// runs once per curve, not once per evaluation.
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
  double w[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};   // centred sums: t, t^2, l, l^2, f^2
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
