// Block-level FP64 GP algebra: one CTA owns one light curve's kernel matrix, packed as 8x8 tiles in shared memory.
//
// Functions here are written against a small "block context" (tid/lane/warp, sync(), warp_sync(), mma(), shfl_xor()) so the
// very same code is executed by a CUDA thread block (DevCtx in gp_kernels.cu: __syncthreads, mma.sync.m8n8k4.f64, shuffles)
// and, for tests without a GPU, by host threads with barriers standing in for the warp/block primitives
// (tests/_host/host_gp_block.cpp).  Nothing in here is a CPU fallback of the product: the host build exists only under tests/.
//
// What is computed (reference lines: sklearn/gaussian_process/_gpr.py:583-651, kernels.py:1558-1585):
//   K   = c * exp(-1/2 |x_i/l - x_j/l|^2) + (s2 + alpha_i) I            (a4)
//   L   = chol(K)                       blocked left-looking, NB = 8     (a5; LAPACK dpotrf in the reference)
//   X   = L^-1                          blocked in-place triangular inverse (replaces the identity-RHS dpotrs)
//   z   = L^-1 y, a = X^T z;  LML = -1/2 z.z - sum log L_ii - N/2 log 2pi
//   K^-1 = X^T X is never stored: each 8x8 tile is reduced straight into the gradient traces
//   g_k = 1/2 sum_ij (a_i a_j - K^-1_ij) dK_ij/dtheta_k                  (Eq. 5.9)
//
// Every O(N^3) piece runs on 8x8 tiles through the FP64 tensor-core instruction (DMMA m8n8k4; on B200 it sustains the full
// FP64 rate with an eighth of the issue slots of a DFMA loop, measured 37.2 vs 33.9 TFLOP/s).  The diagonal 8x8 blocks are
// factored by one warp holding the block in accumulator layout (two doubles per lane, shuffles for the pivot column), which
// produces the block's inverse X_D = L_D^-1 in the same sweep; X_D is what gets stored on the diagonal, so the panel solve is a
// DMMA product P X_D^T, the triangular inverse finds its diagonal blocks done, and no thread ever holds a whole 8x8 block.
//
// Storage (v5).  Only the lower triangle exists: tile (I, J), J <= I < T lives at 64 (I (I + 1) / 2 + J) doubles, followed by a
// scratch column of T tiles (the triangular inverse's W).  The matrix that is factored is the BORDERED one, of order N + 1,
// padded with an identity block to np = 8 T >= N + 1: row N carries y^T (and, in the last diagonal tile, column N carries y),
// its diagonal entry acts as 1.  So M = [[L, 0], [z^T, 1]] and
//   * the panel solves and trailing updates of the Cholesky turn row N into z^T = (L^-1 y)^T for free,
//   * the triangular inverse of M has -a^T = -(X^T z)^T in row N, so alpha = K^-1 y costs nothing extra either,
// and the border costs no tiles of its own unless N is a multiple of 8 (v3/v4 kept it as a 14th tile-row at N = 100: 12 % of
// all tile operations for one useful row in eight).
// K itself is not kept: the gradient pass recomputes K_ij (one exp) where it meets K^-1_ij.  That halves the footprint
// (N = 100: 55 KB instead of 100 KB), which is what lets four curves be resident per SM instead of two - the evaluation is
// latency bound (serial 8x8 diagonal factorisations, barriers), so residency is throughput.
// Inside a tile element (r, c) sits at 8 c + (r ^ swz(c)), swz(c) = (c & 2 ? 6 : 0) ^ (c & 4): every fragment access pattern used
// below - (g, t), (g, 4+t), (perm g, t), (t, g), (t, perm g) and their +4 variants, lane = 4 g + t - touches 16 distinct
// 64-bit banks per half-warp (checked exhaustively in tests/test_block_host.py::test_tile_swizzle_is_conflict_free).
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define GP_DEV __device__ __forceinline__
#define GP_HD __host__ __device__ __forceinline__
#else
#define GP_DEV inline
#define GP_HD inline
#endif

// The inner DMMA loops run 1 .. T times (mean ~T/3).  Left to the compiler each became a 16/8/4/2/1 unrolling cascade whose branchy
// remainder logic ran once per tile - as many instructions as the products themselves in a kernel whose warps are bound by their own
// instruction stream (63k warp-instructions per evaluation, 55 % of them integer and control).  Rolled: 178.6 -> 169.7 ms for the
// 10,000-curve fit (+5.3 %); unroll 2 / 4: 170.0 / 170.6; rolling the loops over tiles as well: no change; fetching the next tile's
// fragments by hand before issuing the products: 175.1 (profiles/r02_unroll_sweep.md).
#define GP_PRAGMA_(x) _Pragma(#x)
#define GP_PRAGMA(x) GP_PRAGMA_(x)
#ifndef GP_UNROLL_INNER
#define GP_UNROLL_INNER 1
#endif
#define GP_UNROLL_K GP_PRAGMA(unroll GP_UNROLL_INNER)
// Three more instruction savers of the same kind (A/B switches; all on): the trace pass takes PAIRS of neighbouring tiles that share
// their A fragments (-0.6 %), tiles of real observations left of the diagonal skip every mask of the assembly (-1.7 %) and of the trace
// epilogue (-0.7 %): 169.7 -> 164.0 ms.  Dealing the rows of the triangular inverse to the warps in serpentine order: +0.6 %, dropped.
#ifndef GP_LEAN_MAX_P
#define GP_LEAN_MAX_P 5   // families with up to this many hyperparameters get the grouped / mask-free paths (GpKern::kLeanPaths)
#endif
#ifndef GP_TRACE_GROUP
#define GP_TRACE_GROUP 2   // tiles of a tile-row taken together in the trace pass: 2, or 4 (measured: 209 ms against 158 - the sixteen
                           // accumulators and four K pairs of a group of four no longer fit the registers of a 128-thread CTA at 4 CTAs per SM)
#endif
#ifndef GP_TRACE_PAIRS
#define GP_TRACE_PAIRS 1
#endif
#ifndef GP_ASM_FAST
#define GP_ASM_FAST 1
#endif
#ifndef GP_TRACE_FAST
#define GP_TRACE_FAST 1
#endif
#define GP_NB 8
#define GP_LOG_2PI 1.8378770664093453
#define GP_RED 192

// Rolled where the operands come from shared memory; where they come from a slab in global memory (GP_PLACE_SLAB, and the forward
// substitution of the tiled path) the compiler's unrolling keeps a dozen loads in flight and stays: the N <= 303 class of config #4
// ran 610 ms unrolled, 717 ms rolled.  A block context may say which it wants (kRolled: factorisation / inverse / traces,
// kRolledPredict: forward substitution); contexts that say nothing (the host harness) get the rolled loops.
template <class Ctx, class = void>
struct GpLoops { static constexpr bool rolled = true, rolled_predict = true; };
template <class Ctx>
struct GpLoops<Ctx, decltype((void)Ctx::kRolled, void())> { static constexpr bool rolled = Ctx::kRolled, rolled_predict = Ctx::kRolledPredict; };

template <int N>
struct GpInt { static constexpr int value = N; };

template <bool kRolled, class F>
GP_DEV void gp_kloop(int a, int b, F&& body) {
  if constexpr (kRolled) {
    GP_UNROLL_K
    for (int k = a; k < b; ++k) body(k);
  } else {
    for (int k = a; k < b; ++k) body(k);
  }
}

struct GpCurveView {
  int n;               // observations
  const double* xs;    // [n] standardised time              (StandardScaler column 0)
  const int32_t* band; // [n] band index
  const double* y;     // [n] normalised flux
  const double* al;    // [n] noise diagonal (regressor alpha)
  const double* lam;   // [n_bands] standardised log10 wavelength (StandardScaler column 1 per band)
};

struct alignas(16) GpPair {
  double x, y;
};

struct GpSmem {
  double* A;      // packed tiles: T (T + 1) / 2 lower tiles, then (scratch layout) T scratch tiles
  double* W;      // the scratch column (the triangular inverse's W tiles), or null: the scratch-free layout (gp_carve)
  double* stage;  // 2 np doubles where gp_load_curve stages the noise diagonal and y for the assembly: the scratch column, or behind the vectors
  double* u;      // np   x_i0 / l_t
  double* v;      // np   x_i1 / l_lambda
  double* red;    // reduction scratch, GP_RED doubles (32 warps x 6 values)
  double* tab;    // [0..15] c 2^(j/16), [16..31] 2^(j/16): the exponential's table (gp_exp_tab), rebuilt per evaluation
  double* kslab;  // optional: this CTA's slab in global memory (L2) that keeps the assembled K tiles, row-major 8x8, for the gradient
                  // pass (null: K_ij is recomputed there)
  int* flag;      // [0] = Cholesky breakdown
  int np, T;
};

GP_HD int gp_pad(int n) { return (n + GP_NB) / GP_NB * GP_NB; }   // order of the bordered matrix (n + 1), padded to a multiple of 8
GP_HD size_t gp_mat_doubles(int np) { const size_t T = (size_t)np / GP_NB; return 64 * (T * (T + 1) / 2 + T); }
GP_HD size_t gp_vec_doubles(int np) { return 2 * (size_t)np + GP_RED + 32 + 2; }
// The scratch-free layout (length classes where it buys one more resident CTA per SM: N = 104 .. 127, 136 .. 159): no scratch column -
// the triangular inverse forms its W tiles in place and holds a block column's rows in registers across a barrier (gp_trtri) - and the
// staging of the noise diagonal and y moves behind the vectors.
GP_HD size_t gp_mat_doubles_ns(int np) { const size_t T = (size_t)np / GP_NB; return 64 * (T * (T + 1) / 2); }
GP_HD size_t gp_vec_doubles_ns(int np) { return gp_vec_doubles(np) + 2 * (size_t)np; }
// doubles of shared memory needed for a curve of n observations (matrix + vectors)
GP_HD size_t gp_smem_doubles(int n) {
  const int np = gp_pad(n);
  return gp_mat_doubles(np) + gp_vec_doubles(np);
}

// vectors at `vec`, matrix at `mat` (shared memory, or a per-CTA global slab for long curves)
GP_HD GpSmem gp_carve(double* mat, double* vec, int n, bool scratch = true) {
  GpSmem s;
  s.np = gp_pad(n);
  s.T = s.np / GP_NB;
  s.A = mat;
  double* p = vec;
  s.u = p; p += s.np;
  s.v = p; p += s.np;
  s.red = p; p += GP_RED;
  s.tab = p; p += 32;
  s.kslab = nullptr;
  s.flag = (int*)p;
  s.W = scratch ? mat + 64 * (s.T * (s.T + 1) / 2) : nullptr;   // (= gp_tile(T, 0))
  s.stage = scratch ? s.W : p + 2;
  return s;
}

// ---------------------------------------------------------------------------------------------- tile addressing
GP_HD int gp_swz(int c) { return ((c & 2) ? 6 : 0) ^ (c & 4); }
GP_HD int gp_el(int r, int c) { return c * 8 + (r ^ gp_swz(c)); }           // element (r, c) inside a tile
GP_HD int gp_tile(int I, int J) { return 64 * (I * (I + 1) / 2 + J); }      // tile (I, J), J <= I; scratch column: I = T
GP_HD int gp_at(int i, int j) { return gp_tile(i >> 3, j >> 3) + gp_el(i & 7, j & 7); }   // matrix element (i, j), j <= i
// column permutation of the accumulator fragment: register c0 of lane (g, t) holds tile column t, c1 holds column 4 + t
// (instead of 2t, 2t+1), so that accumulator loads/stores are conflict free too.  B-fragment lane g must then supply
// tile column gp_perm(g).
GP_HD int gp_perm(int g) { return (g & 1) ? 4 + (g >> 1) : (g >> 1); }

// per-lane fragment offsets inside a tile (lane = 4 g + t)
struct GpFrag {
  int a0, a1;   // (g, t), (g, 4 + t)          A operand [m = g][k]; also the permuted accumulator
  int b0, b1;   // (perm g, t), (perm g, 4 + t)  B operand = rows of a tile (B[k][n] = tile[n][k])
  int t0, t1;   // (t, g), (4 + t, g)          transposed A / native B = columns of a tile
  int p0, p1;   // (t, perm g), (4 + t, perm g)  B operand = a tile as is, permuted output
};
GP_HD GpFrag gp_frag(int lane) {
  const int g = lane >> 2, t = lane & 3, pg = gp_perm(g);
  GpFrag f;
  f.a0 = gp_el(g, t); f.a1 = gp_el(g, 4 + t);
  f.b0 = gp_el(pg, t); f.b1 = gp_el(pg, 4 + t);
  f.t0 = gp_el(t, g); f.t1 = gp_el(4 + t, g);
  f.p0 = gp_el(t, pg); f.p1 = gp_el(4 + t, pg);
  return f;
}

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
// exp(x) for x <= 0, branch free so that several evaluations interleave.  x = (16 k + j) ln2/16 + r, |r| <= ln2/32:
//   exp(x) = 2^k * 2^(j/16) * exp(r),  n = 16 k + j = rint(x 16/ln2) read from the low word of x 16/ln2 + 1.5 2^52,
//   r = x - n ln2/16 with a two-part constant, exp(r) a degree-7 Taylor polynomial (remainder 1.2e-18), 2^k by exponent arithmetic.
// The 16-entry table sits in shared memory (128 bytes = every bank once: any 32 lookups are conflict free) and may carry a
// constant factor (the kernel amplitude c), which saves the multiplication by it; arguments below xmin give 0 (xmin keeps
// 2^k * tab[j] clear of the subnormal range).  12 FP64 operations where the table-free degree-13 version needed 18 - the
// exponential is a quarter of the FP64 pipe's time in an evaluation.  Error <= 1.6 ulp over [-690, 0]
// (tests/test_block_host.py::test_exp_accuracy; the library exp is 0.5 ulp).
#if defined(__CUDACC__)
#define GP_CONST __constant__
#else
#define GP_CONST static const
#endif
// constants live in constant memory on the device: FP64 instructions take constant-bank operands directly, while immediates
// would each cost two register moves
GP_CONST double gp_exp2_tab[16] = {1.0, 1.0442737824274138, 1.0905077326652577, 1.1387886347566916, 1.189207115002721, 1.241857812073484,
                                   1.2968395546510096, 1.3542555469368927, 1.4142135623730951, 1.4768261459394993, 1.5422108254079407,
                                   1.6104903319492543, 1.681792830507429, 1.7562521603732995, 1.8340080864093424, 1.9152065613971474};
GP_CONST double gp_exp_c[11] = {23.083120654223414, 6755399441055744.0, -0.04332169877307024, -1.1926343307941173e-11,
                                0.0001984126984126984, 0.001388888888888889, 0.008333333333333333, 0.041666666666666664,
                                0.16666666666666666, 0.5, 1.0};
#define GP_EXP_XMIN (-690.0)

GP_DEV double gp_exp_tab(double x, const double* tab, double xmin) {
  const double t = fma(x, gp_exp_c[0], gp_exp_c[1]);   // low word of t = n
  const double kf = t - gp_exp_c[1];
  double r = fma(kf, gp_exp_c[2], x);
  r = fma(kf, gp_exp_c[3], r);
  double p = gp_exp_c[4];
#pragma unroll
  for (int i = 5; i <= 9; ++i) p = fma(p, r, gp_exp_c[i]);
  p = fma(p, r, gp_exp_c[10]);
  p = fma(p, r, gp_exp_c[10]);
#if defined(__CUDA_ARCH__)
  const int n = __double2loint(t);
  p *= tab[n & 15];
  const double res = __hiloint2double(__double2hiint(p) + ((n >> 4) << 20), __double2loint(p));
#else
  int64_t tb, pb;
  memcpy(&tb, &t, 8);
  const int32_t n = (int32_t)(tb & 0xffffffff);
  p *= tab[n & 15];
  memcpy(&pb, &p, 8);
  pb += (int64_t)(n >> 4) << 52;
  double res;
  memcpy(&res, &pb, 8);
#endif
  return x < xmin ? 0.0 : res;
}

// ---------------------------------------------------------------------------------------------- covariance families
// The covariance is   K = c * [M1(r_1)] * RBF([l_t, l_lambda]) + [EX(r_e)] + s2 I   with optional isotropic factors:
//   M1  a Matern factor multiplying the anisotropic RBF   (sklearn: ConstantKernel * Matern * RBF, kernels.py:964-971)
//   EX  an added isotropic Matern or RationalQuadratic    (sklearn: ... + Matern() / + RationalQuadratic(), kernels.py:866-871)
// theta (log-transformed, canonical device order) = [c, (l_1), l_t, l_lambda, (EX: l | alpha, l), s2]; the host permutes between
// sklearn's tree order (kernels.py:739-752) and this one.  Families are compile-time (template arguments), so the default
// family's code is exactly what it was as a dedicated path and the others cost nothing when unused.
// Coordinates live in shared memory divided by the RBF length scales (u = x0 / l_t, v = x1 / l_lambda, kernels.py:1561); an
// isotropic factor with length l sees r^2 = (l_t / l)^2 du^2 + (l_lambda / l)^2 dv^2.
enum { GP_F_NONE = 0, GP_F_MAT05 = 1, GP_F_MAT15 = 2, GP_F_MAT25 = 3, GP_F_RQ = 4 };

GP_DEV double gp_rsqrt(double d);

// sqrt(q) for q >= 0 through the lean reciprocal square root (zero and denormal-range arguments give 0: the factor is 1 there)
GP_DEV double gp_sqrt_pos(double q) { return q > 1.0e-290 ? q * gp_rsqrt(q) : 0.0; }

template <int KIND>
struct GpIso {
  double au, av;   // r^2 = au du^2 + av dv^2
  double alpha;    // RationalQuadratic scale mixture
  const double* tab;   // 2^(j/16) (GpSmem::tab + 16)
  GP_DEV void init(double lt, double ll, double l, double al) {
    const double ru = lt / l, rv = ll / l;
    au = ru * ru; av = rv * rv; alpha = al;
  }
  // value k, dk/dlog(l) (gl) and, for RQ, dk/dlog(alpha) (ga)  (kernels.py:1722-1729,1757-1774 Matern; :1917-1947 RQ)
  template <bool GRAD>
  GP_DEV void eval(double du2, double dv2, double& k, double& gl, double& ga) const {
    const double s = fma(au, du2, av * dv2);
    gl = 0.0; ga = 0.0;
    if constexpr (KIND == GP_F_MAT05) {
      const double r = gp_sqrt_pos(s);
      k = gp_exp_tab(-r, tab, GP_EXP_XMIN);
      if (GRAD) gl = k * r;
    } else if constexpr (KIND == GP_F_MAT15) {
      const double a = gp_sqrt_pos(3.0 * s), e = gp_exp_tab(-a, tab, GP_EXP_XMIN);
      k = (1.0 + a) * e;
      if (GRAD) gl = 3.0 * s * e;
    } else if constexpr (KIND == GP_F_MAT25) {
      const double a = gp_sqrt_pos(5.0 * s), e = gp_exp_tab(-a, tab, GP_EXP_XMIN);
      k = (1.0 + a + a * a / 3.0) * e;
      if (GRAD) gl = (5.0 / 3.0) * s * (a + 1.0) * e;
    } else if constexpr (KIND == GP_F_RQ) {
      const double base = 1.0 + s / (2.0 * alpha), lb = log(base);
      k = gp_exp_tab(-alpha * lb, tab, GP_EXP_XMIN);
      if (GRAD) {
        gl = s * k / base;
        ga = k * (s / (2.0 * base) - alpha * lb);
      }
    } else {
      k = 1.0;
    }
  }
};

template <int F1, int EX>
struct GpKern {
  static_assert(F1 != GP_F_RQ, "the product factor is a Matern");
  static constexpr int NF1 = (F1 == GP_F_NONE) ? 0 : 1;
  static constexpr int NEX = (EX == GP_F_NONE) ? 0 : (EX == GP_F_RQ ? 2 : 1);
  static constexpr int P = 4 + NF1 + NEX;   // hyperparameters
  static constexpr int NTR = P - 1;         // traces over the strictly-lower pairs: every parameter but the noise level
  static constexpr bool kKeepK = (F1 == GP_F_NONE && EX == GP_F_NONE);   // every off-diagonal derivative is K_ij times a weight
  static constexpr bool kLeanPaths = (P <= GP_LEAN_MAX_P);   // mask-free / grouped code paths in the assembly and the trace pass (gp_assemble, gp_inverse_traces)
  double cst, s2, lt, ll;
  double xminc;        // arguments of the amplitude-scaled exponential below this give 0
  const double* tabc;  // c 2^(j/16) (GpSmem::tab)
  GpIso<F1> f1;
  GpIso<EX> ex;

  GP_DEV void init(const double* theta) {
    cst = exp(theta[0]);
    lt = exp(theta[1 + NF1]);
    ll = exp(theta[2 + NF1]);
    s2 = exp(theta[P - 1]);
    xminc = GP_EXP_XMIN + fmax(0.0, -theta[0]);
    if constexpr (NF1 > 0) f1.init(lt, ll, exp(theta[1]), 0.0);
    if constexpr (EX == GP_F_RQ) ex.init(lt, ll, exp(theta[4 + NF1]), exp(theta[3 + NF1]));   // sklearn orders alpha before length_scale
    else if constexpr (NEX > 0) ex.init(lt, ll, exp(theta[3 + NF1]), 0.0);
  }
  // K_ij, i != j, from the squared scaled differences
  GP_DEV double value(double du2, double dv2) const {
    double k = gp_exp_tab(-0.5 * (du2 + dv2), tabc, xminc);   // c exp(.)
    double f, g0, g1;
    if constexpr (NF1 > 0) { f1.template eval<false>(du2, dv2, f, g0, g1); k *= f; }
    if constexpr (NEX > 0) { ex.template eval<false>(du2, dv2, f, g0, g1); k += f; }
    return k;
  }
  // K_ii without the regressor's alpha: every stationary factor is 1 at zero distance, White adds s2 (kernels.py:490,1438-1440)
  GP_DEV double diag() const { return NEX ? (cst + 1.0) + s2 : cst + s2; }
  // tr[k] += w dK_ij / dtheta_k for the NTR parameters that act off the diagonal
  GP_DEV void accumulate(double w, double du2, double dv2, double (&tr)[NTR]) const {
    const double e = gp_exp_tab(-0.5 * (du2 + dv2), tabc, xminc);   // c exp(.)
    double f, gl, ga;
    if constexpr (NF1 > 0) {
      f1.template eval<true>(du2, dv2, f, gl, ga);
      const double wk = w * (e * f);
      tr[0] += wk;
      tr[1] += w * (e * gl);
      tr[2] += wk * du2;
      tr[3] += wk * dv2;
    } else {
      const double wk = w * e;
      tr[0] += wk;
      tr[1] += wk * du2;
      tr[2] += wk * dv2;
    }
    if constexpr (NEX > 0) {
      ex.template eval<true>(du2, dv2, f, gl, ga);
      if constexpr (EX == GP_F_RQ) { tr[3 + NF1] += w * ga; tr[NTR - 1] += w * gl; }
      else tr[3 + NF1] += w * gl;
    }
  }
  // the same with K_ij given (default family only: dK/dlog c = K, dK/dlog l_d = K d_d^2)
  GP_DEV void accumulate_stored(double w, double k, double du2, double dv2, double (&tr)[NTR]) const {
    const double wk = w * k;
    tr[0] += wk;
    tr[1] += wk * du2;
    tr[2] += wk * dv2;
  }
  // gradient of the LML from the pair traces and trd = sum_i W_ii (the 1/2 of Eq. 5.9 cancels against the two triangles)
  GP_DEV void gradient(const double* tr, double trd, double* grad) const {
    grad[0] = tr[0] + 0.5 * cst * trd;
    for (int k = 1; k < NTR; ++k) grad[k] = tr[k];
    grad[P - 1] = 0.5 * s2 * trd;
  }
};
using GpKernDefault = GpKern<GP_F_NONE, GP_F_NONE>;   // C * RBF([l_t, l_lambda]) + White  (fulu/gp_aug.py:29)

// Scaled coordinates and the exponential's tables into shared memory; binds the kernel to the tables.
template <class Ctx, class Kern>
GP_DEV void gp_load_curve(Ctx& c, const GpCurveView& cv, const GpSmem& s, Kern& kn) {
  // X / length_scale (kernels.py:1561) as a multiplication by the reciprocal: an FP64 division is ~25 instructions, two per
  // observation and evaluation; the 1-ulp difference is far below the parity tolerance
  const double ilt = 1.0 / kn.lt, ill = 1.0 / kn.ll;
  kn.tabc = s.tab;
  kn.f1.tab = s.tab + 16;
  kn.ex.tab = s.tab + 16;
  if (c.tid < 32) s.tab[c.tid] = (c.tid < 16 ? kn.cst : 1.0) * gp_exp2_tab[c.tid & 15];
  // the noise diagonal and y ride in the scratch column until the triangular inverse needs it: the assembly then reads them from
  // shared memory instead of waiting on global loads in its diagonal and border tiles
  double* sc = s.stage;
  for (int i = c.tid; i < s.np; i += c.nthr) {
    if (i < cv.n) {
      s.u[i] = cv.xs[i] * ilt;
      s.v[i] = cv.lam[cv.band[i]] * ill;
      sc[i] = cv.al[i];
      sc[s.np + i] = cv.y[i];
    } else {
      s.u[i] = 0.0; s.v[i] = 0.0;
      sc[i] = 0.0; sc[s.np + i] = 0.0;
    }
  }
  if (c.tid == 0) s.flag[0] = 0;
}

// K into the lower tiles, y^T into row n (and y into column n of the last diagonal tile).  One warp per tile; lane (g, t) writes
// (g, t) and (g, 4 + t).
template <class Ctx, class Kern>
GP_DEV void gp_assemble(Ctx& c, const GpCurveView& cv, const GpSmem& s, const Kern& kn) {
  const int T = s.T, n = cv.n, ntile = T * (T + 1) / 2, g = c.lane >> 2, t = c.lane & 3;
  const GpFrag f = gp_frag(c.lane);
  const double* al = s.stage;   // staged by gp_load_curve (in the scratch column, or behind the vectors in the scratch-free layout)
  const double* yv = al + s.np;
  int I = 0, J = c.warp;
  while (J > I) { J -= I + 1; ++I; }
  for (int q = c.warp; q < ntile; q += c.nwarp) {
    double* tp = s.A + 64 * q;
    const int i = 8 * I + g, j0 = 8 * J + t, j1 = j0 + 4;
    const double ui = s.u[i], vi = s.v[i];
    // diagonal tiles are assembled in full (both triangles): the warp-level factorisation keeps them symmetric
    const double du0 = ui - s.u[j0], dv0 = vi - s.v[j0], du1 = ui - s.u[j1], dv1 = vi - s.v[j1];
    double v0 = kn.value(du0 * du0, dv0 * dv0);
    double v1 = kn.value(du1 * du1, dv1 * dv1);
    // tiles above the last tile-row and left of the diagonal hold nothing but K_ij of real observations (the border row n and the
    // padding live in tile-row T - 1): no masks, no diagonal, no border.  (Only for the 4- and 5-parameter families: every extra
    // path repeats the covariance function's code, and at six parameters the evaluation kernel outgrows the instruction cache -
    // rbf_rq ran 865 ms with the extra paths against 650 ms without, 99 KB of SASS against the default kernel's 53 KB.)
    if (GP_ASM_FAST && Kern::kLeanPaths && I < T - 1 && J < I) {
      tp[f.a0] = v0;
      tp[f.a1] = v1;
      if (Kern::kKeepK && s.kslab != nullptr) {
        s.kslab[64 * q + 8 * g + t] = v0;
        s.kslab[64 * q + 8 * g + 4 + t] = v1;
      }
      J += c.nwarp;
      while (J > I) { J -= I + 1; ++I; }
      continue;
    }
    if (i >= n || j0 >= n) v0 = 0.0;
    if (i >= n || j1 >= n) v1 = 0.0;
    // White on the diagonal, then K[diag] += alpha (_gpr.py:589); identity on the padding; the border's own diagonal entry
    // starts at 0 and collects -z.z during the factorisation (its pivot is taken as 1)
    if (i == j0) v0 = (i < n) ? kn.diag() + al[i] : (i == n ? 0.0 : 1.0);
    if (i == j1) v1 = (i < n) ? kn.diag() + al[i] : (i == n ? 0.0 : 1.0);
    // the border: row n = y^T, and its mirror image in the last diagonal tile
    if (i == n) {
      if (j0 < n) v0 = yv[j0];
      if (j1 < n) v1 = yv[j1];
    } else if (i < n) {
      if (j0 == n) v0 = yv[i];
      if (j1 == n) v1 = yv[i];
    }
    tp[f.a0] = v0;
    tp[f.a1] = v1;
    if (Kern::kKeepK && s.kslab != nullptr) {   // a copy of the tile for the gradient pass, which then needs no second exponential
      s.kslab[64 * q + 8 * g + t] = v0;
      s.kslab[64 * q + 8 * g + 4 + t] = v1;
    }
    J += c.nwarp;
    while (J > I) { J -= I + 1; ++I; }
  }
}

// ---------------------------------------------------------------------------------------------- Cholesky
// Warp-level Cholesky of one 8x8 diagonal block held in accumulator layout (lane (g, t): a0 = A[g][t], a1 = A[g][4 + t], the
// block symmetric), together with the inverse of its factor: on exit z0/z1 hold X = L^-1 in the same layout (exact zeros above
// the diagonal) and rdet has been multiplied by prod_j 1 / L_jj.  Right-looking, one pivot column per step:
//   r = rsqrt(A_jj);  l_ij = A_ij r;  A_ic -= l_ij l_cj;        X_j: = Z_j: r;  Z_i: -= l_ij X_j:   (Z starts as I)
// Same arithmetic as a scalar column Cholesky (division free: L_jj = A_jj r is never needed), but every lane carries two
// matrix entries instead of one thread carrying 36, and the pivot column travels by shuffles.
// 1 / sqrt(d) for a pivot d in (1e-300, 1e300): hardware seed (MUFU.RSQ64H, relative error < 2^-20) and one third-order step
// y (1 + e/2 + 3 e^2 / 8), e = 1 - d y^2: four dependent FP64 operations (8 cycles each on B200) where the library rsqrt
// takes 205 cycles (scripts/ubench/lat.cu) - this sits on the longest dependent chain of the evaluation, N pivots in sequence.
GP_DEV double gp_rsqrt(double d) {
#if defined(__CUDA_ARCH__)
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  const double e = fma(-(d * y), y, 1.0);
  return fma(y * e, fma(0.375, e, 0.5), y);
#else
  return 1.0 / sqrt(d);
#endif
}

// log of a product of positive factors without a log per factor: mantissa and exponent are accumulated separately
// (the library log is ~300 cycles; the factorisation folds its pivots in here and takes ONE log at the end)
struct GpLogProd {
  double m;   // in [1, 2)
  int e;
};
GP_DEV void gp_logprod_mul(GpLogProd& lp, double x) {   // x in (1e-300, 1e300)
  const double p = lp.m * x;
#if defined(__CUDA_ARCH__)
  const int hi = __double2hiint(p);
  lp.e += (hi >> 20) - 1023;
  lp.m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(p));
#else
  int ex;
  const double fr = frexp(p, &ex);   // p = fr 2^ex, fr in [0.5, 1)
  lp.m = 2.0 * fr;
  lp.e += ex - 1;
#endif
}
GP_DEV double gp_logprod_value(const GpLogProd& lp) { return log(lp.m) + (double)lp.e * 0.6931471805599453; }

// jb: local index of the border row when this block holds it (kBorder), 8 otherwise.  The border's pivot is taken as 1; what has
// accumulated in its diagonal entry by then is -z.z (it starts at 0), returned in bq.
// This warp is the critical path of the factorisation (the other warps of the block have a third of its instructions per block
// column), and it is bound by its own instruction count - ~57 instructions per pivot in the first version, ~45 now: the column entry
// of the lane's own row comes by a width-4 shuffle with an immediate lane index, the other lane indices are formed once per block,
// the pivots are tested four at a time through the product of their reciprocal roots (a non-positive or non-finite pivot makes it
// NaN, infinite or non-positive), and only the block that holds the border row carries the border's selects.
template <bool kBorder, class Ctx>
GP_DEV bool gp_chol8_warp(Ctx& c, double a0, double a1, double& z0, double& z1, GpLogProd& rdet, int jb, double& bq) {
  const int g = c.lane >> 2, t = c.lane & 3;
  const int s0 = 4 * t, s1 = 4 * t + 16;   // lanes that hold rows t and 4 + t (column offsets are added per pivot)
  z0 = (g == t) ? 1.0 : 0.0;
  z1 = (g == 4 + t) ? 1.0 : 0.0;
  bool ok = true;
  double rprod = 1.0;   // running product of 1 / L_jj, folded into rdet every four pivots (pivots are within (1e-300, 1e300))
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int jt = j & 3;
    const double aj = (j < 4) ? a0 : a1;           // this lane's entry of block column (j & 3) / (4 + (j & 3))
    const double dj = c.shfl(aj, 4 * j + jt);       // A[j][j]
    double d = dj;
    if (kBorder) {
      if (j == jb) { bq = dj; d = 1.0; }
    }
    const double cg = c.shfl4(aj, jt);              // A[g][j]
    const double c0 = c.shfl(aj, s0 + jt);          // A[t][j]
    const double c1 = c.shfl(aj, s1 + jt);          // A[4 + t][j]
    const double zj0 = c.shfl(z0, 4 * j + t);       // Z[j][t]
    const double zj1 = c.shfl(z1, 4 * j + t);       // Z[j][4 + t]
    const double r = gp_rsqrt(d);
    rprod *= r;
    const double lg = cg * r, l0 = c0 * r, l1 = c1 * r, x0 = zj0 * r, x1 = zj1 * r;
    if (g > j) {
      a0 -= lg * l0;
      a1 -= lg * l1;
      z0 -= lg * x0;
      z1 -= lg * x1;
    } else if (g == j) {
      z0 = x0;
      z1 = x1;
    }
    if ((j & 3) == 3) {
      // a bad pivot is only recorded (the evaluation is discarded as "not positive definite"); nothing waits for the test
      ok = ok && (rprod > 0.0) && (rprod < 1.0e300);
      gp_logprod_mul(rdet, rprod);
      rprod = 1.0;
    }
  }
  return ok;
}

// rows below the diagonal tile kb: P_I <- P_I X_D^T, the panel solve x L_D^T = p as a DMMA product
// with the inverted diagonal block.  One tile per warp at a time.
template <class Ctx>
GP_DEV void gp_panel_mult(Ctx& c, const GpSmem& s, const GpFrag& f, int kb) {
  const int T = s.T;
  if (kb + 1 + c.warp >= T) return;
  const double* D = s.A + gp_tile(kb, kb);
  const double b0 = D[f.b0], b1 = D[f.b1];   // B[k][n] = X_D[n][k]
  for (int I = kb + 1 + c.warp; I < T; I += c.nwarp) {
    double* tp = s.A + gp_tile(I, kb);
    double w0 = 0.0, w1 = 0.0;
    c.mma(w0, w1, tp[f.a0], b0);
    c.mma(w0, w1, tp[f.a1], b1);
    tp[f.a0] = w0;
    tp[f.a1] = w1;
  }
}

// Accumulation of finished block columns into one or two tiles of block column k:
//   C(I, k) = C(I, k) - sum_{J0 <= J < J1} L(I, J) L(k, J)^T,
// left in registers in accumulator layout (acc[r][0] = C[g][t], acc[r][1] = C[g][4 + t]).  The two rows share the B fragments
// (L(k, J), read as rows: B[kk][n] = L(k, J)[n][kk]); each row runs two independent DMMA chains.  I2 < 0: one row only.
template <class Ctx>
GP_DEV void gp_chol_accumulate(Ctx& c, const GpSmem& s, const GpFrag& f, int k, int I1, int I2, int J0, int J1, double (&acc)[2][2]) {
  const double* P1 = s.A + gp_tile(I1, 0);
  const double* Pb = s.A + gp_tile(k, J0);
  double d10 = 0.0, d11 = 0.0;
  acc[0][0] = P1[64 * k + f.a0]; acc[0][1] = P1[64 * k + f.a1];
  P1 += 64 * J0;
  if (I2 < 0) {
    gp_kloop<GpLoops<Ctx>::rolled>(J0, J1, [&](int) {
      c.mma(acc[0][0], acc[0][1], -P1[f.a0], Pb[f.b0]);
      c.mma(d10, d11, -P1[f.a1], Pb[f.b1]);
      P1 += 64; Pb += 64;
    });
  } else {
    const double* P2 = s.A + gp_tile(I2, 0);
    double d20 = 0.0, d21 = 0.0;
    acc[1][0] = P2[64 * k + f.a0]; acc[1][1] = P2[64 * k + f.a1];
    P2 += 64 * J0;
    gp_kloop<GpLoops<Ctx>::rolled>(J0, J1, [&](int) {
      const double b0 = Pb[f.b0], b1 = Pb[f.b1];
      c.mma(acc[0][0], acc[0][1], -P1[f.a0], b0);
      c.mma(d10, d11, -P1[f.a1], b1);
      c.mma(acc[1][0], acc[1][1], -P2[f.a0], b0);
      c.mma(d20, d21, -P2[f.a1], b1);
      P1 += 64; P2 += 64; Pb += 64;
    });
    acc[1][0] += d20; acc[1][1] += d21;
  }
  acc[0][0] += d10; acc[0][1] += d11;
}

// rows I = first, first + stride, ... < T of block column k: apply the terms [J0, J1) and store (two rows per pass)
template <class Ctx>
GP_DEV void gp_chol_update_rows(Ctx& c, const GpSmem& s, const GpFrag& f, int k, int first, int stride, int J0, int J1) {
  const int T = s.T;
  double acc[2][2];
  for (int I = first; I < T; I += 2 * stride) {
    const int I2 = (I + stride < T) ? I + stride : -1;
    gp_chol_accumulate(c, s, f, k, I, I2, J0, J1, acc);
    double* C1 = s.A + gp_tile(I, k);
    C1[f.a0] = acc[0][0];
    C1[f.a1] = acc[0][1];
    if (I2 >= 0) {
      double* C2 = s.A + gp_tile(I2, k);
      C2[f.a0] = acc[1][0];
      C2[f.a1] = acc[1][1];
    }
  }
}

// Blocked Cholesky of the bordered matrix, left-looking with one column of lookahead.  At the top of step kb the block columns
// < kb are final and block column kb already carries every term but the last (J = kb - 1).  Then
//   warp 0        applies that term to the diagonal tile and factors it, all in registers (the serial chain of 8 pivots);
//   other warps   apply it to the rows below, then run ahead: block column kb + 1 receives the terms J < kb in one
//                 accumulation per tile (long independent DMMA chains, one load and one store of C);
//   barrier; the rows of block column kb are multiplied by the inverted diagonal block; barrier.
// So the critical path of a step is one tile update + the 8x8 factorisation, as in a right-looking sweep, while each tile of
// the matrix is loaded and stored twice instead of once per step.  Sets flag on breakdown.  On exit the strictly-lower tiles
// hold L, the diagonal tiles hold the INVERSES of L's diagonal blocks, row n holds z^T = (L^-1 y)^T (left of the last diagonal
// tile), and logdet = sum_i log L_ii, quad = z.z = y^T K^-1 y (both valid in thread 0, zero elsewhere).
// border = false: the matrix is a diagonal block of a larger one that does not hold the border row (gp_tiled.cuh).
template <class Ctx>
GP_DEV void gp_cholesky(Ctx& c, const GpSmem& s, int n, double& logdet, double& quad, bool border = true) {
  const int T = s.T;
  double bq = 0.0;
  const GpFrag f = gp_frag(c.lane);
  GpLogProd rdet;
  rdet.m = 1.0;
  rdet.e = 0;
  long long tdiag = 0;
  const int nrow = c.nwarp > 1 ? c.nwarp - 1 : 1;          // warps that take the rows below the diagonal
  const int wrow = c.nwarp > 1 ? c.warp - 1 : 0;
  for (int kb = 0; kb < T; ++kb) {
    const int Jlast = kb > 0 ? kb - 1 : 0;
    if (c.warp == 0) {
      double acc[2][2], z0, z1;
      const long long t0 = c.clock();
      gp_chol_accumulate(c, s, f, kb, kb, -1, Jlast, kb, acc);
      const bool pd = (border && kb == T - 1) ? gp_chol8_warp<true>(c, acc[0][0], acc[0][1], z0, z1, rdet, n & 7, bq)
                                              : gp_chol8_warp<false>(c, acc[0][0], acc[0][1], z0, z1, rdet, 8, bq);
      if (!pd && c.lane == 0) s.flag[0] = 1;
      double* D = s.A + gp_tile(kb, kb);
      D[f.a0] = z0;
      D[f.a1] = z1;
      tdiag += c.clock() - t0;
    }
    if (c.warp > 0 || c.nwarp == 1) {
      if (kb > 0) gp_chol_update_rows(c, s, f, kb, kb + 1 + wrow, nrow, Jlast, kb);
      if (kb > 0 && kb + 1 < T) gp_chol_update_rows(c, s, f, kb + 1, kb + 1 + wrow, nrow, 0, kb);
    }
    c.sync();
    gp_panel_mult(c, s, f, kb);
    c.sync();
  }
  c.mark_value(9, tdiag);   // diagnostics: cycles warp 0 spent on the diagonal tiles
  logdet = (c.tid == 0) ? -gp_logprod_value(rdet) : 0.0;
  quad = (c.tid == 0) ? -bq : 0.0;
}

// ---------------------------------------------------------------------------------------------- triangular inverse
// M = [[L, 0], [z^T, 1]] -> M^-1 = [[X, 0], [-a^T, 1]], block columns from the last to the first.  For block column kb with
// rows I in (kb, T):   W_I = -M_I,kb X_kb,kb goes to the scratch column, then X_I,kb = sum_{kb < K <= I} X_I,K W_K replaces
// M_I,kb.  Reading W from scratch instead of in place removes every hazard between rows: two barriers per block column
// (v4 finished the rows in waves of one per warp with a barrier each - 58 barriers at T = 13 instead of 26).  Forming every W in
// place in one pass up front and holding a block column's rows in registers across a barrier (no W pass in front of each block
// column, same products in the same order) was measured too: 161.0 against 157.9 ms per 10,000-curve fit - dropped.
template <class Ctx>
GP_DEV void gp_trtri(Ctx& c, const GpSmem& s) {
  const int T = s.T;
  double* A = s.A;
  double* W = s.W;   // scratch column: W_I at 64 I
  const GpFrag f = gp_frag(c.lane);
  // (the diagonal tiles already hold the inverted diagonal blocks: gp_cholesky)
  if (W == nullptr) {
    // The scratch-free layout (every warp takes at most four rows of a block column: plan() in gp_kernels.cu).  ALL the W_I,kb =
    // -M_I,kb X_kb,kb are formed first, in place, in one pass without barriers between them (the factor itself is not needed any
    // more); a block column then is: every warp computes its rows into registers, reading W where it lies - barrier - the rows are
    // written - barrier.  The same products in the same order as the scratch version (bit-identical); at equal residency it is 2 %
    // slower - it exists for the classes where giving up the scratch column buys a resident CTA.
    {
      int I = 1, J = c.warp;   // strictly-lower tiles (I, J), J < I: tile-row I holds I of them
      while (I < T && J >= I) { J -= I; ++I; }
      while (I < T) {
        const double* X11 = A + gp_tile(J, J);
        double* tp = A + gp_tile(I, J);
        double w0 = 0.0, w1 = 0.0;
        c.mma(w0, w1, -tp[f.a0], X11[f.p0]);
        c.mma(w0, w1, -tp[f.a1], X11[f.p1]);
        tp[f.a0] = w0;   // (every lane's operands were consumed by the warp-wide products above)
        tp[f.a1] = w1;
        J += c.nwarp;
        while (I < T && J >= I) { J -= I; ++I; }
      }
    }
    c.sync();
    for (int kb = T - 2; kb >= 0; --kb) {
      double res[4][2];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int I = T - 1 - c.warp - r * c.nwarp;
        res[r][0] = 0.0; res[r][1] = 0.0;
        if (I > kb) {
          double p0 = 0.0, p1 = 0.0, q0 = 0.0, q1 = 0.0;
          const double* Xr = A + gp_tile(I, kb + 1);     // X_I,K for K = kb + 1 .. I
          const double* Wk = A + gp_tile(kb + 1, kb);    // W_K,kb down block column kb
          gp_kloop<GpLoops<Ctx>::rolled>(kb + 1, I + 1, [&](int K) {
            c.mma(p0, p1, Xr[f.a0], Wk[f.p0]);
            c.mma(q0, q1, Xr[f.a1], Wk[f.p1]);
            Xr += 64;
            Wk += 64 * (K + 1);
          });
          res[r][0] = p0 + q0;
          res[r][1] = p1 + q1;
        }
      }
      c.sync();   // every W of this block column has been read
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int I = T - 1 - c.warp - r * c.nwarp;
        if (I > kb) {
          double* O = A + gp_tile(I, kb);
          O[f.a0] = res[r][0];
          O[f.a1] = res[r][1];
        }
      }
      c.sync();
    }
    return;
  }
  for (int kb = T - 2; kb >= 0; --kb) {
    // (1) W_I = -(M_I,kb X11)
    {
      const double* X11 = A + gp_tile(kb, kb);
      const double b0 = X11[f.p0], b1 = X11[f.p1];
      for (int I = kb + 1 + c.warp; I < T; I += c.nwarp) {
        const double* tp = A + gp_tile(I, kb);
        double w0 = 0.0, w1 = 0.0;
        c.mma(w0, w1, -tp[f.a0], b0);
        c.mma(w0, w1, -tp[f.a1], b1);
        W[64 * I + f.a0] = w0;
        W[64 * I + f.a1] = w1;
      }
    }
    c.sync();
    // (2) the rows, longest first (row I costs I - kb tile products)
    for (int I = T - 1 - c.warp; I > kb; I -= c.nwarp) {
      double p0 = 0.0, p1 = 0.0, q0 = 0.0, q1 = 0.0;
      const double* Xr = A + gp_tile(I, kb + 1);   // X_I,K for K = kb + 1 .. I
      const double* Wk = W + 64 * (kb + 1);
      gp_kloop<GpLoops<Ctx>::rolled>(kb + 1, I + 1, [&](int) {
        c.mma(p0, p1, Xr[f.a0], Wk[f.p0]);
        c.mma(q0, q1, Xr[f.a1], Wk[f.p1]);
        Xr += 64;
        Wk += 64;
      });
      double* O = A + gp_tile(I, kb);
      O[f.a0] = p0 + q0;
      O[f.a1] = p1 + q1;
    }
    c.sync();
  }
}

// Fused K^-1 = X^T X (8x8 DMMA tiles), recomputed K and gradient traces.  tr[k] = sum_{i>j} W_ij dK_ij/dtheta_k for the
// parameters that act off the diagonal (default family: tr[0] = sum W_ij K_ij, tr[1]/tr[2] the same weighted by the squared
// scaled time / wavelength differences), trd = sum_i W_ii  (W = alpha alpha^T - K^-1).
template <class Ctx, class Kern>
GP_DEV void gp_inverse_traces(Ctx& c, const GpSmem& s, int n, const Kern& kn, double (&tr)[Kern::NTR], double& trd) {
  const int T = s.T, ntile = T * (T + 1) / 2, g = c.lane >> 2, t = c.lane & 3, rb = n & 7;
  const double* A = s.A;
  const GpFrag f = gp_frag(c.lane);
  const double* bord = A + gp_tile(T - 1, 0);   // row n of M^-1 holds -alpha
  const int ob_g = gp_el(rb, g), ob_0 = gp_el(rb, 2 * t), ob_1 = gp_el(rb, 2 * t + 1);
  const bool brow0 = (t == rb), brow1 = (4 + t == rb);   // this lane's k-rows of a tile that are the border row
#pragma unroll
  for (int k = 0; k < Kern::NTR; ++k) tr[k] = 0.0;
  trd = 0.0;
  const bool stored = Kern::kKeepK && s.kslab != nullptr;
  // epilogue of one tile (I, J) given its K^-1 entries: lane (g, t) owns (i, j) = (i0 + g, j0 + 2t) and (i0 + g, j0 + 2t + 1)
  auto finish = [&](int I, int J, double k0, double k1, const GpPair& kst) {
    const int i = 8 * I + g, j = 8 * J + 2 * t;
    const double ai = bord[64 * I + ob_g], ui = s.u[i], vi = s.v[i];                 // (the border row holds -alpha)
    const double w0 = ai * bord[64 * J + ob_0] - k0, w1 = ai * bord[64 * J + ob_1] - k1;
    const GpPair uj = *reinterpret_cast<const GpPair*>(s.u + j), vj = *reinterpret_cast<const GpPair*>(s.v + j);   // j is even
    const double du0 = ui - uj.x, dv0 = vi - vj.x, du1 = ui - uj.y, dv1 = vi - vj.y;
    const double du0s = du0 * du0, dv0s = dv0 * dv0, du1s = du1 * du1, dv1s = dv1 * dv1;
    // strictly-lower entries inside the curve count (K_ij and its derivatives are recomputed, not stored); on a diagonal tile
    // the diagonal feeds trd
    if (GP_TRACE_FAST && Kern::kLeanPaths && I < T - 1 && J < I) {   // a tile of real observations left of the diagonal: every entry counts, nothing feeds trd
      if (stored) {
        kn.accumulate_stored(w0, kst.x, du0s, dv0s, tr);
        kn.accumulate_stored(w1, kst.y, du1s, dv1s, tr);
      } else {
        kn.accumulate(w0, du0s, dv0s, tr);
        kn.accumulate(w1, du1s, dv1s, tr);
      }
      return;
    }
    const bool in = i < n;
    if (stored) {
      kn.accumulate_stored((in && i > j) ? w0 : 0.0, kst.x, du0s, dv0s, tr);
      kn.accumulate_stored((in && i > j + 1) ? w1 : 0.0, kst.y, du1s, dv1s, tr);
    } else {
      kn.accumulate((in && i > j) ? w0 : 0.0, du0s, dv0s, tr);
      kn.accumulate((in && i > j + 1) ? w1 : 0.0, du1s, dv1s, tr);
    }
    if (I == J) trd += (in && i == j) ? w0 : ((in && i == j + 1) ? w1 : 0.0);
  };
  if constexpr (GP_TRACE_PAIRS && Kern::kLeanPaths) {
  // Work items are GROUPS of up to GP_TRACE_GROUP neighbouring tiles (I, J), (I, J + 1), ... of a tile-row: they share the A fragments
  // (X_K,I), the later tiles' fragments sit at fixed offsets from the first's - per k-tile 2 + 2 n loads and one set of address / loop
  // instructions for 2 n products, and 2 n independent DMMA chains per warp.
  constexpr int NG = GP_TRACE_GROUP;
  auto run = [&](auto cnt_c, int I, int J) {
    constexpr int N = decltype(cnt_c)::value;
    const int q = I * (I + 1) / 2 + J;
    double acc[N][4];
    GpPair kst[N];
#pragma unroll
    for (int e = 0; e < N; ++e) {
      acc[e][0] = acc[e][1] = acc[e][2] = acc[e][3] = 0.0;
      kst[e] = GpPair{0.0, 0.0};
      if (stored) kst[e] = *reinterpret_cast<const GpPair*>(s.kslab + 64 * (q + e) + 8 * g + 2 * t);   // K[i][j], K[i][j + 1]: in flight during the products
    }
    // K^-1[i][j] = sum_{k >= i, k != n} X[k][i] X[k][j]:  A-fragment (m = g, k = t) = X_K,I[t][g], B (k = t, n = g) = X_K,J[t][g];
    // the diagonal tiles of X carry explicit zeros above the diagonal; the border row (in the last tile-row) is left out
    const double* Xi = A + gp_tile(I, I);
    const double* Xj = A + gp_tile(I, J);
    gp_kloop<GpLoops<Ctx>::rolled>(I, T - 1, [&](int K) {
      const double a0 = Xi[f.t0], a1 = Xi[f.t1];
#pragma unroll
      for (int e = 0; e < N; ++e) {
        c.mma(acc[e][0], acc[e][1], a0, Xj[64 * e + f.t0]);
        c.mma(acc[e][2], acc[e][3], a1, Xj[64 * e + f.t1]);
      }
      Xi += 64 * (K + 1);
      Xj += 64 * (K + 1);
    });
    const double a0 = brow0 ? 0.0 : Xi[f.t0], a1 = brow1 ? 0.0 : Xi[f.t1];
#pragma unroll
    for (int e = 0; e < N; ++e) {
      c.mma(acc[e][0], acc[e][1], a0, Xj[64 * e + f.t0]);
      c.mma(acc[e][2], acc[e][3], a1, Xj[64 * e + f.t1]);
    }
#pragma unroll
    for (int e = 0; e < N; ++e) finish(I, J + e, acc[e][0] + acc[e][2], acc[e][1] + acc[e][3], kst[e]);
  };
  int I = 0, Jp = c.warp;   // Jp: item index inside tile-row I, which holds (I + NG) / NG items
  while (Jp >= (I + NG) / NG) { Jp -= (I + NG) / NG; ++I; }
  while (I < T) {
    const int J = NG * Jp;
    const int cnt = I + 1 - J < NG ? I + 1 - J : NG;
    if (cnt == 1) run(GpInt<1>{}, I, J);
    else if (cnt == 2) run(GpInt<2>{}, I, J);
#if GP_TRACE_GROUP > 2
    else if (cnt == 3) run(GpInt<3>{}, I, J);
    else run(GpInt<4>{}, I, J);
#endif
    Jp += c.nwarp;
    while (I < T && Jp >= (I + NG) / NG) { Jp -= (I + NG) / NG; ++I; }
  }
  } else {
  int I = 0, J = c.warp;
  while (J > I) { J -= I + 1; ++I; }
  for (int q = c.warp; q < ntile; q += c.nwarp) {
    double p0 = 0.0, p1 = 0.0, q0 = 0.0, q1 = 0.0;
    GpPair kst = {0.0, 0.0};
    if (stored) kst = *reinterpret_cast<const GpPair*>(s.kslab + 64 * q + 8 * g + 2 * t);   // K[i][j], K[i][j + 1]: in flight during the products
    // K^-1[i][j] = sum_{k >= i, k != n} X[k][i] X[k][j]:  A-fragment (m = g, k = t) = X_K,I[t][g], B (k = t, n = g) = X_K,J[t][g];
    // the diagonal tiles of X carry explicit zeros above the diagonal; the border row (in the last tile-row) is left out
    const double* Xi = A + gp_tile(I, I);
    const double* Xj = A + gp_tile(I, J);
    gp_kloop<GpLoops<Ctx>::rolled>(I, T - 1, [&](int K) {
      c.mma(p0, p1, Xi[f.t0], Xj[f.t0]);
      c.mma(q0, q1, Xi[f.t1], Xj[f.t1]);
      Xi += 64 * (K + 1);
      Xj += 64 * (K + 1);
    });
    c.mma(p0, p1, brow0 ? 0.0 : Xi[f.t0], Xj[f.t0]);
    c.mma(q0, q1, brow1 ? 0.0 : Xi[f.t1], Xj[f.t1]);
    finish(I, J, p0 + q0, p1 + q1, kst);
    J += c.nwarp;
    while (J > I) { J -= I + 1; ++I; }
  }
  }
}

// ---------------------------------------------------------------------------------------------- one LML + gradient evaluation
// theta = the family's log hyperparameters (default: log[c, l_t, l_lambda, s2]).  Result valid in every thread.
// ok=false <=> K not positive definite.  grad has Kern::P entries.
template <class Kern, class Ctx>
GP_DEV void gp_eval_lml_grad(Ctx& c, const GpCurveView& cv, const double* theta, const GpSmem& s, double& lml, double* grad, bool& ok) {
  Kern kn;
  kn.init(theta);
  c.mark(0);
  gp_load_curve(c, cv, s, kn);
  c.sync();
  c.mark(1);
  // GP_WHATIF (profiling builds only, scripts/gpu_whatif.sh): bit k set = phase k left out, to measure its marginal cost under load
#ifndef GP_WHATIF
#define GP_WHATIF 0
#endif
  if (!(GP_WHATIF & 1)) gp_assemble(c, cv, s, kn);
  c.sync();
  c.mark(2);
  constexpr int NT = Kern::NTR;
  double v[NT + 3];   // traces, trd, quadratic form, log-determinant
  if (!(GP_WHATIF & 2)) gp_cholesky(c, s, cv.n, v[NT + 2], v[NT + 1]);
  else { v[NT + 2] = 0.0; v[NT + 1] = 0.0; }
  c.mark(3);
  ok = (s.flag[0] == 0);
  if (!ok) {
    lml = -INFINITY;
    for (int k = 0; k < Kern::P; ++k) grad[k] = 0.0;
    c.sync();
    return;
  }
  c.mark(4);
  if (!(GP_WHATIF & 4)) gp_trtri(c, s);
  c.mark(5);
  c.mark(6);
  {
    double tr[NT];
    if (!(GP_WHATIF & 8)) gp_inverse_traces(c, s, cv.n, kn, tr, v[NT]);
    else { for (int k = 0; k < NT; ++k) tr[k] = 0.0; v[NT] = 0.0; }
#pragma unroll
    for (int k = 0; k < NT; ++k) v[k] = tr[k];
  }
  c.mark(7);
  gp_block_sum(c, v, s.red);
  c.mark(8);
  lml = -0.5 * v[NT + 1] - v[NT + 2] - 0.5 * (double)cv.n * GP_LOG_2PI;
  kn.gradient(v, v[NT], grad);
}

// ---------------------------------------------------------------------------------------------- factor for prediction
// L_ = chol(K(theta)) and z = L^-1 y (_gpr.py:349-367; alpha_ = L^-T z is never needed: mean* = v.z with v = L^-1 k*).
// Returns LML as a by-product.
template <class Ctx, class Kern>
GP_DEV void gp_factor(Ctx& c, const GpCurveView& cv, Kern& kn, const GpSmem& s, double& lml, bool& ok) {
  gp_load_curve(c, cv, s, kn);
  c.sync();
  gp_assemble(c, cv, s, kn);
  c.sync();
  double v[2];
  gp_cholesky(c, s, cv.n, v[1], v[0]);
  ok = (s.flag[0] == 0);
  if (!ok) { lml = -INFINITY; c.sync(); return; }
  gp_block_sum(c, v, s.red);
  lml = -0.5 * v[0] - v[1] - 0.5 * (double)cv.n * GP_LOG_2PI;
}

// ---------------------------------------------------------------------------------------------- prediction
// Query points against a factored curve (strictly-lower tiles = L, diagonal tiles = inverted diagonal blocks, row n = z^T), eight
// query points per warp at a time, as a blocked triangular solve on the FP64 tensor-core instruction:
//   k*_i = c exp(-1/2 |x*/l - x_i/l|^2)          (kernels.py:1569-1570; White contributes 0 off the training set)
//   V    = L^-1 K*^T  by block forward substitution  (_gpr.py:460-462: solve_triangular, NOT an explicit inverse - at noise
//                                                  levels of 1e-5 the explicit-inverse product loses ~40x in the variance)
//        V_I = X_II (K*_I - sum_{J<I} L_IJ V_J),  X_II = L_II^-1 from gp_cholesky (only the 8x8 diagonal blocks are inverted)
//   mean* = k*.alpha = v.z                        (_gpr.py:446-447): the border row of the bordered system M [v; mu] = [k*; 0]
//                                                  solves to mu = -z.v, so the mean needs no pass of its own
//   |v|^2 -> var = (c + s2) - |v|^2               (_gpr.py:480-481)
// A warp keeps its eight columns of V as T tiles of [row][query] doubles (`Vw`: shared memory, or the warp's global slab for
// long curves); nothing of V goes to DRAM for curves that fit shared memory (v4 wrote 223 KB per curve).
// DMMA operands: A = a tile of the matrix (lane (g, t): [g][t], [g][4 + t]), B = a tile of V (lane (g, t): [t][g], [4 + t][g]),
// C = rows g, queries 2t and 2t + 1.
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

// Eight query points q0 .. q0 + 7 (those >= qd.m are padding) by one warp.  Returns 1.0 if a negative variance was clamped.
template <class Ctx, class Kern>
GP_DEV double gp_predict_tile(Ctx& c, const GpCurveView& cv, const Kern& kn, const double* stats, const GpQuery& qd, const GpSmem& s,
                              double* Vw, int q0, double* mean, double* stdv) {
  const int n = cv.n, T = s.T, g = c.lane >> 2, t = c.lane & 3, rb = n & 7;
  const GpFrag f = gp_frag(c.lane);
  const double* A = s.A;
  // this lane's two queries (columns 2t, 2t + 1 of the tile): scaled coordinates
  double uq[2], vq[2];
  bool valid[2];
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const int q = q0 + 2 * t + e;
    valid[e] = q < qd.m;
    double tq = 0.0;
    int bq = 0;
    if (valid[e]) {
      if (qd.n_obs > 0) { bq = q / qd.n_obs; tq = gp_linspace(qd.t_min, qd.t_max, qd.n_obs, q - bq * qd.n_obs); }
      else { tq = qd.qt[q]; bq = qd.qband[q]; }
    }
    uq[e] = ((tq - stats[0]) / stats[1]) * (1.0 / kn.lt);   // ss.transform then X / length_scale (scaled like the training points)
    vq[e] = cv.lam[bq] * (1.0 / kn.ll);
  }
  // a V tile holds element (row k, query n) at 8 k + (n ^ (k & 2 ? 4 : 0)): the 16-byte stores in accumulator layout
  // ([row g][queries 2t, 2t + 1]) and the 8-byte loads in B-operand layout ([k = t][n = g], [k = 4 + t][n = g]) are then both
  // free of bank conflicts (plain 8 k + n made rows k and k + 2 collide: 31 % conflicting wavefronts in the first version)
  const int oc = g * 8 + ((2 * t) ^ ((g & 2) ? 4 : 0));
  const int ob0 = t * 8 + (g ^ ((t & 2) ? 4 : 0)), ob1 = ob0 + 32;
  double sq0 = 0.0, sq1 = 0.0, m0 = 0.0, m1 = 0.0;
  for (int I = 0; I < T; ++I) {
    const int i = 8 * I + g;
    const double ui = s.u[i], vi = s.v[i];
    const double du0 = uq[0] - ui, dv0 = vq[0] - vi, du1 = uq[1] - ui, dv1 = vq[1] - vi;
    double p0 = (i < n && valid[0]) ? kn.value(du0 * du0, dv0 * dv0) : 0.0;
    double p1 = (i < n && valid[1]) ? kn.value(du1 * du1, dv1 * dv1) : 0.0;
    double r0 = 0.0, r1 = 0.0;
    const double* L = A + gp_tile(I, 0);
    const double* Vj = Vw;
    gp_kloop<GpLoops<Ctx>::rolled_predict>(0, I, [&](int) {
      c.mma(p0, p1, -L[f.a0], Vj[ob0]);
      c.mma(r0, r1, -L[f.a1], Vj[ob1]);
      L += 64;
      Vj += 64;
    });
    // right-hand side through the warp's own tile I (accumulator layout -> B layout), then V_I = X_II rhs
    double* Vi = Vw + 64 * I;
    *reinterpret_cast<GpPair*>(Vi + oc) = GpPair{p0 + r0, p1 + r1};   // (one 16-byte store: oc is even)
    c.warp_sync();
    const double b0 = Vi[ob0], b1 = Vi[ob1];
    double v0 = 0.0, v1 = 0.0;
    c.mma(v0, v1, L[f.a0], b0);   // L points at the diagonal tile now: X_II, lower triangular with explicit zeros above
    c.mma(v0, v1, L[f.a1], b1);
    c.warp_sync();
    if (i < n) {
      sq0 += v0 * v0;
      sq1 += v1 * v1;
    } else {
      if (i == n) { m0 = -v0; m1 = -v1; }   // the border row: -z.v
      v0 = 0.0; v1 = 0.0;                   // rows past the curve never feed later rows (there are none), keep V clean
    }
    *reinterpret_cast<GpPair*>(Vi + oc) = GpPair{v0, v1};
    c.warp_sync();
  }
  // column sums over the rows (lanes with the same t), the mean from the lanes that held the border row
#pragma unroll
  for (int o = 4; o < 32; o <<= 1) {
    sq0 += c.shfl_xor(sq0, o);
    sq1 += c.shfl_xor(sq1, o);
  }
  m0 = c.shfl(m0, 4 * rb + t);
  m1 = c.shfl(m1, 4 * rb + t);
  double neg = 0.0;
  if (g == 0) {
    const double kss = kn.diag();   // kernel_.diag(X*) (kernels.py:490,1315-1319,1438-1440)
    const double ymean = stats[4], ystd = stats[5];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      if (!valid[e]) continue;
      const int q = q0 + 2 * t + e;
      double var = kss - (e ? sq1 : sq0);
      if (var < 0.0) { var = 0.0; neg = 1.0; }            // _gpr.py:485-491
      mean[q] = ystd * (e ? m1 : m0) + ymean;             // _gpr.py:450
      stdv[q] = sqrt(var * (ystd * ystd));                // _gpr.py:494,500
    }
  }
  return neg;
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
// mean/std [m] of one curve at hyperparameters theta.  stats = per-curve (t mean, t scale, ., ., flux mean, flux std).
// Vw: nwarp x T tiles of scratch (64 doubles each), one slab of T tiles per warp.
// Returns status bits (valid in all threads): 2 = not PD, 16 = a negative variance was clamped.
template <class Kern, class Ctx>
GP_DEV int gp_predict_curve(Ctx& c, const GpCurveView& cv, const double* theta, const double* stats, const GpQuery& qd,
                            const GpSmem& s, double* Vw, double* mean, double* stdv, double& lml) {
  bool ok;
  Kern kn;
  kn.init(theta);
  gp_factor(c, cv, kn, s, lml, ok);
  if (!ok) {
    for (int q = c.tid; q < qd.m; q += c.nthr) { mean[q] = NAN; stdv[q] = NAN; }
    return 2;
  }
  c.sync();   // the factor is complete for every warp
  double neg[1] = {0.0};
  double* mine = Vw + (size_t)c.warp * s.T * 64;
  for (int q0 = 8 * c.warp; q0 < qd.m; q0 += 8 * c.nwarp) neg[0] += gp_predict_tile(c, cv, kn, stats, qd, s, mine, q0, mean, stdv);
  gp_block_sum(c, neg, s.red);
  return neg[0] > 0.0 ? 16 : 0;
}
