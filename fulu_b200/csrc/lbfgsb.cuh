// Batched bound-constrained quasi-Newton optimiser: one L-BFGS-B instance per thread.
//
// Replaces scipy.optimize.minimize(method="L-BFGS-B") as driven by
// sklearn/gaussian_process/_gpr.py:658-668 (reference call site fulu/gp_aug.py:68-77).  SciPy's
// numerical core is a binary (_lbfgsb.so, L-BFGS-B 3.0); this is a from-scratch implementation of
// the published algorithm (Byrd, Lu, Nocedal, Zhu 1995; Zhu et al. 1997; Morales & Nocedal 2011;
// More & Thuente 1994 line search) as a *reverse-communication state machine*: lbfgsb_advance()
// consumes (f, g) at the current x and runs until it needs the next (f, g) or terminates.  That
// shape lets one GPU thread own one optimisation while the expensive objective (an N x N Cholesky)
// is evaluated by a whole CTA in a separate kernel.
//
// Design difference from the Fortran: the number of hyperparameters n is tiny (4..6) while the
// history length m is 10, so instead of the compact 2m x 2m representation the limited-memory matrix
//     B = theta*I - W M W^T
// is materialised as a dense n x n matrix by replaying the m stored BFGS updates on B0 = theta*I
// (Byrd-Nocedal-Schnabel 1994, Thm 2.3: the two are the same matrix).  The generalised Cauchy point
// and the subspace minimisation then work on dense B; iterates are the same up to rounding.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define FG_HD __host__ __device__ __forceinline__
// The big pieces are kept out of line and their outer loops rolled on the device: fully inlined and unrolled the optimiser step was
// 370 KB of SASS for 5.5k executed instructions - a third of its time went to instruction fetch, and inside the fused fit kernel
// (gp_device.cuh) it pushed the evaluation code out of the instruction cache.
#define FG_NI __host__ __device__ __noinline__ inline
#define FG_ROLLED _Pragma("unroll 1")
#ifndef FG_ROLL_INNER
#define FG_ROLL_INNER 0
#endif
#if FG_ROLL_INNER
#define FG_ROLLED_IN _Pragma("unroll 1")   // the loops over the (4 .. 6) variables as well
#else
#define FG_ROLLED_IN
#endif
#else
#define FG_HD inline
#define FG_NI inline
#define FG_ROLLED
#define FG_ROLLED_IN
#endif

#define FG_MAXP 6    // max hyperparameters
#define FG_LBFGS_M 10  // SciPy default maxcor

enum FgTask : int32_t {
  FG_TASK_START = 0,      // waiting for f,g at the (projected) start point
  FG_TASK_LNSRCH = 1,     // waiting for f,g at a line-search trial point
  FG_TASK_CONV_PG = 2,    // converged: projected gradient <= pgtol
  FG_TASK_CONV_F = 3,     // converged: relative reduction of f <= factr*epsmch
  FG_TASK_ABNORMAL = 4,   // abnormal termination in line search
  FG_TASK_MAXITER = 5,    // iteration / evaluation limit
  FG_TASK_IDLE = 6        // slot unused
};

struct FgOptions {
  int32_t n;        // number of variables
  int32_t m;        // history (<= FG_LBFGS_M)
  int32_t maxiter;  // 15000
  int32_t maxfun;   // 15000
  int32_t maxls;    // 20
  int32_t pad;
  double factr;     // 1e7
  double pgtol;     // 1e-5
  double lo[FG_MAXP];
  double hi[FG_MAXP];
};

struct alignas(16) FgState {
  double x[FG_MAXP];   // current point handed to the objective
  double g[FG_MAXP];
  double f;
  double t[FG_MAXP];   // x at the start of the line search
  double r[FG_MAXP];   // g at the start of the line search
  double d[FG_MAXP];   // search direction
  double z[FG_MAXP];   // subspace minimiser (x + d)
  double S[FG_LBFGS_M][FG_MAXP];
  double Y[FG_LBFGS_M][FG_MAXP];
  double theta;
  double fold, dnorm, stp, gd, gdold, stpmx, sbgnrm;
  // dcsrch state
  double ginit, gtest, gx, gy, finit, fx, fy, stx, sty, stmin, stmax, width, width1;
  int32_t brackt, stage;
  int32_t col, head;       // stored pairs (ring buffer: oldest at head)
  int32_t iter, nfgv, ifun, iback;
  int32_t task;
  int32_t nskip;
  int32_t pad_[2];      // size a multiple of 16 bytes: the state is copied as 16-byte words
};

FG_HD double fg_dot(const double* a, const double* b, int n) {
  double s = 0.0;
  FG_ROLLED_IN
  for (int i = 0; i < n; ++i) s += a[i] * b[i];
  return s;
}

// Infinity norm of the projected gradient (projgr).
template <int NV = 0>
FG_HD double fg_projgr(const FgState& s, const FgOptions& o) {
  const int n = NV > 0 ? NV : o.n;
  double nrm = 0.0;
  FG_ROLLED_IN
  for (int i = 0; i < n; ++i) {
    double gi = s.g[i];
    if (gi < 0.0) gi = fmax(s.x[i] - o.hi[i], gi);
    else gi = fmin(s.x[i] - o.lo[i], gi);
    nrm = fmax(nrm, fabs(gi));
  }
  return nrm;
}

// Dense B from the stored pairs: B0 = theta I, then BFGS updates oldest -> newest.
FG_HD void fg_form_B(const FgState& s, int n, int m, double B[FG_MAXP][FG_MAXP]) {
  FG_ROLLED_IN
  for (int i = 0; i < n; ++i)
    FG_ROLLED_IN
    for (int j = 0; j < n; ++j) B[i][j] = (i == j) ? s.theta : 0.0;
  FG_ROLLED
  for (int k = 0; k < s.col; ++k) {
    int p = (s.head + k) % m;
    const double* sv = s.S[p];
    const double* yv = s.Y[p];
    double Bs[FG_MAXP];
    double sBs = 0.0, ys = 0.0;
    FG_ROLLED_IN
    for (int i = 0; i < n; ++i) {
      double a = 0.0;
      FG_ROLLED_IN
      for (int j = 0; j < n; ++j) a += B[i][j] * sv[j];
      Bs[i] = a;
    }
    FG_ROLLED_IN
    for (int i = 0; i < n; ++i) { sBs += sv[i] * Bs[i]; ys += yv[i] * sv[i]; }
    double ib = 1.0 / sBs, iy = 1.0 / ys;
    FG_ROLLED_IN
    for (int i = 0; i < n; ++i)
      FG_ROLLED_IN
      for (int j = 0; j < n; ++j) B[i][j] += yv[i] * yv[j] * iy - Bs[i] * Bs[j] * ib;
  }
}

// Generalised Cauchy point (Algorithm CP).  Returns xcp in s.z and the free set (vars not fixed at a bound
// along the projected steepest-descent path) in freev[]; returns false if the model is not convex.
template <int NV = 0>
FG_HD bool fg_cauchy(FgState& s, const FgOptions& o, const double B[FG_MAXP][FG_MAXP], bool freev[FG_MAXP]) {
  const int n = NV > 0 ? NV : o.n;
  const double epsmch = 2.220446049250313e-16;
  double d[FG_MAXP], tb[FG_MAXP], zz[FG_MAXP];
  bool moving[FG_MAXP];
  int nbreak = 0, nmove = 0;
  FG_ROLLED_IN
  for (int i = 0; i < n; ++i) {
    double neggi = -s.g[i];
    double tl = s.x[i] - o.lo[i], tu = o.hi[i] - s.x[i];
    bool fixed = (tl <= 0.0 && neggi <= 0.0) || (tu <= 0.0 && neggi >= 0.0) || (neggi == 0.0);
    s.z[i] = s.x[i];
    zz[i] = 0.0;
    if (fixed) {
      d[i] = 0.0; moving[i] = false; tb[i] = 0.0;
      // a variable with zero gradient strictly inside the box stays free for the subspace step
      freev[i] = (tl > 0.0 && tu > 0.0);
    } else {
      d[i] = neggi; moving[i] = true; freev[i] = true; ++nmove;
      tb[i] = (neggi < 0.0) ? tl / (-neggi) : tu / neggi;
      ++nbreak;
    }
  }
  if (nmove == 0) return true;  // xcp = x
  double f1 = 0.0;
  FG_ROLLED_IN
  for (int i = 0; i < n; ++i) f1 -= d[i] * d[i];
  double f2 = 0.0;
  FG_ROLLED_IN
  for (int i = 0; i < n; ++i) {
    double a = 0.0;
    FG_ROLLED_IN
    for (int j = 0; j < n; ++j) a += B[i][j] * d[j];
    f2 += d[i] * a;
  }
  const double f2_org = s.theta * (-f1);
  if (!(f2 > 0.0)) return false;
  double dtm = -f1 / f2;
  double tsum = 0.0, tj0 = 0.0;
  int nleft = nbreak;
  bool all_fixed = false;
  while (nleft > 0) {
    int ib = -1;
    double tj = 0.0;
    FG_ROLLED_IN
    for (int i = 0; i < n; ++i)
      if (moving[i] && (ib < 0 || tb[i] < tj)) { ib = i; tj = tb[i]; }
    double dt = tj - tj0;
    if (dtm < dt) break;
    // variable ib reaches its bound: fix it there
    tsum += dt;
    tj0 = tj;
    --nleft;
    double dib = d[ib];
    FG_ROLLED_IN
    for (int i = 0; i < n; ++i) if (moving[i]) zz[i] = tsum * d[i];
    d[ib] = 0.0; moving[ib] = false; freev[ib] = false;
    if (dib > 0.0) { zz[ib] = o.hi[ib] - s.x[ib]; s.z[ib] = o.hi[ib]; }
    else { zz[ib] = o.lo[ib] - s.x[ib]; s.z[ib] = o.lo[ib]; }
    if (nleft == 0) { all_fixed = true; dtm = 0.0; break; }
    // f' = g.d + d^T B z ; f'' = d^T B d on the next segment
    f1 = 0.0; f2 = 0.0;
    FG_ROLLED_IN
    for (int i = 0; i < n; ++i) {
      if (d[i] == 0.0) continue;
      double bz = 0.0, bd = 0.0;
      FG_ROLLED_IN
      for (int j = 0; j < n; ++j) { bz += B[i][j] * zz[j]; bd += B[i][j] * d[j]; }
      f1 += d[i] * (s.g[i] + bz);
      f2 += d[i] * bd;
    }
    f2 = fmax(epsmch * f2_org, f2);
    dtm = -f1 / f2;
  }
  if (!all_fixed) {
    if (dtm <= 0.0) dtm = 0.0;
    tsum += dtm;
  }
  FG_ROLLED_IN
  for (int i = 0; i < n; ++i)
    if (moving[i]) s.z[i] = s.x[i] + tsum * d[i];
  return true;
}

// Subspace minimisation over the free variables at the Cauchy point (direct primal method with the
// projection step of L-BFGS-B 3.0).  s.z holds xcp on entry, the subspace minimiser on exit.
template <int NV = 0>
FG_HD bool fg_subsm(FgState& s, const FgOptions& o, const double B[FG_MAXP][FG_MAXP], const bool freev[FG_MAXP]) {
  const int n = NV > 0 ? NV : o.n;
  int idx[FG_MAXP];
  int nf = 0;
  FG_ROLLED_IN
  for (int i = 0; i < n; ++i) if (freev[i]) idx[nf++] = i;
  if (nf == 0) return true;
  // r = -(g + B (xcp - x)) on the free variables
  double rr[FG_MAXP], A[FG_MAXP][FG_MAXP];
  FG_ROLLED_IN
  for (int a = 0; a < nf; ++a) {
    int i = idx[a];
    double bz = 0.0;
    FG_ROLLED_IN
    for (int j = 0; j < n; ++j) bz += B[i][j] * (s.z[j] - s.x[j]);
    rr[a] = -(s.g[i] + bz);
    FG_ROLLED_IN
    for (int b = 0; b < nf; ++b) A[a][b] = B[i][idx[b]];
  }
  // Cholesky solve A du = rr
  FG_ROLLED_IN
  for (int j = 0; j < nf; ++j) {
    double dj = A[j][j];
    FG_ROLLED_IN
    for (int k = 0; k < j; ++k) dj -= A[j][k] * A[j][k];
    if (!(dj > 0.0)) return false;
    dj = sqrt(dj);
    A[j][j] = dj;
    FG_ROLLED_IN
    for (int i = j + 1; i < nf; ++i) {
      double v = A[i][j];
      FG_ROLLED_IN
      for (int k = 0; k < j; ++k) v -= A[i][k] * A[j][k];
      A[i][j] = v / dj;
    }
  }
  FG_ROLLED_IN
  for (int i = 0; i < nf; ++i) {
    double v = rr[i];
    FG_ROLLED_IN
    for (int k = 0; k < i; ++k) v -= A[i][k] * rr[k];
    rr[i] = v / A[i][i];
  }
  FG_ROLLED_IN
  for (int i = nf - 1; i >= 0; --i) {
    double v = rr[i];
    FG_ROLLED_IN
    for (int k = i + 1; k < nf; ++k) v -= A[k][i] * rr[k];
    rr[i] = v / A[i][i];
  }
  // projection of the Newton point onto the box
  double xp[FG_MAXP];
  FG_ROLLED_IN
  for (int i = 0; i < n; ++i) xp[i] = s.z[i];
  bool hit = false;
  FG_ROLLED_IN
  for (int a = 0; a < nf; ++a) {
    int k = idx[a];
    double xk = fmin(o.hi[k], fmax(o.lo[k], s.z[k] + rr[a]));
    if (xk == o.lo[k] || xk == o.hi[k]) hit = true;
    s.z[k] = xk;
  }
  if (!hit) return true;
  double ddp = 0.0;
  FG_ROLLED_IN
  for (int i = 0; i < n; ++i) ddp += (s.z[i] - s.x[i]) * s.g[i];
  if (ddp > 0.0) {
    // projected point is not a descent direction: fall back to truncating the Newton step
    FG_ROLLED_IN
    for (int i = 0; i < n; ++i) s.z[i] = xp[i];
    double alpha = 1.0, temp1 = 1.0;
    int ibd = -1;
    FG_ROLLED_IN
    for (int a = 0; a < nf; ++a) {
      int k = idx[a];
      double dk = rr[a];
      if (dk < 0.0) { double t2 = o.lo[k] - s.z[k]; if (t2 >= 0.0) temp1 = 0.0; else if (dk * alpha < t2) temp1 = t2 / dk; }
      else if (dk > 0.0) { double t2 = o.hi[k] - s.z[k]; if (t2 <= 0.0) temp1 = 0.0; else if (dk * alpha > t2) temp1 = t2 / dk; }
      if (temp1 < alpha) { alpha = temp1; ibd = a; }
    }
    if (alpha < 1.0 && ibd >= 0) {
      int k = idx[ibd];
      double dk = rr[ibd];
      if (dk > 0.0) { s.z[k] = o.hi[k]; rr[ibd] = 0.0; }
      else if (dk < 0.0) { s.z[k] = o.lo[k]; rr[ibd] = 0.0; }
    }
    FG_ROLLED_IN
    for (int a = 0; a < nf; ++a) s.z[idx[a]] += alpha * rr[a];
  }
  return true;
}

// ------------------------------------------------------------------ More-Thuente line search (dcsrch/dcstep)
FG_NI void fg_dcstep(double& stx, double& fx, double& dx, double& sty, double& fy, double& dy, double& stp, double fp, double dp,
                     int32_t& brackt, double stpmin, double stpmax) {
  double sgnd = dp * (dx / fabs(dx));
  double stpf, stpc, stpq, theta, s, gamma, p, q, r;
  if (fp > fx) {
    theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
    s = fmax(fabs(theta), fmax(fabs(dx), fabs(dp)));
    gamma = s * sqrt((theta / s) * (theta / s) - (dx / s) * (dp / s));
    if (stp < stx) gamma = -gamma;
    p = (gamma - dx) + theta;
    q = ((gamma - dx) + gamma) + dp;
    r = p / q;
    stpc = stx + r * (stp - stx);
    stpq = stx + ((dx / ((fx - fp) / (stp - stx) + dx)) / 2.0) * (stp - stx);
    if (fabs(stpc - stx) < fabs(stpq - stx)) stpf = stpc;
    else stpf = stpc + (stpq - stpc) / 2.0;
    brackt = 1;
  } else if (sgnd < 0.0) {
    theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
    s = fmax(fabs(theta), fmax(fabs(dx), fabs(dp)));
    gamma = s * sqrt((theta / s) * (theta / s) - (dx / s) * (dp / s));
    if (stp > stx) gamma = -gamma;
    p = (gamma - dp) + theta;
    q = ((gamma - dp) + gamma) + dx;
    r = p / q;
    stpc = stp + r * (stx - stp);
    stpq = stp + (dp / (dp - dx)) * (stx - stp);
    if (fabs(stpc - stp) > fabs(stpq - stp)) stpf = stpc;
    else stpf = stpq;
    brackt = 1;
  } else if (fabs(dp) < fabs(dx)) {
    theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
    s = fmax(fabs(theta), fmax(fabs(dx), fabs(dp)));
    gamma = s * sqrt(fmax(0.0, (theta / s) * (theta / s) - (dx / s) * (dp / s)));
    if (stp > stx) gamma = -gamma;
    p = (gamma - dp) + theta;
    q = (gamma + (dx - dp)) + gamma;
    r = p / q;
    if (r < 0.0 && gamma != 0.0) stpc = stp + r * (stx - stp);
    else if (stp > stx) stpc = stpmax;
    else stpc = stpmin;
    stpq = stp + (dp / (dp - dx)) * (stx - stp);
    if (brackt) {
      if (fabs(stpc - stp) < fabs(stpq - stp)) stpf = stpc;
      else stpf = stpq;
      if (stp > stx) stpf = fmin(stp + 0.66 * (sty - stp), stpf);
      else stpf = fmax(stp + 0.66 * (sty - stp), stpf);
    } else {
      if (fabs(stpc - stp) > fabs(stpq - stp)) stpf = stpc;
      else stpf = stpq;
      stpf = fmin(stpmax, stpf);
      stpf = fmax(stpmin, stpf);
    }
  } else {
    if (brackt) {
      theta = 3.0 * (fp - fy) / (sty - stp) + dy + dp;
      s = fmax(fabs(theta), fmax(fabs(dy), fabs(dp)));
      gamma = s * sqrt((theta / s) * (theta / s) - (dy / s) * (dp / s));
      if (stp > sty) gamma = -gamma;
      p = (gamma - dp) + theta;
      q = ((gamma - dp) + gamma) + dy;
      r = p / q;
      stpc = stp + r * (sty - stp);
      stpf = stpc;
    } else if (stp > stx) stpf = stpmax;
    else stpf = stpmin;
  }
  if (fp > fx) { sty = stp; fy = fp; dy = dp; }
  else {
    if (sgnd < 0.0) { sty = stx; fy = fx; dy = dx; }
    stx = stp; fx = fp; dx = dp;
  }
  stp = stpf;
}

enum { FG_LS_FG = 0, FG_LS_CONV = 1, FG_LS_WARN = 2, FG_LS_ERROR = 3 };

// One call of dcsrch after the START call.  Returns FG_LS_*; stp updated in s.stp.
FG_HD int fg_dcsrch_start(FgState& s, double f, double g, double stpmax) {
  const double ftol = 1e-3;
  if (s.stp < 0.0 || s.stp > stpmax || g >= 0.0) return FG_LS_ERROR;
  s.brackt = 0;
  s.stage = 1;
  s.finit = f;
  s.ginit = g;
  s.gtest = ftol * g;
  s.width = stpmax - 0.0;
  s.width1 = s.width / 0.5;
  s.stx = 0.0; s.fx = f; s.gx = g;
  s.sty = 0.0; s.fy = f; s.gy = g;
  s.stmin = 0.0;
  s.stmax = s.stp + 4.0 * s.stp;
  return FG_LS_FG;
}

FG_HD int fg_dcsrch_step(FgState& s, double f, double g, double stpmin, double stpmax) {
  const double ftol = 1e-3, gtol = 0.9, xtol = 0.1;
  double stp = s.stp;
  double ftest = s.finit + stp * s.gtest;
  if (s.stage == 1 && f <= ftest && g >= 0.0) s.stage = 2;
  int res = FG_LS_FG;
  if (s.brackt && (stp <= s.stmin || stp >= s.stmax)) res = FG_LS_WARN;
  if (s.brackt && s.stmax - s.stmin <= xtol * s.stmax) res = FG_LS_WARN;
  if (stp == stpmax && f <= ftest && g <= s.gtest) res = FG_LS_WARN;
  if (stp == stpmin && (f > ftest || g >= s.gtest)) res = FG_LS_WARN;
  if (f <= ftest && fabs(g) <= gtol * (-s.ginit)) res = FG_LS_CONV;
  if (res != FG_LS_FG) return res;
  if (s.stage == 1 && f <= s.fx && f > ftest) {
    double fm = f - stp * s.gtest;
    double fxm = s.fx - s.stx * s.gtest;
    double fym = s.fy - s.sty * s.gtest;
    double gm = g - s.gtest;
    double gxm = s.gx - s.gtest;
    double gym = s.gy - s.gtest;
    fg_dcstep(s.stx, fxm, gxm, s.sty, fym, gym, stp, fm, gm, s.brackt, s.stmin, s.stmax);
    s.fx = fxm + s.stx * s.gtest;
    s.fy = fym + s.sty * s.gtest;
    s.gx = gxm + s.gtest;
    s.gy = gym + s.gtest;
  } else {
    fg_dcstep(s.stx, s.fx, s.gx, s.sty, s.fy, s.gy, stp, f, g, s.brackt, s.stmin, s.stmax);
  }
  if (s.brackt) {
    if (fabs(s.sty - s.stx) >= 0.66 * s.width1) stp = s.stx + 0.5 * (s.sty - s.stx);
    s.width1 = s.width;
    s.width = fabs(s.sty - s.stx);
  }
  if (s.brackt) {
    s.stmin = fmin(s.stx, s.sty);
    s.stmax = fmax(s.stx, s.sty);
  } else {
    s.stmin = stp + 1.1 * (stp - s.stx);
    s.stmax = stp + 4.0 * (stp - s.stx);
  }
  stp = fmax(stp, stpmin);
  stp = fmin(stp, stpmax);
  if ((s.brackt && (stp <= s.stmin || stp >= s.stmax)) || (s.brackt && s.stmax - s.stmin <= xtol * s.stmax)) stp = s.stx;
  s.stp = stp;
  return FG_LS_FG;
}

// ------------------------------------------------------------------ driver
FG_HD void fg_reset_memory(FgState& s) {
  s.col = 0;
  s.head = 0;
  s.theta = 1.0;
}

FG_HD void lbfgsb_init(FgState& s, const FgOptions& o, const double* x0) {
  FG_ROLLED_IN
  for (int i = 0; i < FG_MAXP; ++i) {
    double v = (i < o.n) ? fmin(o.hi[i], fmax(o.lo[i], x0[i])) : 0.0;  // _lbfgsb_py.py:359 clips x0 into the box
    s.x[i] = v; s.g[i] = 0.0; s.t[i] = v; s.r[i] = 0.0; s.d[i] = 0.0; s.z[i] = v;
  }
  s.f = 0.0;
  fg_reset_memory(s);
  s.fold = 0.0; s.dnorm = 0.0; s.stp = 0.0; s.gd = 0.0; s.gdold = 0.0; s.stpmx = 0.0; s.sbgnrm = 0.0;
  s.ginit = s.gtest = s.gx = s.gy = s.finit = s.fx = s.fy = s.stx = s.sty = s.stmin = s.stmax = s.width = s.width1 = 0.0;
  s.brackt = 0; s.stage = 0;
  s.iter = 0; s.nfgv = 0; s.ifun = 0; s.iback = 0; s.nskip = 0;
  s.task = FG_TASK_START;
}

// Starts an iteration at the current (x, f, g): Cauchy point, subspace minimisation, line-search set-up.
// Leaves s.x at the first trial point and s.task = FG_TASK_LNSRCH, or a terminal task.
template <int NV = 0>
FG_NI void fg_begin_iteration(FgState& s, const FgOptions& o) {
  const int n = NV > 0 ? NV : o.n;
  FG_ROLLED
  for (int attempt = 0; attempt < 3; ++attempt) {
    double B[FG_MAXP][FG_MAXP];
    bool freev[FG_MAXP];
    fg_form_B(s, n, o.m, B);
    if (!fg_cauchy<NV>(s, o, B, freev)) { fg_reset_memory(s); continue; }
    if (s.col > 0) {
      if (!fg_subsm<NV>(s, o, B, freev)) { fg_reset_memory(s); continue; }
    }
    // line search set-up (lnsrlb)
    double dtd = 0.0;
    FG_ROLLED_IN
    for (int i = 0; i < n; ++i) { s.d[i] = s.z[i] - s.x[i]; dtd += s.d[i] * s.d[i]; }
    s.dnorm = sqrt(dtd);
    double stpmx = 1e10;
    if (s.iter == 0) stpmx = 1.0;
    else {
      FG_ROLLED_IN
      for (int i = 0; i < n; ++i) {
        double a1 = s.d[i];
        if (a1 < 0.0) { double a2 = o.lo[i] - s.x[i]; if (a2 >= 0.0) stpmx = 0.0; else if (a1 * stpmx < a2) stpmx = a2 / a1; }
        else if (a1 > 0.0) { double a2 = o.hi[i] - s.x[i]; if (a2 <= 0.0) stpmx = 0.0; else if (a1 * stpmx > a2) stpmx = a2 / a1; }
      }
    }
    s.stpmx = stpmx;
    s.stp = 1.0;  // all variables are boxed
    FG_ROLLED_IN
    for (int i = 0; i < n; ++i) { s.t[i] = s.x[i]; s.r[i] = s.g[i]; }
    s.fold = s.f;
    s.ifun = 0;
    s.iback = 0;
    s.gd = fg_dot(s.g, s.d, n);
    s.gdold = s.gd;
    bool bad = !(s.gd < 0.0);  // not a descent direction (info = -4)
    if (!bad) bad = (fg_dcsrch_start(s, s.f, s.gd, s.stpmx) == FG_LS_ERROR);
    if (bad) {
      if (s.col == 0) { s.task = FG_TASK_ABNORMAL; return; }
      fg_reset_memory(s);
      continue;
    }
    s.ifun = 1; s.nfgv += 1; s.iback = 0;
    if (s.stp == 1.0) for (int i = 0; i < n; ++i) s.x[i] = s.z[i];
    else for (int i = 0; i < n; ++i) s.x[i] = s.stp * s.d[i] + s.t[i];
    s.task = FG_TASK_LNSRCH;
    return;
  }
  s.task = FG_TASK_ABNORMAL;
}

// Consume (f, g) evaluated at s.x.  On return either s.task is FG_TASK_LNSRCH (evaluate at the new s.x) or terminal
// (then s.x / s.f / s.g hold the result).
// NV: the number of variables as a compile-time constant (the loops over the variables then unroll exactly and the small vectors and the
// dense B stay in registers), or 0 to take it from the options.
template <int NV = 0>
FG_HD void lbfgsb_advance(FgState& s, const FgOptions& o, double f, const double* g) {
  const int n = NV > 0 ? NV : o.n;
  const double epsmch = 2.220446049250313e-16;
  if (!(f == f) || fabs(f) > 1e100) {  // -inf LML (non-PD K, _gpr.py:592-593) or NaN: treat as a very bad trial point
    f = 1e100;
    FG_ROLLED_IN
    for (int i = 0; i < n; ++i) s.g[i] = 0.0;
  } else {
    FG_ROLLED_IN
    for (int i = 0; i < n; ++i) s.g[i] = g[i];
  }
  s.f = f;
  if (s.task == FG_TASK_START) {
    s.nfgv = 1;
    s.sbgnrm = fg_projgr<NV>(s, o);
    if (s.sbgnrm <= o.pgtol) { s.task = FG_TASK_CONV_PG; return; }
    fg_begin_iteration<NV>(s, o);
    return;
  }
  if (s.task != FG_TASK_LNSRCH) return;
  // ---- line search continues (lnsrlb label 556)
  s.gd = fg_dot(s.g, s.d, n);
  int ls = fg_dcsrch_step(s, s.f, s.gd, 0.0, s.stpmx);
  if (ls == FG_LS_FG) {
    s.ifun += 1; s.nfgv += 1; s.iback = s.ifun - 1;
    if (s.iback >= o.maxls) {
      // restore the previous iterate
      FG_ROLLED_IN
      for (int i = 0; i < n; ++i) { s.x[i] = s.t[i]; s.g[i] = s.r[i]; }
      s.f = s.fold;
      s.nfgv -= 1;
      if (s.col == 0) { s.ifun -= 1; s.iback -= 1; s.iter += 1; s.task = FG_TASK_ABNORMAL; return; }
      fg_reset_memory(s);
      fg_begin_iteration<NV>(s, o);
      return;
    }
    if (s.nfgv > o.maxfun) {  // evaluation budget (checked by SciPy's driver at NEW_X; here also mid-search as a safety net)
      FG_ROLLED_IN
      for (int i = 0; i < n; ++i) { s.x[i] = s.t[i]; s.g[i] = s.r[i]; }
      s.f = s.fold;
      s.task = FG_TASK_MAXITER;
      return;
    }
    if (s.stp == 1.0) for (int i = 0; i < n; ++i) s.x[i] = s.z[i];
    else for (int i = 0; i < n; ++i) s.x[i] = s.stp * s.d[i] + s.t[i];
    return;  // task stays FG_TASK_LNSRCH
  }
  // ---- new iterate accepted (CONVERGENCE or WARNING from dcsrch)
  s.iter += 1;
  s.sbgnrm = fg_projgr<NV>(s, o);
  if (s.sbgnrm <= o.pgtol) { s.task = FG_TASK_CONV_PG; return; }
  double ddum = fmax(fmax(fabs(s.fold), fabs(s.f)), 1.0);
  if ((s.fold - s.f) <= epsmch * o.factr * ddum) { s.task = FG_TASK_CONV_F; return; }
  // BFGS pair
  double rr = 0.0, dr, dd;
  FG_ROLLED_IN
  for (int i = 0; i < n; ++i) { s.r[i] = s.g[i] - s.r[i]; rr += s.r[i] * s.r[i]; }
  if (s.stp == 1.0) { dr = s.gd - s.gdold; dd = -s.gdold; }
  else {
    dr = (s.gd - s.gdold) * s.stp;
    FG_ROLLED_IN
    for (int i = 0; i < n; ++i) s.d[i] *= s.stp;
    dd = -s.gdold * s.stp;
  }
  if (dr <= epsmch * dd) {
    s.nskip += 1;
  } else {
    int pos;
    if (s.col < o.m) { pos = (s.head + s.col) % o.m; s.col += 1; }
    else { pos = s.head; s.head = (s.head + 1) % o.m; }
    FG_ROLLED_IN
    for (int i = 0; i < n; ++i) { s.S[pos][i] = s.d[i]; s.Y[pos][i] = s.r[i]; }
    s.theta = rr / dr;
  }
  if (s.iter >= o.maxiter || s.nfgv > o.maxfun) { s.task = FG_TASK_MAXITER; return; }
  fg_begin_iteration<NV>(s, o);
}

FG_HD bool fg_task_done(int32_t task) { return task != FG_TASK_START && task != FG_TASK_LNSRCH; }
