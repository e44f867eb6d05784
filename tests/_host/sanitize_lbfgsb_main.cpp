// TEST HARNESS ONLY: the optimiser state machine (fulu_b200/csrc/lbfgsb.cuh) under -fsanitize=address,undefined: the state is
// fixed-size arrays inside a struct (m = 10 pairs, up to FG_MAXP variables), so an index past an array's end is a UBSan "bounds"
// report.  Objectives that drive it through its branches: interior optimum, optimum on the bounds, a start outside the bounds,
// an objective that is +inf in part of the box (line-search backtracking), a flat objective, tiny iteration / evaluation limits.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "host_lbfgsb.cpp"

static int g_n = 4, g_kind = 0;

static void objective(const double* x, double* f, double* g) {
  const int n = g_n;
  double v = 0.0;
  for (int i = 0; i < n; ++i) g[i] = 0.0;
  if (g_kind == 0) {   // Rosenbrock chain
    for (int i = 0; i + 1 < n; ++i) {
      const double a = x[i + 1] - x[i] * x[i], b = 1.0 - x[i];
      v += 100.0 * a * a + b * b;
      g[i] += -400.0 * a * x[i] - 2.0 * b;
      g[i + 1] += 200.0 * a;
    }
  } else if (g_kind == 1) {   // quadratic whose minimiser lies outside the box on every axis
    for (int i = 0; i < n; ++i) { const double d = x[i] - (i % 2 ? 20.0 : -20.0); v += (1.0 + i) * d * d; g[i] = 2.0 * (1.0 + i) * d; }
  } else if (g_kind == 2) {   // +inf where x0 > 1 (a failed Cholesky looks like this to the optimiser, _gpr.py:590-593)
    for (int i = 0; i < n; ++i) { const double d = x[i] - 0.9; v += d * d * d * d + d * d; g[i] = 4.0 * d * d * d + 2.0 * d; }
    if (x[0] > 1.0) { v = INFINITY; for (int i = 0; i < n; ++i) g[i] = 0.0; }
  } else {   // flat
    v = 3.0;
  }
  *f = v;
}

int main(int argc, char** argv) {
  int bad = 0;
  for (int n = 1; n <= FG_MAXP; ++n)
    for (int kind = 0; kind < 4; ++kind)
      for (int lim = 0; lim < 3; ++lim) {
        g_n = n; g_kind = kind;
        double x0[FG_MAXP], lo[FG_MAXP], hi[FG_MAXP], x[FG_MAXP], f = 0.0;
        for (int i = 0; i < n; ++i) { lo[i] = -11.5; hi[i] = 11.5; x0[i] = (kind == 2 ? 5.0 : -1.2 + 0.7 * i) + (lim == 2 ? 100.0 : 0.0); }
        int nfev = 0, niter = 0;
        std::vector<double> traj((size_t)64 * (n + 1));
        const int maxiter = lim == 1 ? 2 : 15000, maxfun = lim == 1 ? 3 : 15000, maxls = lim == 1 ? 1 : 20;
        const int task = host_lbfgsb_minimize(n, x0, lo, hi, 1e7, 1e-5, maxiter, maxfun, maxls, objective, x, &f, &nfev, &niter, traj.data(), 64);
        bool ok = std::isfinite(f) || kind == 2;
        for (int i = 0; i < n; ++i) ok = ok && x[i] >= lo[i] && x[i] <= hi[i];
        if (kind == 1 && lim != 1) for (int i = 0; i < n; ++i) ok = ok && std::fabs(x[i] - (i % 2 ? hi[i] : lo[i])) < 1e-8;
        if (kind == 0 && lim == 0 && n >= 2) {   // a stationary point (the chain has a second local minimum near x0 = -1 for n >= 4)
          double fx, gx[FG_MAXP];
          objective(x, &fx, gx);
          for (int i = 0; i < n; ++i) ok = ok && std::fabs(gx[i]) < 1e-2;
        }
        if (!ok) { printf("BAD n=%d kind=%d lim=%d task=%d f=%g nfev=%d\n", n, kind, lim, task, f, nfev); ++bad; }
      }
  printf("lbfgsb sanitize: %d bad\n", bad);
  return bad ? 1 : 0;
}
