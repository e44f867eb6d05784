// TEST HARNESS ONLY: compiles fulu_b200/csrc/lbfgsb.cuh for the host so the optimiser state machine can be
// compared with scipy.optimize.minimize(L-BFGS-B) on the same objective without a GPU.  Not part of the product.
#include "../../fulu_b200/csrc/lbfgsb.cuh"

typedef void (*objective_fn)(const double* x, double* f, double* g);

extern "C" int host_lbfgsb_minimize(int n, const double* x0, const double* lo, const double* hi, double factr, double pgtol,
                                    int maxiter, int maxfun, int maxls, objective_fn fn, double* x_out, double* f_out,
                                    int* nfev, int* niter, double* traj, int max_traj) {
  FgOptions o;
  o.n = n; o.m = FG_LBFGS_M; o.maxiter = maxiter; o.maxfun = maxfun; o.maxls = maxls; o.pad = 0;
  o.factr = factr; o.pgtol = pgtol;
  for (int i = 0; i < FG_MAXP; ++i) { o.lo[i] = i < n ? lo[i] : 0.0; o.hi[i] = i < n ? hi[i] : 0.0; }
  FgState s;
  lbfgsb_init(s, o, x0);
  int evals = 0;
  while (!fg_task_done(s.task)) {
    double f, g[FG_MAXP];
    fn(s.x, &f, g);
    if (evals < max_traj) { for (int i = 0; i < n; ++i) traj[evals * (n + 1) + i] = s.x[i]; traj[evals * (n + 1) + n] = f; }
    ++evals;
    lbfgsb_advance(s, o, f, g);
    if (evals > 100000) break;
  }
  for (int i = 0; i < n; ++i) x_out[i] = s.x[i];
  *f_out = s.f;
  *nfev = evals;
  *niter = s.iter;
  return s.task;
}
