// TEST HARNESS ONLY: runs fulu_b200/csrc/gp_block.cuh (the code a CUDA thread block executes) on a pool of host threads with
// a barrier standing in for __syncthreads, so the block algorithm can be checked against the oracle without a GPU.
#include <barrier>
#include <cstring>
#include <thread>
#include <vector>

#include "../../fulu_b200/csrc/gp_block.cuh"

struct HostCtx {
  int tid, nthr;
  std::barrier<>* bar;
  void sync() { bar->arrive_and_wait(); }
};

template <class F>
static void run_block(int nthr, F&& body) {
  std::barrier<> bar(nthr);
  std::vector<std::thread> th;
  for (int t = 0; t < nthr; ++t) th.emplace_back([&, t] { HostCtx c{t, nthr, &bar}; body(c); });
  for (auto& t : th) t.join();
}

extern "C" int host_gp_prepare(int n, const double* t, const double* flux, const double* ferr, const int32_t* band, const double* loglam,
                               int n_bands, int use_err, double alpha0, int nthr, double* xs, double* y, double* al, double* lam,
                               double* stats) {
  std::vector<double> red(nthr > 64 ? nthr : 64);
  int status = 0;
  run_block(nthr, [&](HostCtx& c) {
    int st = gp_prepare_curve(c, n, t, flux, ferr, band, loglam, n_bands, use_err, alpha0, xs, y, al, lam, stats, red.data());
    if (c.tid == 0) status = st;
  });
  return status;
}

extern "C" int host_gp_lml(int n, const double* xs, const int32_t* band, const double* y, const double* al, const double* lam,
                           const double* theta, int nthr, double* lml, double* grad) {
  std::vector<double> smem(gp_smem_doubles(n, nthr));
  GpCurveView cv{n, xs, band, y, al, lam};
  int okout = 0;
  run_block(nthr, [&](HostCtx& c) {
    GpSmem s = gp_carve(smem.data(), n, nthr);
    double l, g[4];
    bool ok;
    gp_eval_lml_grad(c, cv, theta, s, l, g, ok);
    if (c.tid == 0) { *lml = l; std::memcpy(grad, g, sizeof(g)); okout = ok; }
  });
  return okout;
}

extern "C" int host_gp_predict(int n, const double* xs, const int32_t* band, const double* y, const double* al, const double* lam,
                               const double* theta, const double* stats, int m, int n_obs, double t_min, double t_max,
                               const double* qt, const int32_t* qband, int nthr, double* mean, double* stdv, double* lml) {
  std::vector<double> smem(gp_smem_doubles(n, nthr));
  std::vector<double> V((size_t)gp_pad(n) * GP_QC);
  GpCurveView cv{n, xs, band, y, al, lam};
  GpQuery qd{m, n_obs, t_min, t_max, qt, qband};
  int status = 0;
  run_block(nthr, [&](HostCtx& c) {
    GpSmem s = gp_carve(smem.data(), n, nthr);
    double l;
    int st = gp_predict_curve(c, cv, theta, stats, qd, s, V.data(), mean, stdv, l);
    if (c.tid == 0) { status = st; *lml = l; }
  });
  return status;
}
