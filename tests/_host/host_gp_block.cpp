// TEST HARNESS ONLY: runs fulu_b200/csrc/gp_block.cuh (the code a CUDA thread block executes) on host threads.  A block barrier
// stands in for __syncthreads, a per-warp barrier for __syncwarp, and the warp-collective DMMA / shuffle instructions are
// emulated through a per-warp scratch area, so the block algorithm can be checked against the oracle without a GPU.
#include <atomic>
#include <barrier>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

#include <atomic>

#include "../../fulu_b200/csrc/gp_block.cuh"
#include "../../fulu_b200/csrc/gp_tiled.cuh"

static std::atomic<long long> g_mma_count{0};   // DMMA instructions issued (one per warp-wide call), for work accounting
extern "C" long long host_gp_mma_count() { return g_mma_count.exchange(0); }

struct WarpShared {
  std::barrier<> bar{32};
  double a[32], b[32], x[32];
};

#ifdef SAN_SKIP_SYNC
static int san_block_index = -1;   // written between run_block calls only
#endif

struct HostCtx {
  int tid, nthr, lane, warp, nwarp;
  std::barrier<>* block_bar;
  WarpShared* ws;
#ifdef SAN_SKIP_SYNC   // self-test of the sanitizer run (sanitize_main.cpp): every thread drops its $SAN_SKIP_SYNC-th block barrier
  int nsync = 0;
  void sync() {   // $SAN_SKIP_BLOCK: which run_block call of the process (0 = prepare, 1 = lml, 2 = predict in sanitize_main)
    static const int skip = getenv("SAN_SKIP_SYNC") ? atoi(getenv("SAN_SKIP_SYNC")) : 0;
    static const int blk = getenv("SAN_SKIP_BLOCK") ? atoi(getenv("SAN_SKIP_BLOCK")) : 1;
    if (++nsync != skip || san_block_index != blk) block_bar->arrive_and_wait();
  }
#else
  void sync() { block_bar->arrive_and_wait(); }
#endif
  void mark(int) {}
  long long clock() { return 0; }
  void mark_value(int, long long) {}
  void warp_sync() { ws->bar.arrive_and_wait(); }
  // D = A(8x4, row) * B(4x8, col) + C; lane (g,t): a = A[g][t], b = B[t][g], c0 = C[g][2t], c1 = C[g][2t+1]
  void mma(double& c0, double& c1, double a, double b) {
    if (lane == 0) g_mma_count.fetch_add(1, std::memory_order_relaxed);
    ws->a[lane] = a;
    ws->b[lane] = b;
    ws->bar.arrive_and_wait();
    const int g = lane >> 2, t = lane & 3;
    for (int k = 0; k < 4; ++k) {
      c0 = fma(ws->a[4 * g + k], ws->b[4 * (2 * t) + k], c0);
      c1 = fma(ws->a[4 * g + k], ws->b[4 * (2 * t + 1) + k], c1);
    }
    ws->bar.arrive_and_wait();
  }
  double shfl(double v, int src) {
    ws->x[lane] = v;
    ws->bar.arrive_and_wait();
    double r = ws->x[src & 31];
    ws->bar.arrive_and_wait();
    return r;
  }
  double shfl4(double v, int src) { return shfl(v, (lane & ~3) | (src & 3)); }   // __shfl_sync(.., width = 4)
  double shfl_xor(double v, int m) {
    ws->x[lane] = v;
    ws->bar.arrive_and_wait();
    double r = ws->x[lane ^ m];
    ws->bar.arrive_and_wait();
    return r;
  }
  // ---- emulation of mbarrier + bulk copies (gp_tiled.cuh).  A barrier is 16 bytes here (GP_MBAR_WORDS = 2): the phase counter, the
  // arrival count it is re-armed with, and one 64-bit state word = pending arrivals << 48 | outstanding bytes (biased), so that
  // "last arrival and no byte outstanding" is one atomic observation, whoever makes it true.  A bulk copy is a memcpy by the issuing
  // thread followed by the subtraction of its bytes: a consumer can never see a stage before its data.
  struct HostMbar { std::atomic<uint32_t> phase; uint32_t count; std::atomic<uint64_t> state; };
  static_assert(sizeof(HostMbar) == 16, "a barrier is 16 bytes of (emulated) shared memory");
  static constexpr uint64_t kBias = 1ull << 40;
  static HostMbar* mb(uint64_t* bar) { return reinterpret_cast<HostMbar*>(bar); }
  static void mb_update(HostMbar* m, uint64_t delta) {
    const uint64_t v = m->state.fetch_add(delta, std::memory_order_acq_rel) + delta;
    if (v == kBias) {   // no arrival pending, no byte outstanding: the phase completes, the barrier re-arms
      m->state.store(((uint64_t)m->count << 48) | kBias, std::memory_order_relaxed);
      m->phase.fetch_add(1, std::memory_order_release);
    }
  }
  void mbar_init(uint64_t* bar, int count) {
    HostMbar* m = new (bar) HostMbar;
    m->phase.store(0);
    m->count = (uint32_t)count;
    m->state.store(((uint64_t)count << 48) | kBias);
  }
  void mbar_fence_init() {}
  void mbar_arrive(uint64_t* bar) { mb_update(mb(bar), (uint64_t)0 - (1ull << 48)); }
  void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { mb_update(mb(bar), (uint64_t)bytes - (1ull << 48)); }
  void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    if (bytes == 0 || (bytes & 15) || ((uintptr_t)dst & 15) || ((uintptr_t)src & 15)) abort();
#ifdef SAN_SKIP_SYNC
    if (late_bar != nullptr) {
      std::memcpy(dst, src, bytes);   // too early: the slower warps may still be reading this stage
      mbar_wait(late_bar, late_parity);
      late_bar = nullptr;
    }
#endif
    std::memcpy(dst, src, bytes);
    mb_update(mb(bar), (uint64_t)0 - (uint64_t)bytes);
  }
  void mbar_wait(uint64_t* bar, uint32_t parity) {
    while ((mb(bar)->phase.load(std::memory_order_acquire) & 1u) == parity) std::this_thread::yield();
  }
#ifdef SAN_SKIP_SYNC   // self-test of the sanitizer run: the producer writes into a stage BEFORE waiting for its release (the wait
  uint64_t* late_bar = nullptr;   // itself still happens, right after that premature write, so the barriers' phases stay in step)
  uint32_t late_parity = 0;
#endif
  void mbar_wait_reuse(uint64_t* bar, uint32_t parity) {
#ifdef SAN_SKIP_SYNC
    static const bool skip = getenv("SAN_SKIP_REUSE_WAIT") != nullptr;
    if (skip) { late_bar = bar; late_parity = parity; return; }
#endif
    mbar_wait(bar, parity);
  }
  void fence_async() { std::atomic_thread_fence(std::memory_order_seq_cst); }
};

// family code = F1 | EX << 4 (include/fulu_gp.h FULU_GP_KERNEL)
#define HOST_FAMILIES(X) X(0, 0) X(0, 1) X(0, 2) X(0, 3) X(0, 4) X(2, 0) X(2, 2)

template <class F>
static void run_block(int nthr, F&& body) {
  const int nwarp = nthr / 32;
#ifdef SAN_SKIP_SYNC
  ++san_block_index;
#endif
  std::barrier<> bar(nthr);
  std::vector<std::unique_ptr<WarpShared>> ws;
  for (int w = 0; w < nwarp; ++w) ws.emplace_back(new WarpShared);
  std::vector<std::thread> th;
  for (int t = 0; t < nthr; ++t)
    th.emplace_back([&, t] {
      HostCtx c{t, nthr, t & 31, t >> 5, nwarp, &bar, ws[t >> 5].get()};
      body(c);
    });
  for (auto& t : th) t.join();
}

extern "C" int host_gp_prepare(int n, const double* t, const double* flux, const double* ferr, const int32_t* band, const double* loglam,
                               int n_bands, int use_err, double alpha0, int nthr, double* xs, double* y, double* al, double* lam,
                               double* stats) {
  std::vector<double> red(GP_RED);
  int status = 0;
  run_block(nthr, [&](HostCtx& c) {
    int st = gp_prepare_curve(c, n, t, flux, ferr, band, loglam, n_bands, use_err, alpha0, xs, y, al, lam, stats, red.data());
    if (c.tid == 0) status = st;
  });
  return status;
}


template <class Kern>
static int lml_impl(int family, int n, const double* xs, const int32_t* band, const double* y, const double* al, const double* lam,
                    const double* theta, int nthr, double* lml, double* grad) {
  const int np = gp_pad(n);
  const bool ns = (family & 0x400) != 0;   // bit 10: the scratch-free layout (exactly its own size: an access to a scratch column would be out of bounds)
  std::vector<double> smem(ns ? gp_mat_doubles_ns(np) + gp_vec_doubles_ns(np) : gp_smem_doubles(n));
  std::vector<double> kslab(gp_mat_doubles(gp_pad(n)));
  GpCurveView cv{n, xs, band, y, al, lam};
  int okout = 0;
  run_block(nthr, [&](HostCtx& c) {
    GpSmem s = ns ? gp_carve(smem.data(), smem.data() + gp_mat_doubles_ns(np), n, false) : gp_carve(smem.data(), smem.data() + gp_mat_doubles(np), n);
    if (family & 0x200) s.kslab = kslab.data();   // bit 9: the gradient pass reads K back from a slab instead of recomputing it
    double l, g[Kern::P];
    bool ok;
    gp_eval_lml_grad<Kern>(c, cv, theta, s, l, g, ok);
    if (c.tid == 0) { *lml = l; std::memcpy(grad, g, sizeof(g)); okout = ok; }
  });
  return okout;
}

// the long-curve path (gp_tiled.cuh): matrix in a "global" slab, operands streamed through "shared-memory" stages
template <class Kern>
static int lml_tiled_impl(int n, const double* xs, const int32_t* band, const double* y, const double* al, const double* lam,
                          const double* theta, int nthr, int reps, double* lml, double* grad) {
  using Cfg = GpTiledCfgDefault;
  const int nwarp = nthr / 32;
  std::vector<double> smem(gp_tiled_smem_doubles<Cfg>(n, nwarp));
  std::vector<double> slab(gp_tiled_slab_doubles(gp_pad(n)), std::nan(""));   // a read of something never written shows up as NaN
  GpCurveView cv{n, xs, band, y, al, lam};
  int okout = 0;
  run_block(nthr, [&](HostCtx& c) {
    GpTiled g0 = gp_tiled_carve<Cfg>(slab.data(), smem.data(), 1, nwarp, 0);
    gp_tiled_init<Cfg>(c, g0);
    uint32_t it = 0;
    for (int rep = 0; rep < reps; ++rep) {   // (several evaluations in a row: the ring of stages and its barriers carry on)
      GpTiled g = gp_tiled_carve<Cfg>(slab.data(), smem.data(), n, nwarp, it);
      double l, gr[Kern::P];
      bool ok;
      gp_tiled_eval_lml_grad<Kern, Cfg>(c, cv, theta, g, l, gr, ok);
      it = g.it;
      if (c.tid == 0) { *lml = l; std::memcpy(grad, gr, sizeof(gr)); okout = ok; }
      c.sync();
    }
  });
  return okout;
}

extern "C" int host_gp_lml_tiled(int family, int n, const double* xs, const int32_t* band, const double* y, const double* al, const double* lam,
                                 const double* theta, int nthr, int reps, double* lml, double* grad) {
#define X(F1, EX) if ((family & 0xff) == ((F1) | ((EX) << 4))) return lml_tiled_impl<GpKern<F1, EX>>(n, xs, band, y, al, lam, theta, nthr, reps, lml, grad);
  HOST_FAMILIES(X)
#undef X
  return -1;
}

template <class Kern>
static int predict_tiled_impl(int n, const double* xs, const int32_t* band, const double* y, const double* al, const double* lam,
                              const double* theta, const double* stats, int m, int n_obs, double t_min, double t_max,
                              const double* qt, const int32_t* qband, int nthr, double* mean, double* stdv, double* lml) {
  using Cfg = GpTiledCfgDefault;
  const int nwarp = nthr / 32;
  std::vector<double> smem(gp_tiled_smem_doubles<Cfg>(n, nwarp));
  std::vector<double> slab(gp_tiled_slab_doubles(gp_pad(n)), std::nan(""));
  std::vector<double> V((size_t)nwarp * (gp_pad(n) / 8) * 64);
  GpCurveView cv{n, xs, band, y, al, lam};
  GpQuery qd{m, n_obs, t_min, t_max, qt, qband};
  int status = 0;
  run_block(nthr, [&](HostCtx& c) {
    GpTiled g0 = gp_tiled_carve<Cfg>(slab.data(), smem.data(), 1, nwarp, 0);
    gp_tiled_init<Cfg>(c, g0);
    GpTiled g = gp_tiled_carve<Cfg>(slab.data(), smem.data(), n, nwarp, 0);
    double l;
    int st = gp_tiled_predict_curve<Kern, Cfg>(c, cv, theta, stats, qd, g, V.data(), mean, stdv, l);
    if (c.tid == 0) { status = st; *lml = l; }
  });
  return status;
}

extern "C" int host_gp_predict_tiled(int family, int n, const double* xs, const int32_t* band, const double* y, const double* al,
                                     const double* lam, const double* theta, const double* stats, int m, int n_obs, double t_min,
                                     double t_max, const double* qt, const int32_t* qband, int nthr, double* mean, double* stdv, double* lml) {
#define X(F1, EX) \
  if (family == ((F1) | ((EX) << 4))) \
    return predict_tiled_impl<GpKern<F1, EX>>(n, xs, band, y, al, lam, theta, stats, m, n_obs, t_min, t_max, qt, qband, nthr, mean, stdv, lml);
  HOST_FAMILIES(X)
#undef X
  return -1;
}

extern "C" int host_gp_lml(int family, int n, const double* xs, const int32_t* band, const double* y, const double* al, const double* lam,
                           const double* theta, int nthr, double* lml, double* grad) {
#define X(F1, EX) if ((family & 0xff) == ((F1) | ((EX) << 4))) return lml_impl<GpKern<F1, EX>>(family, n, xs, band, y, al, lam, theta, nthr, lml, grad);
  HOST_FAMILIES(X)
#undef X
  return -1;
}

template <class Kern>
static int predict_impl(int n, const double* xs, const int32_t* band, const double* y, const double* al, const double* lam,
                        const double* theta, const double* stats, int m, int n_obs, double t_min, double t_max,
                        const double* qt, const int32_t* qband, int nthr, double* mean, double* stdv, double* lml) {
  std::vector<double> smem(gp_smem_doubles(n));
  std::vector<double> V((size_t)(nthr / 32) * (gp_pad(n) / 8) * 64);
  GpCurveView cv{n, xs, band, y, al, lam};
  GpQuery qd{m, n_obs, t_min, t_max, qt, qband};
  int status = 0;
  const int np = gp_pad(n);
  run_block(nthr, [&](HostCtx& c) {
    GpSmem s = gp_carve(smem.data(), smem.data() + gp_mat_doubles(np), n);
    double l;
    int st = gp_predict_curve<Kern>(c, cv, theta, stats, qd, s, V.data(), mean, stdv, l);
    if (c.tid == 0) { status = st; *lml = l; }
  });
  return status;
}

extern "C" int host_gp_predict(int family, int n, const double* xs, const int32_t* band, const double* y, const double* al, const double* lam,
                               const double* theta, const double* stats, int m, int n_obs, double t_min, double t_max,
                               const double* qt, const int32_t* qband, int nthr, double* mean, double* stdv, double* lml) {
#define X(F1, EX) \
  if (family == ((F1) | ((EX) << 4))) \
    return predict_impl<GpKern<F1, EX>>(n, xs, band, y, al, lam, theta, stats, m, n_obs, t_min, t_max, qt, qband, nthr, mean, stdv, lml);
  HOST_FAMILIES(X)
#undef X
  return -1;
}

extern "C" void host_gp_exp_neg(int n, const double* x, double* out) {
  for (int i = 0; i < n; ++i) out[i] = gp_exp_tab(x[i], gp_exp2_tab, GP_EXP_XMIN);
}
