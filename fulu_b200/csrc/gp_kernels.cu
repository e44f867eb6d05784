// CUDA kernels and C-ABI entry points of libfulu_gp (sm_100a).  See include/fulu_gp.h for the contract and
// gp_block.cuh / lbfgsb.cuh for the algorithms.
//
// Execution model
//   prep kernel      one CTA per curve: StandardScaler statistics, y normalisation, noise diagonal  (a1-a3)
//   eval kernel      persistent CTAs; each takes one (curve, start) optimiser slot, builds K in shared memory,
//                    factors it and returns -LML and its gradient                                   (a4, a5)
//   step kernel      one thread per active slot: advances its L-BFGS-B state machine, and when a run finishes pulls the
//                    next (curve, start) task into the freed slot ("continuous batching")            (a6)
//   tail kernel      once fewer slots are live than CTAs (and for every small fit): one launch, CTA b owns slot b and
//                    alternates evaluation and optimiser step until the slot runs dry                 (a4-a6)
//   select kernel    best of the S runs per curve                                                    (a6)
//   predict kernel   one CTA per curve: final factorisation + augmentation grid / query points       (a7-a9)
// There is no collective anywhere: curves are independent, multi-GPU runs shard the batch.
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "gp_device.cuh"

namespace {

inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// ------------------------------------------------------------------------------------------------ workspace
struct WsLayout {
  size_t xs, y, al, lam, stats, pstatus;                                      // prepared curves
  size_t slots, slot_task, slot_nev, f_out, g_out, list, counts, next_task;   // optimiser window
  size_t run_f, run_x, run_nev, run_end;                                      // per (curve, start) results
  size_t gA, gV;                                                              // per-CTA global scratch
  size_t total;
  int window, max_rounds, grid_eval, grid_predict, threads_eval, occ_eval;
  bool stream_tail;   // the default window of a long-curve class: the whole fit in fit_tail_kernel
  size_t gA_stride, gV_stride;   // doubles per CTA
  size_t slabs;                  // per-CTA slabs allocated at gA
  bool globalA;   // the matrix lives in a global slab (place != GP_PLACE_SMEM)
  bool noscratch; // shared-memory placement in the scratch-free layout (gp_block.cuh): one more CTA per SM for N = 104 .. 127, 136 .. 159
  int place;      // GP_PLACE_*
  size_t smem_eval, smem_predict;
};

struct DeviceInfo {
  int sms;
};

int device_info(DeviceInfo& di) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return (int)e;
  e = cudaDeviceGetAttribute(&di.sms, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) return (int)e;
  return 0;
}

constexpr int kDefaultWindow = 65536;   // measured on the 10,000-curve fit (60,000 tasks): 4096 -> 210.6 ms, 8192 -> 194.6, 16384 -> 188.0, 32768 -> 184.3, >= tasks -> 183.2
constexpr int kDefaultMaxRounds = 200000;

void plan(const fulu_gp_batch& b, int P, int S, int window, int max_rounds, int sms, WsLayout& w) {
  const size_t B = (size_t)b.n_curves, NT = (size_t)b.n_obs_total;
  const int nmax = b.n_max < 1 ? 1 : b.n_max;
  const size_t np = (size_t)gp_pad(nmax);
  const size_t vec = gp_vec_doubles((int)np) * sizeof(double) + 16;
  const size_t full = gp_mat_doubles((int)np) * sizeof(double) + vec;
  w.globalA = full > kMaxSmem;
  w.noscratch = false;
  // Residency: 228 KB of shared memory per SM, 1 KB reserved per CTA; the evaluation of a curve that fits shared memory is latency
  // bound, so as many curves per SM as fit, and the 512 threads the register file allows are split between them.
  // Longer curves keep the matrix in a per-CTA slab in global memory: up to N = 303 two 256-thread CTAs per SM run the 8x8 block code
  // on it (the slabs of a wave stay in L2), above that one 512-thread CTA per SM streams it through shared-memory stages (gp_tiled.cuh).
  int occ;
  // Boundary between the slab and the tiled placement, measured on B200 with 296 evaluations in flight (algorithmic TFLOP/s, slab /
  // tiled): N = 240: 7.9 / 6.1, 300: 10.3 / 8.2, 400: 8.2 / 10.9, 500: 6.9 / 14.5 - the class N <= 303 (config #4's longest) stays on
  // the slab, everything above streams.  (FULU_GP_TILED_ABOVE=<np>: A/B switch.)
  const char* tiled_above = getenv("FULU_GP_TILED_ABOVE");
  const size_t np_slab_max = tiled_above ? (size_t)atoi(tiled_above) : 304;
  w.place = !w.globalA ? GP_PLACE_SMEM : (np <= np_slab_max ? GP_PLACE_SLAB : GP_PLACE_TILED);
  if (w.place == GP_PLACE_TILED) {
    occ = 1;
    w.smem_eval = gp_tiled_smem_doubles<TiledCfg>(nmax, kEvalMaxThreads / 32) * sizeof(double) + 16;
    w.smem_predict = gp_tiled_smem_doubles<TiledCfg>(nmax, kPredictThreads / 32) * sizeof(double) + 16;
  } else if (w.place == GP_PLACE_SLAB) {
    occ = 2;
    w.smem_eval = vec;
  } else {
    w.smem_eval = full;
    occ = (int)((228 * 1024) / (w.smem_eval + 1024));
    // The scratch-free layout where it turns two resident CTAs into three or one into two and no warp gets more than four rows of a
    // block column of the inverse (N = 104 .. 127: 3 x 160 threads, N = 136 .. 159: 2 x 256).  FULU_GP_NO_SCRATCH=0 / 1: A/B switch.
    const size_t full_ns = gp_mat_doubles_ns((int)np) * sizeof(double) + gp_vec_doubles_ns((int)np) * sizeof(double) + 16;
    const int occ_ns = (int)((228 * 1024) / (full_ns + 1024));
    const int warps_ns = occ_ns > 0 ? kEvalMaxThreads / occ_ns / 32 : 0;
    const int T = (int)(np / GP_NB);
    const char* e = getenv("FULU_GP_NO_SCRATCH");
    const bool allowed = e ? atoi(e) != 0 : true;
    if (allowed && occ >= 1 && occ <= 2 && occ_ns > occ && warps_ns >= 1 && (T - 2 + warps_ns) / warps_ns <= 4) {
      w.noscratch = true;
      w.smem_eval = full_ns;
      occ = occ_ns;
    }
  }
  // prediction: eight warps, each with its eight columns of V as T tiles (shared memory, or the CTA's global slab for long curves)
  const size_t vtiles = (size_t)(kPredictThreads / 32) * (np / GP_NB) * 64 * sizeof(double);
  // (the prediction kernel always uses the scratch layout: `full`, or the vectors alone when the matrix is in a slab)
  const size_t smem_factor = w.place == GP_PLACE_SMEM ? full : w.smem_eval;
  const bool globalV = w.globalA || smem_factor + vtiles > kMaxSmem;
  if (w.place != GP_PLACE_TILED) w.smem_predict = globalV ? smem_factor : smem_factor + vtiles;
  if (occ < 1) occ = 1;
  if (occ > 16) occ = 16;
  int threads = kEvalMaxThreads / occ / 32 * 32;
  if (threads < 32) threads = 32;
  w.occ_eval = occ;
  w.threads_eval = threads;
  w.grid_eval = sms * occ;
  int occ_p = (int)((228 * 1024) / (w.smem_predict + 1024));   // 256-thread prediction CTAs, at most two per SM (registers)
  if (occ_p > 2) occ_p = 2;
  if (occ_p < 1) occ_p = 1;
  w.grid_predict = sms * occ_p;
  // Long curves (one CTA per SM, matrix in a global slab: an evaluation is 3 .. 190 ms): by default the window is one slot per CTA
  // and the whole fit runs in the fused kernel, in which a CTA that finishes a run pulls the next task at once.  Rounds would make
  // every slot wait for the slowest evaluation of the round (lengths are mixed in these classes) and run a second, partly empty
  // wave whenever more slots are live than CTAs; the optimiser step on one thread (~80 us) is nothing next to such an evaluation.
  w.stream_tail = false;
  if (window <= 0) {
    if (w.globalA && occ == 1) { window = w.grid_eval; w.stream_tail = true; }
    else window = kDefaultWindow;
  }
  size_t tasks = (size_t)(b.order ? b.n_order : b.n_curves) * (size_t)(S > 0 ? S : 1);
  if ((size_t)window > tasks) window = (int)(tasks > 0 ? tasks : 1);
  w.window = window;
  // (a caller's cap can only lower the default: the workspace query does not know it, so the round counters are sized for the default)
  w.max_rounds = (max_rounds > 0 && max_rounds < kDefaultMaxRounds) ? max_rounds : kDefaultMaxRounds;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o = align_up(o + bytes); return r; };
  w.xs = take(NT * 8); w.y = take(NT * 8); w.al = take(NT * 8);
  w.lam = take(B * (size_t)b.n_bands * 8);
  w.stats = take(B * 6 * 8);
  w.pstatus = take(B * 4);
  // (the fused fit addresses its slots as CTA x lane of the stepper warp, whatever the window)
  const size_t nslot = (size_t)window > (size_t)sms * kFusedMaxSlots ? (size_t)window : (size_t)sms * kFusedMaxSlots;
  w.slots = take(nslot * sizeof(FgState));
  w.slot_task = take(nslot * 8);
  w.slot_nev = take(nslot * 4);
  w.f_out = take((size_t)window * 8);
  w.g_out = take((size_t)window * FG_MAXP * 8);
  w.list = take((size_t)window * 2 * 4);
  w.counts = take(((size_t)w.max_rounds + 2) * 4);
  w.next_task = take(8);
  w.run_f = take(tasks * 8);
  w.run_x = take(tasks * (size_t)FG_MAXP * 8);
  w.run_nev = take(tasks * 4);
  w.run_end = take(tasks * 4);
  // per-CTA slab: the matrix itself (+ the inverse's scratch panel and the inverted diagonal blocks) for long curves, a copy of the
  // K tiles for the gradient pass otherwise (FULU_GP_NO_KSLAB=1: none).  One slab per CTA that can be at work: never more than the
  // tasks (evaluation kernels) or the curves (prediction) of the call.
  w.gA_stride = w.place == GP_PLACE_TILED ? gp_tiled_slab_doubles((int)np)
                : w.place == GP_PLACE_SLAB ? gp_mat_doubles((int)np)
                                           : (getenv("FULU_GP_NO_KSLAB") ? 0 : 64 * ((np / GP_NB) * (np / GP_NB + 1) / 2));
  const size_t curves = (size_t)(b.order ? b.n_order : b.n_curves);
  size_t slabs = (size_t)w.grid_eval < tasks ? (size_t)w.grid_eval : tasks;
  if (w.globalA) {
    const size_t ps = (size_t)w.grid_predict < curves ? (size_t)w.grid_predict : curves;
    if (ps > slabs) slabs = ps;
  }
  if (slabs < 1) slabs = 1;
  w.slabs = slabs;
  w.gA = take(slabs * w.gA_stride * 8);
  w.gV_stride = globalV ? vtiles / sizeof(double) : 0;
  w.gV = take((size_t)w.grid_predict * w.gV_stride * 8);
  w.total = o;
}

template <class T>
T* at(void* ws, size_t off) { return reinterpret_cast<T*>(reinterpret_cast<char*>(ws) + off); }

// ------------------------------------------------------------------------------------------------ kernels
struct PrepArgs {
  int64_t B;              // curves to prepare (entries of order, or all)
  const int32_t* order;
  const int64_t* offsets;
  const double *t, *flux, *ferr, *loglam;
  const int32_t* band;
  int n_bands, use_err;
  int n_max;              // the caller's bound on the curve length: it sized the shared memory / slab of every later kernel
  double alpha0;
  double *xs, *y, *al, *lam, *stats;
  int32_t* pstatus;
};

__global__ void __launch_bounds__(kPrepThreads) prep_kernel(PrepArgs a) {
  __shared__ double red[GP_RED];
  DevCtx c;
  for (int64_t k = blockIdx.x; k < a.B; k += gridDim.x) {
    const int64_t cu = a.order ? (int64_t)a.order[k] : k;
    const int64_t o = a.offsets[cu];
    const int n = (int)(a.offsets[cu + 1] - o);
    int st = gp_prepare_curve(c, n, a.t + o, a.flux + o, a.ferr + o, a.band + o, a.loglam, a.n_bands, a.use_err, a.alpha0, a.xs + o,
                              a.y + o, a.al + o, a.lam + cu * a.n_bands, a.stats + cu * 6, red);
    if (n > a.n_max) st = FULU_GP_STATUS_BAD_INPUT;   // longer than the caller said: flagged and skipped by every consumer, never run
    if (threadIdx.x == 0) a.pstatus[cu] = st;
    __syncthreads();
  }
}

__global__ void fit_init_kernel(FitDev d) {
  const int slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= d.window) return;
  if (acquire_task(d, slot, d.slots[slot])) {
    const int pos = atomicAdd(&d.counts[0], 1);
    d.list[pos] = slot;
  }
}

// One optimiser state per thread.  The advance is ~5.5k dependent, branchy instructions per slot (a third of them the replay of the
// stored BFGS pairs into the dense B) at one warp per scheduler: 75-84 us per launch whatever the layout - ncu: 31 % long-scoreboard,
// 30 % instruction-fetch, 25 % fixed-latency stalls, issue slots 7.6 % used (profiles/r01_v5_step_kernel_ncu.md).  Measured on the
// 10,000-curve fit (188 launches): state advanced in place in global memory, 128-thread blocks 14.2 ms; on a thread-local copy
// 14.0 ms, with 64-thread blocks 13.5 ms (kept), 32-thread blocks 13.1 ms; staged through shared memory 19.3 ms; slots spread over
// 2-32x more warps 14.0-28 ms.
#ifndef GP_STEP_THREADS
#define GP_STEP_THREADS 64
#endif
template <int NV>
__global__ void fit_step_kernel(FitDev d, int round) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (*d.ended || i >= d.counts[round]) return;
  const int slot = d.list[(size_t)(round & 1) * d.window + i];
  if (advance_slot<NV>(d, slot, d.f_out[slot], d.g_out + slot * FG_MAXP)) {
    const int pos = atomicAdd(&d.counts[round + 1], 1);
    d.list[(size_t)((round + 1) & 1) * d.window + pos] = slot;
  }
}

// Runs that were still alive when the round cap hit: record where they are.
__global__ void fit_flush_kernel(FitDev d) {
  const int slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= d.window) return;
  const int64_t task = d.slot_task[slot];
  FgState& st = d.slots[slot];
  if (task < 0 || fg_task_done(st.task)) return;
  const bool in_search = (st.task == FG_TASK_LNSRCH);
  d.run_f[task] = in_search ? st.fold : INFINITY;
  for (int k = 0; k < FG_MAXP; ++k) d.run_x[task * FG_MAXP + k] = in_search ? st.t[k] : st.x[k];
  d.run_nev[task] = d.slot_nev[slot];
  d.run_end[task] = FG_TASK_MAXITER;
}

struct SelectArgs {
  int64_t B;              // curves processed
  const int32_t* order;
  int S, P;
  const double *run_f, *run_x;
  const int32_t *run_nev, *run_end, *pstatus;
  double lo[FG_MAXP], hi[FG_MAXP];
  double *theta_opt, *lml_opt, *run_lml, *run_theta;
  int32_t *n_eval, *status, *run_n_eval;
};

__global__ void fit_select_kernel(SelectArgs a) {
  const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (k >= a.B) return;
  const int64_t cu = a.order ? (int64_t)a.order[k] : k;   // runs are stored by position, results by curve id
  int best = 0, nev = 0;
  double fb = a.run_f[k * a.S];
  for (int s = 0; s < a.S; ++s) {
    const double f = a.run_f[k * a.S + s];
    if (f < fb) { fb = f; best = s; }   // np.argmin: first minimum (_gpr.py:336-340)
    nev += a.run_nev[k * a.S + s];
    if (a.run_lml) a.run_lml[cu * a.S + s] = (f < 1.0e99) ? -f : -INFINITY;
    if (a.run_n_eval) a.run_n_eval[cu * a.S + s] = a.run_nev[k * a.S + s];
    if (a.run_theta)
      for (int q = 0; q < a.P; ++q) a.run_theta[(cu * a.S + s) * a.P + q] = a.run_x[(k * a.S + s) * FG_MAXP + q];
  }
  int st = 0;
  if (a.pstatus[cu] != 0) st |= FULU_GP_STATUS_BAD_INPUT;
  const int end = a.run_end[k * a.S + best];
  if (end == FG_TASK_ABNORMAL || end == FG_TASK_MAXITER) st |= FULU_GP_STATUS_NOT_CONVERGED;
  for (int q = 0; q < a.P; ++q) {
    const double th = a.run_x[(k * a.S + best) * FG_MAXP + q];
    a.theta_opt[cu * a.P + q] = th;
    const double tol = 1e-8 + 1e-5 * fabs(th);   // np.isclose(bound, theta)  (kernels.py:436-464)
    if (fabs(a.lo[q] - th) <= tol || fabs(a.hi[q] - th) <= tol) st |= FULU_GP_STATUS_AT_BOUND;
  }
  // the optimiser works with f = 1e100 where K is not positive definite (lbfgsb.cuh); a best run that never left that value has
  // no finite LML: report sklearn's -inf (_gpr.py:592-593) and flag it
  const bool never_pd = !(fb < 1.0e99);
  if (never_pd && a.pstatus[cu] == 0) st |= FULU_GP_STATUS_NOT_PD;
  a.lml_opt[cu] = never_pd ? -INFINITY : -fb;
  a.n_eval[cu] = nev;
  a.status[cu] = st;
}

__global__ void export_stats_kernel(int64_t B, const int32_t* order, const double* stats, double* scaler, double* ynorm) {
  const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (k >= B) return;
  const int64_t cu = order ? (int64_t)order[k] : k;
  if (scaler) for (int k = 0; k < 4; ++k) scaler[cu * 4 + k] = stats[cu * 6 + k];
  if (ynorm) { ynorm[cu * 2] = stats[cu * 6 + 4]; ynorm[cu * 2 + 1] = stats[cu * 6 + 5]; }
}

// ------------------------------------------------------------------------------------------------ downstream consumer: band sum + peak
// One warp per curve: the augmented curve summed over the passbands at each grid time (bands added in table order, as pandas'
// groupby("time").sum() meets them in the band-major frame), and its first maximum (Series.argmax; NaNs are skipped).
struct PeakArgs {
  int64_t B;
  const int32_t* order;
  const int64_t* offsets;
  const double* t_raw;
  const double *mean, *t_min, *t_max;
  int n_bands, n_obs;
  double *sum_flux, *peak_t, *peak_flux;
  int32_t* peak_index;
};

__global__ void __launch_bounds__(256) peak_kernel(PeakArgs a) {
  const int lane = threadIdx.x & 31;
  const int64_t k = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (k >= a.B) return;
  const int64_t cu = a.order ? (int64_t)a.order[k] : k;
  const double* m = a.mean + cu * (int64_t)a.n_bands * a.n_obs;
  double best = -INFINITY;
  int bi = 0x7fffffff;
  for (int q = lane; q < a.n_obs; q += 32) {
    double sum = m[q];
    for (int b = 1; b < a.n_bands; ++b) sum += m[(int64_t)b * a.n_obs + q];
    if (a.sum_flux) a.sum_flux[cu * a.n_obs + q] = sum;
    if (sum > best) { best = sum; bi = q; }   // strictly greater: the first maximum of this lane's points (they ascend)
  }
  for (int o = 16; o > 0; o >>= 1) {
    const double ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
  }
  if (lane == 0) {
    double lo, hi;
    if (a.t_min && a.t_max) { lo = a.t_min[cu]; hi = a.t_max[cu]; }
    else {
      const double* tr = a.t_raw + a.offsets[cu];
      const int n = (int)(a.offsets[cu + 1] - a.offsets[cu]);
      lo = hi = n > 0 ? tr[0] : 0.0;
      for (int i = 1; i < n; ++i) { lo = fmin(lo, tr[i]); hi = fmax(hi, tr[i]); }
    }
    const bool found = bi != 0x7fffffff;
    a.peak_index[cu] = found ? bi : -1;
    a.peak_flux[cu] = found ? best : NAN;
    a.peak_t[cu] = found ? gp_linspace(lo, hi, a.n_obs, bi) : NAN;
  }
}

// ------------------------------------------------------------------------------------------------ ingest: object-grouped table -> CSR
// Rows of one object are contiguous (PLAsTiCC / ZTF tables, notebook_examples/object_34299.csv).  A row is a head when its object
// id differs from the previous row's; offsets = positions of the heads.  Three passes over the id column (12 bytes read and 4
// written per row in all, HBM bound): heads per 1024-row chunk, an exclusive scan of the chunk counts by one block, the write.
constexpr int kIngestThreads = 256, kIngestRows = 1024;

__device__ __forceinline__ bool is_head(const int64_t* id, int64_t i) { return i == 0 || id[i] != id[i - 1]; }

__global__ void __launch_bounds__(kIngestThreads) ingest_count_kernel(int64_t n_rows, const int64_t* id, int64_t* chunk_count) {
  __shared__ int wsum[kIngestThreads / 32];
  const int64_t base = (int64_t)blockIdx.x * kIngestRows;
  int c = 0;
  for (int r = threadIdx.x; r < kIngestRows; r += kIngestThreads) {
    const int64_t i = base + r;
    if (i < n_rows && is_head(id, i)) ++c;
  }
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < kIngestThreads / 32; ++w) t += wsum[w];
    chunk_count[blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(1024) ingest_scan_kernel(int64_t n_chunks, int64_t* chunk_count, int64_t* n_objects) {
  // in place: chunk_count[b] <- number of heads before chunk b; one block, 1024 chunks per sweep
  __shared__ int64_t wsum[32];
  __shared__ int64_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int64_t b0 = 0; b0 < n_chunks; b0 += 1024) {
    const int64_t b = b0 + threadIdx.x;
    const int64_t v = b < n_chunks ? chunk_count[b] : 0;
    int64_t x = v;
    for (int o = 1; o < 32; o <<= 1) {
      const int64_t y = __shfl_up_sync(0xffffffffu, x, o);
      if ((threadIdx.x & 31) >= o) x += y;
    }
    if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = x;
    __syncthreads();
    int64_t before = carry;
    for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) before += wsum[w];
    if (b < n_chunks) chunk_count[b] = before + x - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry = before + x;
    __syncthreads();
  }
  if (threadIdx.x == 0) *n_objects = carry;
}

__global__ void __launch_bounds__(kIngestThreads) ingest_write_kernel(int64_t n_rows, const int64_t* id, const int64_t* chunk_base, int64_t max_objects,
                                                                      int64_t* offsets, const int64_t* n_objects, const int32_t* passband,
                                                                      const int32_t* lut, int lut_size, int32_t* band, int32_t* unknown) {
  __shared__ int wsum[kIngestThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t base = (int64_t)blockIdx.x * kIngestRows;
  int64_t rank0 = chunk_base[blockIdx.x];
  int bad = 0;
  for (int r0 = 0; r0 < kIngestRows; r0 += kIngestThreads) {
    const int64_t i = base + r0 + threadIdx.x;
    const bool h = i < n_rows && is_head(id, i);
    const unsigned bal = __ballot_sync(0xffffffffu, h);
    if (lane == 0) wsum[warp] = __popc(bal);
    __syncthreads();
    int before = 0, total = 0;
    for (int w = 0; w < kIngestThreads / 32; ++w) { if (w < warp) before += wsum[w]; total += wsum[w]; }
    if (h) {
      const int64_t pos = rank0 + before + __popc(bal & ((1u << lane) - 1u));
      if (pos < max_objects) offsets[pos] = i;
    }
    if (band && i < n_rows) {   // add_log_lam as a table lookup (fulu/_base_aug.py:9-11); unknown passband -> -1, counted
      const int32_t p = passband[i];
      const int32_t bidx = (p >= 0 && p < lut_size) ? lut[p] : -1;
      band[i] = bidx;
      bad += bidx < 0;
    }
    rank0 += total;
    __syncthreads();
  }
  if (unknown && bad) atomicAdd(unknown, bad);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    const int64_t no = *n_objects;
    if (no <= max_objects) offsets[no] = n_rows;
  }
}

__global__ void copy_status_kernel(int64_t B, const int32_t* order, const int32_t* src, int32_t* dst) {
  const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (k >= B) return;
  const int64_t cu = order ? (int64_t)order[k] : k;
  dst[cu] = src[cu] ? FULU_GP_STATUS_BAD_INPUT : 0;
}

// ------------------------------------------------------------------------------------------------ FP64 peak probes
__global__ void __launch_bounds__(256) dfma_probe(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

__global__ void __launch_bounds__(256) dmma_probe(double* out, int iters, double a, double b) {
  double c0[2] = {0, 0}, c1[2] = {0, 0}, c2[2] = {0, 0}, c3[2] = {0, 0};
  double fa = a + threadIdx.x * 1e-9, fb = b;
  for (int i = 0; i < iters; ++i) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0[0]), "+d"(c0[1]) : "d"(fa), "d"(fb));
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c1[0]), "+d"(c1[1]) : "d"(fa), "d"(fb));
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c2[0]), "+d"(c2[1]) : "d"(fa), "d"(fb));
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c3[0]), "+d"(c3[1]) : "d"(fa), "d"(fb));
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = (c0[0] + c0[1]) + (c1[0] + c1[1]) + (c2[0] + c2[1]) + (c3[0] + c3[1]);
}

// ------------------------------------------------------------------------------------------------ host helpers
#define CU_TRY(x)                          \
  do {                                     \
    cudaError_t e__ = (x);                 \
    if (e__ != cudaSuccess) return (int)e__; \
  } while (0)

// the compiled covariance families (gp_family.cu, one translation unit each)
const GpFamilyOps* family_ops(int kernel) {
  switch (kernel) {
#define GP_FAMILY_CASE(f1, ex) \
  case FULU_GP_KERNEL(f1, ex): return GP_FAMILY_SYMBOL(f1, ex)();
    GP_FAMILY_LIST(GP_FAMILY_CASE)
#undef GP_FAMILY_CASE
    default: return nullptr;
  }
}

int check_batch(const fulu_gp_batch* b, int P) {
  if (!b) return FULU_GP_E_BADARG;
  if (b->n_curves < 0 || b->n_obs_total < 0 || b->n_bands < 1 || b->n_bands > FULU_GP_MAX_BANDS) return FULU_GP_E_BADARG;
  if (b->n_curves > 0 && (!b->offsets || !b->t || !b->flux || !b->flux_err || !b->band || !b->loglam)) return FULU_GP_E_BADARG;
  const GpFamilyOps* ops = family_ops(b->kernel);
  if (!ops) return FULU_GP_E_UNSUPPORTED;
  if (P != ops->n_params) return FULU_GP_E_BADARG;
  if (b->n_max < 1 && b->n_curves > 0) return FULU_GP_E_BADARG;
  if (b->order && (b->n_order < 0 || b->n_order > b->n_curves)) return FULU_GP_E_BADARG;
  return 0;
}

inline int64_t n_active(const fulu_gp_batch& b) { return b.order ? b.n_order : b.n_curves; }

int run_prep(const fulu_gp_batch& b, const WsLayout& w, void* ws, int sms, cudaStream_t st, CurveTable& tab) {
  PrepArgs a;
  a.B = n_active(b); a.order = b.order; a.offsets = b.offsets; a.t = b.t; a.flux = b.flux; a.ferr = b.flux_err; a.loglam = b.loglam; a.band = b.band;
  a.n_bands = b.n_bands; a.use_err = b.use_err; a.alpha0 = b.alpha0; a.n_max = b.n_max;
  a.xs = at<double>(ws, w.xs); a.y = at<double>(ws, w.y); a.al = at<double>(ws, w.al); a.lam = at<double>(ws, w.lam);
  a.stats = at<double>(ws, w.stats); a.pstatus = at<int32_t>(ws, w.pstatus);
  int64_t grid = a.B < (int64_t)sms * 16 ? a.B : (int64_t)sms * 16;
  if (grid > 0) prep_kernel<<<(unsigned)grid, kPrepThreads, 0, st>>>(a);
  CU_TRY(cudaGetLastError());
  tab.offsets = b.offsets; tab.xs = a.xs; tab.y = a.y; tab.al = a.al; tab.lam = a.lam; tab.stats = a.stats; tab.band = b.band;
  tab.pstatus = a.pstatus; tab.n_bands = b.n_bands; tab.order = b.order; tab.n_active = a.B;
  return 0;
}

}  // namespace

// ================================================================================================ C ABI
extern "C" {

int fulu_gp_abi_version(void) { return FULU_GP_ABI_VERSION; }

const char* fulu_gp_strerror(int code) {
  switch (code) {
    case 0: return "ok";
    case FULU_GP_E_BADARG: return "fulu_gp: bad argument";
    case FULU_GP_E_WORKSPACE: return "fulu_gp: workspace too small";
    case FULU_GP_E_UNSUPPORTED: return "fulu_gp: kernel family or size not supported by this build";
    case FULU_GP_E_NOCUDA: return "fulu_gp: no CUDA device";
    default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "fulu_gp: unknown error";
  }
}

int fulu_gp_workspace_bytes(const fulu_gp_batch* batch, int32_t n_params, int32_t n_starts, int32_t window, size_t* bytes) {
  int rc = check_batch(batch, n_params);
  if (rc) return rc;
  if (!bytes) return FULU_GP_E_BADARG;
  DeviceInfo di;
  if (device_info(di)) return FULU_GP_E_NOCUDA;
  WsLayout w;
  plan(*batch, n_params, n_starts, window, 0, di.sms, w);
  *bytes = w.total;
  return 0;
}

static int lml_impl(const fulu_gp_batch* batch, int32_t n_params, const double* theta, double* lml, double* grad, double* scaler,
                    double* ynorm, int32_t* status, void* workspace, size_t workspace_bytes, void* stream, long long* marks);

int fulu_gp_lml_batch(const fulu_gp_batch* batch, int32_t n_params, const double* theta, double* lml, double* grad, double* scaler,
                      double* ynorm, int32_t* status, void* workspace, size_t workspace_bytes, void* stream) {
  return lml_impl(batch, n_params, theta, lml, grad, scaler, ynorm, status, workspace, workspace_bytes, stream, nullptr);
}

int fulu_gp_debug_phase_cycles(const fulu_gp_batch* batch, int32_t n_params, const double* theta, double* lml, double* grad,
                               long long* marks, void* workspace, size_t workspace_bytes, void* stream) {
  return lml_impl(batch, n_params, theta, lml, grad, nullptr, nullptr, nullptr, workspace, workspace_bytes, stream, marks);
}

static int lml_impl(const fulu_gp_batch* batch, int32_t n_params, const double* theta, double* lml, double* grad, double* scaler,
                    double* ynorm, int32_t* status, void* workspace, size_t workspace_bytes, void* stream, long long* marks) {
  int rc = check_batch(batch, n_params);
  if (rc) return rc;
  if (n_active(*batch) == 0) return 0;
  if (!theta || !lml || !grad || !workspace) return FULU_GP_E_BADARG;
  DeviceInfo di;
  if (device_info(di)) return FULU_GP_E_NOCUDA;
  WsLayout w;
  plan(*batch, n_params, 1, 1, 0, di.sms, w);
  if (workspace_bytes < w.total) return FULU_GP_E_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  CurveTable tab;
  rc = run_prep(*batch, w, workspace, di.sms, st, tab);
  if (rc) return rc;
  const int64_t B = n_active(*batch);
  unsigned grid = (unsigned)(B < w.grid_eval ? B : w.grid_eval);
  rc = family_ops(batch->kernel)->lml(w.place, grid, w.threads_eval, w.smem_eval, st, tab, B, n_params, theta, lml, grad,
                                      at<double>(workspace, w.gA), w.gA_stride, marks, w.noscratch ? 1 : 0);
  if (rc) return rc;
  const unsigned g1 = (unsigned)((B + 255) / 256);
  if (scaler || ynorm) export_stats_kernel<<<g1, 256, 0, st>>>(B, tab.order, tab.stats, scaler, ynorm);
  if (status) copy_status_kernel<<<g1, 256, 0, st>>>(B, tab.order, tab.pstatus, status);
  CU_TRY(cudaGetLastError());
  return 0;
}

int fulu_gp_fit_batch(const fulu_gp_batch* batch, const fulu_gp_fit_options* opt, const double* theta_starts, double* theta_opt,
                      double* lml_opt, double* scaler, double* ynorm, int32_t* n_eval, int32_t* status, double* run_lml, double* run_theta,
                      int32_t* run_n_eval, void* workspace, size_t workspace_bytes, void* stream) {
  if (!opt) return FULU_GP_E_BADARG;
  int rc = check_batch(batch, opt->n_params);
  if (rc) return rc;
  if (n_active(*batch) == 0) return 0;
  if (!theta_starts || !theta_opt || !lml_opt || !n_eval || !status || !workspace) return FULU_GP_E_BADARG;
  if (opt->n_starts < 1 || opt->m < 1 || opt->m > FG_LBFGS_M) return FULU_GP_E_BADARG;
  DeviceInfo di;
  if (device_info(di)) return FULU_GP_E_NOCUDA;
  WsLayout w;
  plan(*batch, opt->n_params, opt->n_starts, opt->window, opt->max_rounds, di.sms, w);
  if (workspace_bytes < w.total) return FULU_GP_E_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  CurveTable tab;
  rc = run_prep(*batch, w, workspace, di.sms, st, tab);
  if (rc) return rc;

  const int P = opt->n_params, S = opt->n_starts;
  FitDev d;
  d.tab = tab;
  d.opt.n = P; d.opt.m = opt->m; d.opt.maxiter = opt->maxiter; d.opt.maxfun = opt->maxfun; d.opt.maxls = opt->maxls; d.opt.pad = 0;
  d.opt.factr = opt->factr; d.opt.pgtol = opt->pgtol;
  for (int k = 0; k < FG_MAXP; ++k) { d.opt.lo[k] = k < P ? opt->lower[k] : 0.0; d.opt.hi[k] = k < P ? opt->upper[k] : 0.0; }
  d.n_tasks = n_active(*batch) * (int64_t)S;
  d.S = S; d.P = P; d.window = w.window;
  d.slots = at<FgState>(workspace, w.slots);
  d.slot_task = at<int64_t>(workspace, w.slot_task);
  d.slot_nev = at<int32_t>(workspace, w.slot_nev);
  d.f_out = at<double>(workspace, w.f_out);
  d.g_out = at<double>(workspace, w.g_out);
  d.list = at<int32_t>(workspace, w.list);
  d.counts = at<int32_t>(workspace, w.counts);
  d.next_task = at<unsigned long long>(workspace, w.next_task);
  d.ended = d.counts + w.max_rounds + 1;   // (the last counter is never a round's: rounds end at max_rounds)
  d.ended_mark = d.counts + w.max_rounds + 1;
  d.run_f = at<double>(workspace, w.run_f);
  d.run_x = at<double>(workspace, w.run_x);
  d.run_nev = at<int32_t>(workspace, w.run_nev);
  d.run_end = at<int32_t>(workspace, w.run_end);
  d.starts = theta_starts;
  d.starts_per_curve = opt->starts_per_curve != 0;
  d.gA = at<double>(workspace, w.gA);
  d.gA_stride = w.gA_stride;
  d.noscratch = w.noscratch ? 1 : 0;

  const GpFamilyOps* ops = family_ops(batch->kernel);
  const bool use_tail = getenv("FULU_GP_NO_TAIL") == nullptr;
  const bool streamed = w.stream_tail && use_tail;
  // The fused fit (fit_fused_kernel: one persistent CTA per SM, evaluation groups + stepper warps; no rounds, no optimiser-step kernel,
  // one launch per fit) is OPT-IN (FULU_GP_FUSED=1): bit-identical results, but measured 9-10 % slower than the rounds on B200 at
  // N = 100 (197 ms against 180 ms for the 10,000-curve fit) and on the ragged config #4 - the stepper warps' long branchy code and
  // its local-memory state slow the co-resident evaluation warps by more than the step kernel costs as a launch of its own
  // (DESIGN.md section 6).  It takes fits with more runs than one wave of evaluation CTAs whose matrix is worked on by the 8x8 block
  // code (shared-memory or slab placement); an explicit window or round cap always means the rounds.
  FusedGeom geo;
  geo.groups = w.occ_eval;
  geo.group_threads = w.threads_eval;
  geo.group_smem = (w.smem_eval + 15) / 16 * 2;   // doubles, 16-byte granules
  {
    const char* e = getenv("FULU_GP_FUSED_SLOTS");
    geo.slots = e ? atoi(e) : kFusedDefaultSlots;   // (32 slots, one stepper warp: the best of the sweep in profiles/r02_fused_fit.md)
    if (geo.slots < 1) geo.slots = 1;
    if (geo.slots > kFusedMaxSlots) geo.slots = kFusedMaxSlots;
    e = getenv("FULU_GP_FUSED_ROTATE");
    geo.rotate = e ? atoi(e) : 1;
    e = getenv("FULU_GP_FUSED_STEPPERS");
    geo.steppers = e ? atoi(e) : kFusedDefaultSteppers;
    if (geo.steppers < 1) geo.steppers = 1;
    if (geo.steppers > kFusedMaxSteppers) geo.steppers = kFusedMaxSteppers;
    if (geo.steppers > geo.slots) geo.steppers = geo.slots;
  }
  const size_t fused_smem = (size_t)geo.groups * geo.group_smem * sizeof(double);
  const char* fused_env = getenv("FULU_GP_FUSED");
  const bool fused = use_tail && !streamed && fused_env != nullptr && atoi(fused_env) != 0 && opt->window <= 0 && opt->max_rounds <= 0 &&
                     w.place != GP_PLACE_TILED && d.n_tasks > (int64_t)w.grid_eval && geo.groups <= kFusedMaxGroups &&
                     geo.groups * geo.group_threads <= kEvalMaxThreads && fused_smem + sizeof(FusedShared) + 64 <= kMaxSmem;

  CU_TRY(cudaMemsetAsync(d.counts, 0, ((size_t)w.max_rounds + 2) * 4, st));
  CU_TRY(cudaMemsetAsync(d.next_task, 0, 8, st));
  if (!fused) {
    fit_init_kernel<<<(w.window + 255) / 256, 256, 0, st>>>(d);
    CU_TRY(cudaGetLastError());
    rc = ops->eval_prepare(w.place, w.smem_eval);
    if (rc) return rc;
  }

  const unsigned step_grid = (unsigned)((w.window + GP_STEP_THREADS - 1) / GP_STEP_THREADS);
  const bool prof = opt->profile != nullptr && opt->profile_counts_only == 0;   // per-launch CUDA events (opt-in diagnostics)
  int launches = 3;   // prep, init, select
  std::vector<cudaEvent_t> ev;   // 3 per round when profiling; read back only after the last launch so timing does not perturb it
  int round = 0;
  bool finished = false;   // false only for the A/B mode that forbids the fused end (FULU_GP_NO_TAIL=1)
  std::vector<cudaEvent_t> tev;   // profiling: two events per fused launch (its time counts as evaluation time: the steps are inside)
  auto launch_tail = [&](unsigned ctas, int at_round, int tail_below) -> int {
    if (prof) {
      for (int i = 0; i < 2; ++i) { cudaEvent_t e; CU_TRY(cudaEventCreate(&e)); tev.push_back(e); }
      cudaEventRecord(tev[tev.size() - 2], st);
    }
    ops->tail(w.place, ctas, w.threads_eval, w.smem_eval, st, d, at_round, tail_below);
    if (prof) cudaEventRecord(tev[tev.size() - 1], st);
    CU_TRY(cudaGetLastError());
    ++launches;
    return 0;
  };
  constexpr int kAny = 0x7fffffff;
  if (fused) {
    if (prof) {
      for (int i = 0; i < 2; ++i) { cudaEvent_t e; CU_TRY(cudaEventCreate(&e)); tev.push_back(e); }
      cudaEventRecord(tev[0], st);
    }
    const int64_t want_ctas = (d.n_tasks + geo.slots - 1) / geo.slots;   // (a CTA without a task would exit at once)
    const unsigned ctas = (unsigned)(want_ctas < (int64_t)di.sms ? want_ctas : (int64_t)di.sms);
    if ((rc = ops->fused(w.place, ctas, fused_smem, st, d, geo))) return rc;
    if (prof) cudaEventRecord(tev[1], st);
    --launches;   // (no init kernel)
    ++launches;
    finished = true;
  } else if ((use_tail && d.n_tasks <= (int64_t)w.grid_eval) || streamed) {
    // a small fit (every task has a CTA of its own) or a streaming class: ONE launch, CTA b owns slot b
    const unsigned ctas = (unsigned)(d.n_tasks < (int64_t)w.window ? d.n_tasks : (int64_t)w.window);   // fit_init_kernel filled that many slots at most
    if ((rc = launch_tail(ctas, 0, kAny))) return rc;
    finished = true;
  } else {
    // Rounds of two launches (evaluate every live slot; step every live slot), enqueued WITHOUT the host ever looking at the device:
    // the call is stream-ordered and returns at once.  How many rounds a fit needs is only known on the device (the longest run of the
    // batch: ~80-130 evaluations, more when the window is smaller than the task list), so a generous number is enqueued; a round that
    // finds no live slot costs two empty launches (~2 us each).  Every fourth round is preceded by the fused kernel with "at most one
    // wave of live slots" as its condition: the first time it holds, that launch finishes the fit and every later round is empty.
    // Behind the last round one fused launch runs whatever might still be live to the end, so the result never depends on the number
    // of rounds chosen here.
    const int64_t waves = (d.n_tasks + w.window - 1) / w.window;
    // (the 5- and 6-parameter families run longer: their longest runs of the 10,000-curve fit take 250-270 evaluations against ~125
    // for the default kernel; with only 160 rounds enqueued a third of the rbf_rq fit ended in the one-CTA-per-slot launch behind the
    // last round: 866 ms instead of 650)
    int64_t want = use_tail ? 160 + 100 * (P > 4 ? P - 4 : 0) + 96 * (waves - 1) : 4096;
    if (opt->max_rounds > 0 && want > opt->max_rounds) want = opt->max_rounds;
    if (want > w.max_rounds) want = w.max_rounds;
    for (; round < (int)want; ++round) {
      if (use_tail && round > 0 && (round & 3) == 0 && (rc = launch_tail((unsigned)w.grid_eval, round, w.grid_eval))) return rc;
      if (prof) {
        for (int i = 0; i < 3; ++i) { cudaEvent_t e; CU_TRY(cudaEventCreate(&e)); ev.push_back(e); }
        cudaEventRecord(ev[3 * round], st);
      }
      ops->eval(w.place, w.grid_eval, w.threads_eval, w.smem_eval, st, d, round);
      if (prof) cudaEventRecord(ev[3 * round + 1], st);
      // (one instantiation per number of hyperparameters: its loops over the variables unroll exactly, the dense B lives in registers)
      switch (P) {
        case 4: fit_step_kernel<4><<<step_grid, GP_STEP_THREADS, 0, st>>>(d, round); break;
        case 5: fit_step_kernel<5><<<step_grid, GP_STEP_THREADS, 0, st>>>(d, round); break;
        case 6: fit_step_kernel<6><<<step_grid, GP_STEP_THREADS, 0, st>>>(d, round); break;
        default: fit_step_kernel<0><<<step_grid, GP_STEP_THREADS, 0, st>>>(d, round); break;
      }
      if (prof) cudaEventRecord(ev[3 * round + 2], st);
      launches += 2;
    }
    CU_TRY(cudaGetLastError());
    if (use_tail && opt->max_rounds <= 0) {
      if ((rc = launch_tail((unsigned)w.grid_eval, round, kAny))) return rc;
      finished = true;
    }
  }
  if (opt->profile != nullptr) {
    double eval_ms = 0.0, step_ms = 0.0, tail_ms = 0.0;
    if (prof) {
      for (int r = 0; r < round; ++r) {
        float a = 0.f, b = 0.f;
        cudaEventSynchronize(ev[3 * r + 2]);
        cudaEventElapsedTime(&a, ev[3 * r], ev[3 * r + 1]);
        cudaEventElapsedTime(&b, ev[3 * r + 1], ev[3 * r + 2]);
        eval_ms += a;
        step_ms += b;
      }
      for (cudaEvent_t e : ev) cudaEventDestroy(e);
      for (size_t i = 0; i + 1 < tev.size(); i += 2) {
        float a = 0.f;
        cudaEventSynchronize(tev[i + 1]);
        cudaEventElapsedTime(&a, tev[i], tev[i + 1]);
        eval_ms += a;
        tail_ms += a;
      }
      for (cudaEvent_t e : tev) cudaEventDestroy(e);
    }
    double* pr = opt->profile;
    pr[0] = round; pr[1] = eval_ms; pr[2] = step_ms; pr[3] = launches + (finished ? 0 : 1); pr[4] = w.window; pr[5] = w.grid_eval;
    pr[6] = (double)w.smem_eval; pr[7] = w.globalA ? 1.0 : 0.0; pr[8] = w.threads_eval; pr[9] = w.occ_eval; pr[10] = tail_ms; pr[11] = w.place;
    if (fused) { pr[4] = (double)geo.slots * di.sms; pr[5] = di.sms; pr[6] = (double)fused_smem; }
  }
  if (!finished) {
    fit_flush_kernel<<<(w.window + 255) / 256, 256, 0, st>>>(d);
    CU_TRY(cudaGetLastError());
  }
  SelectArgs sa;
  sa.B = n_active(*batch); sa.order = batch->order; sa.S = S; sa.P = P;
  sa.run_f = d.run_f; sa.run_x = d.run_x; sa.run_nev = d.run_nev; sa.run_end = d.run_end; sa.pstatus = tab.pstatus;
  for (int k = 0; k < FG_MAXP; ++k) { sa.lo[k] = d.opt.lo[k]; sa.hi[k] = d.opt.hi[k]; }
  sa.theta_opt = theta_opt; sa.lml_opt = lml_opt; sa.run_lml = run_lml; sa.run_theta = run_theta; sa.n_eval = n_eval; sa.status = status; sa.run_n_eval = run_n_eval;
  fit_select_kernel<<<(unsigned)((sa.B + 127) / 128), 128, 0, st>>>(sa);
  if (scaler || ynorm) export_stats_kernel<<<(unsigned)((sa.B + 255) / 256), 256, 0, st>>>(sa.B, tab.order, tab.stats, scaler, ynorm);
  CU_TRY(cudaGetLastError());
  return 0;
}

static int predict_common(const fulu_gp_batch* batch, int32_t n_params, PredictArgs& a, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_batch(batch, n_params);
  if (rc) return rc;
  if (n_active(*batch) == 0) return 0;
  if (!a.theta || !a.mean || !a.stdv || !workspace) return FULU_GP_E_BADARG;
  DeviceInfo di;
  if (device_info(di)) return FULU_GP_E_NOCUDA;
  WsLayout w;
  plan(*batch, n_params, 1, 1, 0, di.sms, w);
  if (workspace_bytes < w.total) return FULU_GP_E_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  CurveTable tab;
  rc = run_prep(*batch, w, workspace, di.sms, st, tab);
  if (rc) return rc;
  a.tab = tab; a.B = n_active(*batch); a.P = n_params; a.t_raw = batch->t;
  a.gA = at<double>(workspace, w.gA); a.gV = at<double>(workspace, w.gV);
  a.gA_stride = w.gA_stride; a.gV_stride = w.gV_stride;
  unsigned grid = (unsigned)(a.B < w.grid_predict ? a.B : w.grid_predict);
  return family_ops(batch->kernel)->predict(w.place, grid, w.smem_predict, st, a);
}

int fulu_gp_predict_grid_batch(const fulu_gp_batch* batch, int32_t n_params, const double* theta, const double* t_min, const double* t_max,
                               int32_t n_obs, double* mean, double* std, double* lml, int32_t* status, void* workspace,
                               size_t workspace_bytes, void* stream) {
  if (n_obs < 1) return FULU_GP_E_BADARG;
  PredictArgs a = {};
  a.theta = theta; a.n_obs = n_obs; a.t_min = t_min; a.t_max = t_max; a.mean = mean; a.stdv = std; a.lml = lml; a.status = status;
  return predict_common(batch, n_params, a, workspace, workspace_bytes, stream);
}

int fulu_gp_predict_points_batch(const fulu_gp_batch* batch, int32_t n_params, const double* theta, const int64_t* q_offsets, const double* q_t,
                                 const int32_t* q_band, int32_t q_max, double* mean, double* std, int32_t* status, void* workspace,
                                 size_t workspace_bytes, void* stream) {
  (void)q_max;
  if (!q_offsets || !q_t || !q_band) return FULU_GP_E_BADARG;
  PredictArgs a = {};
  a.theta = theta; a.n_obs = 0; a.q_offsets = q_offsets; a.q_t = q_t; a.q_band = q_band; a.mean = mean; a.stdv = std; a.status = status;
  return predict_common(batch, n_params, a, workspace, workspace_bytes, stream);
}

int fulu_gp_peak_batch(const fulu_gp_batch* batch, const double* mean, const double* t_min, const double* t_max, int32_t n_obs, double* sum_flux,
                       double* peak_t, double* peak_flux, int32_t* peak_index, void* stream) {
  if (!batch || batch->n_curves < 0 || batch->n_bands < 1 || n_obs < 1) return FULU_GP_E_BADARG;
  if (batch->order && (batch->n_order < 0 || batch->n_order > batch->n_curves)) return FULU_GP_E_BADARG;
  const int64_t B = n_active(*batch);
  if (B == 0) return 0;
  if (!mean || !peak_t || !peak_flux || !peak_index || (t_min == nullptr) != (t_max == nullptr)) return FULU_GP_E_BADARG;
  if (!t_min && (!batch->offsets || !batch->t)) return FULU_GP_E_BADARG;
  PeakArgs a;
  a.B = B; a.order = batch->order; a.offsets = batch->offsets; a.t_raw = batch->t; a.mean = mean; a.t_min = t_min; a.t_max = t_max;
  a.n_bands = batch->n_bands; a.n_obs = n_obs; a.sum_flux = sum_flux; a.peak_t = peak_t; a.peak_flux = peak_flux; a.peak_index = peak_index;
  peak_kernel<<<(unsigned)((B * 32 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a);
  CU_TRY(cudaGetLastError());
  return 0;
}

int fulu_gp_ingest_workspace_bytes(int64_t n_rows, size_t* bytes) {
  if (n_rows < 0 || !bytes) return FULU_GP_E_BADARG;
  *bytes = align_up((size_t)((n_rows + kIngestRows - 1) / kIngestRows + 1) * 8);
  return 0;
}

int fulu_gp_ingest_table(int64_t n_rows, const int64_t* object_id, const int32_t* passband, const int32_t* lut, int32_t lut_size,
                         int64_t max_objects, int64_t* offsets, int64_t* n_objects, int32_t* band, int32_t* unknown_bands, void* workspace,
                         size_t workspace_bytes, void* stream) {
  if (n_rows < 0 || max_objects < 0 || !offsets || !n_objects) return FULU_GP_E_BADARG;
  if (n_rows > 0 && !object_id) return FULU_GP_E_BADARG;
  if (band && (!passband || !lut || lut_size < 1)) return FULU_GP_E_BADARG;
  size_t need = 0;
  fulu_gp_ingest_workspace_bytes(n_rows, &need);
  if (!workspace || workspace_bytes < need) return FULU_GP_E_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  if (unknown_bands) CU_TRY(cudaMemsetAsync(unknown_bands, 0, 4, st));
  const int64_t n_chunks = (n_rows + kIngestRows - 1) / kIngestRows;
  int64_t* chunk = reinterpret_cast<int64_t*>(workspace);
  if (n_chunks > 0) ingest_count_kernel<<<(unsigned)n_chunks, kIngestThreads, 0, st>>>(n_rows, object_id, chunk);
  ingest_scan_kernel<<<1, 1024, 0, st>>>(n_chunks, chunk, n_objects);
  if (n_chunks > 0)
    ingest_write_kernel<<<(unsigned)n_chunks, kIngestThreads, 0, st>>>(n_rows, object_id, chunk, max_objects, offsets, n_objects, passband, lut,
                                                                      lut_size, band, unknown_bands);
  else CU_TRY(cudaMemsetAsync(offsets, 0, 8, st));
  CU_TRY(cudaGetLastError());
  return 0;
}

int fulu_gp_fp64_probe(int32_t mode, int32_t iters, double* scratch, size_t scratch_bytes, double* flop, void* stream) {
  if (!flop || !scratch || iters < 1) return FULU_GP_E_BADARG;
  DeviceInfo di;
  if (device_info(di)) return FULU_GP_E_NOCUDA;
  const int blocks = di.sms * 8;
  if (scratch_bytes < (size_t)blocks * 256 * 8) return FULU_GP_E_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  if (mode == 0) dfma_probe<<<blocks, 256, 0, st>>>(scratch, iters, 1.0000001, 1e-9);
  else dmma_probe<<<blocks, 256, 0, st>>>(scratch, iters, 1.0000001, 1e-9);
  CU_TRY(cudaGetLastError());
  // DFMA: 8 FMA per thread per iteration; DMMA m8n8k4: 4 x 256 FMA per warp per iteration
  const double fma_count = mode == 0 ? (double)blocks * 256.0 * 8.0 * iters : (double)blocks * 8.0 * 4.0 * 256.0 * iters;
  *flop = 2.0 * fma_count;
  return 0;
}

}  // extern "C"
