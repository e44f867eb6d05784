// Device-side pieces shared by the per-family translation units (gp_family.cu, one per covariance family) and the C ABI
// (gp_kernels.cu): the CUDA block context, the curve table, the optimiser-window descriptor, and the three kernels whose code
// depends on the covariance family - fixed-theta LML, the evaluation kernel of the fit, and prediction.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#ifdef GP_FUSED_STATS
#include <stdio.h>
#endif

#include "../../include/fulu_gp.h"
#include "gp_block.cuh"
#include "gp_tiled.cuh"
#include "lbfgsb.cuh"

namespace {

struct DevCtx {
  int tid, nthr, lane, warp, nwarp;
  long long* marks;   // optional phase timestamps (fulu_gp_debug_phase_cycles); null in production launches
  // warp index through a shuffle: the compiler then knows it is warp-uniform, so loops and branches on it stay convergent and
  // the mma.sync sites need no WARPSYNC / collective wrappers
  __device__ __forceinline__ DevCtx() : tid((int)threadIdx.x), nthr((int)blockDim.x), lane((int)threadIdx.x & 31),
                                        warp(__shfl_sync(0xffffffffu, (int)threadIdx.x >> 5, 0)), nwarp(((int)blockDim.x + 31) >> 5),
                                        marks(nullptr) {}
  __device__ __forceinline__ void mark(int id) {
    if (marks != nullptr && tid == 0) marks[id] = clock64();
  }
  __device__ __forceinline__ long long clock() { return marks != nullptr ? clock64() : 0; }
  __device__ __forceinline__ void mark_value(int id, long long v) {
    if (marks != nullptr && tid == 0) marks[id] = v;
  }
  __device__ __forceinline__ void sync() { __syncthreads(); }
  __device__ __forceinline__ void warp_sync() { __syncwarp(); }
  // FP64 tensor-core tile: C(8x8) += A(8x4, row) * B(4x8, col); lane (g,t): a = A[g][t], b = B[t][g], c0 = C[g][2t], c1 = C[g][2t+1]
  __device__ __forceinline__ void mma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
  }
  __device__ __forceinline__ double shfl_xor(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
  __device__ __forceinline__ double shfl(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
  __device__ __forceinline__ double shfl4(double v, int src) { return __shfl_sync(0xffffffffu, v, src, 4); }   // from lane `src` of this lane's group of four
  // ---- bulk-asynchronous copies and their barriers (gp_tiled.cuh): SASS UBLKCP / SYNCS
  static __device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
  __device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  }
  __device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
  }
  __device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
  }
  // waits until the phase of the given parity has completed; a wait that lasts ~10 s is a protocol bug: trap instead of hanging the GPU
  __device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t a = smem_u32(bar);
    long long t0 = 0;
    for (int spins = 0;; ++spins) {
      uint32_t done;
      asm volatile(
          "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
          : "=r"(done)
          : "r"(a), "r"(parity)
          : "memory");
      if (done) break;
      if (spins > 4096) {
        if (t0 == 0) t0 = clock64();
        else if (clock64() - t0 > 20000000000LL) __trap();
      }
    }
  }
  // (the producer's wait for a stage to be released; a name of its own so that the host harness can drop it in its self-test)
  __device__ __forceinline__ void mbar_wait_reuse(uint64_t* bar, uint32_t parity) { mbar_wait(bar, parity); }
  // global -> shared, `bytes` a multiple of 16, both addresses 16-byte aligned; completes `bytes` on the barrier
  __device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
  }
  // orders this thread's generic-proxy writes (global or shared) before later asynchronous-proxy accesses (the copy engine)
  __device__ __forceinline__ void fence_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
};

using TiledCfg = GpTiledCfgDefault;

// The block context of a kernel whose matrix lives in placement kPlace (the GP_PLACE_* values below): it also tells the block code
// whether its inner product loops run rolled (operands in shared memory) or as the compiler unrolls them (operands in a global slab:
// gp_block.cuh, GpLoops).
template <int kPlace>
struct DevCtxP : DevCtx {
  static constexpr bool kRolled = kPlace != 1;          // factorisation, inverse, traces: shared memory, or the 64 x 64 blocks of the tiled path
  static constexpr bool kRolledPredict = kPlace == 0;   // forward substitution against the factor
};

// Threads per SM given to the evaluation kernel, split evenly between the resident CTAs; it also fixes the register budget
// (65536 / threads).  Measured on B200 at N = 100 (4 CTAs per SM), 10,000 curves: 1024 threads (64 registers) 229 ms,
// 768 (80) 213 ms, 640 (96) 206 ms, 512 (128) 203 ms, 384 (168) 216 ms, 256 (255) 255 ms per fit - with few registers the
// compiler funnels the fragment loads of the DMMA loops through one register pair and serialises load -> DMMA -> load.
#ifndef GP_EVAL_MAX_THREADS
#define GP_EVAL_MAX_THREADS 512
#endif
constexpr int kEvalMaxThreads = GP_EVAL_MAX_THREADS;
constexpr int kPredictThreads = 256;   // eight warps: each takes eight query points at a time (gp_predict_tile)
constexpr int kPrepThreads = 128;
constexpr size_t kMaxSmem = 227 * 1024;

struct CurveTable {
  const int64_t* offsets;
  const double *xs, *y, *al, *lam, *stats;
  const int32_t* band;
  const int32_t* pstatus;
  const int32_t* order;   // optional processing order (curve ids); null = identity
  int64_t n_active;       // curves this call processes
  int n_bands;
};

__device__ __forceinline__ int64_t curve_of(const CurveTable& t, int64_t k) { return t.order ? (int64_t)t.order[k] : k; }

__device__ __forceinline__ GpCurveView curve_view(const CurveTable& t, int64_t cu) {
  const int64_t o = t.offsets[cu];
  GpCurveView v;
  v.n = (int)(t.offsets[cu + 1] - o);
  v.xs = t.xs + o; v.band = t.band + o; v.y = t.y + o; v.al = t.al + o; v.lam = t.lam + cu * t.n_bands;
  return v;
}

// Where a length class keeps a curve's matrix (plan() in gp_kernels.cu):
//   GP_PLACE_SMEM   packed tiles in shared memory, 1 .. 16 CTAs per SM                                      (N <= 223)
//   GP_PLACE_SLAB   the same 8x8 block code on a per-CTA slab in global memory (L2 resident), two CTAs per SM (N <= 511)
//   GP_PLACE_TILED  slab in global memory streamed through shared-memory stages by bulk copies (gp_tiled.cuh), one CTA per SM
enum { GP_PLACE_SMEM = 0, GP_PLACE_SLAB = 1, GP_PLACE_TILED = 2 };

template <bool kGlobalA>
__device__ __forceinline__ GpSmem carve(double* smem, double* gA, size_t gA_stride, int n, bool noscratch = false) {
  if (!kGlobalA) {
    // (noscratch: the scratch-free layout of the classes where it buys a resident CTA - plan() in gp_kernels.cu, gp_block.cuh)
    if (noscratch) return gp_carve(smem, smem + gp_mat_doubles_ns(gp_pad(n)), n, false);
    return gp_carve(smem, smem + gp_mat_doubles(gp_pad(n)), n);
  }
  // vectors in shared memory, the matrix in this CTA's global slab (stays in L2)
  return gp_carve(gA + (size_t)blockIdx.x * gA_stride, smem, n);
}

// shared-memory classes: the same per-CTA slab region keeps a copy of the assembled K tiles for the gradient pass (GpSmem::kslab)
template <bool kGlobalA>
__device__ __forceinline__ void attach_kslab(GpSmem& s, double* gA, size_t gA_stride) {
  if (!kGlobalA && gA != nullptr && gA_stride > 0) s.kslab = gA + (size_t)blockIdx.x * gA_stride;
}

// One LML + gradient evaluation by this CTA, whichever way the curve's class keeps its matrix: in shared memory (gp_block.cuh) or
// as a slab in global memory streamed through shared-memory stages (gp_tiled.cuh).  pipe_it: the CTA's position in its ring of
// stages (tiled path; carried from evaluation to evaluation).
template <class Kern, int kPlace>
__device__ __forceinline__ void eval_curve(DevCtxP<kPlace>& c, const GpCurveView& cv, const double* theta, double* smem, double* gA, size_t gA_stride,
                                           uint32_t& pipe_it, double& l, double* g, bool& ok, bool noscratch = false) {
  if constexpr (kPlace == GP_PLACE_TILED) {
    GpTiled tl = gp_tiled_carve<TiledCfg>(gA + (size_t)blockIdx.x * gA_stride, smem, cv.n, c.nwarp, pipe_it);
    gp_tiled_eval_lml_grad<Kern, TiledCfg>(c, cv, theta, tl, l, g, ok);
    pipe_it = tl.it;
  } else {
    GpSmem s = carve<kPlace != GP_PLACE_SMEM>(smem, gA, gA_stride, cv.n, noscratch);
    attach_kslab<kPlace != GP_PLACE_SMEM>(s, gA, gA_stride);
    gp_eval_lml_grad<Kern>(c, cv, theta, s, l, g, ok);
  }
}

template <int kPlace>
__device__ __forceinline__ void eval_init(DevCtxP<kPlace>& c, double* smem) {
  if constexpr (kPlace == GP_PLACE_TILED) {
    GpTiled tl = gp_tiled_carve<TiledCfg>(smem, smem, 1, c.nwarp, 0);   // (only the shared-memory part is used)
    gp_tiled_init<TiledCfg>(c, tl);
  }
}

// LML + gradient at one theta per curve (fixed-theta parity entry point)
template <class Kern, int kPlace>
__global__ void __launch_bounds__(kEvalMaxThreads, 1) lml_kernel(CurveTable tab, int64_t B, int P, const double* theta, double* lml, double* grad,
                                                            double* gA, size_t gA_stride, long long* marks, int noscratch) {
  extern __shared__ __align__(16) double smem[];
  DevCtxP<kPlace> c;
  if (blockIdx.x == 0) c.marks = marks;
  uint32_t pipe_it = 0;
  eval_init<kPlace>(c, smem);
  for (int64_t kk = blockIdx.x; kk < B; kk += gridDim.x) {
    const int64_t cu = curve_of(tab, kk);
    if (tab.pstatus[cu] != 0) {
      if (threadIdx.x == 0) { lml[cu] = NAN; for (int k = 0; k < P; ++k) grad[cu * P + k] = NAN; }
      continue;
    }
    GpCurveView cv = curve_view(tab, cu);
    double l, g[Kern::P];
    bool ok;
    eval_curve<Kern, kPlace>(c, cv, theta + cu * P, smem, gA, gA_stride, pipe_it, l, g, ok, noscratch != 0);
    if (threadIdx.x == 0) { lml[cu] = l; for (int k = 0; k < Kern::P; ++k) grad[cu * P + k] = g[k]; }
    __syncthreads();
  }
}

struct FitDev {
  CurveTable tab;
  FgOptions opt;
  int64_t n_tasks;
  int S, P, window;
  FgState* slots;
  int64_t* slot_task;
  int32_t* slot_nev;
  double *f_out, *g_out;
  int32_t* list;     // [2][window]
  int32_t* counts;   // [max_rounds + 2]
  const volatile int32_t* ended;   // nonzero once a fit_tail_kernel launch has taken over what was live (its stamp, round + 1): every later launch has nothing to do
  int32_t* ended_mark;             // the same word, for the launch that sets it
  unsigned long long* next_task;
  double *run_f, *run_x;
  int32_t *run_nev, *run_end;
  const double* starts;   // [S, P], or [B, S, P] by curve id when starts_per_curve
  bool starts_per_curve;
  double* gA;
  size_t gA_stride;
  int noscratch;   // 1: the scratch-free shared-memory layout (plan() in gp_kernels.cu)
};

// pulls the next runnable task into `slot` (whose optimiser state is `state`); returns false when the queue is empty
__device__ __forceinline__ bool acquire_task(const FitDev& d, int slot, FgState& state) {
  for (;;) {
    unsigned long long nt = atomicAdd(d.next_task, 1ULL);
    if (nt >= (unsigned long long)d.n_tasks) { d.slot_task[slot] = -1; state.task = FG_TASK_IDLE; return false; }
    const int64_t cu = curve_of(d.tab, (int64_t)(nt / d.S));
    const int st = (int)(nt % d.S);
    // start points: [S, P] shared by all curves, or [B, S, P] by curve id
    const double* x0 = d.starts + ((d.starts_per_curve ? cu * d.S : 0) + st) * d.P;
    if (d.tab.pstatus[cu] != 0) {   // unusable curve: record and move on
      d.run_f[nt] = INFINITY;
      for (int k = 0; k < d.P; ++k) d.run_x[nt * FG_MAXP + k] = x0[k];
      d.run_nev[nt] = 0;
      d.run_end[nt] = FG_TASK_ABNORMAL;
      continue;
    }
    lbfgsb_init(state, d.opt, x0);
    d.slot_task[slot] = (int64_t)nt;
    d.slot_nev[slot] = 0;
    return true;
  }
}

// One optimiser step of `slot` after its objective evaluation (f, g): advances the L-BFGS-B state machine where the 1.5 KB state
// lies in global memory; a finished run records its result and pulls the next task into the slot.  Returns false when the slot has
// nothing left to do.  (GP_STEP_LOCAL=1: on a thread-local copy, one pipelined copy in and one out.  With run-time loop bounds the
// copy was the faster way, 14.0 against 14.2 ms per fit in round 1; with the number of variables a compile-time constant the small
// vectors are registers anyway and only the stored pairs are touched where they lie: 160.7 -> 159.2 ms.)
#ifndef GP_STEP_LOCAL
#define GP_STEP_LOCAL 0
#endif
// (the state travels as 16-byte words, 23 in flight at a time, in a rolled loop: a member-wise struct copy is ~1,200 instructions in and out)
static_assert(sizeof(FgState) % (16 * 23) == 0, "FgState is copied as 4 x 23 16-byte words");
__device__ __forceinline__ void fg_state_copy(FgState* dst, const FgState* src) {
  double2* o = reinterpret_cast<double2*>(dst);
  const double2* i = reinterpret_cast<const double2*>(src);
#pragma unroll 1
  for (int k0 = 0; k0 < (int)(sizeof(FgState) / 16); k0 += 23) {
    double2 v[23];
#pragma unroll
    for (int k = 0; k < 23; ++k) v[k] = i[k0 + k];
#pragma unroll
    for (int k = 0; k < 23; ++k) o[k0 + k] = v[k];
  }
}

// NV = the family's number of hyperparameters (compile time), 0 = taken from the options
template <int NV>
__device__ __forceinline__ bool advance_slot(const FitDev& d, int slot, double f, const double* g) {
#if GP_STEP_LOCAL
  FgState st;
  fg_state_copy(&st, &d.slots[slot]);
#else
  FgState& st = d.slots[slot];
#endif
#ifdef FG_RUNTIME_N   // (A/B: the number of variables read from the options, loops with run-time bounds)
  lbfgsb_advance<0>(st, d.opt, f, g);
#else
  lbfgsb_advance<NV>(st, d.opt, f, g);
#endif
  d.slot_nev[slot] += 1;
  bool live = true;
  if (fg_task_done(st.task)) {
    const int64_t task = d.slot_task[slot];
    d.run_f[task] = st.f;
    for (int k = 0; k < FG_MAXP; ++k) d.run_x[task * FG_MAXP + k] = st.x[k];
    d.run_nev[task] = d.slot_nev[slot];
    d.run_end[task] = st.task;
    live = acquire_task(d, slot, st);
  }
#if GP_STEP_LOCAL
  fg_state_copy(&d.slots[slot], &st);
#endif
  return live;
}

// One slot run to the end by this CTA: evaluation (all warps) and optimiser step (thread 0) in turn until the slot has nothing left
// to do (a finished run pulls the next task of the queue into the slot).
template <class Kern, int kPlace>
__device__ __forceinline__ void run_slot_to_end(DevCtxP<kPlace>& c, const FitDev& d, int slot, double* smem, uint32_t& pipe_it, int* go) {
  for (;;) {
    const int64_t cu = curve_of(d.tab, d.slot_task[slot] / d.S);
    GpCurveView cv = curve_view(d.tab, cu);
    double l, g[FG_MAXP];
    bool ok;
    eval_curve<Kern, kPlace>(c, cv, d.slots[slot].x, smem, d.gA, d.gA_stride, pipe_it, l, g, ok, d.noscratch != 0);
    if (threadIdx.x == 0) {
      for (int k = 0; k < Kern::P; ++k) g[k] = -g[k];     // obj_func returns (-lml, -grad)  (_gpr.py:302-309)
      for (int k = Kern::P; k < FG_MAXP; ++k) g[k] = 0.0;
      *go = advance_slot<Kern::P>(d, slot, -l, g) ? 1 : 0;
      __threadfence_block();
    }
    __syncthreads();
    if (!*go) break;   // (thread 0 next writes `go` after the barriers of a whole evaluation: no second barrier needed here)
  }
}

// One round of the fit: every live slot's objective is evaluated at its current point.  Rounds are enqueued ahead of time, without
// the host looking at the device (fulu_gp_fit_batch is stream-ordered), so a round may find nothing to do: no live slot, or the fit
// already ended by fit_tail_kernel (d.ended).
template <class Kern, int kPlace>
__global__ void __launch_bounds__(kEvalMaxThreads, 1) fit_eval_kernel(FitDev d, int round) {
  extern __shared__ __align__(16) double smem[];
  DevCtxP<kPlace> c;
  const int count = *d.ended ? 0 : d.counts[round];
  if ((int)blockIdx.x >= count) return;
  const int32_t* list = d.list + (size_t)(round & 1) * d.window;
  uint32_t pipe_it = 0;
  eval_init<kPlace>(c, smem);
  for (int i = blockIdx.x; i < count; i += gridDim.x) {
    const int slot = list[i];
    const int64_t cu = curve_of(d.tab, d.slot_task[slot] / d.S);
    GpCurveView cv = curve_view(d.tab, cu);
    double l, g[Kern::P];
    bool ok;
    eval_curve<Kern, kPlace>(c, cv, d.slots[slot].x, smem, d.gA, d.gA_stride, pipe_it, l, g, ok, d.noscratch != 0);
    if (threadIdx.x == 0) {
      d.f_out[slot] = -l;     // obj_func returns (-lml, -grad)  (_gpr.py:302-309)
      for (int k = 0; k < Kern::P; ++k) d.g_out[slot * FG_MAXP + k] = -g[k];
    }
    __syncthreads();
  }
}

// The END of a fit: if `round`'s list holds at most `tail_below` live slots (and the fit has not ended yet), they are run to the end
// here - CTA b takes the slots list[b], list[b + grid], ..., alternating evaluation (all warps) and optimiser step (thread 0) until a
// slot runs dry - and the fit is marked ended, so that every later round finds nothing to do.  No under-filled rounds, no host round
// trips.  Same block geometry as fit_eval_kernel, so a run's numbers do not depend on which kernel finished it.  Launched
//   (i) as the whole fit when it is small (at most one wave of tasks) or its class streams (long curves: one slot per CTA fed by the
//       task queue),
//  (ii) every few rounds with tail_below = one wave of CTAs: the device-side switch from rounds to the fused end,
// (iii) behind the last enqueued round with tail_below = "any": the fit always ends on the device however many rounds it needed.
template <class Kern, int kPlace>
__global__ void __launch_bounds__(kEvalMaxThreads, 1) fit_tail_kernel(FitDev d, int round, int tail_below) {
  extern __shared__ __align__(16) double smem[];
  __shared__ int go;
  DevCtxP<kPlace> c;
  // the mark is this launch's own stamp (round + 1): CTAs of this launch that start after block 0 has set it must not mistake it for
  // an earlier launch's
  const int seen = *d.ended;
  const int count = (seen != 0 && seen != round + 1) ? 0 : d.counts[round];
  if (count > tail_below || (int)blockIdx.x >= count) return;
  if (blockIdx.x == 0 && threadIdx.x == 0) atomicExch(d.ended_mark, round + 1);   // (block 0 always takes part)
  uint32_t pipe_it = 0;
  eval_init<kPlace>(c, smem);
  for (int b = blockIdx.x; b < count; b += gridDim.x) {
    run_slot_to_end<Kern, kPlace>(c, d, d.list[(size_t)(round & 1) * d.window + b], smem, pipe_it, &go);
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------ the fused fit (round 2)
// OPT-IN (FULU_GP_FUSED=1): bit-identical to the rounds, measured 9-10 % slower on B200 (profiles/r02_fused_fit.md).
// ONE persistent CTA per SM runs a whole fit: no rounds, no optimiser-step kernel, no host in the loop.
//   G evaluation groups   what the rounds path runs as G co-resident CTAs: the same threads per group, the same block code (a named
//                         barrier per group stands in for __syncthreads), the same shared-memory layout per group - a run's numbers are
//                         bit-identical to the rounds path;
//   1-4 stepper warps     each lane owns one optimiser slot of this SM: it advances the L-BFGS-B state machine of its slot whenever an
//                         evaluation of that slot has come back, concurrently with the groups evaluating other slots (an optimiser
//                         step is a 40-75 us serial chain: as a kernel of its own it costs 3-5 % of the fit and a grid-wide barrier
//                         per round; here it was meant to run in issue slots the evaluation leaves idle - what it takes instead is
//                         L1 and instruction cache, and the evaluations around it slow down by more than that).  A finished run
//                         pulls the next (curve, start) task from the global queue into its slot.
// Slot hand-over goes through shared memory: the stepper publishes theta and the curve id and marks the slot READY; a group's leader
// claims the READY slot whose run has gone on longest (long runs are what a fit's tail is made of), the group evaluates, publishes
// (-LML, -gradient) and marks the slot DONE; the stepper picks it up.  No CTA ever waits for another CTA, so the kernel needs no
// co-residency guarantee; inside a CTA every wait is on a thread of the same CTA.  A wait that lasts ~10 s traps instead of hanging.
constexpr int kFusedMaxSlots = 32;
constexpr int kFusedMaxGroups = 16;
constexpr int kFusedDefaultSlots = 32;
constexpr int kFusedDefaultSteppers = 1;
constexpr int kFusedMaxSteppers = 4;   // (warps are allocated four at a time: up to four stepper warps cost no registers)
constexpr int kFusedThreads = kEvalMaxThreads + 32 * kFusedMaxSteppers;

enum { FS_IDLE = 0, FS_READY = 1, FS_BUSY = 2, FS_DONE = 3 };

struct FusedShared {
  double theta[kFusedMaxSlots][FG_MAXP];
  double g[kFusedMaxSlots][FG_MAXP];
  double f[kFusedMaxSlots];
  long long curve[kFusedMaxSlots];
  int state[kFusedMaxSlots];
  int age[kFusedMaxSlots];     // evaluations of the slot's current run so far: the claim priority
  int pick[kFusedMaxGroups];
  int live;                    // slots that still have something to do
};

struct FusedGeom {
  int groups;         // evaluation groups per CTA (= CTAs per SM of the rounds path)
  int group_threads;  // threads per group
  int slots;          // optimiser slots per CTA (<= 32)
  int steppers;       // stepper warps (1 .. 4); each owns slots / steppers slots, one per lane
  int rotate;         // 1: the groups' warp 0 (the serial diagonal chain) on different schedulers
  size_t group_smem;  // doubles of dynamic shared memory per group
};

// block context of one evaluation group: threads [base, base + nthr) of the CTA, barrier `bar` (1 .. 15; warp-sized groups need none)
template <int kPlace>
struct GroupCtx : DevCtxP<kPlace> {
  using DevCtx::tid; using DevCtx::nthr; using DevCtx::lane; using DevCtx::warp; using DevCtx::nwarp;
  int bar;
  __device__ __forceinline__ GroupCtx(int group, int group_threads, int rotate) : DevCtxP<kPlace>() {
    const int hw = (int)threadIdx.x - group * group_threads;   // thread index inside the group
    nthr = group_threads;
    nwarp = group_threads >> 5;
    // which hardware warp of the group plays warp 0: hardware warp h sits on scheduler h % 4
    int j0 = 0;
    if (rotate) {
      if (nwarp >= 4) j0 = (((group * (1 - nwarp)) % 4) + 4) % 4;
      else if (nwarp == 2) j0 = (group >> 1) & 1;
    }
    // (through a shuffle, like DevCtx: the compiler then knows the warp index is warp-uniform and keeps the branches on it convergent)
    const int hwarp = hw >> 5;
    warp = __shfl_sync(0xffffffffu, hwarp - j0 < 0 ? hwarp - j0 + nwarp : hwarp - j0, 0);
    tid = warp * 32 + lane;
    bar = 1 + group;
  }
  __device__ __forceinline__ void sync() {
    if (nthr != 32) asm volatile("bar.sync %0, %1;" ::"r"(bar), "r"(nthr) : "memory");
    __syncwarp();   // (also tells the compiler that the warp is converged here, as __syncthreads does: plain SHFL / DMMA follow)
  }
};

__device__ __forceinline__ void fused_watchdog(unsigned& spins, long long& t0) {
  if (++spins > 4096u) {
    if (t0 == 0) t0 = clock64();
    else if (clock64() - t0 > 20000000000LL) __trap();
  }
  __nanosleep(64);
}

// the stepper hands slot `slot` (lane `lane` of this CTA) to the groups: its next evaluation point, its curve, its age
__device__ __forceinline__ void fused_publish(FusedShared& sh, const FitDev& d, int slot, int lane) {
  const FgState& st = d.slots[slot];
  for (int k = 0; k < FG_MAXP; ++k) sh.theta[lane][k] = st.x[k];
  sh.curve[lane] = (long long)curve_of(d.tab, d.slot_task[slot] / d.S);
  sh.age[lane] = d.slot_nev[slot];
  __threadfence_block();
  *(volatile int*)&sh.state[lane] = FS_READY;
}

template <class Kern, int kPlace>
__global__ void __launch_bounds__(kFusedThreads, 1) fit_fused_kernel(FitDev d, FusedGeom geo) {
  extern __shared__ __align__(16) double smem[];
  __shared__ FusedShared sh;
  const int n_eval_threads = geo.groups * geo.group_threads;
  // (roles and loop exits go through shuffles so that the compiler knows they are warp-uniform: the block code's warp-level
  // exchanges then compile to plain SHFL / DMMA, as in the rounds kernels)
  const bool stepper = __shfl_sync(0xffffffffu, (int)threadIdx.x >= n_eval_threads ? 1 : 0, 0) != 0;
  const int lane = (int)threadIdx.x & 31;
  if (threadIdx.x == 0) sh.live = 0;
  if ((int)threadIdx.x < kFusedMaxSlots) { sh.state[threadIdx.x] = FS_IDLE; sh.age[threadIdx.x] = 0; }
  __syncthreads();
  // ---------------------------------------------------------------- the stepper warps: warp s, lane i <-> slot s * spw + i of this CTA
  const int spw = (geo.slots + geo.steppers - 1) / geo.steppers;   // slots per stepper warp
  const int sw = stepper ? __shfl_sync(0xffffffffu, ((int)threadIdx.x - n_eval_threads) >> 5, 0) : 0;
  const int ls = sw * spw + lane;                                   // this lane's slot (local index)
  const bool has_slot = stepper && lane < spw && ls < geo.slots;
  if (stepper) {
    const int slot = (int)blockIdx.x * kFusedMaxSlots + ls;
    if (has_slot && acquire_task(d, slot, d.slots[slot])) {
      atomicAdd(&sh.live, 1);
      fused_publish(sh, d, slot, ls);
    }
  }
  __syncthreads();   // (the last block-wide barrier: from here on the groups and the stepper meet through shared memory only)
  if (stepper) {
    const int slot = (int)blockIdx.x * kFusedMaxSlots + ls;
    unsigned spins = 0;
    long long t0 = 0;
#ifdef GP_FUSED_STATS
    long long st_begin = clock64(), st_busy = 0, st_pass = 0, st_lanes = 0;
#endif
    for (;;) {
      __syncwarp();
      const bool mine = has_slot && *(volatile int*)&sh.state[ls] == FS_DONE;
#ifdef GP_FUSED_STATS
      const long long tp = clock64();
#endif
      if (mine) {
        __threadfence_block();
        double g[FG_MAXP];
        for (int k = 0; k < FG_MAXP; ++k) g[k] = sh.g[ls][k];
        if (advance_slot<Kern::P>(d, slot, sh.f[ls], g)) fused_publish(sh, d, slot, ls);
        else {
          *(volatile int*)&sh.state[ls] = FS_IDLE;
          __threadfence_block();
          atomicSub(&sh.live, 1);
        }
      }
      __syncwarp();
#ifdef GP_FUSED_STATS
      { const unsigned bal = __ballot_sync(0xffffffffu, mine); if (bal) { st_busy += clock64() - tp; ++st_pass; st_lanes += __popc(bal); } }
#endif
      if (__any_sync(0xffffffffu, mine)) { spins = 0; t0 = 0; continue; }
      const int live = __shfl_sync(0xffffffffu, *(volatile int*)&sh.live, 0);
      if (live == 0) break;
      fused_watchdog(spins, t0);
    }
#ifdef GP_FUSED_STATS
    if (lane == 0 && blockIdx.x < 2)
      printf("cta %d stepper %d: total %lld kcyc, stepping %lld kcyc (%.1f %%), passes %lld, lanes/pass %.2f, kcyc/pass %.1f\n", (int)blockIdx.x, sw,
             (clock64() - st_begin) / 1000, st_busy / 1000, 100.0 * st_busy / (double)(clock64() - st_begin), st_pass,
             (double)st_lanes / (double)(st_pass ? st_pass : 1), (double)st_busy / 1000.0 / (double)(st_pass ? st_pass : 1));
#endif
    return;
  }
  // ------------------------------------------------------------------ an evaluation group
  const int group = __shfl_sync(0xffffffffu, (int)threadIdx.x / geo.group_threads, 0);
  GroupCtx<kPlace> c(group, geo.group_threads, geo.rotate);
  double* gsm = smem + (size_t)group * geo.group_smem;
  double* slab = (d.gA != nullptr && d.gA_stride > 0) ? d.gA + ((size_t)blockIdx.x * geo.groups + group) * d.gA_stride : nullptr;
#ifdef GP_FUSED_STATS
  long long gs_begin = clock64(), gs_wait = 0, gs_n = 0;
#endif
  for (;;) {
    if (c.tid == 0) {
      int pick = -1;
      unsigned spins = 0;
      long long t0 = 0;
#ifdef GP_FUSED_STATS
      const long long tw = clock64();
#endif
      for (;;) {
        int best = -1, best_age = -1;
        for (int s = 0; s < geo.slots; ++s) {
          if (*(volatile int*)&sh.state[s] == FS_READY) {
            const int a = *(volatile int*)&sh.age[s];
            if (a > best_age) { best_age = a; best = s; }
          }
        }
        if (best >= 0) {
          if (atomicCAS(&sh.state[best], FS_READY, FS_BUSY) == FS_READY) { pick = best; break; }
          continue;   // another group was faster: look again
        }
        if (*(volatile int*)&sh.live == 0) break;
        fused_watchdog(spins, t0);
      }
      *(volatile int*)&sh.pick[group] = pick;
      __threadfence_block();
#ifdef GP_FUSED_STATS
      gs_wait += clock64() - tw; ++gs_n;
#endif
    }
    c.sync();
    const int pick = __shfl_sync(0xffffffffu, *(volatile int*)&sh.pick[group], 0);
    if (pick < 0) break;
    __threadfence_block();
    GpCurveView cv = curve_view(d.tab, (int64_t)sh.curve[pick]);
    double theta[FG_MAXP];
    for (int k = 0; k < FG_MAXP; ++k) theta[k] = sh.theta[pick][k];
    double l, g[FG_MAXP];
    bool ok;
    {
      GpSmem s = (kPlace != GP_PLACE_SMEM) ? gp_carve(slab, gsm, cv.n)
                 : d.noscratch   ? gp_carve(gsm, gsm + gp_mat_doubles_ns(gp_pad(cv.n)), cv.n, false)
                                 : gp_carve(gsm, gsm + gp_mat_doubles(gp_pad(cv.n)), cv.n);
      if (kPlace == GP_PLACE_SMEM && slab != nullptr) s.kslab = slab;
      gp_eval_lml_grad<Kern>(c, cv, theta, s, l, g, ok);
    }
    if (c.tid == 0) {
      sh.f[pick] = -l;     // obj_func returns (-lml, -grad)  (_gpr.py:302-309)
      for (int k = 0; k < Kern::P; ++k) sh.g[pick][k] = -g[k];
      for (int k = Kern::P; k < FG_MAXP; ++k) sh.g[pick][k] = 0.0;
      __threadfence_block();
      *(volatile int*)&sh.state[pick] = FS_DONE;
    }
    c.sync();   // (the group's shared memory is free for the next evaluation; pick[group] may be rewritten)
  }
#ifdef GP_FUSED_STATS
  if (c.tid == 0 && blockIdx.x < 2)
    printf("cta %d group %d: total %lld kcyc, waiting for a slot %lld kcyc (%.1f %%), evaluations %lld, kcyc/evaluation %.1f\n", (int)blockIdx.x, group,
           (clock64() - gs_begin) / 1000, gs_wait / 1000, 100.0 * gs_wait / (double)(clock64() - gs_begin), gs_n - 1,
           (double)(clock64() - gs_begin - gs_wait) / 1000.0 / (double)(gs_n > 1 ? gs_n - 1 : 1));
#endif
}

struct PredictArgs {
  CurveTable tab;
  int64_t B;
  int P;
  const double* theta;
  int n_obs;                  // grid mode if > 0
  const double *t_min, *t_max;
  const double* t_raw;        // raw times (for default t_min/t_max)
  const int64_t* q_offsets;   // explicit mode
  const double* q_t;
  const int32_t* q_band;
  double *mean, *stdv, *lml;
  int32_t* status;
  double *gA, *gV;
  size_t gA_stride, gV_stride;
};

template <class Kern, int kPlace>
__global__ void __launch_bounds__(kPredictThreads, 2) predict_kernel(PredictArgs a) {
  extern __shared__ __align__(16) double smem[];
  __shared__ double tlim[2];
  DevCtxP<kPlace> c;
  // the warps' V tiles: behind the matrix and vectors in shared memory, or this CTA's global slab for long curves
  double* Vglobal = a.gV + (size_t)blockIdx.x * a.gV_stride;
  uint32_t pipe_it = 0;
  eval_init<kPlace>(c, smem);
  for (int64_t kk = blockIdx.x; kk < a.B; kk += gridDim.x) {
    const int64_t cu = curve_of(a.tab, kk);
    GpCurveView cv = curve_view(a.tab, cu);
    GpQuery qd;
    int64_t out0;
    if (a.n_obs > 0) {
      qd.n_obs = a.n_obs; qd.m = a.n_obs * a.tab.n_bands; qd.qt = nullptr; qd.qband = nullptr;
      out0 = cu * (int64_t)qd.m;
      if (a.t_min && a.t_max) { qd.t_min = a.t_min[cu]; qd.t_max = a.t_max[cu]; }
      else {
        // default: the curve's own time span
        if (threadIdx.x == 0) {
          const double* tr = a.t_raw + a.tab.offsets[cu];
          double lo = 0.0, hi = 0.0;
          if (cv.n > 0 && a.tab.pstatus[cu] == 0) {   // (an empty or unusable curve has no span: its grid is not evaluated)
            lo = hi = tr[0];
            for (int i = 1; i < cv.n; ++i) { lo = fmin(lo, tr[i]); hi = fmax(hi, tr[i]); }
          }
          tlim[0] = lo; tlim[1] = hi;
        }
        __syncthreads();
        qd.t_min = tlim[0]; qd.t_max = tlim[1];
      }
    } else {
      out0 = a.q_offsets[cu];
      qd.n_obs = 0; qd.m = (int)(a.q_offsets[cu + 1] - out0); qd.qt = a.q_t + out0; qd.qband = a.q_band + out0;
      qd.t_min = qd.t_max = 0.0;
    }
    int st = a.tab.pstatus[cu] != 0 ? FULU_GP_STATUS_BAD_INPUT : 0;
    if (!st && a.n_obs <= 0) {   // explicit points: a query band outside the table is the caller's KeyError - flagged, not indexed
      int bad = 0;
      for (int q = threadIdx.x; q < qd.m; q += blockDim.x) bad |= (qd.qband[q] < 0 || qd.qband[q] >= a.tab.n_bands);
      if (__syncthreads_or(bad)) st = FULU_GP_STATUS_BAD_INPUT;
    }
    double l = NAN;
    if (st) {
      for (int q = threadIdx.x; q < qd.m; q += blockDim.x) { a.mean[out0 + q] = NAN; a.stdv[out0 + q] = NAN; }
    } else if constexpr (kPlace == GP_PLACE_TILED) {
      GpTiled tl = gp_tiled_carve<TiledCfg>(a.gA + (size_t)blockIdx.x * a.gA_stride, smem, cv.n, c.nwarp, pipe_it);
      st = gp_tiled_predict_curve<Kern, TiledCfg>(c, cv, a.theta + cu * a.P, a.tab.stats + cu * 6, qd, tl, Vglobal, a.mean + out0, a.stdv + out0, l);
      pipe_it = tl.it;
    } else {
      GpSmem s = carve<kPlace != GP_PLACE_SMEM>(smem, a.gA, a.gA_stride, cv.n);
      double* V = a.gV_stride ? Vglobal : smem + gp_smem_doubles(cv.n);   // (classes near the shared-memory limit keep V in the slab too)
      st = gp_predict_curve<Kern>(c, cv, a.theta + cu * a.P, a.tab.stats + cu * 6, qd, s, V, a.mean + out0, a.stdv + out0, l);
    }
    if (threadIdx.x == 0) {
      if (a.status) a.status[cu] = st;
      if (a.lml) a.lml[cu] = l;
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------ per-family launchers
// What gp_kernels.cu needs from a family: its parameter count and three launch functions (return a cudaError_t as int).
struct GpFamilyOps {
  int n_params;
  int (*lml)(int place, unsigned grid, int threads, size_t smem, cudaStream_t st, const CurveTable& tab, int64_t B, int P, const double* theta,
             double* lml, double* grad, double* gA, size_t gA_stride, long long* marks, int noscratch);
  int (*eval_prepare)(int place, size_t smem);
  int (*eval)(int place, unsigned grid, int threads, size_t smem, cudaStream_t st, const FitDev& d, int round);
  int (*tail)(int place, unsigned grid, int threads, size_t smem, cudaStream_t st, const FitDev& d, int round, int tail_below);
  int (*fused)(int place, unsigned grid, size_t smem, cudaStream_t st, const FitDev& d, const FusedGeom& geo);   // shared-memory and slab placements
  int (*predict)(int place, unsigned grid, size_t smem, cudaStream_t st, const PredictArgs& a);
};

template <class K>
inline int gp_set_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024) return (int)cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  return 0;
}

// runs `body(std::integral_constant<int, place>)` for the runtime placement
#define GP_PLACE_SWITCH(place, ...)                                               \
  switch (place) {                                                                \
    case GP_PLACE_SMEM: { constexpr int kP = GP_PLACE_SMEM; __VA_ARGS__; } break;         \
    case GP_PLACE_SLAB: { constexpr int kP = GP_PLACE_SLAB; __VA_ARGS__; } break;         \
    default: { constexpr int kP = GP_PLACE_TILED; __VA_ARGS__; } break;                   \
  }

template <class Kern>
struct GpFamilyLaunch {
  static int lml(int place, unsigned grid, int threads, size_t smem, cudaStream_t st, const CurveTable& tab, int64_t B, int P,
                 const double* theta, double* lml, double* grad, double* gA, size_t gA_stride, long long* marks, int noscratch) {
    int rc = 0;
    GP_PLACE_SWITCH(place, {
      if ((rc = gp_set_smem(lml_kernel<Kern, kP>, smem))) return rc;
      lml_kernel<Kern, kP><<<grid, threads, smem, st>>>(tab, B, P, theta, lml, grad, gA, gA_stride, marks, noscratch);
    })
    return (int)cudaGetLastError();
  }
  static int eval_prepare(int place, size_t smem) {
    int rc = 0;
    GP_PLACE_SWITCH(place, {
      if ((rc = gp_set_smem(fit_eval_kernel<Kern, kP>, smem))) return rc;
      rc = gp_set_smem(fit_tail_kernel<Kern, kP>, smem);
    })
    return rc;
  }
  static int eval(int place, unsigned grid, int threads, size_t smem, cudaStream_t st, const FitDev& d, int round) {
    GP_PLACE_SWITCH(place, { fit_eval_kernel<Kern, kP><<<grid, threads, smem, st>>>(d, round); })
    return 0;
  }
  static int tail(int place, unsigned grid, int threads, size_t smem, cudaStream_t st, const FitDev& d, int round, int tail_below) {
    GP_PLACE_SWITCH(place, { fit_tail_kernel<Kern, kP><<<grid, threads, smem, st>>>(d, round, tail_below); })
    return 0;
  }
  static int fused(int place, unsigned grid, size_t smem, cudaStream_t st, const FitDev& d, const FusedGeom& geo) {
    const int threads = geo.groups * geo.group_threads + 32 * geo.steppers;
    int rc = 0;
    if (place == GP_PLACE_SMEM) {
      if ((rc = gp_set_smem(fit_fused_kernel<Kern, GP_PLACE_SMEM>, smem))) return rc;
      fit_fused_kernel<Kern, GP_PLACE_SMEM><<<grid, threads, smem, st>>>(d, geo);
    } else if (place == GP_PLACE_SLAB) {
      if ((rc = gp_set_smem(fit_fused_kernel<Kern, GP_PLACE_SLAB>, smem))) return rc;
      fit_fused_kernel<Kern, GP_PLACE_SLAB><<<grid, threads, smem, st>>>(d, geo);
    } else {
      return FULU_GP_E_UNSUPPORTED;
    }
    return (int)cudaGetLastError();
  }
  static int predict(int place, unsigned grid, size_t smem, cudaStream_t st, const PredictArgs& a) {
    int rc = 0;
    GP_PLACE_SWITCH(place, {
      if ((rc = gp_set_smem(predict_kernel<Kern, kP>, smem))) return rc;
      predict_kernel<Kern, kP><<<grid, kPredictThreads, smem, st>>>(a);
    })
    return (int)cudaGetLastError();
  }
  static const GpFamilyOps* ops() {
    static const GpFamilyOps o = {Kern::P, &lml, &eval_prepare, &eval, &tail, &fused, &predict};
    return &o;
  }
};

}  // namespace

// the compiled families: one translation unit each (gp_family.cu with -DGP_FAM_F1= -DGP_FAM_EX=)
#define GP_FAMILY_LIST(X) X(0, 0) X(0, 2) X(2, 2) X(0, 4) X(0, 1) X(0, 3) X(2, 0)
#define GP_FAMILY_SYMBOL_(f1, ex) gp_family_ops_##f1##_##ex
#define GP_FAMILY_SYMBOL(f1, ex) GP_FAMILY_SYMBOL_(f1, ex)
#define GP_FAMILY_DECL(f1, ex) const GpFamilyOps* GP_FAMILY_SYMBOL(f1, ex)();
GP_FAMILY_LIST(GP_FAMILY_DECL)
