/* fulu_gp.h — C ABI of the B200-native Gaussian-process light-curve augmentation path.
 *
 * Drop-in boundary for the hot path of hse-cs/fulu:
 *   fulu.GaussianProcessesAugmentation.fit / .predict / .augmentation
 *   (reference: fulu/gp_aug.py:37-108, fulu/_base_aug.py:9-20,88-114), whose arithmetic is
 *   sklearn.gaussian_process.GaussianProcessRegressor (sklearn/gaussian_process/_gpr.py:232-368,
 *   444-500, 541-656) driven by SciPy L-BFGS-B.  The reference has no FFI of its own (pure Python);
 *   these entry points are what a ctypes binding inside fulu/gp_aug.py would call instead of
 *   `self.reg.fit(...)` / `self.reg.predict(...)` — see INTEGRATION.md.
 *
 * Conventions
 *   - every pointer in fulu_gp_batch and every in/out array is a DEVICE pointer owned by the caller
 *     (PyTorch tensors); the library allocates nothing and keeps no global state;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*) and never looks at the device: fulu_gp_fit_batch too
 *     enqueues its whole optimisation loop and returns (the loop's length is decided on the device, see there); no call creates a
 *     stream or an event, except fulu_gp_fit_batch's opt-in per-launch timing (fit options `profile`), which also waits for them;
 *   - return value: 0 ok, <0 argument error (FULU_GP_E_*), >0 a cudaError_t;
 *   - per-curve problems never fail the call: they are reported in status[] (FULU_GP_STATUS_* bits).
 *   - batches are ragged CSR: curve i owns observations offsets[i] .. offsets[i+1]-1.
 *   - a few environment variables exist for A/B measurements only (read per call, defaults are the fast paths): FULU_GP_NO_TAIL=1 keeps
 *     the fit on rounds of two launches to the end, FULU_GP_NO_KSLAB=1 recomputes K in the gradient pass instead of reading it back
 *     from the L2 slab, FULU_GP_FUSED=1 runs a fit as ONE persistent launch (evaluation groups + optimiser-stepping warps per SM,
 *     with FULU_GP_FUSED_SLOTS / _STEPPERS / _ROTATE as its geometry; same bits, measured slower: DESIGN.md section 6),
 *     FULU_GP_TILED_ABOVE=<np> moves the slab / tiled boundary, FULU_GP_NO_SCRATCH=0 keeps the N = 104..127 / 136..159 classes on the
 *     layout with a scratch column (one CTA per SM less); none changes which curves are processed or the meaning of any output;
 *   - results are bitwise reproducible for a given curve and launch geometry; the geometry (shared memory, CTAs per SM, threads
 *     per CTA) is a function of n_max only.
 */
#ifndef FULU_GP_H
#define FULU_GP_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FULU_GP_ABI_VERSION 5
#define FULU_GP_MAX_PARAMS 6
#define FULU_GP_MAX_BANDS 16

/* Covariance families:  K = c * [M1] * RBF([lt, ll]) + [EX] + White, a kernel code is FULU_GP_KERNEL(f1, ex) with
 *   f1 (an isotropic Matern factor multiplying the RBF) and ex (an added isotropic term) one of FULU_GP_F_*.
 * theta is log-transformed, in CANONICAL order [c, (l_1), lt, ll, (ex: l | alpha, l), s2]; for kernels written as
 * C*[Matern*]RBF + [Matern|RationalQuadratic] + White this is sklearn's own order (kernels.py:739-752, hyperparameters of one
 * kernel sorted by name :280-288); other orderings of the same sum/product are permuted by the host binding.
 * Families compiled into this build: (0,0) (0,MAT15) (MAT15,MAT15) (0,RQ) (0,MAT05) (0,MAT25) (MAT15,0); others -> E_UNSUPPORTED. */
#define FULU_GP_F_NONE 0
#define FULU_GP_F_MAT05 1                 /* Matern nu = 1/2  (kernels.py:1722-1723) */
#define FULU_GP_F_MAT15 2                 /* Matern nu = 3/2  (kernels.py:1724-1726; sklearn's default nu) */
#define FULU_GP_F_MAT25 3                 /* Matern nu = 5/2  (kernels.py:1727-1729) */
#define FULU_GP_F_RQ 4                    /* RationalQuadratic (kernels.py:1917-1947); ex only */
#define FULU_GP_KERNEL(f1, ex) ((f1) | ((ex) << 4))
#define FULU_GP_KERNEL_RBF_WHITE FULU_GP_KERNEL(0, 0)                 /* C*RBF([lt,ll]) + White, P = 4  (fulu/gp_aug.py:29 default) */
#define FULU_GP_KERNEL_RBF_MATERN_WHITE FULU_GP_KERNEL(0, 2)          /* C*RBF + Matern() + White, P = 5  (fulu/gp_aug.py:24, tests/test_aug.py:37) */
#define FULU_GP_KERNEL_MATERN_RBF_MATERN_WHITE FULU_GP_KERNEL(2, 2)   /* C*Matern()*RBF + Matern() + White, P = 6  (notebook_examples/plotting.ipynb:33) */
#define FULU_GP_KERNEL_RBF_RQ_WHITE FULU_GP_KERNEL(0, 4)              /* C*RBF + RationalQuadratic() + White, P = 6  (fulu/gp_aug.py:5) */

/* argument errors */
#define FULU_GP_E_BADARG (-1)
#define FULU_GP_E_WORKSPACE (-2)     /* workspace too small */
#define FULU_GP_E_UNSUPPORTED (-3)   /* kernel family / size not supported by this build */
#define FULU_GP_E_NOCUDA (-4)

/* per-curve status bits */
#define FULU_GP_STATUS_OK 0
#define FULU_GP_STATUS_BAD_INPUT 1       /* NaN/inf in t/flux/flux_err, band out of range (sklearn: ValueError, _gpr.py:257-265), or longer than n_max */
#define FULU_GP_STATUS_NOT_PD 2          /* K(theta*) not positive definite (sklearn: LinAlgError, _gpr.py:349-361) */
#define FULU_GP_STATUS_NOT_CONVERGED 4   /* best run ended ABNORMAL / at the iteration limit (ConvergenceWarning, utils/optimize.py:436-458) */
#define FULU_GP_STATUS_AT_BOUND 8        /* theta* numerically on a bound (ConvergenceWarning, kernels.py:436-464) */
#define FULU_GP_STATUS_NEG_VARIANCE 16   /* a predictive variance was clamped to 0 (_gpr.py:485-491) */

typedef struct fulu_gp_batch {
  int64_t n_curves;          /* B */
  int64_t n_obs_total;       /* offsets[B] */
  int32_t n_max;             /* upper bound on the length of the curves this call processes (host-known); it fixes the launch geometry.
                                A listed curve longer than n_max is flagged FULU_GP_STATUS_BAD_INPUT and skipped, never run */
  int32_t n_bands;           /* entries of loglam */
  const int64_t* offsets;    /* [B+1] */
  const double* t;           /* [n_obs_total] observation times        (fulu fit arg `t`) */
  const double* flux;        /* [n_obs_total]                          (`flux`) */
  const double* flux_err;    /* [n_obs_total]                          (`flux_err`) */
  const int32_t* band;       /* [n_obs_total] index into loglam        (`passband` through passband2lam, _base_aug.py:9-11) */
  const double* loglam;      /* [n_bands] log10 wavelength table        (passband2lam values in dict order) */
  int32_t kernel;            /* FULU_GP_KERNEL_* */
  int32_t use_err;           /* fulu use_err flag (gp_aug.py:72-73) */
  double alpha0;             /* regressor alpha when !use_err: 1e-10 (_gpr.py:215) */
  const int32_t* order;      /* optional [n_order]: the curves (indices into the CSR arrays) this call processes, in processing
                                order; NULL = all n_curves.  Lets a caller run one launch per length class (n_max = the class
                                bound) over a batch that stays where it is in HBM - no repacking.  Inputs and outputs are always
                                indexed by curve id; rows of curves that are not listed are left untouched. */
  int64_t n_order;
} fulu_gp_batch;

typedef struct fulu_gp_fit_options {
  int32_t n_params;          /* P of the batch's kernel family (4 for the default kernel, up to FULU_GP_MAX_PARAMS) */
  int32_t n_starts;          /* S = 1 + n_restarts_optimizer = 6 */
  int32_t m;                 /* L-BFGS history, 10 */
  int32_t maxiter;           /* 15000 */
  int32_t maxfun;            /* 15000 */
  int32_t maxls;             /* 20 */
  int32_t window;            /* optimiser instances resident at once (0 = default: one per task up to 65,536; for curves whose matrix
                                streams from global memory with one CTA per SM, n_max >= 304, one per CTA - the fit then runs in one fused launch
                                whose CTAs pull the next run as soon as theirs ends; results do not depend on the window) */
  int32_t max_rounds;        /* 0 (default): the fit always runs to the end on the device.  > 0: hard cap on evaluation rounds - runs still alive
                                after that many rounds are recorded where they stand, flagged FULU_GP_STATUS_NOT_CONVERGED */
  int32_t starts_per_curve;  /* 0: theta_starts is [S,P], shared by all curves (sklearn's restarts); 1: [B,S,P], indexed by curve id
                                (e.g. S = 1 continuation runs from each curve's current optimum) */
  int32_t profile_counts_only; /* with `profile`: fill only the counts ([0], [3]..[9]) - no per-launch events, same launches as an
                                unprofiled call (per-launch timing makes every round two separately timed kernels) */
  double factr;              /* 1e7  (ftol = factr * eps) */
  double pgtol;              /* 1e-5 */
  double lower[FULU_GP_MAX_PARAMS];   /* log-bounds, ln(1e-5) */
  double upper[FULU_GP_MAX_PARAMS];   /* ln(1e5) */
  double* profile;           /* optional HOST pointer to 12 doubles, filled on return: [0] evaluation rounds, [1] ms spent in the
                                evaluation kernels (CUDA events on `stream`; includes [10]), [2] ms in the optimiser-step kernel, [3] kernels
                                this call launched, [4] window, [5] evaluation grid (CTAs), [6] dynamic shared memory per CTA (bytes), [7] 1 if
                                K lives in global memory, [8] threads per evaluation CTA, [9] evaluation CTAs per SM, [10] ms of the fused
                                launch that finishes the fit (evaluations and optimiser steps of the last live slots), [11] placement of
                                the matrix (0 shared memory, 1 global slab with the 8x8 block code, 2 global slab streamed through
                                shared-memory stages).  NULL = no per-launch timing. */
} fulu_gp_fit_options;

int fulu_gp_abi_version(void);
const char* fulu_gp_strerror(int code);

/* Bytes of device workspace needed by any of the calls below for this batch shape. */
int fulu_gp_workspace_bytes(const fulu_gp_batch* batch, int32_t n_params, int32_t n_starts, int32_t window, size_t* bytes);

/* Log marginal likelihood and gradient at given hyperparameters, one theta per curve.
 * Replaces GaussianProcessRegressor.log_marginal_likelihood(theta, eval_gradient=True) (_gpr.py:541-656).
 * theta [B,P] in; lml [B], grad [B,P] out; non-PD K gives lml=-inf, grad=0 (_gpr.py:590-593).
 * scaler [B,4] = (t mean, t scale, loglam mean, loglam scale), ynorm [B,2] = (flux mean, flux std) out (may be NULL). */
int fulu_gp_lml_batch(const fulu_gp_batch* batch, int32_t n_params, const double* theta, double* lml, double* grad,
                      double* scaler, double* ynorm, int32_t* status, void* workspace, size_t workspace_bytes, void* stream);

/* Hyperparameter optimisation: S bound-constrained L-BFGS-B runs per curve from theta_starts [S,P], best LML wins.
 * Replaces GaussianProcessRegressor.fit's optimiser loop (_gpr.py:299-340,658-668).
 * theta_starts is read in place (see starts_per_curve); points outside the bounds are clipped (_lbfgsb_py.py:359).
 * Stream-ordered: rounds of two launches (evaluate every live run's objective; advance every live run's optimiser) are enqueued ahead
 * of time - ~160 for a batch whose runs all fit the window - and a round that finds nothing left to do costs two empty launches; the
 * round that finds at most one wave of live runs finishes them in place, and one fused launch behind the last round runs whatever
 * might still be alive to the end, so the outputs never depend on how many rounds were enqueued.  Small fits and the long-curve
 * classes are a single fused launch.
 * theta_opt [B,P], lml_opt [B], n_eval [B] (LML+gradient evaluations spent on the curve), status [B] out.
 * scaler [B,4] = (t mean, t scale, loglam mean, loglam scale) and ynorm [B,2] = (flux mean, flux std) out (may be NULL): the fitted
 * StandardScaler (fulu/gp_aug.py:65-66) and sklearn's y normalisation (_gpr.py:275-280), i.e. aug.ss and reg._y_train_mean/_std.
 * run_lml [B,S] (final LML of every run), run_theta [B,S,P] and run_n_eval [B,S] (evaluations of every run) may be NULL.
 * A curve whose covariance was never positive definite along its best run gets lml_opt = -inf and FULU_GP_STATUS_NOT_PD
 * (sklearn: log_marginal_likelihood_value_ = -inf, then LinAlgError in fit, _gpr.py:349-361). */
int fulu_gp_fit_batch(const fulu_gp_batch* batch, const fulu_gp_fit_options* opt, const double* theta_starts, double* theta_opt,
                      double* lml_opt, double* scaler, double* ynorm, int32_t* n_eval, int32_t* status, double* run_lml,
                      double* run_theta, int32_t* run_n_eval, void* workspace, size_t workspace_bytes, void* stream);

/* Augmentation: predictive mean / std on the band-major grid linspace(t_min[i], t_max[i], n_obs) x n_bands.
 * Replaces BaseAugmentation.augmentation -> predict -> reg.predict(return_std=True)
 * (_base_aug.py:88-114, gp_aug.py:99-108, _gpr.py:444-500), including the final factorisation (_gpr.py:349-367).
 * theta [B,P] in; t_min/t_max [B] in (NULL = the curve's own min/max); mean/std [B, n_bands, n_obs] out. */
int fulu_gp_predict_grid_batch(const fulu_gp_batch* batch, int32_t n_params, const double* theta, const double* t_min,
                               const double* t_max, int32_t n_obs, double* mean, double* std, double* lml, int32_t* status,
                               void* workspace, size_t workspace_bytes, void* stream);

/* Prediction at arbitrary points: curve i is queried at q_t/q_band[q_offsets[i] .. q_offsets[i+1]-1]
 * (fulu predict(t, passband), gp_aug.py:80-108). mean/std [q_offsets[B]] out. */
int fulu_gp_predict_points_batch(const fulu_gp_batch* batch, int32_t n_params, const double* theta, const int64_t* q_offsets,
                                 const double* q_t, const int32_t* q_band, int32_t q_max, double* mean, double* std,
                                 int32_t* status, void* workspace, size_t workspace_bytes, void* stream);

/* Downstream consumer on the device (reference: LcPlotter.plot_sum_passbands / plot_peak, fulu/plotting.py:78-115):
 * the augmented curve summed over the passbands at each grid time (groupby("time").sum() of the band-major frame, :93,114) and the
 * grid time of its first maximum (curve["time"][curve["flux"].argmax()], :115).  A pipeline that only needs the peak leaves the
 * [B, n_bands, n_obs] curves on the device and reads back 20 bytes per curve.
 * mean [B, n_bands, n_obs] in (as written by fulu_gp_predict_grid_batch); t_min / t_max [B] in (both NULL = the curve's own span, taken
 * from batch->t); sum_flux [B, n_obs] out (may be NULL); peak_t [B], peak_flux [B], peak_index [B] out (NaN / -1 when every sum is
 * NaN).  Of `batch` only n_curves, n_bands, offsets, t and the optional index list are used. */
int fulu_gp_peak_batch(const fulu_gp_batch* batch, const double* mean, const double* t_min, const double* t_max, int32_t n_obs,
                       double* sum_flux, double* peak_t, double* peak_flux, int32_t* peak_index, void* stream);

/* Upstream ingest on the device: a survey table whose rows are grouped by object (PLAsTiCC / ZTF tables,
 * notebook_examples/object_34299.csv) -> the ragged CSR layout of fulu_gp_batch.  object_id [n_rows] in; offsets [max_objects + 1]
 * and n_objects (one int64 on the device) out: offsets[k] = first row of the k-th object, offsets[n_objects] = n_rows (objects beyond
 * max_objects are counted in n_objects but not written).  If band != NULL: band[i] = lut[passband[i]] (the reference's add_log_lam
 * dict lookup, fulu/_base_aug.py:9-11, as a table: lut[id] = index into loglam or -1), rows with an unknown passband get -1 and are
 * counted in *unknown_bands (the reference raises KeyError).  t / flux / flux_err are used where they lie: no column is copied. */
int fulu_gp_ingest_workspace_bytes(int64_t n_rows, size_t* bytes);
int fulu_gp_ingest_table(int64_t n_rows, const int64_t* object_id, const int32_t* passband, const int32_t* lut, int32_t lut_size,
                         int64_t max_objects, int64_t* offsets, int64_t* n_objects, int32_t* band, int32_t* unknown_bands,
                         void* workspace, size_t workspace_bytes, void* stream);

/* FP64 roofline probe: enqueues a register-resident DFMA (mode 0) or DMMA m8n8k4 (mode 1) loop on every SM of the current device
 * and returns the floating-point operations it performs in *flop (host).  The CALLER times it (events on `stream`) and owns the
 * scratch (device, >= 8 x 256 x 8 bytes per SM; the kernel's sink).  Used by bench.py for the roofline denominator because
 * MEASURED_PEAKS.json has no FP64 entry.  Like every other entry point it allocates nothing and does not synchronise. */
int fulu_gp_fp64_probe(int32_t mode, int32_t iters, double* scratch, size_t scratch_bytes, double* flop, void* stream);

/* Diagnostics: same as fulu_gp_lml_batch, and CTA 0 writes clock64() at nine phase boundaries of its first... every evaluation into
 * marks[0..8] (device pointer to >= 16 int64): load, assemble, cholesky, z, inverse, alpha, traces, reduce.  Used by profiles/. */
int fulu_gp_debug_phase_cycles(const fulu_gp_batch* batch, int32_t n_params, const double* theta, double* lml, double* grad,
                               long long* marks, void* workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FULU_GP_H */
