/*
 * epi_b200.h -- C ABI of libepi_b200.so: the B200 (sm_100a) batched engine for
 * the data-parallel hot path of alphanumericslab/EpidemicModeling.
 *
 * The reference is interpreted MATLAB with NO FFI layer; the boundary a
 * replacement must honour is its MATLAB function signatures.
 * Each entry point below names the reference function(s) it replaces
 * (file:line relative to the reference root).  The MEX gateway
 * (matlab/epi_mex.cpp), the Python mirror (epidemicmodeling_b200/api.py), the
 * tests and the benchmark all call exactly these symbols.
 *
 * Conventions
 *  - plain C: POD structs, pointers and sizes; no C++/torch types.
 *  - every entry returns 0 on success, <0 on error (EPI_ERR_*); the message is
 *    retrievable with epi_last_error().  No exception crosses the boundary.
 *  - the CALLER allocates every output; the library owns only its context
 *    (stream, scratch/tape buffers).  No pointer into library memory is returned.
 *  - `mem` selects where ALL arrays named in an args struct live:
 *      EPI_MEM_HOST   : host pointers; the call stages H2D, runs, copies D2H and
 *                       returns after the stream is synchronised (the MATLAB-style
 *                       blocking call; this is the `e2e` path of bench.py);
 *      EPI_MEM_DEVICE : device pointers (e.g. torch tensors); kernels are
 *                       enqueued on the context's stream and the call returns
 *                       without synchronising (call epi_sync()).
 *  - all floating point is IEEE binary64.  Batched arrays are
 *    "trajectory-minor" structure-of-arrays:  X[t][field][b]  (b fastest), so
 *    that a warp's accesses coalesce; for B == 1 these coincide with MATLAB's
 *    column-major m x T / m x m x T / L x T arrays.
 *  - trajectories are grouped: trajectory b belongs to group b / G.  Arrays
 *    marked "per group" are shared by the G trajectories of a group (a region's
 *    NPI history, observation series, model parameters ...).
 *  - there is NO CPU fallback: without a CUDA device epi_create() fails.
 */
#ifndef EPI_B200_H
#define EPI_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EPI_LMAX 12 /* NPIs per day (OxCGRT-style 12-NPI inputs) */

enum { EPI_MEM_HOST = 0, EPI_MEM_DEVICE = 1 };

enum {
  EPI_OK = 0,
  EPI_ERR_ARG = -1,        /* invalid argument */
  EPI_ERR_ORDER = -2,      /* 'Undefined order'  (GenericExtendedKalmanFilter.m:111,151) */
  EPI_ERR_OBS_TYPE = -3,   /* 'unknown observation type' (SIAlphaModelEKF.m:57,87) */
  EPI_ERR_QR_SHAPE = -4,   /* Q/R covariance shape mismatch (GenericExtendedKalmanFilter.m:75,90) */
  EPI_ERR_CUDA = -10,      /* CUDA runtime error (message has the detail) */
  EPI_ERR_NO_DEVICE = -11, /* no CUDA device: the library has no CPU fallback */
  EPI_ERR_NOMEM = -12
};

/* model variants = the reference wrappers that bind callbacks to the filter */
enum {
  EPI_MODEL_SIALPHA = 0,         /* Tools/SIAlphaModelEKF.m:1                       m = 3 */
  EPI_MODEL_SIALPHA_FLIPPED = 1, /* Tools/SIAlphaModelBackwardEKF.m:1               m = 3 */
  EPI_MODEL_OPTCTRL = 2,         /* Tools/SIAlphaModelEKFOptControlled.m:1          m = 6 */
  EPI_MODEL_OPTCTRL_FLIPPED = 3, /* Tools/SIAlphaModelBackwardEKFOptControlled.m:1  m = 6 */
  EPI_MODEL_LEGACY_TOOLS = 4,    /* Tools/NewCaseEKFEstimatorWithOptimalNPI.m:1     m = 6 */
  EPI_MODEL_LEGACY_CODEGEN = 5   /* MatlabCodeGenerator/NewCaseEKFEstimatorWithOptimalNPI.m:1
                                    + NlinStateUpdate.m, StateJacobians.m, ObsJacobian.m,
                                    NlinObsUpdate.m, *HessianTerms.m, *HardMargins.m   m = 6 */
};
enum { EPI_OBS_NEWCASES = 0, EPI_OBS_TOTALCASES = 1 };
enum { EPI_Q_CONST = 0, EPI_Q_PERDAY_SCALAR = 1, EPI_Q_PERDAY_FULL = 2 };
enum { EPI_R_CONST = 0, EPI_R_PERDAY = 1 };

/* The reference's `params` struct: fields of
 * MatlabCodeGenerator/NewCaseEKFEstimatorWithOptimalNPI.prj:1105-1116 plus
 * s_min, i_min, obs_type (Tools/TrainPredictPrescribeNPI.m:202-224). */
typedef struct {
  double dt, beta, gamma, b;
  double alpha_min, alpha_max, s_min, i_min;
  double epsilon, sigma;
  double a[EPI_LMAX], u_min[EPI_LMAX], u_max[EPI_LMAX], w[EPI_LMAX];
  int L;        /* number of NPIs, <= EPI_LMAX */
  int obs_type; /* EPI_OBS_* */
} epi_model_params;

typedef struct epi_ctx epi_ctx;

/* context ------------------------------------------------------------------ */
int epi_create(int device, epi_ctx **ctx);
void epi_destroy(epi_ctx *ctx);
const char *epi_last_error(const epi_ctx *ctx); /* ctx may be NULL: last create error */
/* cudaStream_t to enqueue on; NULL = the context's own (non-blocking) stream.  To use the
 * legacy default stream pass the cudaStreamLegacy handle ((void *)0x1). */
int epi_set_stream(epi_ctx *ctx, void *cuda_stream);
int epi_sync(epi_ctx *ctx);
/* cap on library-owned scratch (tape) bytes; larger batches are processed in
 * waves of trajectories.  0 = default (60 % of free device memory). */
int epi_set_scratch_limit(epi_ctx *ctx, size_t bytes);
/* The context caches the device blocks it allocated (staging, scratch) for reuse by later
 * calls; this hands them all back to the driver (synchronises the stream first). */
int epi_release_cache(epi_ctx *ctx);
/* kernel launches issued by this context since creation (bench `gpu_launches`) */
long long epi_launch_count(const epi_ctx *ctx);
/* device time (ms, CUDA events on the context's stream) of the kernels of the
 * most recent batched call, by phase; returns the number of phases written. */
int epi_last_kernel_times(epi_ctx *ctx, float *ms, const char **names, int max);
/* measured FP64 FMA throughput of the device (TFLOP/s): the FP64 roof that
 * bench.py's roofline uses (MEASURED_PEAKS.json carries no FP64 figure). */
int epi_fp64_probe(epi_ctx *ctx, int iters, double *tflops);

/* SEIRP ensembles ------------------------------------------------------------
 * replaces  [s,e,i,r,p] = SEIRP(alpha_e, alpha_i, kappa, rho, beta, mu, gamma,
 *                               s0, e0, i0, r0, p0, T, dt)      Tools/SEIRP.m:1
 *      and  SEIRPSaturatedResource(...)        Tools/SEIRPSaturatedResource.m:1 */
enum { EPI_RATES_CONST = 0, EPI_RATES_SHARED_SERIES = 1, EPI_RATES_SERIES = 2 };
enum { EPI_SEIRP_OUT_FULL = 0, EPI_SEIRP_OUT_FINAL = 1 };
typedef struct {
  int mem;
  int B, K;            /* trajectories; samples K = round(T/dt) (SEIRP.m:13) */
  double dt;
  int rate_mode;       /* CONST: rates[7][B]; SHARED_SERIES: rates[7][K] (the MATLAB
                          1xK vectors, shared by all B); SERIES: rates[K][7][B].
                          Row order: alpha_e, alpha_i, kappa, rho, beta, mu, gamma */
  const double *rates;
  const double *ic;    /* [5][B]: s0, e0, i0, r0, p0 */
  int saturated;       /* 1: SEIRPSaturatedResource (rows beta, mu of `rates` ignored) */
  double beta_0, beta_s, mu_0, mu_s, sigma, i_0;
  int out_mode;        /* FULL: out[5][K][B] ; FINAL: out[5][B] */
  double *out;
} epi_seirp_args;
int epi_seirp_batch(epi_ctx *ctx, const epi_seirp_args *a);

/* SI-alpha rollout (+ fused NPICost) -------------------------------------------
 * replaces  [s,i,alpha] = SIalpha_Controlled(u, s0, i0, alpha0, u_max, alpha_min,
 *              alpha_max, gamma, a, b, beta, s_noise_std, i_noise_std,
 *              alpha_noise_std, K, dt)               Tools/SIalpha_Controlled.m:1
 *      and  [J0,J1] = NPICost(newcases, inputs, weights)       Tools/NPICost.m:1
 *      as chained in Tools/TrainPredictPrescribeNPI.m:481-493,512-519.
 * The reference's in-line randn draws become the explicit `noise` input. */
enum { EPI_U_F64 = 0, EPI_U_U8 = 1, EPI_U_PHILOX = 3 };
/* EPI_U_PHILOX: the schedules are not supplied but generated in the kernel with the rule of
 * Tools/TrainPredictPrescribeNPI.m:499-510 -- trajectory b of a group is Monte-Carlo scenario
 * sc = (b mod G) + 1 of region b / G; scenarios with sc < G/2 draw ONE level per NPI and hold it
 * over time, the others draw one level per NPI per day; every draw is uniform on the integers
 * [prm.u_min(j), prm.u_max(j)] (MATLAB randi).  The reference's global Mersenne-Twister stream
 * becomes a counter-based one: word (j mod 4) of Philox4x32-10(counter = {day, j / 4, sc - 1,
 * region}, key = seed), day = 0 for the held scenarios, mapped to a level by
 * lo + floor(word * (hi - lo + 1) / 2^32).  Same schedule on any GPU count and wave split. */
typedef struct {
  int mem;
  int B, K, L, G;
  const epi_model_params *prm; /* per group: dt, beta, gamma, a, b, u_max, alpha_min, alpha_max */
  const double *x0;            /* per group [3]: s0, i0, alpha0 */
  const double *noise_std;     /* per group [3] or NULL (zeros) */
  int u_kind;                  /* EPI_U_F64 / EPI_U_U8 */
  const void *u;               /* [K][L][B] NPI schedule (K forecast days) */
  const double *noise;         /* [K][3][B] standard normals (s, i, alpha order) or NULL */
  double *s, *i, *alpha;       /* optional trajectories, each [K][B] */
  /* fused NPICost over [history, forecast] (all optional when J0 == NULL): */
  int T_total;                 /* T_hist + K, the divisor of both means */
  const double *j0_prefix;     /* per group: sum of the historic new cases */
  const double *j1_prefix;     /* per group: sum of weights.*inputs over the history */
  const double *w;             /* per group [K][L]: day-wise weights of the forecast block */
  double *J0, *J1;             /* [B] */
  unsigned long long seed;     /* EPI_U_PHILOX: the Philox key (u is ignored) */
  long long first;             /* EPI_U_PHILOX: global index of trajectory 0 of this call (sharded batches) */
} epi_rollout_args;
int epi_rollout_cost_batch(epi_ctx *ctx, const epi_rollout_args *a);

/* The schedules EPI_U_PHILOX integrates, written out: u [K][L][B] uint8 (e.g. to recover the
 * schedule of the Pareto knee).  Same (seed, first, G, prm) as the rollout call. */
typedef struct {
  int mem;
  int B, K, L, G;
  const epi_model_params *prm; /* per group: u_min, u_max */
  unsigned long long seed;
  long long first;
  unsigned char *u;            /* [K][L][B] */
} epi_schedules_args;
int epi_random_schedules(epi_ctx *ctx, const epi_schedules_args *a);

/* replaces  [J0,J1] = NPICost(newcases, inputs, weights)          Tools/NPICost.m:1
 * as a stand-alone call (the rollout above fuses it).  newcases [T][B],
 * inputs [T][L][B], weights per group [T][L]; J0, J1 [B].  Summation order:
 * per-day sums over the L inputs in row order, then over days. */
typedef struct {
  int mem;
  int B, T, L, G;
  const double *newcases, *inputs, *weights;
  double *J0, *J1;
} epi_npicost_args;
int epi_npicost_batch(epi_ctx *ctx, const epi_npicost_args *a);

/* replaces  [s,i] = SI_Controlled(alpha, beta, s0, i0, K, dt)   Tools/SI_Controlled.m:1
 * alpha [K][B]; beta, s0, i0 per trajectory [B]; outputs s, i [K][B]. */
typedef struct {
  int mem;
  int B, K;
  double dt;
  const double *alpha, *beta, *s0, *i0;
  double *s, *i;
} epi_si_args;
int epi_si_controlled_batch(epi_ctx *ctx, const epi_si_args *a);

/* Per-region preprocessing ------------------------------------------------------
 * replaces the data-cleaning block of Tools/TrainPredictPrescribeNPI.m for B regions at once:
 * :121-128 NPI forward fill (N/A -> previous day's level, else 0); :162-172 new cases =
 * diff of the cumulative series, negatives clipped, trailing NaN = last valid value, other NaN = 0;
 * :173 causal W-day moving average (filter(ones(1,W), W, .)); :174 zero-phase round(W/2)-tap moving
 * average (filtfilt); :175-180 normalisation by the population and the cumulative smoothed series;
 * :200-201 I0 = max(min_cases, mean of the first n_first positive smoothed days); :240
 * R_v = 0.1 ((zero-phase - refined)/N)^2.  cc [T][B], population [B], ip [T][L][B]; outputs [T][B]
 * (ip_filled [T][L][B], I0 [B]), each optional except none.  T must exceed 3 (round(W/2) - 1)
 * (filtfilt's data-length rule) and be >= 2 (:166); 1 <= W <= 32. */
typedef struct {
  int mem;
  int B, T, L, W, n_first;
  double min_cases;
  const double *cc, *population, *ip;
  double *ip_filled, *refined, *smoothed, *zerolag, *normalized, *confirmed_norm, *R_v, *I0;
} epi_preprocess_args;
int epi_preprocess_batch(epi_ctx *ctx, const epi_preprocess_args *a);

/* Non-negative regression between the EKF rounds ---------------------------------
 * replaces  reg_coef_a = lsqnonneg(X, y)  followed by the alternating-intercept loop of
 * Tools/TrainPredictPrescribeNPI.m:264-278 (and :326-339) for B regions at once: X [n][p][B]
 * (n regression days, p <= 12 NPIs: NPI_MAXES - InterventionPlans), y [n][B] (the smoothed alpha);
 * a [p][B] >= 0, b [B], n_alt [B] (accepted alternations, optional).  max_alt = 100 in the reference
 * (NONNEGATIVELS_IRERATIONS); 0 = plain lsqnonneg with b = 0. */
typedef struct {
  int mem;
  int B, n, p, max_alt;
  const double *X, *y;
  double *a, *b;
  int *n_alt;
} epi_nnls_args;
int epi_nnls_affine_batch(epi_ctx *ctx, const epi_nnls_args *a);

/* Exponential-fit EKF / smoother -------------------------------------------------
 * replaces  [S_MINUS, S_PLUS, P_MINUS, P_PLUS, K_GAIN, S_SMOOTH, P_SMOOTH, innovations, rho] =
 *              Rt_ExpFitEKF(x, s_init, params, w_bar, v_bar, Ps_init, Q_w, R_v, beta, gamma,
 *              inv_monitor_len, order)                          Tools/Rt_ExpFitEKF.m:1
 * (2 states: new cases and growth exponent; order 2 adds the second-order Hessian terms of
 * Rt_ExpFitEKF.m:153-196; order outside {1,2} returns EPI_ERR_ORDER, :47).  x [T][B] (NaN =
 * missing), s_init [2][B]; params = {time_scale, alpha, sigma} [3], w_bar [2], Ps_init / Q_w
 * [4] column-major and R_v [1] per group of G trajectories.  Outputs per trajectory, each
 * optional: S_* [T][2][B], P_* [T][4][B] (column-major pages), K_GAIN [T][2][B],
 * innovations [T][B], rho [T][B]. */
typedef struct {
  int mem;
  int B, T, G;
  const double *x, *s_init;
  const double *params, *w_bar, *Ps_init, *Q, *R;
  double v_bar, beta, gamma;
  int W, order;
  double *S_MINUS, *S_PLUS, *P_MINUS, *P_PLUS, *K_GAIN, *S_SMOOTH, *P_SMOOTH, *innovations, *rho;
} epi_rt_expfit_args;
int epi_rt_expfit_batch(epi_ctx *ctx, const epi_rt_expfit_args *a);

/* EKF + fixed-interval smoother ------------------------------------------------
 * replaces  GenericExtendedKalmanFilter(u, x, handles, params, s_init, Ps_init,
 *              s_final, Ps_final, w_bar, v_bar, Q_w, R_v, beta, gamma,
 *              inv_monitor_len, order)    Tools/GenericExtendedKalmanFilter.m:1
 *  for the four known handle sets (EPI_MODEL_SIALPHA .. OPTCTRL_FLIPPED), and
 *  NewCaseEKFEstimatorWithOptimalNPI(...)  (EPI_MODEL_LEGACY_*).
 * w_bar is accepted by the reference but read by no callback; it has no field.
 * order 1 and 2 are accepted (every model's Hessian terms are identically
 * zero); anything else returns EPI_ERR_ORDER. */
typedef struct {
  int mem;
  int model;
  int B, T, L, G;
  const epi_model_params *prm;  /* per group */
  const double *epsilon;        /* [B] per-trajectory override of prm.epsilon, or NULL */
  int u_per_traj;  const double *u;   /* per group [T][L] | per trajectory [T][L][B]; NaN = optimise */
  int x_per_traj;  const double *x;   /* per group [T]    | per trajectory [T][B];    NaN = missing */
  int r_mode, fixed_R, r_per_traj;
  const double *R;              /* CONST: per group [1] | per traj [B];  PERDAY: per group [T] | [T][B] */
  int q_mode;
  const double *Q;              /* per group: CONST [m*m] col-major | PERDAY_SCALAR [T] | PERDAY_FULL [T][m*m] */
  int init_per_traj;            /* 0: per group [m] / [m*m] col-major; 1: [m][B] / [m*m][B] */
  const double *s_init, *Ps_init, *s_final, *Ps_final;
  double v_bar, beta, gamma;
  int W, order;                 /* inv_monitor_len, order */
  /* outputs (any may be NULL), trajectory-minor: */
  double *u_opt, *u_opt_smooth;        /* [T][L][B]  (u_opt_smooth NULL-only for LEGACY) */
  double *S_MINUS, *S_PLUS, *S_SMOOTH; /* [T][m][B] */
  double *P_MINUS, *P_PLUS, *P_SMOOTH; /* [T][m*m][B], each m x m column-major */
  double *K_GAIN;                      /* [T][m][B] */
  double *innovations, *rho;           /* [T][B] */
  int *status;                         /* [B] optional: bit0 NaN/Inf guard hit, bits 8.. min pinv rank */
} epi_ekf_args;
int epi_ekf_eks_batch(epi_ctx *ctx, const epi_ekf_args *a);

/* Pareto front + knee -----------------------------------------------------------
 * replaces the filter of Tools/TrainPredictPrescribeNPI.m:624-627 and the knee
 * point of :633.  J0, J1 [n_sets][n]; on_front [n_sets][n]; I_opt [n_sets] (0-based). */
typedef struct {
  int mem;
  int n_sets, n;
  const double *J0, *J1;
  unsigned char *on_front;
  int *I_opt;
} epi_pareto_args;
int epi_pareto_batch(epi_ctx *ctx, const epi_pareto_args *a);

/* The fused optimal-NPI Pareto sweep ---------------------------------------------
 * replaces the loop body + epilogue of Tools/TrainPredictPrescribeNPI.m:421-495,
 * 624-633 for all regions x all epsilon at once: 6-state EKF/EKS
 * (SIAlphaModelEKFOptControlled), SIalpha_Controlled rollout of the smoothed
 * schedule, NPICost, Pareto mask and knee.  Trajectory b = region*n_eps + e. */
typedef struct {
  int mem;
  int n_regions, n_eps, T, T_hist, L;
  const epi_model_params *prm; /* per region (w = NPI cost weights; epsilon ignored) */
  const double *eps;           /* [n_eps] shared grid */
  const double *u;             /* per region [T][L], NaN on days to optimise (:458) */
  const double *x;             /* per region [T], NaN on forecast days (:364) */
  const double *R;             /* per region [T] (:360) */
  const double *s_init, *Ps_init, *s_final, *Ps_final, *Q; /* per region [6] / [36] col-major */
  double beta_ekf, gamma_ekf;
  int W;
  const double *x0;            /* per region [3]: rollout start s,i,alpha (:481) */
  const double *newcases_hist; /* per region [T_hist] (:493 s_historic.*i_historic.*alpha_historic) */
  const double *weights;       /* per region [T][L] day-wise weights (:390) */
  const double *noise_std;     /* per region [3] or NULL */
  const double *noise;         /* [T-T_hist][3][B] or NULL */
  double *J0, *J1;             /* [n_regions][n_eps] */
  unsigned char *on_front;     /* [n_regions][n_eps] or NULL */
  int *I_opt;                  /* [n_regions] or NULL */
  double *u_knee;              /* [n_regions][T-T_hist][L] schedule at the knee, or NULL */
  double *u_fore;              /* [T-T_hist][L][B] every smoothed schedule, or NULL */
  double *P_first;             /* [36][B] P_SMOOTH(:,:,1), or NULL */
  int lean;                    /* 0: run the smoother over all T days, as the reference does.
                                  1: every output of this call except P_first depends on the smoothed
                                     states of the days to optimise only (on history days u is given, so
                                     u_opt_smooth == u: GenericExtendedKalmanFilter.m:229 passes it through),
                                     hence the smoother gains / backward recursion run for days >= T_hist
                                     only and the forward tape keeps those days only.  Same J0, J1, front,
                                     knee and schedules, bit for bit.  Requires a NaN-free history block of
                                     u (else J1 = NaN) and P_first == NULL. */
} epi_sweep_args;
int epi_sweep(epi_ctx *ctx, const epi_sweep_args *a);

/* The same sweep on several GPUs from ONE blocking host call ----------------------
 * replaces the region loop of Tools/TrainPredictPrescribeNPI.m:93 around the call site :460 for a host
 * (MATLAB/Octave through matlab/epi_mex.cpp, or any single process) that owns every GPU of the box:
 * ctxs[i] is a context created on GPU i (epi_create(i, ...)); regions are sharded in contiguous blocks of
 * ceil(n_regions / n_ctx), shard i runs as one epi_sweep on ctxs[i] from its own host thread, and every
 * output lands in the caller's arrays exactly where the single-GPU call puts it (J0/J1/on_front [n_regions]
 * [n_eps], I_opt [n_regions], u_knee, u_fore, P_first) -- bit-identical to epi_sweep on one context.
 * a->mem must be EPI_MEM_HOST (device pointers belong to one GPU).  No collective is involved: the gather is
 * each shard's device-to-host copy.  On error the code of the first failing shard is returned and
 * epi_last_error(ctxs[0]) names it.  n_ctx == 1 is epi_sweep. */
int epi_sweep_multi(epi_ctx *const *ctxs, int n_ctx, const epi_sweep_args *a);

#ifdef __cplusplus
}
#endif
#endif /* EPI_B200_H */
