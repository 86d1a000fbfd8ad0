/*
 * epi_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the reference's MATLAB hot path
 * (alphanumericslab/EpidemicModeling).  Every function cites the reference
 * file:line it follows.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library, and only as the
 * checker / CPU baseline -- never as the shipped path.
 *
 * PARITY STATUS: "parity unpinned".  The reference ships no numeric golden
 * vectors and neither MATLAB nor Octave exists in this image, so the oracle
 * cannot be checked against reference outputs; it is pinned only against the
 * structural invariants and the closed-form SEIRP solution the reference
 * scripts contain (tests/test_oracle_*.py) and against an independently
 * written NumPy/LAPACK twin (oracle/numpy_twin.py).
 *
 * Arithmetic contract (shared with the CUDA kernels so that agreement is
 * bit-for-bit, not "close"):
 *   - IEEE-754 binary64, round-to-nearest-even, no implicit FMA contraction
 *     (gcc -ffp-contract=off / nvcc --fmad=false);
 *   - scalar expressions are evaluated exactly as MATLAB parses them
 *     (left-to-right, unary minus before '*');
 *   - matrix products (BLAS in MATLAB, order unspecified there) are DEFINED
 *     here as: first term a*b, every further term fma(a,b,acc), index
 *     ascending, structural zeros of C (cols 4..6), A (see orc_A_pattern)
 *     and B = I, D = 1 skipped;
 *   - min/max follow MATLAB (the non-NaN operand wins);
 *   - pinv is DEFINED as a threshold-cyclic-Jacobi eigendecomposition of the
 *     symmetric argument followed by MATLAB's rank truncation
 *     tol = max(size)*eps(max|lambda|)  (see orc_pinv_sym);
 *   - mrdivide (legacy model) is DEFINED as LU with partial pivoting of the
 *     transposed system (see orc_mrdivide).
 */
#ifndef EPI_ORACLE_H
#define EPI_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_LMAX 12

/* model variants (the Tools wrappers that bind callbacks to the generic filter) */
enum {
  ORC_SIALPHA = 0,          /* Tools/SIAlphaModelEKF.m                         m=3 */
  ORC_SIALPHA_FLIPPED = 1,  /* Tools/SIAlphaModelBackwardEKF.m                 m=3 */
  ORC_OPTCTRL = 2,          /* Tools/SIAlphaModelEKFOptControlled.m            m=6 */
  ORC_OPTCTRL_FLIPPED = 3,  /* Tools/SIAlphaModelBackwardEKFOptControlled.m    m=6 */
  ORC_LEGACY_TOOLS = 4,     /* Tools/NewCaseEKFEstimatorWithOptimalNPI.m       m=6 */
  ORC_LEGACY_CODEGEN = 5    /* MatlabCodeGenerator/NewCaseEKF...+callbacks     m=6 */
};
enum { ORC_OBS_NEWCASES = 0, ORC_OBS_TOTALCASES = 1 };
enum { ORC_Q_CONST = 0, ORC_Q_PERDAY_SCALAR = 1, ORC_Q_PERDAY_FULL = 2 };
enum { ORC_R_CONST = 0, ORC_R_PERDAY = 1 };

/* the reference's `params` struct (MatlabCodeGenerator/...prj:1105-1116 plus
 * s_min, i_min, obs_type of Tools/TrainPredictPrescribeNPI.m:202-224) */
typedef struct {
  double dt, beta, gamma, b;
  double alpha_min, alpha_max, s_min, i_min;
  double epsilon, sigma;
  double a[ORC_LMAX], u_min[ORC_LMAX], u_max[ORC_LMAX], w[ORC_LMAX];
  int L;
  int obs_type;
} orc_params;

/* Tools/SEIRP.m:13-32.  rate_stride = 1: rates are 1xK vectors (only the first
 * K-1 entries are read, as in the reference); 0: scalars. */
void orc_seirp(const double *alpha_e, const double *alpha_i, const double *kappa,
               const double *rho, const double *beta, const double *mu,
               const double *gamma, int rate_stride, double s0, double e0, double i0,
               double r0, double p0, int K, double dt, double *s, double *e, double *i,
               double *r, double *p);

/* Tools/SEIRPSaturatedResource.m:13-36 */
void orc_seirp_saturated(const double *alpha_e, const double *alpha_i, const double *kappa,
                         const double *rho, const double *gamma, int rate_stride, double s0,
                         double e0, double i0, double r0, double p0, int K, double dt,
                         double beta_0, double beta_s, double mu_0, double mu_s, double sigma,
                         double i_0, double *s, double *e, double *i, double *r, double *p);

/* Tools/SIalpha_Controlled.m:15-32.  u is LxK column-major; noise is 3xK
 * column-major standard-normal draws consumed in the reference's randn call
 * order (s, i, alpha) or NULL for zero noise.  Outputs are 1xK (initial
 * condition dropped, as :30-32). */
void orc_sialpha_controlled(const double *u, int L, double s0, double i0, double alpha0,
                            const double *u_max, double alpha_min, double alpha_max,
                            double gamma, const double *a, double b, double beta,
                            double s_noise_std, double i_noise_std, double alpha_noise_std,
                            int K, double dt, const double *noise, double *s, double *i,
                            double *alpha);

/* Tools/SI_Controlled.m:12-22 */
void orc_si_controlled(const double *alpha, double beta, double s0, double i0, int K, double dt,
                       double *s, double *i);

/* Tools/NPICost.m:6-10.  inputs/weights are LxT column-major; sums run in
 * a defined order: per-day column sums (row order), then over days. */
void orc_npicost(const double *newcases, int T, const double *inputs, const double *weights,
                 int L, double *J0, double *J1);

/* Tools/TrainPredictPrescribeNPI.m:624-633: strict-dominance Pareto mask and
 * knee index (0-based; first minimum, NaN skipped as MATLAB's min does). */
void orc_pareto(const double *J0, const double *J1, int n, unsigned char *on_front, int *I_opt);

/* Tools/Rt_ExpFitEKF.m: 2-state exponential-fit EKF/EKS with second-order (Hessian) terms */
int orc_rt_expfit_ekf(const double *x, int T, const double *s_init, const double *params, const double *w_bar,
                      double v_bar, const double *Ps_init, const double *Q, double R, double beta, double gamma,
                      int W, int order, double *S_MINUS, double *S_PLUS, double *P_MINUS, double *P_PLUS,
                      double *K_GAIN, double *S_SMOOTH, double *P_SMOOTH, double *innovations, double *rho);

/* per-region preprocessing: Tools/TrainPredictPrescribeNPI.m:121-128, 162-187, 200-201, 240 */
int orc_preprocess_region(const double *cc, int T, double N, int W, int n_first, double min_cases, double *ip,
                          int L, double *refined, double *smoothed, double *zerolag, double *normalized,
                          double *confirmed_norm, double *R_v, double *I0);

/* non-negative regression with alternating intercept: TrainPredictPrescribeNPI.m:264-278 */
int orc_nnls_affine(const double *X, const double *y, int n, int p, int max_alt, double *a, double *b);
void orc_lsqnonneg(const double *X, const double *y, int n, int p, double *a);

/* random NPI schedules (TrainPredictPrescribeNPI.m:499-510) on a Philox4x32-10 counter stream */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
void orc_random_schedule(uint64_t seed, uint32_t region, uint32_t scenario, int n_scenarios, int L, int K,
                         const double *u_min, const double *u_max, unsigned char *u);

/* pinv as DEFINED by this oracle for the generic smoother
 * (Tools/GenericExtendedKalmanFilter.m:215).  A, X are mxm column-major
 * (A symmetric, only the upper triangle is read).  Returns the number of
 * Jacobi sweeps; *rank receives the number of retained eigenvalues. */
int orc_pinv_sym(const double *A, int m, double *X, int *rank);

/* mrdivide as DEFINED by this oracle for the legacy smoother
 * (Tools/NewCaseEKFEstimatorWithOptimalNPI.m:132):  X = B / A. */
void orc_mrdivide(const double *B, const double *A, int m, double *X);

/* The EKF + fixed-interval smoother.
 * generic:  Tools/GenericExtendedKalmanFilter.m:41-233 with the callbacks of
 *           the selected model;
 * legacy:   Tools/NewCaseEKFEstimatorWithOptimalNPI.m:9-143 (+ :150-290).
 * The *_FLIPPED models include the time flip of u, x, the init/final swap and
 * the flip back of every output except rho (Tools/SIAlphaModelBackwardEKF.m:19-40).
 * Arrays are MATLAB column-major: u LxT, x 1xT, S_* mxT, P_* mxmxT, K_GAIN mxT.
 * Any output pointer may be NULL.  Returns 0, or <0 on an argument the
 * reference would `error()` on. */
int orc_ekf_eks(int model, const orc_params *prm, int T, const double *u, const double *x,
                const double *s_init, const double *Ps_init, const double *s_final,
                const double *Ps_final, double v_bar, int q_mode, const double *Q, int r_mode,
                int fixed_R, const double *R, double beta, double gamma, int W, int order,
                double *u_opt, double *u_opt_smooth, double *S_MINUS, double *S_PLUS,
                double *S_SMOOTH, double *P_MINUS, double *P_PLUS, double *P_SMOOTH,
                double *K_GAIN, double *innovations, double *rho);

/* One region of the optimal-NPI Pareto sweep
 * (Tools/TrainPredictPrescribeNPI.m:415-495 and :624-633):  for each epsilon
 * run the 6-state EKF/EKS, roll the smoothed schedule out with
 * SIalpha_Controlled from (s_h, i_h, alpha_h), and cost it with NPICost. */
typedef struct {
  orc_params prm;            /* epsilon field ignored (taken from eps[]) */
  int T, T_hist;             /* T = T_hist + forecast days */
  const double *u_hist;      /* L x T_hist */
  const double *x;           /* 1 x T (NaN on forecast days) */
  const double *R;           /* 1 x T */
  double s_init[6], Ps_init[36], s_final[6], Ps_final[36], Q[36];
  double beta_ekf, gamma_ekf;
  int W;
  double s_h, i_h, alpha_h;  /* rollout start = last historic smoothed state */
  const double *newcases_hist; /* 1 x T_hist : s.*i.*alpha of the historic estimate */
  const double *weights;     /* L x T day-wise NPI weights */
  double noise_std[3];
  const double *noise;       /* n_eps x (3 x T_f) or NULL */
} orc_sweep_region;

void orc_sweep_region_run(const orc_sweep_region *rg, const double *eps, int n_eps, double *J0,
                          double *J1, double *u_fore /* n_eps x (L x T_f) or NULL */,
                          unsigned char *on_front, int *I_opt);

/* batch drivers used ONLY for CPU-baseline timing (OpenMP over trajectories) */
void orc_sweep_batch(const orc_sweep_region *rg, int n_regions, const double *eps, int n_eps,
                     double *J0, double *J1, unsigned char *on_front, int *I_opt, int n_threads);
int orc_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif
