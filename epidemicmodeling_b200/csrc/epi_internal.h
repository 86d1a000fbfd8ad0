// epi_internal.h -- kernel parameter blocks and launcher prototypes shared
// between the kernel translation units and the C-ABI layer (capi.cu).
#pragma once
#include <cstddef>
#include <cuda_runtime.h>

#include "../../include/epi_b200.h"

namespace epi {

// A per-trajectory array in trajectory-minor layout: element (t, f, b) lives at
// p[((size_t)t * F + f) * stride + off + b].  p == nullptr means "absent".
// Caller-owned device arrays have stride = caller's B and off = first
// trajectory of the current wave; library scratch and host-staged buffers have
// stride = wave size and off = 0.
struct TArr {
  double *p;
  long long stride, off;
};
struct CArr {
  const double *p;
  long long stride, off;
};

// All pointers are DEVICE pointers (see include/epi_b200.h for the layouts).
// `B` = trajectories of THIS launch (one wave of the caller's batch), `b0` =
// index of its first trajectory in the caller's batch (group = (b0 + b) / G).
struct EkfParams {
  int model;
  int B, T, L, G, W;
  long long b0;
  const epi_model_params *prm;     // per group
  CArr epsilon;                    // [B] or absent
  const double *eps_grid;          // sweep: shared grid, epsilon = eps_grid[(b0 + b) % eps_mod]
  int eps_mod;
  const double *u_grp; CArr u_trj; // per group [T][L]  |  per trajectory [T][L][B]
  const double *x_grp; CArr x_trj; // per group [T]     |  per trajectory [T][B]
  int r_mode, fixed_R;
  const double *R_grp; CArr R_trj; // CONST: [1] | [B];  PERDAY: [T] | [T][B]
  int q_mode;
  const double *Q;                 // per group
  int init_per_traj;
  const double *s_init_g, *Ps_init_g, *s_final_g, *Ps_final_g;  // per group [m] / [m*m]
  CArr s_init_t, Ps_init_t, s_final_t, Ps_final_t;              // per trajectory [m][B] / [m*m][B]
  double v_bar, beta, gamma;
  // the forward tape (always present: caller outputs or scratch)
  TArr S_MINUS, S_PLUS, P_MINUS, P_PLUS;
  int tiled;                       // 1: all four tape arrays are library scratch in the tile layout
                                   //    [b/32][T][F][32] (ekf_common.cuh); P pages of the generic
                                   //    models then hold the packed upper triangle (F = m(m+1)/2)
  int k0;                          // first day the smoother needs (0 = all).  Lean sweeps (tiled only):
                                   // the tape holds days k0..T-1, gains/backward run for k >= k0
  int fwd_segments;                // > 1: time-segmented persistent forward launch (m = 6 generic, tiled, no monitor)
  int bwd_prefetch;                // > 0: the backward recursion pulls the tape page of day k - bwd_prefetch into L2 (small batches)
  int bwd_stages;                  // > 0: staged backward recursion (small batches): shared-memory ring of this many tape days per tile
  int *fwd_sync;                   //      [1 + tiles] ints, zeroed: item counter, per-tile finished segments
  // forward || gains for small batches ("piped schedule"): the forward kernel counts the tiles that have
  // finished time chunk c in pipe_sync[c]; the gains of the days [gk_lo, gk_hi) are launched on a second stream behind
  // a stream wait on that word.  gk_lo == gk_hi == 0: the gain launch covers all days k0 .. T-2.
  int pipe_chunks;
  int day_sync;                    //      1: the warps of a forward CTA meet at a barrier every day (instruction lines shared)
  unsigned *pipe_sync;             //      [pipe_chunks] words, zeroed
  int gk_lo, gk_hi;
  TArr J;                          // scratch smoother gains, always tiled: [b/32][T-1-k0][m*m][32]
  const double *dot_grp;           // per group [T]: input term of days without NaN inputs (NaN = per trajectory)
  const double *cost_grp;          // per group [T]: that day's sum_j w*u (sweep), or null
  // optional outputs
  TArr u_opt, u_opt_smooth, S_SMOOTH, P_SMOOTH, K_GAIN, innov, rho;
  int *status;                     // [B] of this wave, or null
  // sweep extras: per-day scalars consumed by the fused rollout
  TArr dot_day;                    // tiled [b/32][T][1][32]: gamma*a'*(u_max - u_opt_smooth(:,t))
  TArr cost_day;                   // tiled [b/32][T][1][32]: sum_j w(j,t)*u_opt_smooth(j,t)
  const double *weights;           // per group [T][L]
  TArr u_fore;                     // [T-T_hist][L][B]
  int T_hist;
  TArr P_first;                    // [m*m][B]
};

void launch_group_day(const epi_model_params *prm, const double *u, const double *weights, int n_groups,
                      int T, int L, double *dot_grp, double *cost_grp, cudaStream_t st);
void launch_ekf_forward(const EkfParams &p, cudaStream_t st);
int forward_segments(long long tiles, int slots);  // segment count minimising the idle tail (1 = plain launch)
int forward_resident_slots6();                     // resident one-warp CTAs of the m = 6 forward kernel on this device
void launch_eks_gain(const EkfParams &p, cudaStream_t st);
// piped schedule: four tiles per CTA, one warp per SM sub-partition, the SM kept to itself (csrc/ekf_forward.cu)
bool forward_piped_ok(const EkfParams &p, int n_sms);
void launch_ekf_forward_piped(const EkfParams &p, cudaStream_t st);
void pipe_chunk_days(const EkfParams &p, int c, int &kb, int &ke);  // days [kb, ke) of time chunk c
void launch_eks_backward(const EkfParams &p, cudaStream_t st);
// lane-group forms (csrc/ekf_rows.cu): six lanes per trajectory, for small batches of the sweep's call shape
void launch_ekf_forward_rows(const EkfParams &p, cudaStream_t st);
void launch_eks_backward_rows(const EkfParams &p, cudaStream_t st);
bool rows_forward_ok(const EkfParams &p);
bool rows_backward_ok(const EkfParams &p);
bool rows_wanted(long long B, bool forward);
// two warps per tile (csrc/ekf_pair.cu): the forward pass of small batches of the sweep's call shape
void launch_ekf_forward_pair(const EkfParams &p, cudaStream_t st);
bool pair_forward_wanted(const EkfParams &p);

struct SeirpParams {
  int B, K, rate_mode, saturated, out_mode;
  double dt, beta_0, beta_s, mu_0, mu_s, sigma, i_0;
  CArr rates;                  // CONST [7][B]; SERIES [K][7][B]
  const double *rates_shared;  // SHARED_SERIES [7][K]
  CArr ic;                     // [5][B]
  TArr out;                    // FULL [5][K][B]; FINAL [5][B]
};
void launch_seirp(const SeirpParams &p, cudaStream_t st);

struct RolloutParams {
  int B, K, L, G;
  long long b0;
  const epi_model_params *prm;
  const double *x0, *noise_std;   // per group [3]
  int u_kind;  // EPI_U_F64, EPI_U_U8, 2 = precomputed per-day scalars (sweep), EPI_U_PHILOX = generated
  unsigned long long seed;        // EPI_U_PHILOX
  long long first;                // EPI_U_PHILOX: global index of the caller's trajectory 0
  const void *u; long long u_stride, u_off;   // [K][L][B]
  CArr noise;                     // [K][3][B]
  TArr s, i, alpha;               // [K][B]
  int T_total, T_hist;
  const double *j0_prefix, *j1_prefix, *w;    // per group
  const double *newcases_hist;    // sweep: per group [T_hist] (summed in-kernel)
  const double *dot_day, *cost_day;  // sweep: tiled [b/32][T_total][1][32] of this wave
  const double *hist_cost_grp;       // sweep: per group [T_total] day costs of the history days whose inputs are all given (NaN otherwise)
  int hist_cost_per_traj;            // 1 (full sweep): a NaN entry falls back to the per-trajectory cost_day written by eks_backward
  TArr J0, J1;                    // [B]
};
void launch_rollout(const RolloutParams &p, cudaStream_t st);
// sweep: the NPICost sums over the (given) history are the same for every trajectory of a region: one thread per region
// adds them in day order; pre1[g] = NaN marks a region whose history has a day with missing NPIs (per-trajectory sums)
void launch_hist_prefix(const double *newcases_hist, const double *cost_grp, int n_groups, int T_hist, int T_total,
                        double *pre0, double *pre1, cudaStream_t st);
// write the EPI_U_PHILOX schedules: u [K][L][stride] uint8, trajectories first + b0 + (0..B-1)
void launch_random_schedules(const epi_model_params *prm, unsigned long long seed, long long first, int B, int K,
                             int L, int G, unsigned char *u, long long stride, long long off, cudaStream_t st);

// Tools/Rt_ExpFitEKF.m (rt_expfit.cu).  Per-trajectory arrays [T][F][B]; per-group tables.
struct RtParams {
  int B, T, G, W, order;
  long long b0;
  CArr x;                      // [T][B], NaN = missing
  CArr s_init;                 // [2][B]
  const double *params;        // per group [3]: time_scale, alpha, sigma
  const double *w_bar;         // per group [2]
  const double *Ps_init, *Q;   // per group [4] column-major 2x2
  const double *R;             // per group [1]
  double v_bar, beta, gamma;
  TArr S_MINUS, S_PLUS, P_MINUS, P_PLUS;   // the tape: always present (caller outputs or scratch)
  TArr K_GAIN, S_SMOOTH, P_SMOOTH, innov, rho;  // optional
};
void launch_rt_expfit(const RtParams &p, cudaStream_t st);

// per-region preprocessing (preprocess.cu).  Every per-region array is [rows][B], B = regions.
struct PreprocParams {
  int B, T, L, W, n_first;
  double min_cases;
  const double *cc;          // [T][B] cumulative confirmed cases (NaN allowed)
  const double *population;  // [B]
  const double *ip_in;       // [T][L][B] NPI levels (NaN allowed)
  double *ip_out;            // [T][L][B]
  double *refined, *smoothed, *zerolag, *normalized, *confirmed_norm, *R_v;  // [T][B]
  double *I0;                // [B]
  double *scratch;           // [2][T + 2*nfact][B]
};
void launch_preprocess(const PreprocParams &p, cudaStream_t st);

// non-negative regression with alternating intercept (nnls.cu); every array [rows][B]
struct NnlsParams {
  int B, n, p, max_alt;
  const double *X;   // [n][p][B]
  const double *y;   // [n][B]
  double *a;         // [p][B]
  double *b;         // [B]
  int *n_alt;        // [B] accepted alternations, or null
};
void launch_nnls_affine(const NnlsParams &q, cudaStream_t st);

struct SiParams {
  int B, K;
  double dt;
  CArr alpha, beta, s0, i0;
  TArr s, i;
};
void launch_si(const SiParams &p, cudaStream_t st);

struct CostParams {
  int B, T, L, G;
  long long b0;
  CArr newcases, inputs;    // [T][B], [T][L][B]
  const double *weights;    // per group [T][L]
  TArr J0, J1;
};
void launch_npicost(const CostParams &p, cudaStream_t st);

struct ParetoParams {
  int n_sets, n;
  const double *J0, *J1;
  unsigned char *on_front;
  int *I_opt;
};
void launch_pareto(const ParetoParams &p, cudaStream_t st);
// large point sets: sort-based evaluation of the same predicate (pareto_sorted.cu)
constexpr int kParetoBruteMax = 8192;  // points per set up to which the O(n^2) CTA kernel is used
size_t pareto_sorted_scratch_bytes(int n_sets, int n);
int launch_pareto_sorted(const ParetoParams &p, void *scratch, size_t scratch_bytes, cudaStream_t st);

// gather the knee schedule: u_knee[r][t][j] = u_fore[t][j][r*n_eps + I_opt[r]]
void launch_gather_knee(const double *u_fore, const int *I_opt, double *u_knee, int n_regions,
                        int n_eps, int Tf, int L, cudaStream_t st);

// FP64 FMA throughput probe (bench.py's measured FP64 roof): each thread runs
// `iters` rounds of 8 independent dependent-FMA chains; writes one value/thread.
void launch_fp64_probe(double *out, int blocks, int threads, int iters, cudaStream_t st);

}  // namespace epi
