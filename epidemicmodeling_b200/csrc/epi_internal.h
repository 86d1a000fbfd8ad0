// epi_internal.h -- kernel parameter blocks and launcher prototypes shared
// between the kernel translation units and the C-ABI layer (capi.cu).
#pragma once
#include <cstddef>
#include <cuda_runtime.h>

#include "../../include/epi_b200.h"

namespace epi {

// A per-trajectory array in trajectory-minor layout: element (t, f, b) lives at
// p[((size_t)t * F + f) * stride + off + b].  p == nullptr means "absent".
// Caller-owned arrays have stride = caller's B and off = first trajectory of the
// current wave; library scratch has stride = wave size and off = 0.
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
// index of its first trajectory in the caller's batch.
struct EkfParams {
  int model;
  int B, T, L, G, W;
  long long b0;
  const epi_model_params *prm;     // per group
  CArr epsilon;                    // [B] or absent
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
  TArr J;                          // scratch smoother gains [T-1][m*m][B]
  // optional outputs
  TArr u_opt, u_opt_smooth, S_SMOOTH, P_SMOOTH, K_GAIN, innov, rho;
  int *status;                     // [caller B] (+b0) or null
  // sweep extras: per-day scalars consumed by the fused rollout
  TArr dot_day;                    // [T][B] gamma*a'*(u_max - u_opt_smooth(:,t))
  TArr cost_day;                   // [T][B] sum_j w(j,t)*u_opt_smooth(j,t)
  const double *weights;           // per group [T][L]
  TArr u_fore;                     // [T-T_hist][L][B]
  int T_hist;
  TArr P_first;                    // [m*m][B]
};

void launch_ekf_forward(const EkfParams &p, cudaStream_t st);
void launch_eks_gain(const EkfParams &p, cudaStream_t st);
void launch_eks_backward(const EkfParams &p, cudaStream_t st);

struct SeirpParams {
  int B, K, rate_mode, saturated, out_mode;
  double dt, beta_0, beta_s, mu_0, mu_s, sigma, i_0;
  const double *rates, *ic;
  double *out;
};
void launch_seirp(const SeirpParams &p, cudaStream_t st);

struct RolloutParams {
  int B, K, L, G;
  const epi_model_params *prm;
  const double *x0, *noise_std;
  int u_kind;  // EPI_U_F64, EPI_U_U8, or 2 = precomputed per-day scalars (sweep)
  const void *u;
  const double *noise;
  double *s, *i, *alpha;
  int T_total, T_hist;
  const double *j0_prefix, *j1_prefix, *w;
  const double *newcases_hist;  // sweep: per group [T_hist] (summed in-kernel)
  const double *dot_day, *cost_day;  // sweep: [T][B]
  double *J0, *J1;
};
void launch_rollout(const RolloutParams &p, cudaStream_t st);

struct SiParams {
  int B, K;
  double dt;
  const double *alpha, *beta, *s0, *i0;
  double *s, *i;
};
void launch_si(const SiParams &p, cudaStream_t st);

struct ParetoParams {
  int n_sets, n;
  const double *J0, *J1;
  unsigned char *on_front;
  int *I_opt;
};
void launch_pareto(const ParetoParams &p, cudaStream_t st);

// gather the knee schedule: u_knee[r][t][j] = u_fore[t][j][r*n_eps + I_opt[r]]
void launch_gather_knee(const double *u_fore, const int *I_opt, double *u_knee, int n_regions,
                        int n_eps, int Tf, int L, cudaStream_t st);

}  // namespace epi
