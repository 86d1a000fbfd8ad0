// ekf_common.cuh -- addressing helpers shared by the EKF forward, smoother-gain
// and smoother-backward kernels.
#pragma once
#include "epi_internal.h"
#include "epi_linalg.cuh"

namespace epi {

static __device__ const double kZeroInputs[EPI_LMAX] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};

// Per-thread view of a per-trajectory array X[t][f][b].
//   strided (caller-visible layout): element (t,f) at p[(t*F + f) * S], S = batch stride
//   tiled   (library scratch)      : [tile = b/32][t][f][32]; S == 32 is a compile-time
//                                    constant, so every field offset folds into the
//                                    instruction's immediate and a warp's access is one
//                                    contiguous 256-byte row of a contiguous (tile, day) page.
template <bool TILED>
struct Tape {
  double *p;   // element (t = 0, f = 0) of this thread's trajectory
  size_t S;    // elements between consecutive fields
  size_t day;  // elements between consecutive days
  EPI_DI double *at_day(int t) const { return p + (size_t)t * day; }
  EPI_DI size_t f(int field) const { return TILED ? (size_t)field * 32 : (size_t)field * S; }
};
// `days` = number of days the array holds, `t0` = absolute index of its first day (lean
// sweeps keep only the days the smoother needs); at_day() takes absolute day indices.
template <bool TILED>
EPI_DI Tape<TILED> make_tape(const TArr &a, int F, int days, int b, int t0 = 0) {
  Tape<TILED> t;
  if (TILED) {
    t.S = 32;
    t.day = (size_t)F * 32;
    t.p = a.p + ((size_t)(b >> 5) * (size_t)days * (size_t)F) * 32 + (size_t)(b & 31) - (size_t)t0 * t.day;
  } else {
    t.p = a.p + (size_t)a.off + (size_t)b;
    t.S = (size_t)a.stride;
    t.day = (size_t)F * (size_t)a.stride;
  }
  return t;
}
// bytes-free element count of a tiled scratch array for a wave of B trajectories
__host__ __device__ inline size_t tiled_elems(long long B, size_t rows) {
  return (size_t)((B + 31) / 32) * 32 * rows;
}

// covariance page <-> registers.  Tiled scratch of the generic (exactly symmetric) models
// holds the packed upper triangle (21 doubles for m = 6); everything else the full
// column-major m x m page (field j*M + i).
template <int M, bool SYM, bool TILED>
EPI_DI void tape_store_cov(const Mat<M, SYM> &Pm, const Tape<TILED> &tp, double *__restrict__ day) {
  if (SYM && TILED) {
#pragma unroll
    for (int i = 0; i < M; ++i)
#pragma unroll
      for (int j = i; j < M; ++j) day[tp.f(Mat<M, true>::idx(i, j))] = Pm(i, j);
  } else {
#pragma unroll
    for (int i = 0; i < M; ++i)
#pragma unroll
      for (int j = 0; j < M; ++j) day[tp.f(j * M + i)] = Pm(i, j);
  }
}
// PACKED: the page holds the packed upper triangle
template <int M, bool SYM, bool TILED, bool PACKED>
EPI_DI void tape_load_cov(Mat<M, SYM> &Pm, const Tape<TILED> &tp, const double *__restrict__ day) {
#pragma unroll
  for (int i = 0; i < M; ++i)
#pragma unroll
    for (int j = 0; j < M; ++j)
      if (!SYM || j >= i) Pm.at(i, j) = day[tp.f(PACKED ? Mat<M, true>::idx(i, j) : (j * M + i))];
}
template <int M, bool TILED>
EPI_DI constexpr int cov_fields(bool sym) { return (sym && TILED) ? M * (M + 1) / 2 : M * M; }

// per-thread view of the per-group / per-trajectory inputs
struct TrajIn {
  const epi_model_params *prm;
  long long g;
  double eps;
  const double *u;   size_t u_js, u_ts;   // u(j, t) = u[t*u_ts + j*u_js]
  const double *x;   size_t x_ts;
  const double *R;   size_t R_ts;         // PERDAY only
  double R_const;
  const double *Q;
  const double *dot_grp;                  // per-group per-day precomputed input term (or null)
  const double *cost_grp;
};
EPI_DI TrajIn traj_inputs(const EkfParams &P, int b, int M) {
  TrajIn t;
  const long long gb = P.b0 + b;
  const long long g = gb / P.G;
  t.g = g;
  t.prm = P.prm + g;
  t.eps = P.epsilon.p ? P.epsilon.p[P.epsilon.off + b]
          : P.eps_grid ? P.eps_grid[(int)(gb % P.eps_mod)] : t.prm->epsilon;
  if (P.u_trj.p) { t.u = P.u_trj.p + P.u_trj.off + b; t.u_js = (size_t)P.u_trj.stride; t.u_ts = (size_t)P.L * P.u_trj.stride; }
  else           { t.u = P.u_grp + (size_t)g * P.T * P.L; t.u_js = 1; t.u_ts = (size_t)P.L; }
  if (P.x_trj.p) { t.x = P.x_trj.p + P.x_trj.off + b; t.x_ts = (size_t)P.x_trj.stride; }
  else           { t.x = P.x_grp + (size_t)g * P.T; t.x_ts = 1; }
  t.R = nullptr; t.R_ts = 0; t.R_const = 0.0;
  if (P.r_mode == EPI_R_CONST) {
    t.R_const = P.R_trj.p ? P.R_trj.p[P.R_trj.off + b] : P.R_grp[g];
  } else {
    if (P.R_trj.p) { t.R = P.R_trj.p + P.R_trj.off + b; t.R_ts = (size_t)P.R_trj.stride; }
    else           { t.R = P.R_grp + (size_t)g * P.T; t.R_ts = 1; }
  }
  const size_t mm = (size_t)M * M;
  t.Q = P.Q + (P.q_mode == EPI_Q_CONST ? (size_t)g * mm
               : P.q_mode == EPI_Q_PERDAY_SCALAR ? (size_t)g * P.T : (size_t)g * P.T * mm);
  t.dot_grp = P.dot_grp ? P.dot_grp + (size_t)g * P.T : nullptr;
  t.cost_grp = P.cost_grp ? P.cost_grp + (size_t)g * P.T : nullptr;
  return t;
}
EPI_DI double q_elem(const double *__restrict__ Q, int q_mode, int M, int k, int i, int j) {
  // read-only path: the loads may then be scheduled across the tape stores of the same day
  if (q_mode == EPI_Q_CONST) return __ldg(Q + j * M + i);
  if (q_mode == EPI_Q_PERDAY_FULL) return __ldg(Q + (size_t)k * M * M + j * M + i);
  return (i == j) ? __ldg(Q + k) : 0.0;  // B*q*B' with B = I
}

#define EPI_DISPATCH_MODEL(model, CALL)                                   \
  switch (model) {                                                        \
    case EPI_MODEL_SIALPHA: CALL(EPI_MODEL_SIALPHA); break;               \
    case EPI_MODEL_SIALPHA_FLIPPED: CALL(EPI_MODEL_SIALPHA_FLIPPED); break; \
    case EPI_MODEL_OPTCTRL: CALL(EPI_MODEL_OPTCTRL); break;               \
    case EPI_MODEL_OPTCTRL_FLIPPED: CALL(EPI_MODEL_OPTCTRL_FLIPPED); break; \
    case EPI_MODEL_LEGACY_TOOLS: CALL(EPI_MODEL_LEGACY_TOOLS); break;     \
    case EPI_MODEL_LEGACY_CODEGEN: CALL(EPI_MODEL_LEGACY_CODEGEN); break; \
    default: break;                                                       \
  }

}  // namespace epi
