// ekf_kernels.cu -- EKF forward pass, smoother-gain pass and backward pass for
// batches of independent trajectories (sm_100a, FP64, --fmad=false).
//
// Reference behaviour: Tools/GenericExtendedKalmanFilter.m:41-233 (generic
// models) and Tools/NewCaseEKFEstimatorWithOptimalNPI.m:9-143 (legacy models).
//
// B200 mapping (DESIGN.md "Kernels"):
//   ekf_forward  : one thread per trajectory, strictly sequential in time
//                  (:98-186).  State, covariance (packed symmetric for the
//                  generic models), gain and Jacobian stay in registers; the
//                  per-day tape (S_MINUS, S_PLUS, P_MINUS, P_PLUS) is written
//                  trajectory-minor so every store instruction of a warp is
//                  one 256-byte coalesced transaction.
//   eks_gain     : the smoother gain J_k = (P+_k A_k') pinv(P-_{k+1}) (:206-217)
//                  depends only on the forward tape, NOT on the backward
//                  recursion, so it is computed for every (trajectory, day)
//                  pair in parallel: (T-1)*B threads.  This is where >80 % of
//                  the FP64 work is (the 6x6 Jacobi pinv) and it has all the
//                  parallelism a B200 wants even for a single-region sweep.
//   eks_backward : one thread per trajectory, the cheap sequential recursion
//                  (:218-229) + the bang-bang schedule of the smoothed costate.
#include "epi_internal.h"
#include "epi_linalg.cuh"

namespace epi {

__device__ const double kZeroInputs[EPI_LMAX] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};

EPI_DI size_t tidx(const TArr &a, int t, int f, int F, int b) {
  return ((size_t)t * F + f) * (size_t)a.stride + (size_t)a.off + b;
}
EPI_DI size_t cidx(const CArr &a, int t, int f, int F, int b) {
  return ((size_t)t * F + f) * (size_t)a.stride + (size_t)a.off + b;
}

// covariance tape pages: full column-major m x m (caller-visible outputs) or,
// for library scratch of the generic (exactly symmetric) models, the packed
// upper triangle -- 21 instead of 36 doubles per page for m = 6.
template <int M, bool SYM>
EPI_DI void store_tape(const Mat<M, SYM> &Pm, const TArr &a, int t, int b, bool packed) {
  if (SYM && packed) {
    constexpr int N = Mat<M, true>::N;
    double *base = a.p + ((size_t)t * N) * (size_t)a.stride + (size_t)a.off + b;
#pragma unroll
    for (int i = 0; i < M; ++i)
#pragma unroll
      for (int j = i; j < M; ++j) base[(size_t)Mat<M, true>::idx(i, j) * (size_t)a.stride] = Pm(i, j);
  } else {
    store_mat<M, SYM>(Pm, a.p + tidx(a, t, 0, M * M, b), (size_t)a.stride);
  }
}
template <int M, bool SYM>
EPI_DI void load_tape(Mat<M, SYM> &Pm, const TArr &a, int t, int b, bool packed) {
  if (packed) {
    constexpr int N = Mat<M, true>::N;
    const double *base = a.p + ((size_t)t * N) * (size_t)a.stride + (size_t)a.off + b;
#pragma unroll
    for (int i = 0; i < M; ++i)
#pragma unroll
      for (int j = 0; j < M; ++j)
        if (!SYM || j >= i) Pm.at(i, j) = base[(size_t)Mat<M, true>::idx(i, j) * (size_t)a.stride];
  } else {
    load_mat<M, SYM>(Pm, a.p + tidx(a, t, 0, M * M, b), (size_t)a.stride);
  }
}

// per-thread view of the per-group / per-trajectory inputs
struct TrajIn {
  const epi_model_params *prm;
  double eps;
  const double *u;   size_t u_js, u_ts;   // u(j, t) = u[t*u_ts + j*u_js]
  const double *x;   size_t x_ts;
  const double *R;   size_t R_ts;         // PERDAY only
  double R_const;
  const double *Q;
};
EPI_DI TrajIn traj_inputs(const EkfParams &P, int b, int M) {
  TrajIn t;
  const long long gb = P.b0 + b;
  const long long g = gb / P.G;
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
  return t;
}
EPI_DI double q_elem(const double *__restrict__ Q, int q_mode, int M, int k, int i, int j) {
  if (q_mode == EPI_Q_CONST) return Q[j * M + i];
  if (q_mode == EPI_Q_PERDAY_FULL) return Q[(size_t)k * M * M + j * M + i];
  return (i == j) ? Q[k] : 0.0;  // B*q*B' with B = I
}

// ===========================================================================
// forward pass
// ===========================================================================
template <int MODEL, bool MONITOR>
__global__ void __launch_bounds__(64) ekf_forward_kernel(const __grid_constant__ EkfParams P) {
  constexpr int M = model_dim(MODEL);
  constexpr bool LEG = model_legacy(MODEL);
  constexpr bool SYM = !LEG;
  constexpr bool REV = model_flipped(MODEL);
  constexpr int MM = M * M;
  extern __shared__ double win[];  // MONITOR: [3][W][blockDim.x]

  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= P.B) return;
  const TrajIn in = traj_inputs(P, b, M);
  const epi_model_params *__restrict__ prm = in.prm;
  const int T = P.T, L = P.L, W = P.W;
  const int obs_type = prm->obs_type;
  const double gamma = P.gamma, beta = P.beta, v_bar = P.v_bar;

  double s[M];
  Mat<M, SYM> Pm;  // P(k|k-1)
  if (P.init_per_traj) {
#pragma unroll
    for (int i = 0; i < M; ++i) s[i] = P.s_init_t.p[cidx(P.s_init_t, 0, i, M, b)];
    load_mat<M, SYM>(Pm, P.Ps_init_t.p + P.Ps_init_t.off + b, (size_t)P.Ps_init_t.stride);
  } else {
    const long long g = (P.b0 + b) / P.G;
#pragma unroll
    for (int i = 0; i < M; ++i) s[i] = P.s_init_g[g * M + i];
    load_mat<M, SYM>(Pm, P.Ps_init_g + (size_t)g * MM, 1);
  }

  if (MONITOR) {
    for (int j = 0; j < 3 * W; ++j) win[(size_t)j * blockDim.x + threadIdx.x] = 0.0;
  }
  int head = 0;               // ring position of the newest window slot
  double R_over = 0.0;        // adapted R for the next step (:184)
  bool has_over = false;

  for (int k = 0; k < T; ++k) {
    const int pos = REV ? (T - 1 - k) : k;
    // :100-101 store the a-priori estimate
#pragma unroll
    for (int i = 0; i < M; ++i) P.S_MINUS.p[tidx(P.S_MINUS, pos, i, M, b)] = s[i];
    store_tape<M, SYM>(Pm, P.P_MINUS, pos, b, P.tape_packed != 0);

    double Rk;
    if (LEG) {
      Rk = (k == 0) ? in.R_const : R_over;  // scalar R adapted in place (:31,:111)
    } else {
      const double base = (P.r_mode == EPI_R_CONST) ? in.R_const : in.R[(size_t)k * in.R_ts];
      Rk = has_over ? R_over : base;
      has_over = false;
    }
    double C[3];
    const double xhat = obs_model<MODEL>(obs_type, s, v_bar, C);  // :115-119
    const double xk = in.x[(size_t)pos * in.x_ts];
    const bool valid = !(xk != xk);                               // :122

    double K[M], sp[M], innov;
    Mat<M, SYM> Pp;  // P(k|k)
    if (valid) {
      innov = xk - xhat;  // :123
      double PCt[M], CP[3];
#pragma unroll
      for (int i = 0; i < M; ++i)
        PCt[i] = fma(Pm(i, 2), C[2], fma(Pm(i, 1), C[1], Pm(i, 0) * C[0]));
#pragma unroll
      for (int j = 0; j < 3; ++j)
        CP[j] = fma(C[2], Pm(2, j), fma(C[1], Pm(1, j), C[0] * Pm(0, j)));
      const double S0 = fma(CP[2], C[2], fma(CP[1], C[1], CP[0] * C[0]));
      const double denom = S0 + gamma * Rk;  // :124 (+ Gsp + Gvp = 0)
#pragma unroll
      for (int i = 0; i < M; ++i) K[i] = PCt[i] / denom;
      double Mx[M][3];  // I - K*C, columns 0..2 (columns 3.. are identity)
#pragma unroll
      for (int i = 0; i < M; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) Mx[i][j] = ((i == j) ? 1.0 : 0.0) - K[i] * C[j];
      Mat<M, false> MP;
#pragma unroll
      for (int i = 0; i < M; ++i)
#pragma unroll
        for (int j = 0; j < M; ++j) {
          double acc = fma(Mx[i][2], Pm(2, j), fma(Mx[i][1], Pm(1, j), Mx[i][0] * Pm(0, j)));
          if (i >= 3) acc = acc + Pm(i, j);
          MP.at(i, j) = acc;
        }
      if (LEG) {
#pragma unroll
        for (int i = 0; i < M; ++i)
#pragma unroll
          for (int j = 0; j < M; ++j) Pp.at(i, j) = MP(i, j) / gamma;  // legacy :64
      } else {
        // :127 Joseph form, :138 symmetrisation
#pragma unroll
        for (int i = 0; i < M; ++i)
#pragma unroll
          for (int j = i; j < M; ++j) {
            double mij = fma(MP(i, 2), Mx[j][2], fma(MP(i, 1), Mx[j][1], MP(i, 0) * Mx[j][0]));
            if (j >= 3) mij = mij + MP(i, j);
            double mji = fma(MP(j, 2), Mx[i][2], fma(MP(j, 1), Mx[i][1], MP(j, 0) * Mx[i][0]));
            if (i >= 3) mji = mji + MP(j, i);
            const double pij = (mij + (K[i] * Rk) * K[j]) / gamma;
            const double pji = (mji + (K[j] * Rk) * K[i]) / gamma;
            Pp.at(i, j) = (pij + pji) / 2.0;
          }
      }
#pragma unroll
      for (int i = 0; i < M; ++i) sp[i] = s[i] + K[i] * innov;  // :129
    } else {  // :131-134
      innov = 0.0;
#pragma unroll
      for (int i = 0; i < M; ++i) { K[i] = 0.0; sp[i] = s[i]; }
      if (LEG) {
        Pp = Pm;
      } else {
#pragma unroll
        for (int i = 0; i < M; ++i)
#pragma unroll
          for (int j = i; j < M; ++j) Pp.at(i, j) = (Pm(i, j) + Pm(j, i)) / 2.0;  // :138
      }
    }
    state_margins<MODEL>(prm, sp);  // :141

    // :155-157 state update + Jacobian at s(k|k) (one pass over the NPI inputs)
    double *uo = P.u_opt.p ? P.u_opt.p + tidx(P.u_opt, pos, 0, L, b) : nullptr;
    const InputPass ip = input_pass<MODEL, true, false>(
        prm, in.eps, (M == 6) ? sp[M - 1] : 0.0, in.u + (size_t)pos * in.u_ts, in.u_js, L, uo,
        (size_t)P.u_opt.stride, nullptr);
    double sn[M];
    state_eqs<MODEL>(prm, in.eps, sp, ip.dot, sn);
    Mat<M, false> A;
    state_jacobian<MODEL>(prm, in.eps, sp, ip.a25, A);
    Mat<M, false> AP;
    mul_A_P<M, SYM>(A, Pp, AP);
    // :158 P(k+1|k) = A P A' + Q, :161 symmetrisation
    if (LEG) {
#pragma unroll
      for (int i = 0; i < M; ++i)
#pragma unroll
        for (int j = 0; j < M; ++j)
          Pm.at(i, j) = mul_X_At_ij<M, false>(AP, A, i, j) + q_elem(in.Q, P.q_mode, M, k, i, j);
    } else {
#pragma unroll
      for (int i = 0; i < M; ++i)
#pragma unroll
        for (int j = i; j < M; ++j) {
          const double pij = mul_X_At_ij<M, false>(AP, A, i, j) + q_elem(in.Q, P.q_mode, M, k, i, j);
          const double pji = mul_X_At_ij<M, false>(AP, A, j, i) + q_elem(in.Q, P.q_mode, M, k, j, i);
          Pm.at(i, j) = (pij + pji) / 2.0;
        }
    }
    state_margins<MODEL>(prm, sn);  // :164
#pragma unroll
    for (int i = 0; i < M; ++i) s[i] = sn[i];

    // :167-169
#pragma unroll
    for (int i = 0; i < M; ++i) P.S_PLUS.p[tidx(P.S_PLUS, pos, i, M, b)] = sp[i];
    store_tape<M, SYM>(Pp, P.P_PLUS, pos, b, P.tape_packed != 0);
    if (P.K_GAIN.p) {
#pragma unroll
      for (int i = 0; i < M; ++i) P.K_GAIN.p[tidx(P.K_GAIN, pos, i, M, b)] = K[i];
    }
    if (P.innov.p) P.innov.p[tidx(P.innov, pos, 0, 1, b)] = innov;

    if (MONITOR) {
      // :172-185 innovation whiteness monitor.  The three W-long windows live in
      // shared memory as ring buffers; sums run newest -> oldest over all W
      // slots (leading zeros included), as the reference's cat() windows do.
      const int cnt = (k + 1 < W) ? (k + 1) : W;
      head = (head == 0) ? (W - 1) : (head - 1);
      double *wm = win + threadIdx.x;
      const size_t bs = blockDim.x;
      wm[(size_t)(0 * W + head) * bs] = innov;
      double sm = 0.0;
      for (int j = 0, q = head; j < W; ++j) { sm += wm[(size_t)(0 * W + q) * bs]; q = (q + 1 == W) ? 0 : q + 1; }
      const double mu = sm / (double)cnt;
      const double cc = (innov - mu) * (innov - mu);
      wm[(size_t)(1 * W + head) * bs] = cc;
      wm[(size_t)(2 * W + head) * bs] = LEG ? (cc / Rk) : (cc / (Rk + kEps));  // :178 / legacy :108
      double sn_ = 0.0;
      for (int j = 0, q = head; j < W; ++j) { sn_ += wm[(size_t)(2 * W + q) * bs]; q = (q + 1 == W) ? 0 : q + 1; }
      if (P.rho.p) P.rho.p[tidx(P.rho, k, 0, 1, b)] = sn_ / (double)cnt;  // rho is NOT time-flipped
      const bool adapt = LEG ? (beta != 1.0 && valid)
                             : (beta != 1.0 && valid && P.fixed_R && (k + 1 < T));  // :180 / :110
      if (adapt) {
        double sc = 0.0;
        for (int j = 0, q = head; j < W; ++j) { sc += wm[(size_t)(1 * W + q) * bs]; q = (q + 1 == W) ? 0 : q + 1; }
        if (LEG) R_over = beta * Rk + ((1.0 - beta) * sc) / (double)cnt;   // legacy :111
        else     R_over = beta * Rk + (1.0 - beta) * (sc / (double)cnt);   // :182-184
        has_over = true;
      } else if (LEG) {
        R_over = Rk;
      }
    } else if (LEG) {
      R_over = Rk;
    }
  }
}

// ===========================================================================
// smoother gain pass: one thread per (day k, trajectory b)
// ===========================================================================
template <int MODEL>
__global__ void __launch_bounds__(128) eks_gain_kernel(const __grid_constant__ EkfParams P) {
  constexpr int M = model_dim(MODEL);
  constexpr bool LEG = model_legacy(MODEL);
  constexpr bool SYM = !LEG;
  constexpr bool REV = model_flipped(MODEL);
  constexpr int MM = M * M;
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)(P.T - 1) * P.B;
  if (tid >= total) return;
  const int k = (int)(tid / P.B);
  const int b = (int)(tid % P.B);
  const int T = P.T, L = P.L;
  const int pos = REV ? (T - 1 - k) : k;
  const int posn = REV ? (T - 2 - k) : (k + 1);
  const TrajIn in = traj_inputs(P, b, M);
  const epi_model_params *__restrict__ prm = in.prm;

  double sp[M];
#pragma unroll
  for (int i = 0; i < M; ++i) sp[i] = P.S_PLUS.p[tidx(P.S_PLUS, pos, i, M, b)];
  const InputPass ip = input_pass<MODEL, true, false>(prm, in.eps, (M == 6) ? sp[M - 1] : 0.0,
                                                      in.u + (size_t)pos * in.u_ts, in.u_js, L,
                                                      nullptr, 0, nullptr);
  Mat<M, false> A;
  state_jacobian<MODEL>(prm, in.eps, sp, ip.a25, A);  // :206

  Mat<M, false> Jm;
  int rank = M;
  bool bad = false;
  if (!LEG) {
    Mat<M, true> Pn, X;
    load_tape<M, true>(Pn, P.P_MINUS, posn, b, P.tape_packed != 0);
#pragma unroll
    for (int q = 0; q < Mat<M, true>::N; ++q) bad |= !(fabs(Pn.v[q]) <= 1.79769313486231570815e308);  // :211
    if (bad) {
#pragma unroll
      for (int q = 0; q < MM; ++q) Jm.v[q] = 0.0;  // :213
    } else {
      rank = pinv_sym<M>(Pn, X);
      Mat<M, true> Pp;
      load_tape<M, true>(Pp, P.P_PLUS, pos, b, P.tape_packed != 0);
      Mat<M, false> PAt;
#pragma unroll
      for (int i = 0; i < M; ++i)
#pragma unroll
        for (int j = 0; j < M; ++j) PAt.at(i, j) = mul_X_At_ij<M, true>(Pp, A, i, j);
#pragma unroll
      for (int i = 0; i < M; ++i)
#pragma unroll
        for (int j = 0; j < M; ++j) {  // :215
          double acc = PAt(i, 0) * X(0, j);
#pragma unroll
          for (int l = 1; l < M; ++l) acc = fma(PAt(i, l), X(l, j), acc);
          Jm.at(i, j) = acc;
        }
    }
  } else {
    // legacy :132   J = (P+ A') / P-   via LU with partial pivoting of (P-)'
    Mat<M, false> Pp, lu, rhs;
    load_mat<M, false>(Pp, P.P_PLUS.p + tidx(P.P_PLUS, pos, 0, MM, b), (size_t)P.P_PLUS.stride);
    {
      Mat<M, false> Pn;
      load_mat<M, false>(Pn, P.P_MINUS.p + tidx(P.P_MINUS, posn, 0, MM, b), (size_t)P.P_MINUS.stride);
#pragma unroll
      for (int i = 0; i < M; ++i)
#pragma unroll
        for (int j = 0; j < M; ++j) lu.at(i, j) = Pn(j, i);
    }
#pragma unroll
    for (int i = 0; i < M; ++i)
#pragma unroll
      for (int j = 0; j < M; ++j) rhs.at(j, i) = mul_X_At_ij<M, false>(Pp, A, i, j);  // rhs = (P+ A')'
    lu_solve_inplace<M>(lu, rhs);
#pragma unroll
    for (int i = 0; i < M; ++i)
#pragma unroll
      for (int j = 0; j < M; ++j) Jm.at(i, j) = rhs(j, i);
  }
  store_mat<M, false>(Jm, P.J.p + tidx(P.J, k, 0, MM, b), (size_t)P.J.stride);
  if (P.status && (bad || rank < M)) {
    // max over days of ((M - rank) << 8 | guard-hit); 0 == clean, full rank everywhere
    atomicMax(P.status + b, ((M - rank) << 8) | (bad ? 1 : 0));
  }
}

// ===========================================================================
// backward pass: one thread per trajectory
// ===========================================================================
template <int MODEL, bool WANT_P>
__global__ void __launch_bounds__(64) eks_backward_kernel(const __grid_constant__ EkfParams P) {
  constexpr int M = model_dim(MODEL);
  constexpr bool LEG = model_legacy(MODEL);
  constexpr bool SYM = !LEG;
  constexpr bool REV = model_flipped(MODEL);
  constexpr int MM = M * M;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= P.B) return;
  const int T = P.T, L = P.L;
  const TrajIn in = traj_inputs(P, b, M);
  const epi_model_params *__restrict__ prm = in.prm;
  const long long g = (P.b0 + b) / P.G;
  const bool want_cost = P.cost_day.p != nullptr;
  const double *wts = want_cost ? P.weights + (size_t)g * T * L : nullptr;

  // :189-202 terminal conditions
  const int posT = REV ? 0 : (T - 1);
  double ss[M];
  Mat<M, false> Ps;
#pragma unroll
  for (int i = 0; i < M; ++i) ss[i] = P.S_PLUS.p[tidx(P.S_PLUS, posT, i, M, b)];
  if (WANT_P) load_tape<M, false>(Ps, P.P_PLUS, posT, b, SYM && P.tape_packed != 0);
  {
    double sf[M];
    Mat<M, false> Pf;
    if (P.init_per_traj) {
#pragma unroll
      for (int i = 0; i < M; ++i) sf[i] = P.s_final_t.p[cidx(P.s_final_t, 0, i, M, b)];
      if (WANT_P) load_mat<M, false>(Pf, P.Ps_final_t.p + P.Ps_final_t.off + b, (size_t)P.Ps_final_t.stride);
    } else {
#pragma unroll
      for (int i = 0; i < M; ++i) sf[i] = P.s_final_g[g * M + i];
      if (WANT_P) load_mat<M, false>(Pf, P.Ps_final_g + (size_t)g * MM, 1);
    }
#pragma unroll
    for (int i = 0; i < M; ++i)
      if (!(sf[i] != sf[i])) ss[i] = sf[i];
    if (WANT_P) {
      if (!LEG) {
#pragma unroll
        for (int q = 0; q < MM; ++q)
          if (!(Pf.v[q] != Pf.v[q])) Ps.v[q] = Pf.v[q];  // :198-202 element-wise
      } else {
        // legacy :125-127  P_SMOOTH(row, col, T) = Ps_final(row, col) (sub-matrix)
        bool rowset[M], colset[M];
#pragma unroll
        for (int i = 0; i < M; ++i) { rowset[i] = false; colset[i] = false; }
#pragma unroll
        for (int i = 0; i < M; ++i)
#pragma unroll
          for (int j = 0; j < M; ++j)
            if (!(Pf(i, j) != Pf(i, j))) { rowset[i] = true; colset[j] = true; }
#pragma unroll
        for (int i = 0; i < M; ++i)
#pragma unroll
          for (int j = 0; j < M; ++j)
            if (rowset[i] && colset[j]) Ps.at(i, j) = Pf(i, j);
      }
    }
  }
  if (P.S_SMOOTH.p) {
#pragma unroll
    for (int i = 0; i < M; ++i) P.S_SMOOTH.p[tidx(P.S_SMOOTH, posT, i, M, b)] = ss[i];
  }
  if (WANT_P && P.P_SMOOTH.p)
    store_mat<M, false>(Ps, P.P_SMOOTH.p + tidx(P.P_SMOOTH, posT, 0, MM, b), (size_t)P.P_SMOOTH.stride);
  if (!LEG) {
    // u_opt_smooth(:, T) is never written by the reference => zeros (:95,:204)
    double *uo = nullptr;
    size_t uo_s = 0;
    if (P.u_opt_smooth.p) { uo = P.u_opt_smooth.p + tidx(P.u_opt_smooth, posT, 0, L, b); uo_s = (size_t)P.u_opt_smooth.stride; }
    else if (P.u_fore.p && posT >= P.T_hist) { uo = P.u_fore.p + tidx(P.u_fore, posT - P.T_hist, 0, L, b); uo_s = (size_t)P.u_fore.stride; }
    if (uo || P.dot_day.p) {
      const InputPass z = want_cost
          ? input_pass<MODEL, false, true>(prm, in.eps, 0.0, kZeroInputs, 1, L, uo, uo_s, wts + (size_t)posT * L)
          : input_pass<MODEL, false, false>(prm, in.eps, 0.0, kZeroInputs, 1, L, uo, uo_s, nullptr);
      if (P.dot_day.p) P.dot_day.p[tidx(P.dot_day, posT, 0, 1, b)] = z.dot;
      if (want_cost) P.cost_day.p[tidx(P.cost_day, posT, 0, 1, b)] = z.cost;
    }
  }

  for (int k = T - 2; k >= 0; --k) {  // :204
    const int pos = REV ? (T - 1 - k) : k;
    const int posn = REV ? (T - 2 - k) : (k + 1);
    Mat<M, false> Jm;
    load_mat<M, false>(Jm, P.J.p + tidx(P.J, k, 0, MM, b), (size_t)P.J.stride);
    double ds[M], sk[M];
#pragma unroll
    for (int l = 0; l < M; ++l) ds[l] = ss[l] - P.S_MINUS.p[tidx(P.S_MINUS, posn, l, M, b)];
#pragma unroll
    for (int i = 0; i < M; ++i) {
      double acc = Jm(i, 0) * ds[0];
#pragma unroll
      for (int l = 1; l < M; ++l) acc = fma(Jm(i, l), ds[l], acc);
      sk[i] = P.S_PLUS.p[tidx(P.S_PLUS, pos, i, M, b)] + acc;  // :218
    }
    state_margins<MODEL>(prm, sk);  // :221
    if (WANT_P) {
      Mat<M, SYM> Pp, Pn;
      load_tape<M, SYM>(Pp, P.P_PLUS, pos, b, SYM && P.tape_packed != 0);
      load_tape<M, SYM>(Pn, P.P_MINUS, posn, b, SYM && P.tape_packed != 0);
      Mat<M, false> D, JD;
#pragma unroll
      for (int i = 0; i < M; ++i)
#pragma unroll
        for (int j = 0; j < M; ++j) D.at(i, j) = Pn(i, j) - Ps(i, j);
#pragma unroll
      for (int i = 0; i < M; ++i)
#pragma unroll
        for (int j = 0; j < M; ++j) {
          double acc = Jm(i, 0) * D(0, j);
#pragma unroll
          for (int l = 1; l < M; ++l) acc = fma(Jm(i, l), D(l, j), acc);
          JD.at(i, j) = acc;
        }
      Mat<M, false> Pnew;
#pragma unroll
      for (int i = 0; i < M; ++i)
#pragma unroll
        for (int j = 0; j < M; ++j) {
          double acc = JD(i, 0) * Jm(j, 0);
#pragma unroll
          for (int l = 1; l < M; ++l) acc = fma(JD(i, l), Jm(j, l), acc);
          Pnew.at(i, j) = Pp(i, j) - acc;  // :223
        }
      if (LEG) {
        Ps = Pnew;
      } else {
#pragma unroll
        for (int i = 0; i < M; ++i)
#pragma unroll
          for (int j = 0; j < M; ++j) Ps.at(i, j) = (Pnew(i, j) + Pnew(j, i)) / 2.0;  // :226
      }
      if (P.P_SMOOTH.p)
        store_mat<M, false>(Ps, P.P_SMOOTH.p + tidx(P.P_SMOOTH, pos, 0, MM, b), (size_t)P.P_SMOOTH.stride);
    }
#pragma unroll
    for (int i = 0; i < M; ++i) ss[i] = sk[i];
    if (P.S_SMOOTH.p) {
#pragma unroll
      for (int i = 0; i < M; ++i) P.S_SMOOTH.p[tidx(P.S_SMOOTH, pos, i, M, b)] = ss[i];
    }
    if (!LEG) {
      // :229 re-run the state equation's input stage on the smoothed state
      double *uo = nullptr;
      size_t uo_s = 0;
      if (P.u_opt_smooth.p) { uo = P.u_opt_smooth.p + tidx(P.u_opt_smooth, pos, 0, L, b); uo_s = (size_t)P.u_opt_smooth.stride; }
      else if (P.u_fore.p && pos >= P.T_hist) { uo = P.u_fore.p + tidx(P.u_fore, pos - P.T_hist, 0, L, b); uo_s = (size_t)P.u_fore.stride; }
      if (uo || P.dot_day.p) {
        const double s5 = (M == 6) ? ss[M - 1] : 0.0;
        const InputPass z = want_cost
            ? input_pass<MODEL, false, true>(prm, in.eps, s5, in.u + (size_t)pos * in.u_ts, in.u_js, L, uo, uo_s, wts + (size_t)pos * L)
            : input_pass<MODEL, false, false>(prm, in.eps, s5, in.u + (size_t)pos * in.u_ts, in.u_js, L, uo, uo_s, nullptr);
        if (P.dot_day.p) P.dot_day.p[tidx(P.dot_day, pos, 0, 1, b)] = z.dot;
        if (want_cost) P.cost_day.p[tidx(P.cost_day, pos, 0, 1, b)] = z.cost;
      }
    }
  }
  if (WANT_P && P.P_first.p) {
    store_mat<M, false>(Ps, P.P_first.p + tidx(P.P_first, 0, 0, MM, b), (size_t)P.P_first.stride);
  }
}

// ===========================================================================
// launchers
// ===========================================================================
template <int MODEL>
static void launch_fwd_model(const EkfParams &p, cudaStream_t st, bool monitor) {
  const int block = (model_dim(MODEL) == 6) ? 32 : 64;
  const int grid = (p.B + block - 1) / block;
  if (monitor) {
    const size_t smem = (size_t)3 * p.W * block * sizeof(double);
    cudaFuncSetAttribute(ekf_forward_kernel<MODEL, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)smem);
    ekf_forward_kernel<MODEL, true><<<grid, block, smem, st>>>(p);
  } else {
    ekf_forward_kernel<MODEL, false><<<grid, block, 0, st>>>(p);
  }
}
template <int MODEL>
static void launch_gain_model(const EkfParams &p, cudaStream_t st) {
  if (p.T < 2) return;
  const size_t total = (size_t)(p.T - 1) * p.B;
  const int block = 128;
  eks_gain_kernel<MODEL><<<(unsigned)((total + block - 1) / block), block, 0, st>>>(p);
}
template <int MODEL>
static void launch_bwd_model(const EkfParams &p, cudaStream_t st, bool want_p) {
  const int block = (model_dim(MODEL) == 6) ? 32 : 64;
  const int grid = (p.B + block - 1) / block;
  if (want_p) eks_backward_kernel<MODEL, true><<<grid, block, 0, st>>>(p);
  else        eks_backward_kernel<MODEL, false><<<grid, block, 0, st>>>(p);
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

void launch_ekf_forward(const EkfParams &p, cudaStream_t st) {
  // the monitor is dead code unless rho is wanted or R adapts (beta != 1)
  const bool monitor = (p.rho.p != nullptr) || (p.beta != 1.0);
#define CALL(MDL) launch_fwd_model<MDL>(p, st, monitor)
  EPI_DISPATCH_MODEL(p.model, CALL)
#undef CALL
}
void launch_eks_gain(const EkfParams &p, cudaStream_t st) {
#define CALL(MDL) launch_gain_model<MDL>(p, st)
  EPI_DISPATCH_MODEL(p.model, CALL)
#undef CALL
}
void launch_eks_backward(const EkfParams &p, cudaStream_t st) {
  const bool want_p = (p.P_SMOOTH.p != nullptr) || (p.P_first.p != nullptr);
#define CALL(MDL) launch_bwd_model<MDL>(p, st, want_p)
  EPI_DISPATCH_MODEL(p.model, CALL)
#undef CALL
}

}  // namespace epi
