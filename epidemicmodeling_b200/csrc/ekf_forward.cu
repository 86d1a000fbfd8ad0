// ekf_forward.cu -- EKF forward pass for batches of independent trajectories
// (sm_100a, FP64, --fmad=false) + the per-group/per-day input pre-pass.
//
// Reference behaviour: Tools/GenericExtendedKalmanFilter.m:98-186 (generic models) and
// Tools/NewCaseEKFEstimatorWithOptimalNPI.m:40-114 (legacy models).
//
// One thread per trajectory, strictly sequential in time.  State, covariance (packed
// symmetric for the generic models), gain and Jacobian stay in registers; loop-invariant
// model scalars are read once; the per-day tape (S_MINUS, S_PLUS, P_MINUS, P_PLUS) is
// written so that every store instruction of a warp is one contiguous 256-byte row
// (tiled scratch: all offsets are instruction immediates).
#include <type_traits>

#include <cstdlib>
#include "ekf_common.cuh"
#ifdef EPI_FWD_NO_STORE  // experiment: cost of the tape stores (results are garbage)
#define EPI_FWD_STORE_COND (pos >= k0 && s[0] == 123.456)
#else
#define EPI_FWD_STORE_COND (pos >= k0)
#endif

namespace epi {

// ===========================================================================
// per-(group, day) input pre-pass
// ===========================================================================
// On a day whose NPI inputs are all given (no NaN to optimise), the input term
// gamma*a'*(u_max - u) (SIAlphaModelEKF.m:46 / ...OptControlled.m:67), A(3,6) = 0 and the
// day's weighted cost sum_j w*u do not depend on the trajectory: evaluate them once per
// (group, day) with the SAME operation sequence the per-trajectory code uses.  A NaN result
// means "evaluate per trajectory" (a NaN input, or NaN parameters -- the per-trajectory
// path then reproduces the same NaN), so the shortcut can never change a result.
__global__ void __launch_bounds__(128) group_day_kernel(const epi_model_params *__restrict__ prm,
                                                        const double *__restrict__ u,
                                                        const double *__restrict__ weights, int n_groups,
                                                        int T, int L, double *__restrict__ dot_grp,
                                                        double *__restrict__ cost_grp) {
  const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= (size_t)n_groups * T) return;
  const int g = (int)(q / T);
  const epi_model_params *__restrict__ p = prm + g;
  const double *__restrict__ ud = u + q * L;
  const double gamma = p->gamma;
  double dot = 0.0, cost = 0.0;
  bool has_nan = false;
  for (int j = 0; j < L; ++j) {
    const double uj = ud[j];
    has_nan |= (uj != uj);
    const double gj = gamma * p->a[j];
    const double d = p->u_max[j] - uj;
    dot = (j == 0) ? gj * d : fma(gj, d, dot);
    if (weights) {
      const double wu = weights[q * L + j] * uj;
      cost = (j == 0) ? wu : (cost + wu);
    }
  }
  const double nan = __longlong_as_double(0x7ff8000000000000ll);
  dot_grp[q] = has_nan ? nan : dot;
  if (cost_grp) cost_grp[q] = has_nan ? nan : cost;
}
void launch_group_day(const epi_model_params *prm, const double *u, const double *weights, int n_groups,
                      int T, int L, double *dot_grp, double *cost_grp, cudaStream_t st) {
  const size_t total = (size_t)n_groups * T;
  if (!total) return;
  group_day_kernel<<<(unsigned)((total + 127) / 128), 128, 0, st>>>(prm, u, weights, n_groups, T, L, dot_grp,
                                                                     cost_grp);
}

// ===========================================================================
// forward pass
// ===========================================================================
// Days [k_begin, k_end) of trajectory b.  k_begin == 0 starts from the initial conditions, a later
// start resumes from the a-priori page (S_MINUS, P_MINUS) of day k_begin that the previous segment
// left on the tape -- without the innovation monitor that page IS the whole filter state.
// PLAIN = the common call shape fixed at compile time: constant Q (EPI_Q_CONST) and none of the optional
// per-day outputs (u_opt, K_GAIN, innov).  The run-time dispatch on these inside the day loop cost 8 % of the
// loop's instructions and 10 % of the pass (4.08 -> 3.65 ms on the sweep): the loop body is larger than the
// instruction cache can hold for two warps per scheduler.
template <int MODEL, bool MONITOR, bool TILED, bool PLAIN = false>
EPI_DI void forward_days(const EkfParams &P, const int b, const int k_begin, const int k_end, double *win) {
  constexpr int M = model_dim(MODEL);
  constexpr bool LEG = model_legacy(MODEL);
  constexpr bool SYM = !LEG;
  constexpr bool REV = model_flipped(MODEL);
  constexpr int MM = M * M;
  constexpr int PF = (SYM && TILED) ? M * (M + 1) / 2 : MM;
  using TmpMat = Mat<M, false>;  // (a shared-memory backing, SMat, was tried: no register relief)

  if (b >= P.B) return;
  const TrajIn in = traj_inputs(P, b, M);
  const ModelConsts mc = load_consts(in.prm);
  const int T = P.T, L = P.L, W = P.W;
  const int q_mode = PLAIN ? (int)EPI_Q_CONST : P.q_mode;
  const double gamma = P.gamma, beta = P.beta, v_bar = P.v_bar, eps = in.eps;
  const InvDiv by_gamma = make_invdiv(gamma);

  const int k0 = P.k0;  // lean sweeps: days before k0 are never read by the smoother
  const Tape<TILED> tSm = make_tape<TILED>(P.S_MINUS, M, T - k0, b, k0), tSp = make_tape<TILED>(P.S_PLUS, M, T - k0, b, k0);
  const Tape<TILED> tPm = make_tape<TILED>(P.P_MINUS, PF, T - k0, b, k0), tPp = make_tape<TILED>(P.P_PLUS, PF, T - k0, b, k0);

  double s[M];
  Mat<M, SYM> Pm;  // P(k|k-1)
  if (k_begin > 0) {
    // resume: written by another CTA of this launch (possibly on another SM) => bypass L1
    const int pos = REV ? (T - 1 - k_begin) : k_begin;
    const double *__restrict__ d = tSm.at_day(pos);
#pragma unroll
    for (int i = 0; i < M; ++i) s[i] = __ldcg(d + tSm.f(i));
    const double *__restrict__ dP = tPm.at_day(pos);
#pragma unroll
    for (int i = 0; i < M; ++i)
#pragma unroll
      for (int j = 0; j < M; ++j)
        if (!SYM || j >= i) Pm.at(i, j) = __ldcg(dP + tPm.f((SYM && TILED) ? Mat<M, true>::idx(i, j) : (j * M + i)));
  } else if (P.init_per_traj) {
    const double *si = P.s_init_t.p + P.s_init_t.off + b;
#pragma unroll
    for (int i = 0; i < M; ++i) s[i] = si[(size_t)i * P.s_init_t.stride];
    load_mat<M, SYM>(Pm, P.Ps_init_t.p + P.Ps_init_t.off + b, (size_t)P.Ps_init_t.stride);
  } else {
#pragma unroll
    for (int i = 0; i < M; ++i) s[i] = P.s_init_g[in.g * M + i];
    load_mat<M, SYM>(Pm, P.Ps_init_g + (size_t)in.g * MM, 1);
  }

  if (MONITOR) {
    for (int j = 0; j < 3 * W; ++j) win[(size_t)j * blockDim.x + threadIdx.x] = 0.0;
  }
  int head = 0;               // ring position of the newest window slot
  double R_over = 0.0;        // adapted R for the next step (:184)
  bool has_over = false;

  // the three per-day scalars (observation, R, per-group input term) are fetched one day ahead:
  // they come from L2/HBM (each is read once) and were the exposed latency of every step
  auto day_x = [&](int kk) { return __ldg(in.x + (size_t)(REV ? (T - 1 - kk) : kk) * in.x_ts); };
  auto day_R = [&](int kk) { return (P.r_mode == EPI_R_CONST) ? in.R_const : __ldg(in.R + (size_t)kk * in.R_ts); };
  auto day_pre = [&](int kk) {
    return in.dot_grp ? __ldg(in.dot_grp + (REV ? (T - 1 - kk) : kk)) : __longlong_as_double(0x7ff8000000000000ll);
  };
  double x_nxt = 0.0, R_nxt = 0.0, pre_nxt = 0.0;
  if (k_begin < k_end) { x_nxt = day_x(k_begin); R_nxt = LEG ? 0.0 : day_R(k_begin); pre_nxt = day_pre(k_begin); }
  const bool day_sync = P.day_sync != 0;
  for (int k = k_begin; k < k_end; ++k) {
    if (day_sync) __syncthreads();  // piped schedule: the four warps of a CTA walk the loop body in step
    const int pos = REV ? (T - 1 - k) : k;
    const double x_cur = x_nxt, R_cur = R_nxt, pre_cur = pre_nxt;
    if (k + 1 < k_end) { x_nxt = day_x(k + 1); R_nxt = LEG ? 0.0 : day_R(k + 1); pre_nxt = day_pre(k + 1); }
    // :100-101 store the a-priori estimate
    if (EPI_FWD_STORE_COND) {
      double *__restrict__ d = tSm.at_day(pos);
#pragma unroll
      for (int i = 0; i < M; ++i) d[tSm.f(i)] = s[i];
      tape_store_cov<M, SYM, TILED>(Pm, tPm, tPm.at_day(pos));
    }

    double Rk;
    if (LEG) {
      Rk = (k == 0) ? in.R_const : R_over;  // scalar R adapted in place (:31,:111)
    } else {
      const double base = R_cur;
      Rk = has_over ? R_over : base;
      has_over = false;
    }
    double C[3];
    const double xhat = obs_model<MODEL>(mc.obs_type, s, v_bar, C);  // :115-119
    const double xk = x_cur;
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
      const InvDiv by_denom = make_invdiv(denom);
      {
        ExpRange rng;
#pragma unroll
        for (int i = 0; i < M; ++i) K[i] = div_fast(PCt[i], by_denom, rng);
        if (!(by_denom.ok && rng.safe())) {
#pragma unroll
          for (int i = 0; i < M; ++i) K[i] = slow_div(PCt[i], denom);  // (out of line: cold, and the loop body must stay small)
        }
      }
      double Mx[M][3];  // I - K*C, columns 0..2 (columns 3.. are identity)
#pragma unroll
      for (int i = 0; i < M; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) Mx[i][j] = ((i == j) ? 1.0 : 0.0) - K[i] * C[j];
      TmpMat MP;
#pragma unroll
      for (int i = 0; i < M; ++i)
#pragma unroll
        for (int j = 0; j < M; ++j) {
          double acc = fma(Mx[i][2], Pm(2, j), fma(Mx[i][1], Pm(1, j), Mx[i][0] * Pm(0, j)));
          if (i >= 3) acc = acc + Pm(i, j);
          MP.at(i, j) = acc;
        }
      // the m^2 quotients by gamma: fast exact-division path for the whole block, true divisions
      // only if some numerator left the safe exponent range (checked once, after the block)
      auto covariance_update = [&](auto exact_tag) {
        constexpr bool EXACT = decltype(exact_tag)::value;
        ExpRange rng;
        if (LEG) {
#pragma unroll
          for (int i = 0; i < M; ++i)
#pragma unroll
            for (int j = 0; j < M; ++j)
              Pp.at(i, j) = EXACT ? slow_div(MP(i, j), gamma) : div_fast(MP(i, j), by_gamma, rng);  // legacy :64
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
              const double nij = mij + (K[i] * Rk) * K[j];
              const double nji = mji + (K[j] * Rk) * K[i];
              const double pij = EXACT ? slow_div(nij, gamma) : div_fast(nij, by_gamma, rng);
              const double pji = EXACT ? slow_div(nji, gamma) : div_fast(nji, by_gamma, rng);
              Pp.at(i, j) = (pij + pji) / 2.0;
            }
        }
        return by_gamma.ok && rng.safe();
      };
      if (!covariance_update(std::false_type{})) covariance_update(std::true_type{});
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
    state_margins<MODEL>(mc, sp);  // :141

    // :155-157 state update + Jacobian at s(k|k) (one pass over the NPI inputs, or the
    // per-group value when the day has no input to optimise)
    double *uo = (!PLAIN && P.u_opt.p) ? P.u_opt.p + (size_t)P.u_opt.off + b + (size_t)pos * L * P.u_opt.stride : nullptr;
    const double *ud = in.u + (size_t)pos * in.u_ts;
    double dotv, a25 = 0.0;
    const double pre = pre_cur;
    if (pre == pre) {
      dotv = pre;
      if (uo) {
#pragma unroll
        for (int j = 0; j < EPI_LMAX; ++j)
          if (j < L) uo[(size_t)j * P.u_opt.stride] = ud[(size_t)j * in.u_js];
      }
    } else {
      const InputPass ip = input_pass<MODEL, true, false>(mc, eps, (M == 6) ? sp[M - 1] : 0.0, ud, in.u_js, L, uo,
                                                          (size_t)P.u_opt.stride, nullptr);
      dotv = ip.dot;
      a25 = ip.a25;
    }
    double sn[M];
    state_eqs<MODEL>(mc, eps, sp, dotv, sn);
    Mat<M, false> A;
    state_jacobian<MODEL>(mc, eps, sp, a25, A);
    TmpMat AP;
    mul_A_P<M, SYM>(A, Pp, AP);
    // :158 P(k+1|k) = A P A' + Q, :161 symmetrisation
    if (LEG) {
#pragma unroll
      for (int i = 0; i < M; ++i)
#pragma unroll
        for (int j = 0; j < M; ++j)
          Pm.at(i, j) = mul_X_At_ij<M>(AP, A, i, j) + q_elem(in.Q, q_mode, M, k, i, j);
    } else {
#pragma unroll
      for (int i = 0; i < M; ++i)
#pragma unroll
        for (int j = i; j < M; ++j) {
          const double pij = mul_X_At_ij<M>(AP, A, i, j) + q_elem(in.Q, q_mode, M, k, i, j);
          const double pji = mul_X_At_ij<M>(AP, A, j, i) + q_elem(in.Q, q_mode, M, k, j, i);
          Pm.at(i, j) = (pij + pji) / 2.0;
        }
    }
    state_margins<MODEL>(mc, sn);  // :164
#pragma unroll
    for (int i = 0; i < M; ++i) s[i] = sn[i];

    // :167-169
    if (EPI_FWD_STORE_COND) {
      double *__restrict__ d = tSp.at_day(pos);
#pragma unroll
      for (int i = 0; i < M; ++i) d[tSp.f(i)] = sp[i];
      tape_store_cov<M, SYM, TILED>(Pp, tPp, tPp.at_day(pos));
    }
    if (!PLAIN && P.K_GAIN.p) {
      double *d = P.K_GAIN.p + (size_t)P.K_GAIN.off + b + (size_t)pos * M * P.K_GAIN.stride;
#pragma unroll
      for (int i = 0; i < M; ++i) d[(size_t)i * P.K_GAIN.stride] = K[i];
    }
    if (!PLAIN && P.innov.p) P.innov.p[(size_t)pos * P.innov.stride + P.innov.off + b] = innov;

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
      if (P.rho.p) P.rho.p[(size_t)k * P.rho.stride + P.rho.off + b] = sn_ / (double)cnt;  // rho is NOT time-flipped
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
  if (k_end < T) {  // hand the a-priori page of day k_end to the segment that continues
    const int pos = REV ? (T - 1 - k_end) : k_end;
    double *__restrict__ d = tSm.at_day(pos);
#pragma unroll
    for (int i = 0; i < M; ++i) d[tSm.f(i)] = s[i];
    tape_store_cov<M, SYM, TILED>(Pm, tPm, tPm.at_day(pos));
  }
}

#ifndef EPI_FWD_MIN_BLOCKS6
#define EPI_FWD_MIN_BLOCKS6 8  // m = 6: 32-thread CTAs, 8 per SM = 255 registers (measured best)
#endif
template <int MODEL, bool MONITOR, bool TILED, bool PLAIN = false>
__global__ void __launch_bounds__((model_dim(MODEL) == 6) ? 32 : 64, (model_dim(MODEL) == 6) ? EPI_FWD_MIN_BLOCKS6 : 4)
ekf_forward_kernel(const __grid_constant__ EkfParams P) {
  extern __shared__ double win[];  // MONITOR: [3][W][blockDim.x]
  forward_days<MODEL, MONITOR, TILED, PLAIN>(P, blockIdx.x * blockDim.x + threadIdx.x, 0, P.T, win);
}
// the PLAIN call shape (see forward_days)
static bool forward_plain(const EkfParams &p) {
  return p.q_mode == EPI_Q_CONST && !p.u_opt.p && !p.K_GAIN.p && !p.innov.p;
}

// Persistent, time-segmented form for batches of only a few waves (the 236 x 250 sweep is
// 1844 warps on 1184 resident slots: 1.56 waves, i.e. the second wave leaves 44 % of the slots
// idle).  The T days of every 32-trajectory tile are cut into P.fwd_segments segments; CTAs (one
// warp each) draw (segment, tile) items segment-major from an atomic counter, so a finished slot
// continues with the next segment of some other tile instead of idling.  An item waits for its
// predecessor (same tile, previous segment) through a per-tile progress word; predecessors
// always have a lower item number, so they are running or done and the wait cannot deadlock.
// sync[0] = item counter, sync[1 + tile] = segments of that tile finished (zeroed by the host).
template <int MODEL, bool TILED, bool PLAIN = false>
__global__ void __launch_bounds__(32, EPI_FWD_MIN_BLOCKS6) ekf_forward_segmented_kernel(const __grid_constant__ EkfParams P) {
  const int S = P.fwd_segments, T = P.T, k0 = P.k0;
  const int n_tiles = (P.B + 31) / 32;
  const int lane = threadIdx.x;
  for (;;) {
    int item = 0;
    if (lane == 0) item = atomicAdd(P.fwd_sync, 1);
    item = __shfl_sync(0xffffffffu, item, 0);
    if (item >= n_tiles * S) return;
    const int seg = item / n_tiles, tile = item - seg * n_tiles;
    // boundaries at or after k0: every resume page is on the (possibly lean) tape
    const int kb = (seg == 0) ? 0 : k0 + (int)(((long long)(T - k0) * seg) / S);
    const int ke = (seg + 1 == S) ? T : k0 + (int)(((long long)(T - k0) * (seg + 1)) / S);
    if (seg > 0) {
      if (lane == 0) {
        while (atomicAdd(P.fwd_sync + 1 + tile, 0) < seg) __nanosleep(256);
      }
      __syncwarp();
      __threadfence();
    }
    if (ke > kb || seg == 0) forward_days<MODEL, false, TILED, PLAIN>(P, tile * 32 + lane, kb, ke, nullptr);
    __threadfence();
    __syncwarp();
    if (lane == 0) atomicExch(P.fwd_sync + 1 + tile, seg + 1);
  }
}

template <int MODEL, bool TILED>
static void launch_fwd_model(const EkfParams &p, cudaStream_t st, bool monitor) {
  const int block = (model_dim(MODEL) == 6) ? 32 : 64;
  const int grid = (p.B + block - 1) / block;
  if (monitor) {
    const size_t smem = (size_t)3 * p.W * block * sizeof(double);
    cudaFuncSetAttribute(ekf_forward_kernel<MODEL, true, TILED>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)smem);
    ekf_forward_kernel<MODEL, true, TILED><<<grid, block, smem, st>>>(p);
  } else if (TILED && forward_plain(p)) {
    ekf_forward_kernel<MODEL, false, TILED, TILED><<<grid, block, 0, st>>>(p);  // (PLAIN exists for TILED only)
  } else {
    ekf_forward_kernel<MODEL, false, TILED><<<grid, block, 0, st>>>(p);
  }
}

// Number of time segments for `tiles` one-warp CTAs on `slots` resident CTAs (1 = plain launch).
// The kernel saturates the SM from ~6 of its 8 warps on, so segmenting only trims the idle tail of the
// last wave, and every extra segment adds waits on predecessors.  Measured with the PLAIN kernel (B200,
// 1184 slots; forward ms for S = 1 / 2 / 3):  1422 tiles (1.20 waves) 2.85 / 3.03 / 3.29 - 1844 tiles
// (1.56, the 236 x 250 sweep) 3.52 / 3.04 / 3.83 - 2344 tiles (1.98) 3.66 / 3.51 / 3.53 - 2961 tiles (2.50)
// 5.00 / 4.63 / 4.47 - 4141 tiles (3.50) 6.20 / 6.36 / 6.09.
int forward_segments(long long tiles, int slots) {
  if (tiles <= slots || tiles > 6LL * slots || tiles % slots == 0) return 1;
  const double waves = (double)tiles / (double)slots;
  if (waves <= 1.3) return 1;
  if (waves <= 2.25) return 2;
  return 3;
}

template <int MODEL>
static bool launch_fwd_segmented(const EkfParams &p, cudaStream_t st) {
  if constexpr (model_dim(MODEL) == 6 && !model_legacy(MODEL)) {
    static int slots = 0;
    if (!slots) {
      int dev = 0, sms = 0, per_sm = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ekf_forward_segmented_kernel<MODEL, true>, 32, 0);
      slots = sms * (per_sm > 0 ? per_sm : 1);
    }
    const int tiles = (p.B + 31) / 32;
    int grid = tiles < slots ? tiles : slots;
    if (const char *e = getenv("EPI_FWD_GRID")) grid = atoi(e);  // tuning experiments
    if (forward_plain(p)) ekf_forward_segmented_kernel<MODEL, true, true><<<grid, 32, 0, st>>>(p);
    else ekf_forward_segmented_kernel<MODEL, true><<<grid, 32, 0, st>>>(p);
    return true;
  } else {
    return false;
  }
}

// ---- piped schedule (small batches of the sweep: a region shard of the strong-scaling form) ------------------------
// A shard of a few hundred tiles is latency-bound in this pass (one warp walks the T days of its tile at ~2 us a day
// whatever else the GPU does) and leaves most SMs idle, while the gains of day k only need the tape up to day k + 1.
// Here the pass runs as CTAs of 3 to 6 warps (one tile each) that ask for so much dynamic shared memory that no gain
// CTA fits beside them: a forward warp that shares its scheduler with gain warps takes 5 us a day instead of 2
// (measured).  The FP64 pipe is shared by the sub-partitions of an SM, so the pass itself slows down with
// the warps per SM (1.11 ms spread thin, 1.42 at 4 per SM, 1.8 at 6): pipe_warps picks the smallest CTA that leaves
// about half of the SMs to the gains.  After every time chunk a warp publishes its progress (tape stores, fence,
// one atomic per tile); the host has queued the gains of that chunk on a second stream behind a stream wait on the
// counter, so they fill the SMs the forward pass does not occupy.  Same forward_days, same bits.
constexpr int kPipeMaxWarps = 6;
constexpr size_t kPipeSmem = 200 * 1024;  // + one gain CTA's 48 KB exceeds an SM's 227 KB

void pipe_chunk_days(const EkfParams &p, int c, int &kb, int &ke) {
  const int S = p.pipe_chunks, T = p.T, k0 = p.k0;
  kb = (c == 0) ? 0 : k0 + (int)(((long long)(T - k0) * c) / S);
  ke = (c + 1 == S) ? T : k0 + (int)(((long long)(T - k0) * (c + 1)) / S);
}

template <int MODEL>
__global__ void __launch_bounds__(32 * kPipeMaxWarps, 1) ekf_forward_piped_kernel(const __grid_constant__ EkfParams P) {
  const int S = P.pipe_chunks, T = P.T, k0 = P.k0;
  const int tile = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (tile >= (P.B + 31) / 32) return;
  // lanes without a trajectory leave the KERNEL here (not just forward_days): the per-day CTA barrier counts the
  // threads that have not exited
  const unsigned mask = __ballot_sync(0xffffffffu, tile * 32 + lane < P.B);
  if (tile * 32 + lane >= P.B) return;
  for (int c = 0; c < S; ++c) {
    const int kb = (c == 0) ? 0 : k0 + (int)(((long long)(T - k0) * c) / S);
    const int ke = (c + 1 == S) ? T : k0 + (int)(((long long)(T - k0) * (c + 1)) / S);
    if (ke > kb || c == 0) forward_days<MODEL, false, true, true>(P, tile * 32 + lane, kb, ke, nullptr);
    __threadfence();
    __syncwarp(mask);
    if (lane == 0) atomicAdd(P.pipe_sync + c, 1u);
  }
}

// warps (tiles) per CTA: measured on 148 SMs, ms per step for 3 / 4 / 6 warps (one-stream schedule beside it):
// 235 tiles 2.45 / 2.48 / 2.62 (2.70), 461 tiles - / 3.62 / 3.53 (3.70).  0 = the batch is too large to pipe.
static int pipe_warps(int tiles, int n_sms) {
  if (const char *e = getenv("EPI_PIPE_WARPS")) {
    const int w = atoi(e);
    return (w >= 1 && w <= kPipeMaxWarps && (tiles + w - 1) / w <= n_sms - n_sms / 10) ? w : 0;
  }
  for (int w : {3, 4, 6})
    if ((tiles + w - 1) / w <= (n_sms * 11) / 20) return w;
  return 0;
}

bool forward_piped_ok(const EkfParams &p, int n_sms) {
  const bool monitor = (p.rho.p != nullptr) || (p.beta != 1.0);
  // the call shape of the sweep, enough days to cut, and about half of the SMs left for the gains
  return p.model == EPI_MODEL_OPTCTRL && p.tiled && !monitor && forward_plain(p) && p.T - p.k0 >= 64 && p.B > 0 &&
         pipe_warps((p.B + 31) / 32, n_sms) > 0;
}

void launch_ekf_forward_piped(const EkfParams &p, cudaStream_t st) {
  int dev = 0, n_sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, dev);
  const int tiles = (p.B + 31) / 32, kPipeWarps = pipe_warps(tiles, n_sms), ctas = (tiles + kPipeWarps - 1) / kPipeWarps;
  auto kern = ekf_forward_piped_kernel<EPI_MODEL_OPTCTRL>;
  size_t smem = kPipeSmem;
  if (const char *e = getenv("EPI_PIPE_SMEM_KB")) smem = (size_t)atoi(e) * 1024;  // tuning experiments
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  EkfParams q = p;
  q.day_sync = 0;  // a CTA barrier per day, so that the warps share their instruction lines: 1.50 against 1.42 ms
  if (const char *e = getenv("EPI_PIPE_DAYSYNC")) q.day_sync = atoi(e);
  kern<<<ctas, 32 * kPipeWarps, smem, st>>>(q);
}

int forward_resident_slots6() {
  int dev = 0, sms = 0, per_sm = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ekf_forward_segmented_kernel<EPI_MODEL_OPTCTRL, true>, 32, 0);
  return sms * (per_sm > 0 ? per_sm : 1);
}

void launch_ekf_forward(const EkfParams &p, cudaStream_t st) {
  // the monitor is dead code unless rho is wanted or R adapts (beta != 1)
  const bool monitor = (p.rho.p != nullptr) || (p.beta != 1.0);
  if (rows_forward_ok(p) && rows_wanted(p.B, true)) {   // small batch: six lanes per trajectory (csrc/ekf_rows.cu)
    launch_ekf_forward_rows(p, st);
    return;
  }
  if (pair_forward_wanted(p)) {   // small batch: two warps per tile (csrc/ekf_pair.cu)
    launch_ekf_forward_pair(p, st);
    return;
  }
  if (p.fwd_segments > 1 && p.fwd_sync && p.tiled && !monitor) {
    bool done = false;
#define CALL(MDL) done = launch_fwd_segmented<MDL>(p, st)
    EPI_DISPATCH_MODEL(p.model, CALL)
#undef CALL
    if (done) return;
  }
#define CALL(MDL)                                                  \
  if (p.tiled) launch_fwd_model<MDL, true>(p, st, monitor);        \
  else launch_fwd_model<MDL, false>(p, st, monitor)
  EPI_DISPATCH_MODEL(p.model, CALL)
#undef CALL
}

}  // namespace epi
