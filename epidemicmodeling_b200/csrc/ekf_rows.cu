// ekf_rows.cu -- lane-group forms of the two time-sequential passes (sm_100a, FP64, --fmad=false):
// SIX LANES PER TRAJECTORY, one covariance row per lane, for the 6-state state+costate model
// (EPI_MODEL_OPTCTRL) in the call shape of the fused sweep (tiled scratch tape, constant Q, no innovation
// monitor, none of the optional per-day outputs).
//
// Why: the forward EKF (GenericExtendedKalmanFilter.m:98-186) and the smoother recursion (:218-229) walk the
// T days of a trajectory one after the other.  With one thread per trajectory a day is a ~1800-instruction
// dependency chain, so a region-sharded sweep (TrainPredictPrescribeNPI.m:93; 7375 trajectories = 231 warps per GPU
// at 8 GPUs) cannot go below ~1.5 ms per pass however few trajectories a GPU holds.  Here lane r of a
// 6-lane group owns row r of P (and computes the scalar model callbacks redundantly): a day is ~4x
// shorter and there are 6.4x the warps to hide its latencies.  It costs ~1.6x the issue slots of the
// one-thread form, so the host picks it for SMALL batches only (launch_ekf_forward / launch_eks_backward).
//
// Bit-identical to ekf_forward.cu / eks_backward.cu: every matrix element is produced by ONE lane with the
// operation sequence of the one-thread kernels (first term a*b, then fma, index ascending; structural zeros
// skipped).  Where the one-thread code reads the packed symmetric P as P(l,j) a lane reads its own row as
// P(j,l) -- the same stored value.  Rows / columns move between the lanes of a group by shuffles (broadcasts)
// and through a per-warp shared-memory tile (transpositions).
#include <cstdlib>

#include "ekf_common.cuh"

namespace epi {

namespace {

constexpr int kRowsGroup = 6;   // lanes per trajectory
constexpr int kRowsPerWarp = 5; // trajectories per warp (lanes 30, 31 shadow group 4)
constexpr int kRowsBlock = 128; // 4 warps per CTA
#ifndef EPI_ROWS_MAX_BATCH
#define EPI_ROWS_MAX_BATCH 8192
#endif
constexpr long long kRowsMaxBatch = EPI_ROWS_MAX_BATCH;  // trajectories per launch up to which the lane-group kernels are used

struct RowLane {
  int b;        // trajectory (clamped to a valid one for shadow lanes)
  int row;      // covariance row owned by this lane
  int gbase;    // first lane of the group
  bool store;   // this lane's results go to the tape
  double *tile; // this group's 6 x 6 transposition tile in shared memory
};

EPI_DI RowLane row_lane(int B, double *smem_warp) {
  const int lane = threadIdx.x & 31;
  const int warp = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
  RowLane L;
  int grp = lane / kRowsGroup;
  int row = lane - grp * kRowsGroup;
  bool live = true;
  if (grp >= kRowsPerWarp) { grp = kRowsPerWarp - 1; live = false; }  // lanes 30, 31: rows 0, 1 of group 4 again
  int b = warp * kRowsPerWarp + grp;
  if (b >= B) { b = B - 1; live = false; }
  L.b = b; L.row = row; L.gbase = grp * kRowsGroup; L.store = live;
  L.tile = smem_warp + ((lane / kRowsGroup) < kRowsPerWarp ? grp : kRowsPerWarp) * 36;  // shadow lanes: a scratch tile
  return L;
}

// v of lane (group base + j) for every lane of the warp (all 32 lanes take part)
EPI_DI double from_row(double v, const RowLane &L, int j) { return __shfl_sync(0xffffffffu, v, L.gbase + j); }

// out[j] = in_of_lane_j[row]: transposition of the group's 6 x 6 value matrix (in[] = this lane's row)
EPI_DI void transpose6(const double (&in)[6], double (&out)[6], const RowLane &L) {
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 6; ++j) L.tile[L.row * 6 + j] = in[j];
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 6; ++j) out[j] = L.tile[j * 6 + L.row];
}

// offset (in fields) of packed entry (row, j), j >= row: Mat<6, true>::idx
EPI_DI int packed_row_base(int row) { return (row * (11 - row)) / 2; }

}  // namespace

// ---------------------------------------------------------------------------------------------------
// forward pass, days [0, T)
// ---------------------------------------------------------------------------------------------------
template <int MODEL>
__global__ void __launch_bounds__(kRowsBlock) ekf_forward_rows_kernel(const __grid_constant__ EkfParams P) {
  constexpr int M = 6;
  static_assert(model_dim(MODEL) == 6 && !model_legacy(MODEL) && !model_flipped(MODEL), "generic forward 6-state model");
  __shared__ double tiles[(kRowsBlock / 32) * (kRowsPerWarp + 1) * 36];
  const RowLane L_ = row_lane(P.B, tiles + (threadIdx.x >> 5) * (kRowsPerWarp + 1) * 36);
  const int b = L_.b, row = L_.row;
  const TrajIn in = traj_inputs(P, b, M);
  const ModelConsts mc = load_consts(in.prm);
  const int T = P.T, L = P.L, k0 = P.k0;
  const double gamma = P.gamma, v_bar = P.v_bar, eps = in.eps;
  const InvDiv by_gamma = make_invdiv(gamma);
  const Tape<true> tSm = make_tape<true>(P.S_MINUS, M, T - k0, b, k0), tSp = make_tape<true>(P.S_PLUS, M, T - k0, b, k0);
  const Tape<true> tPm = make_tape<true>(P.P_MINUS, 21, T - k0, b, k0), tPp = make_tape<true>(P.P_PLUS, 21, T - k0, b, k0);
  const int pbase = packed_row_base(row);

  double s[M], prow[M], qrow[M];
  if (P.init_per_traj) {
    const double *si = P.s_init_t.p + P.s_init_t.off + b;
#pragma unroll
    for (int i = 0; i < M; ++i) s[i] = si[(size_t)i * P.s_init_t.stride];
    const double *pi = P.Ps_init_t.p + P.Ps_init_t.off + b;
#pragma unroll
    for (int j = 0; j < M; ++j) {   // load_mat<M, true> keeps the upper triangle: P(i,j), i <= j, = field j*M + i
      const int lo = row < j ? row : j, hi = row < j ? j : row;
      prow[j] = pi[(size_t)(hi * M + lo) * P.Ps_init_t.stride];
    }
  } else {
#pragma unroll
    for (int i = 0; i < M; ++i) s[i] = P.s_init_g[in.g * M + i];
    const double *pi = P.Ps_init_g + (size_t)in.g * 36;
#pragma unroll
    for (int j = 0; j < M; ++j) {
      const int lo = row < j ? row : j, hi = row < j ? j : row;
      prow[j] = pi[hi * M + lo];
    }
  }
  double drow[3];   // row `row` of the identity, columns 0..2
#pragma unroll
  for (int j = 0; j < 3; ++j) drow[j] = (row == j) ? 1.0 : 0.0;
  // constant Q, row `row`: q_elem(i, j) = Q[j*M + i]
#pragma unroll
  for (int j = 0; j < M; ++j) qrow[j] = __ldg(in.Q + j * M + row);

  auto day_x = [&](int kk) { return __ldg(in.x + (size_t)kk * in.x_ts); };
  auto day_R = [&](int kk) { return (P.r_mode == EPI_R_CONST) ? in.R_const : __ldg(in.R + (size_t)kk * in.R_ts); };
  auto day_pre = [&](int kk) { return in.dot_grp ? __ldg(in.dot_grp + kk) : __longlong_as_double(0x7ff8000000000000ll); };
  double x_nxt = day_x(0), R_nxt = day_R(0), pre_nxt = day_pre(0);

#pragma unroll 1
  for (int k = 0; k < T; ++k) {
    const double x_cur = x_nxt, Rk = R_nxt, pre = pre_nxt;
    if (k + 1 < T) { x_nxt = day_x(k + 1); R_nxt = day_R(k + 1); pre_nxt = day_pre(k + 1); }
    // :100-101 the a-priori estimate (dynamic register index: select the lane's own state entry)
    double s_own = s[0];
#pragma unroll
    for (int i = 1; i < M; ++i) s_own = (row == i) ? s[i] : s_own;
    if (L_.store && k >= k0) {
      tSm.at_day(k)[tSm.f(row)] = s_own;
      double *__restrict__ d = tPm.at_day(k);
#pragma unroll
      for (int j = 0; j < M; ++j)
        if (j >= row) d[tPm.f(pbase + j)] = prow[j];
    }

    double C[3];
    const double xhat = obs_model<MODEL>(mc.obs_type, s, v_bar, C);  // :115-119
    const bool valid = !(x_cur != x_cur);                               // :122
    const double innov = x_cur - xhat;                                  // :123

    // (:131-134) a day without observation keeps the a-priori estimate; days without any observation in the
    // warp (the forecast horizon of a sweep) skip the update altogether -- the branch is warp-uniform, so the
    // shuffles and transpositions inside it are executed by all 32 lanes or by none
    double pprow[M], sp[M];
#pragma unroll
    for (int j = 0; j < M; ++j) pprow[j] = (prow[j] + prow[j]) / 2.0;  // :138 on the a-priori page
#pragma unroll
    for (int i = 0; i < M; ++i) sp[i] = s[i];
    if (__any_sync(0xffffffffu, valid)) {
      // PCt[row] (== CP[row]: P is exactly symmetric), the innovation variance and the gain
      const double pct = fma(prow[2], C[2], fma(prow[1], C[1], prow[0] * C[0]));
      const double pc0 = from_row(pct, L_, 0), pc1 = from_row(pct, L_, 1), pc2 = from_row(pct, L_, 2);
      const double S0 = fma(pc2, C[2], fma(pc1, C[1], pc0 * C[0]));
      const double denom = S0 + gamma * Rk;                               // :124
      const InvDiv by_denom = make_invdiv(denom);
      const double k_own = div_by(pct, by_denom);
      double K[M];
#pragma unroll
      for (int i = 0; i < M; ++i) K[i] = from_row(k_own, L_, i);
      // I - K*C, columns 0..2: this lane's row, and every row for the Joseph form
      double mxr[3];
#pragma unroll
      for (int j = 0; j < 3; ++j) mxr[j] = drow[j] - k_own * C[j];
      // MP(row, :) = (I - K C) P with rows 0..2 of P(k|k-1) from their lanes
      double MP[M];
#pragma unroll
      for (int j = 0; j < M; ++j) {
        const double p0 = from_row(prow[j], L_, 0), p1 = from_row(prow[j], L_, 1), p2 = from_row(prow[j], L_, 2);
        const double acc = fma(mxr[2], p2, fma(mxr[1], p1, mxr[0] * p0));
        MP[j] = (row >= 3) ? (acc + prow[j]) : acc;
      }
      // :127 Joseph form, row `row` before the symmetrisation
      double pu[M], puT[M];
      const double kr = k_own * Rk;
#pragma unroll
      for (int j = 0; j < M; ++j) {
        const double mj0 = ((j == 0) ? 1.0 : 0.0) - K[j] * C[0], mj1 = ((j == 1) ? 1.0 : 0.0) - K[j] * C[1],
                     mj2 = ((j == 2) ? 1.0 : 0.0) - K[j] * C[2];
        double mij = fma(MP[2], mj2, fma(MP[1], mj1, MP[0] * mj0));
        if (j >= 3) mij = mij + MP[j];
        const double nij = mij + kr * K[j];
        pu[j] = div_by(nij, by_gamma);
      }
      transpose6(pu, puT, L_);
      if (valid) {
#pragma unroll
        for (int j = 0; j < M; ++j) pprow[j] = (pu[j] + puT[j]) / 2.0;   // :138 (p_ij + p_ji == p_ji + p_ij)
#pragma unroll
        for (int i = 0; i < M; ++i) sp[i] = s[i] + K[i] * innov;         // :129
      }
    }
    state_margins<MODEL>(mc, sp);  // :141

    // :155-157 state update + Jacobian at s(k|k)
    const double *ud = in.u + (size_t)k * in.u_ts;
    double dotv, a25 = 0.0;
    if (pre == pre) {
      dotv = pre;
    } else {
      const InputPass ip = input_pass<MODEL, true, false>(mc, eps, sp[M - 1], ud, in.u_js, L, nullptr, 0, nullptr);
      dotv = ip.dot;
      a25 = ip.a25;
    }
    double sn[M];
    state_eqs<MODEL>(mc, eps, sp, dotv, sn);
    Mat<M, false> A;
    state_jacobian<MODEL>(mc, eps, sp, a25, A);
    // column `row` of A P(k|k): AP(i, row) = sum_l A(i,l) P(l,row), P(l,row) = this lane's P(row,l)
    double apc[M], apr[M];
#pragma unroll
    for (int i = 0; i < M; ++i) {
      double acc = 0.0;
      bool first = true;
#pragma unroll
      for (int l = 0; l < M; ++l)
        if (a_nz(M, i, l)) {
          acc = first ? A(i, l) * pprow[l] : fma(A(i, l), pprow[l], acc);
          first = false;
        }
      apc[i] = acc;
    }
    transpose6(apc, apr, L_);   // apr[l] = AP(row, l)
    // :158 row `row` of A P A' + Q before the symmetrisation, and its transposed partner
    double qu[M], quT[M];
#pragma unroll
    for (int j = 0; j < M; ++j) {
      double acc = 0.0;
      bool first = true;
#pragma unroll
      for (int l = 0; l < M; ++l)
        if (a_nz(M, j, l)) {
          acc = first ? apr[l] * A(j, l) : fma(apr[l], A(j, l), acc);
          first = false;
        }
      qu[j] = acc + qrow[j];
    }
    transpose6(qu, quT, L_);
    // :167-169 the a-posteriori estimate
    double sp_own = sp[0];
#pragma unroll
    for (int i = 1; i < M; ++i) sp_own = (row == i) ? sp[i] : sp_own;
    if (L_.store && k >= k0) {
      tSp.at_day(k)[tSp.f(row)] = sp_own;
      double *__restrict__ d = tPp.at_day(k);
#pragma unroll
      for (int j = 0; j < M; ++j)
        if (j >= row) d[tPp.f(pbase + j)] = pprow[j];
    }
#pragma unroll
    for (int j = 0; j < M; ++j) prow[j] = (qu[j] + quT[j]) / 2.0;  // :161 (p_ij + p_ji == p_ji + p_ij)
    state_margins<MODEL>(mc, sn);  // :164
#pragma unroll
    for (int i = 0; i < M; ++i) s[i] = sn[i];
  }
}

// ---------------------------------------------------------------------------------------------------
// smoother recursion (:189-230) given the gains J_k of eks_gain, states only (no P_SMOOTH)
// ---------------------------------------------------------------------------------------------------
// Lane r holds row r of J_k: S_SMOOTH(r,k) = S_PLUS(r,k) + J(r,:) (S_SMOOTH(:,k+1) - S_MINUS(:,k+1)), then the
// six lanes exchange their entries.  A day is a ~30-instruction chain, so the pass is bound by the latency of
// the tape loads: the next day's operands are in registers one day ahead and the lines of the day
// kBwdPrefetch days ahead are pulled into L2.
constexpr int kBwdPrefetch = 8;

template <int MODEL>
__global__ void __launch_bounds__(kRowsBlock) eks_backward_rows_kernel(const __grid_constant__ EkfParams P) {
  constexpr int M = 6, MM = 36;
  static_assert(model_dim(MODEL) == 6 && !model_legacy(MODEL) && !model_flipped(MODEL), "generic forward 6-state model");
  const RowLane L_ = row_lane(P.B, nullptr);
  const int b = L_.b, row = L_.row;
  const bool lead = L_.store && row == 0;   // the lane that writes a trajectory's per-day scalars / schedule
  const int T = P.T, L = P.L, k0 = P.k0;
  const TrajIn in = traj_inputs(P, b, M);
  const ModelConsts mc = load_consts(in.prm);
  const long long g = in.g;
  const bool want_cost = P.cost_day.p != nullptr;
  const double *wts = want_cost ? P.weights + (size_t)g * T * L : nullptr;
  const Tape<true> tSm = make_tape<true>(P.S_MINUS, M, T - k0, b, k0), tSp = make_tape<true>(P.S_PLUS, M, T - k0, b, k0);
  const Tape<true> tJ = make_tape<true>(P.J, MM, T - 1 - k0 > 0 ? T - 1 - k0 : 1, b, k0);
  const Tape<true> tDot = make_tape<true>(P.dot_day, 1, T, b), tCost = make_tape<true>(P.cost_day, 1, T, b);

  // eks_backward.cu: emit_inputs -- evaluated by every lane of the group (same values), written by `lead`
  const double nan = __longlong_as_double(0x7ff8000000000000ll);
  // the per-group values of a day (NaN = evaluate per trajectory), fetched one day ahead by the day loop
  auto group_pre = [&](int pos) { return in.dot_grp ? __ldg(in.dot_grp + pos) : nan; };
  auto group_cost = [&](int pos) { return (want_cost && in.cost_grp) ? __ldg(in.cost_grp + pos) : (want_cost ? nan : 0.0); };
  auto emit_inputs = [&](int pos, const double *u_day, size_t u_js, double s5, double pre, double prec) {
    double *uo = nullptr;
    size_t uo_s = 0;
    if (P.u_opt_smooth.p) {
      uo = P.u_opt_smooth.p + (size_t)P.u_opt_smooth.off + b + (size_t)pos * L * P.u_opt_smooth.stride;
      uo_s = (size_t)P.u_opt_smooth.stride;
    } else if (P.u_fore.p && pos >= P.T_hist) {
      uo = P.u_fore.p + (size_t)P.u_fore.off + b + (size_t)(pos - P.T_hist) * L * P.u_fore.stride;
      uo_s = (size_t)P.u_fore.stride;
    }
    if (!uo && !P.dot_day.p) return;
    if (!lead) uo = nullptr;
    double dotv, costv;
    if (pre == pre && prec == prec) {
      dotv = pre;
      costv = prec;
      if (uo) {
#pragma unroll
        for (int j = 0; j < EPI_LMAX; ++j)
          if (j < L) uo[(size_t)j * uo_s] = u_day[(size_t)j * u_js];
      }
    } else {
      const InputPass z = want_cost
          ? input_pass<MODEL, false, true>(mc, in.eps, s5, u_day, u_js, L, uo, uo_s, wts + (size_t)pos * L)
          : input_pass<MODEL, false, false>(mc, in.eps, s5, u_day, u_js, L, uo, uo_s, nullptr);
      dotv = z.dot;
      costv = z.cost;
    }
    if (lead) {
      if (P.dot_day.p) tDot.at_day(pos)[0] = dotv;
      if (want_cost) tCost.at_day(pos)[0] = costv;
    }
  };
  auto own = [&](const double (&v)[M]) {
    double o = v[0];
#pragma unroll
    for (int i = 1; i < M; ++i) o = (row == i) ? v[i] : o;
    return o;
  };

  // :189-202 terminal conditions
  const int posT = T - 1;
  double ss[M];
  {
    const double *__restrict__ d = tSp.at_day(posT);
#pragma unroll
    for (int i = 0; i < M; ++i) ss[i] = d[tSp.f(i)];
    double sf[M];
    if (P.init_per_traj) {
      const double *si = P.s_final_t.p + P.s_final_t.off + b;
#pragma unroll
      for (int i = 0; i < M; ++i) sf[i] = si[(size_t)i * P.s_final_t.stride];
    } else {
#pragma unroll
      for (int i = 0; i < M; ++i) sf[i] = P.s_final_g[g * M + i];
    }
#pragma unroll
    for (int i = 0; i < M; ++i)
      if (!(sf[i] != sf[i])) ss[i] = sf[i];
  }
  if (P.S_SMOOTH.p && L_.store)
    P.S_SMOOTH.p[(size_t)P.S_SMOOTH.off + b + ((size_t)posT * M + row) * P.S_SMOOTH.stride] = own(ss);
  emit_inputs(posT, kZeroInputs, 1, 0.0, nan, want_cost ? nan : 0.0);   // u_opt_smooth(:, T) is never written by the reference (:95,:204)

  struct Day { double J[M]; double sm[M]; double sp; };
  auto load_day = [&](int k, Day &d) {
    const double *__restrict__ j = tJ.at_day(k) + tJ.f(row * M);
#pragma unroll
    for (int q = 0; q < M; ++q) d.J[q] = j[tJ.f(q)];
    const double *__restrict__ a = tSm.at_day(k + 1);
#pragma unroll
    for (int i = 0; i < M; ++i) d.sm[i] = a[tSm.f(i)];
    d.sp = tSp.at_day(k)[tSp.f(row)];
  };
  auto prefetch_day = [&](int k) {
    const double *j = tJ.at_day(k) + tJ.f(row * M);
#pragma unroll
    for (int q = 0; q < M; ++q) asm volatile("prefetch.global.L2 [%0];" ::"l"(j + tJ.f(q)));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(tSm.at_day(k + 1) + tSm.f(row)));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(tSp.at_day(k) + tSp.f(row)));
  };
  for (int k = T - 2; k >= k0 && k > T - 2 - kBwdPrefetch; --k) prefetch_day(k);
  Day cur;
  double pre_nxt = nan, prec_nxt = nan;
  if (T - 2 >= k0) { load_day(T - 2, cur); pre_nxt = group_pre(T - 2); prec_nxt = group_cost(T - 2); }
#pragma unroll 1
  for (int k = T - 2; k >= k0; --k) {  // :204
    if (k - kBwdPrefetch >= k0) prefetch_day(k - kBwdPrefetch);
    const double pre_cur = pre_nxt, prec_cur = prec_nxt;
    if (k > k0) { pre_nxt = group_pre(k - 1); prec_nxt = group_cost(k - 1); }
    double ds[M];
#pragma unroll
    for (int l = 0; l < M; ++l) ds[l] = ss[l] - cur.sm[l];
    double acc = cur.J[0] * ds[0];
#pragma unroll
    for (int l = 1; l < M; ++l) acc = fma(cur.J[l], ds[l], acc);
    const double sk_own = cur.sp + acc;  // :218
    if (k > k0) load_day(k - 1, cur);
    double sk[M];
#pragma unroll
    for (int i = 0; i < M; ++i) sk[i] = from_row(sk_own, L_, i);
    state_margins<MODEL>(mc, sk);  // :221
#pragma unroll
    for (int i = 0; i < M; ++i) ss[i] = sk[i];
    if (P.S_SMOOTH.p && L_.store)
      P.S_SMOOTH.p[(size_t)P.S_SMOOTH.off + b + ((size_t)k * M + row) * P.S_SMOOTH.stride] = own(ss);
    // :229 re-run the state equation's input stage on the smoothed state
    emit_inputs(k, in.u + (size_t)k * in.u_ts, in.u_js, ss[M - 1], pre_cur, prec_cur);
  }
}

void launch_ekf_forward_rows(const EkfParams &p, cudaStream_t st) {
  const long long warps = ((long long)p.B + kRowsPerWarp - 1) / kRowsPerWarp;
  const unsigned grid = (unsigned)((warps * 32 + kRowsBlock - 1) / kRowsBlock);
  if (p.model == EPI_MODEL_OPTCTRL) ekf_forward_rows_kernel<EPI_MODEL_OPTCTRL><<<grid, kRowsBlock, 0, st>>>(p);
}
void launch_eks_backward_rows(const EkfParams &p, cudaStream_t st) {
  const long long warps = ((long long)p.B + kRowsPerWarp - 1) / kRowsPerWarp;
  const unsigned grid = (unsigned)((warps * 32 + kRowsBlock - 1) / kRowsBlock);
  if (p.model == EPI_MODEL_OPTCTRL) eks_backward_rows_kernel<EPI_MODEL_OPTCTRL><<<grid, kRowsBlock, 0, st>>>(p);
}

// the call shapes the lane-group kernels implement
bool rows_forward_ok(const EkfParams &p) {
  const bool monitor = (p.rho.p != nullptr) || (p.beta != 1.0);
  return p.model == EPI_MODEL_OPTCTRL && p.tiled && !monitor && p.q_mode == EPI_Q_CONST && !p.u_opt.p && !p.K_GAIN.p &&
         !p.innov.p && p.B > 0 && p.T > 0;
}
bool rows_backward_ok(const EkfParams &p) {
  return p.model == EPI_MODEL_OPTCTRL && p.tiled && !p.P_SMOOTH.p && !p.P_first.p && p.B > 0 && p.T > 0;
}
// 0 = never, 1 = always (when the shape allows), otherwise by batch size: the lane-group form pays once the
// one-thread kernels cannot fill the machine (measured crossover)
// The lane-group kernels are opt-in (EPI_ROWS=1, where the call shape allows): with the measurements in DESIGN.md 4
// neither form beats the tuned one-thread kernels (forward 1.27 vs 1.11 ms, recursion 0.72 vs 0.58 ms at 7500
// trajectories x 561 days; further behind at larger batches).
bool rows_wanted(long long B, bool forward) {
  (void)B; (void)forward;
  if (const char *e = getenv("EPI_ROWS")) return atoi(e) != 0;   // read per call: tests and tuning runs toggle it
  return false;
}

}  // namespace epi
