// eks_backward.cu -- smoother backward recursion (sm_100a, FP64, --fmad=false).
//
// Tools/GenericExtendedKalmanFilter.m:189-230 (terminal conditions :189-202, recursion
// :218-229) and Tools/NewCaseEKFEstimatorWithOptimalNPI.m:118-139 given the gains J_k
// computed by eks_gain.  One thread per trajectory; the recursion itself is cheap and
// strictly sequential, so the kernel is bound by the latency of streaming J/S-/S+ from
// HBM: the next day's tape page is prefetched into registers while the current day is
// being processed.
#include "ekf_common.cuh"

namespace epi {

template <int M>
struct BwdDay {
  double J[M * M];
  double sm[M];  // S_MINUS(:, k+1)
  double sp[M];  // S_PLUS(:, k)
};

// ROOMY (the sweep's call shape, chosen by the host through P.bwd_prefetch > 0): a 255-register budget instead of 128.
// With 128 registers the kernel spills 200 bytes, and the spill STORE of a prefetched J entry has to wait for that
// load -- the one-day-ahead register prefetch is exposed again (ncu, 7500 trajectories: 20 % of the stall samples on
// two STL instructions).  Without spills, with the tape page three days ahead pulled into L2 and the per-group
// scalars fetched a day ahead: 0.57 / 0.63 / 1.07 / 2.15 ms at 7.5k / 14.7k / 29.5k / 59k trajectories x 561 days
// against 1.02 / 1.09 / 1.38 / 2.31 ms.  (Round 1 measured a register cap of 152 ALONE as a loss: 2.44 ms.)
template <int MODEL, bool WANT_P, bool TILED, bool ROOMY = false>
__global__ void __launch_bounds__(64, (WANT_P || ROOMY) ? 4 : 8) eks_backward_kernel(const __grid_constant__ EkfParams P) {
  constexpr int M = model_dim(MODEL);
  constexpr bool LEG = model_legacy(MODEL);
  constexpr bool SYM = !LEG;
  constexpr bool REV = model_flipped(MODEL);
  constexpr int MM = M * M;
  constexpr int PF = (SYM && TILED) ? M * (M + 1) / 2 : MM;
  constexpr bool PACKED = SYM && TILED;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= P.B) return;
  const int T = P.T, L = P.L;
  const TrajIn in = traj_inputs(P, b, M);
  const ModelConsts mc = load_consts(in.prm);
  const long long g = in.g;
  const bool want_cost = P.cost_day.p != nullptr;
  const double *wts = want_cost ? P.weights + (size_t)g * T * L : nullptr;

  const int k0 = P.k0;  // lean sweeps: the recursion stops at the first day whose schedule is needed
  const Tape<TILED> tSm = make_tape<TILED>(P.S_MINUS, M, T - k0, b, k0), tSp = make_tape<TILED>(P.S_PLUS, M, T - k0, b, k0);
  const Tape<TILED> tPm = make_tape<TILED>(P.P_MINUS, PF, T - k0, b, k0), tPp = make_tape<TILED>(P.P_PLUS, PF, T - k0, b, k0);
  const Tape<true> tJ = make_tape<true>(P.J, MM, T - 1 - k0 > 0 ? T - 1 - k0 : 1, b, k0);
  const Tape<true> tDot = make_tape<true>(P.dot_day, 1, T, b), tCost = make_tape<true>(P.cost_day, 1, T, b);

  // writes the schedule of day `pos` implied by state `s5` (and the per-day scalars of the sweep)
  const double nan = __longlong_as_double(0x7ff8000000000000ll);
  // the per-group values of a day (NaN = evaluate per trajectory); the day loop fetches them one day ahead
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
    if (P.dot_day.p) tDot.at_day(pos)[0] = dotv;
    if (want_cost) tCost.at_day(pos)[0] = costv;
  };

  // :189-202 terminal conditions
  const int posT = REV ? 0 : (T - 1);
  double ss[M];
  Mat<M, false> Ps;
  {
    const double *__restrict__ d = tSp.at_day(posT);
#pragma unroll
    for (int i = 0; i < M; ++i) ss[i] = d[tSp.f(i)];
  }
  if (WANT_P) tape_load_cov<M, false, TILED, PACKED>(Ps, tPp, tPp.at_day(posT));
  {
    double sf[M];
    Mat<M, false> Pf;
    if (P.init_per_traj) {
      const double *si = P.s_final_t.p + P.s_final_t.off + b;
#pragma unroll
      for (int i = 0; i < M; ++i) sf[i] = si[(size_t)i * P.s_final_t.stride];
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
    double *d = P.S_SMOOTH.p + (size_t)P.S_SMOOTH.off + b + (size_t)posT * M * P.S_SMOOTH.stride;
#pragma unroll
    for (int i = 0; i < M; ++i) d[(size_t)i * P.S_SMOOTH.stride] = ss[i];
  }
  if (WANT_P && P.P_SMOOTH.p)
    store_mat<M, false>(Ps, P.P_SMOOTH.p + (size_t)P.P_SMOOTH.off + b + (size_t)posT * MM * P.P_SMOOTH.stride,
                        (size_t)P.P_SMOOTH.stride);
  // u_opt_smooth(:, T) is never written by the reference => zeros (:95,:204)
  if (!LEG) emit_inputs(posT, kZeroInputs, 1, 0.0, nan, want_cost ? nan : 0.0);

  auto load_day = [&](int k, BwdDay<M> &d) {
    const int pos = REV ? (T - 1 - k) : k;
    const int posn = REV ? (T - 2 - k) : (k + 1);
    const double *__restrict__ j = tJ.at_day(k);
#pragma unroll
    for (int q = 0; q < MM; ++q) d.J[q] = j[tJ.f(q)];
    const double *__restrict__ a = tSm.at_day(posn);
    const double *__restrict__ c = tSp.at_day(pos);
#pragma unroll
    for (int i = 0; i < M; ++i) { d.sm[i] = a[tSm.f(i)]; d.sp[i] = c[tSp.f(i)]; }
  };

  // the recursion is bound by the latency of the next day's tape page (one thread walks T days): pull the page of
  // day k - P.bwd_prefetch into L2 ahead of the one-day-ahead register loads
  const int pfd = WANT_P ? 0 : P.bwd_prefetch;
  auto prefetch_day = [&](int k) {
    const int pos = REV ? (T - 1 - k) : k;
    const int posn = REV ? (T - 2 - k) : (k + 1);
    const double *j = tJ.at_day(k);
#pragma unroll
    for (int q = 0; q < MM; ++q) asm volatile("prefetch.global.L2 [%0];" ::"l"(j + tJ.f(q)));
    const double *a = tSm.at_day(posn), *c = tSp.at_day(pos);
#pragma unroll
    for (int i = 0; i < M; ++i) {
      asm volatile("prefetch.global.L2 [%0];" ::"l"(a + tSm.f(i)));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(c + tSp.f(i)));
    }
  };
  if (pfd > 0)
    for (int k = T - 3; k >= k0 && k > T - 2 - pfd; --k) prefetch_day(k);
  BwdDay<M> cur;
  if (!WANT_P && T - 2 >= k0) load_day(T - 2, cur);
  double pre_nxt = nan, prec_nxt = nan;
  if (T - 2 >= k0) { const int p0 = REV ? 1 : (T - 2); pre_nxt = group_pre(p0); prec_nxt = group_cost(p0); }
#pragma unroll 1
  for (int k = T - 2; k >= k0; --k) {  // :204
    const int pos = REV ? (T - 1 - k) : k;
    const int posn = REV ? (T - 2 - k) : (k + 1);
    const double pre_cur = pre_nxt, prec_cur = prec_nxt;
    if (k > k0) { const int pp = REV ? (T - k) : (k - 1); pre_nxt = group_pre(pp); prec_nxt = group_cost(pp); }
    if (pfd > 0 && k - pfd >= k0) prefetch_day(k - pfd);
    if (WANT_P) load_day(k, cur);
    double ds[M], sk[M];
#pragma unroll
    for (int l = 0; l < M; ++l) ds[l] = ss[l] - cur.sm[l];
#pragma unroll
    for (int i = 0; i < M; ++i) {
      double acc = cur.J[i * M] * ds[0];
#pragma unroll
      for (int l = 1; l < M; ++l) acc = fma(cur.J[i * M + l], ds[l], acc);
      sk[i] = cur.sp[i] + acc;  // :218
    }
    // the tape page of this day is consumed: start streaming the previous day's page into
    // the same registers now, so its HBM latency overlaps the rest of this iteration
    if (!WANT_P && k > k0) load_day(k - 1, cur);
    state_margins<MODEL>(mc, sk);  // :221
    if (WANT_P) {
      Mat<M, SYM> Pp, Pn;
      tape_load_cov<M, SYM, TILED, PACKED>(Pp, tPp, tPp.at_day(pos));
      tape_load_cov<M, SYM, TILED, PACKED>(Pn, tPm, tPm.at_day(posn));
      Mat<M, false> D, JD;
#pragma unroll
      for (int i = 0; i < M; ++i)
#pragma unroll
        for (int j = 0; j < M; ++j) D.at(i, j) = Pn(i, j) - Ps(i, j);
#pragma unroll
      for (int i = 0; i < M; ++i)
#pragma unroll
        for (int j = 0; j < M; ++j) {
          double acc = cur.J[i * M] * D(0, j);
#pragma unroll
          for (int l = 1; l < M; ++l) acc = fma(cur.J[i * M + l], D(l, j), acc);
          JD.at(i, j) = acc;
        }
      Mat<M, false> Pnew;
#pragma unroll
      for (int i = 0; i < M; ++i)
#pragma unroll
        for (int j = 0; j < M; ++j) {
          double acc = JD(i, 0) * cur.J[j * M];
#pragma unroll
          for (int l = 1; l < M; ++l) acc = fma(JD(i, l), cur.J[j * M + l], acc);
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
        store_mat<M, false>(Ps, P.P_SMOOTH.p + (size_t)P.P_SMOOTH.off + b + (size_t)pos * MM * P.P_SMOOTH.stride,
                            (size_t)P.P_SMOOTH.stride);
    }
#pragma unroll
    for (int i = 0; i < M; ++i) ss[i] = sk[i];
    if (P.S_SMOOTH.p) {
      double *d = P.S_SMOOTH.p + (size_t)P.S_SMOOTH.off + b + (size_t)pos * M * P.S_SMOOTH.stride;
#pragma unroll
      for (int i = 0; i < M; ++i) d[(size_t)i * P.S_SMOOTH.stride] = ss[i];
    }
    // :229 re-run the state equation's input stage on the smoothed state
    if (!LEG) emit_inputs(pos, in.u + (size_t)pos * in.u_ts, in.u_js, (M == 6) ? ss[M - 1] : 0.0, pre_cur, prec_cur);
  }
  if (WANT_P && P.P_first.p)
    store_mat<M, false>(Ps, P.P_first.p + (size_t)P.P_first.off + b, (size_t)P.P_first.stride);
}

template <int MODEL, bool TILED>
static void launch_bwd_model(const EkfParams &p, cudaStream_t st, bool want_p) {
  const int block = (model_dim(MODEL) == 6) ? 32 : 64;
  const int grid = (p.B + block - 1) / block;
  if (want_p) eks_backward_kernel<MODEL, true, TILED><<<grid, block, 0, st>>>(p);
  else if (p.bwd_prefetch > 0 && model_dim(MODEL) == 6) eks_backward_kernel<MODEL, false, TILED, true><<<grid, block, 0, st>>>(p);
  else        eks_backward_kernel<MODEL, false, TILED><<<grid, block, 0, st>>>(p);
}

void launch_eks_backward(const EkfParams &p, cudaStream_t st) {
  const bool want_p = (p.P_SMOOTH.p != nullptr) || (p.P_first.p != nullptr);
  if (rows_backward_ok(p) && rows_wanted(p.B, false)) {   // small batch: six lanes per trajectory (csrc/ekf_rows.cu)
    launch_eks_backward_rows(p, st);
    return;
  }
#define CALL(MDL)                                               \
  if (p.tiled) launch_bwd_model<MDL, true>(p, st, want_p);      \
  else launch_bwd_model<MDL, false>(p, st, want_p)
  EPI_DISPATCH_MODEL(p.model, CALL)
#undef CALL
}

}  // namespace epi
