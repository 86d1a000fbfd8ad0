// eks_backward.cu -- smoother backward recursion (sm_100a, FP64, --fmad=false).
//
// Tools/GenericExtendedKalmanFilter.m:189-230 (terminal conditions :189-202, recursion
// :218-229) and Tools/NewCaseEKFEstimatorWithOptimalNPI.m:118-139 given the gains J_k
// computed by eks_gain.  One thread per trajectory; the recursion itself is cheap and
// strictly sequential, so the kernel is bound by the latency of streaming J/S-/S+ from
// HBM: the next day's tape page is prefetched into registers while the current day is
// being processed.
#include "ekf_common.cuh"
#include "epi_async.cuh"

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
//
// STAGED (region shards of the strong-scaling sweep, a few hundred tiles): the day of a small batch is ~150 cycles of
// arithmetic behind an L2 / HBM round trip that a one-day-ahead register prefetch cannot cover (1 us a day measured,
// 0.56 ms for 561 days whatever the batch).  The three tape pages of a (tile, day) are contiguous -- J 9216 B, S- and
// S+ 1536 B each -- so lane 0 streams them with three TMA bulk copies per day into a shared-memory ring that runs
// P.bwd_stages days ahead (mbarrier completion); the warp reads its columns from the ring.  One 32-thread CTA per tile.
constexpr int kBwdMaxStages = 16;
template <int MODEL, bool WANT_P, bool TILED, bool ROOMY = false, bool STAGED = false>
__global__ void __launch_bounds__(64, (WANT_P || ROOMY) ? 4 : 8) eks_backward_kernel(const __grid_constant__ EkfParams P) {
  constexpr int M = model_dim(MODEL);
  constexpr bool LEG = model_legacy(MODEL);
  constexpr bool SYM = !LEG;
  constexpr bool REV = model_flipped(MODEL);
  constexpr int MM = M * M;
  constexpr int PF = (SYM && TILED) ? M * (M + 1) / 2 : MM;
  constexpr bool PACKED = SYM && TILED;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  extern __shared__ __align__(128) double bwd_ring[];  // STAGED: [stages][J 36 | S- 6 | S+ 6][32]
  __shared__ unsigned long long bwd_bars[kBwdMaxStages];
  unsigned wmask = 0xffffffffu;
  if (STAGED) wmask = __ballot_sync(0xffffffffu, b < P.B);  // (one warp per CTA)
  if (b >= P.B) return;
  const int T = P.T, L = P.L;
  const TrajIn in = traj_inputs(P, b, M);
  const ModelConsts mc = load_consts(in.prm);
  const long long g = in.g;
  const bool want_cost = STAGED || P.cost_day.p != nullptr;
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
    if (!STAGED && P.u_opt_smooth.p) {  // (STAGED = the sweep's call shape: u_fore, the per-day scalars, no u_opt_smooth)
      uo = P.u_opt_smooth.p + (size_t)P.u_opt_smooth.off + b + (size_t)pos * L * P.u_opt_smooth.stride;
      uo_s = (size_t)P.u_opt_smooth.stride;
    } else if (P.u_fore.p && pos >= P.T_hist) {
      uo = P.u_fore.p + (size_t)P.u_fore.off + b + (size_t)(pos - P.T_hist) * L * P.u_fore.stride;
      uo_s = (size_t)P.u_fore.stride;
    }
    if (!STAGED && !uo && !P.dot_day.p) return;
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
          ? input_pass<MODEL, false, true, STAGED>(mc, in.eps, s5, u_day, u_js, L, uo, uo_s, wts + (size_t)pos * L)
          : input_pass<MODEL, false, false, STAGED>(mc, in.eps, s5, u_day, u_js, L, uo, uo_s, nullptr);
      dotv = z.dot;
      costv = z.cost;
    }
    if (STAGED || P.dot_day.p) tDot.at_day(pos)[0] = dotv;
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
  const int pfd = (WANT_P || STAGED) ? 0 : P.bwd_prefetch;
  // the schedule of a day to optimise reads that day's input and weight rows (per group): pull the rows of the NEXT
  // day to process into L1 while this one is computed (their latency was 17 % of the small-batch recursion, ncu)
  auto prefetch_rows = [&](int pos) {
    const double *ur = in.u + (size_t)pos * in.u_ts;
    asm volatile("prefetch.global.L1 [%0];" ::"l"(ur));
    asm volatile("prefetch.global.L1 [%0];" ::"l"(ur + (size_t)(L - 1) * in.u_js));
    if (want_cost) {
      asm volatile("prefetch.global.L1 [%0];" ::"l"(wts + (size_t)pos * L));
      asm volatile("prefetch.global.L1 [%0];" ::"l"(wts + (size_t)pos * L + (L - 1)));
    }
  };
  if constexpr (STAGED) {
    // day k is item n = T - 2 - k of the ring: slot n % D, phase (n / D) & 1 (kept as running counters); lane 0 of the
    // warp holds the page starts of the tile and issues the copies, D days ahead of the recursion
    constexpr int kPage = (MM + 2 * M) * 32;  // doubles per ring slot
    const int D = P.bwd_stages;
    const bool leader = (threadIdx.x & 31) == 0;
    // pages of the next day to issue, from block-uniform values only (the copy instruction takes uniform registers: a
    // lane-derived address costs a broadcast loop per copy)
    const size_t tile = blockIdx.x;
    const int Tn = T - k0;
    const double *gj = P.J.p + (tile * (size_t)(Tn - 1) + (size_t)(T - 2 - k0)) * (MM * 32);
    const double *gsm = P.S_MINUS.p + (tile * (size_t)Tn + (size_t)(T - 1 - k0)) * (M * 32);
    const double *gsp = P.S_PLUS.p + (tile * (size_t)Tn + (size_t)(T - 2 - k0)) * (M * 32);
    int islot = 0;
    // the three pages of one day into slot `islot`: every lane keeps the (uniform) bookkeeping, the leader copies
    auto issue_next = [&]() {
      if (leader) {
        double *dst = bwd_ring + (size_t)islot * kPage;
        mbar_expect_tx(&bwd_bars[islot], (unsigned)(kPage * sizeof(double)));
        bulk_load_row(dst, gj, (unsigned)(MM * 32 * sizeof(double)), &bwd_bars[islot]);
        bulk_load_row(dst + MM * 32, gsm, (unsigned)(M * 32 * sizeof(double)), &bwd_bars[islot]);
        bulk_load_row(dst + (MM + M) * 32, gsp, (unsigned)(M * 32 * sizeof(double)), &bwd_bars[islot]);
      }
      gj -= MM * 32; gsm -= M * 32; gsp -= M * 32;
      islot = (islot + 1 == D) ? 0 : islot + 1;
    };
    if (leader) {
      for (int q = 0; q < D; ++q) mbar_init(&bwd_bars[q], 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp(wmask);
    int issued = T - 2;  // next day to issue
    for (; issued >= k0 && issued > T - 2 - D; --issued) issue_next();
    int slot = 0;
    unsigned phase = 0;
    double pre_nxt = nan, prec_nxt = nan;
    if (T - 2 >= k0) { pre_nxt = group_pre(T - 2); prec_nxt = group_cost(T - 2); }
#pragma unroll 1
    for (int k = T - 2; k >= k0; --k) {  // :204
      const double pre_cur = pre_nxt, prec_cur = prec_nxt;
      if (k > k0) {
        pre_nxt = group_pre(k - 1); prec_nxt = group_cost(k - 1);
        if (k - 1 >= P.T_hist) prefetch_rows(k - 1);
      }
      mbar_wait(&bwd_bars[slot], phase);
      const double *__restrict__ sv = bwd_ring + (size_t)slot * kPage + (threadIdx.x & 31);
      double ds[M], sk[M], sp[M];
#pragma unroll
      for (int l = 0; l < M; ++l) ds[l] = ss[l] - sv[(MM + l) * 32];
#pragma unroll
      for (int i = 0; i < M; ++i) sp[i] = sv[(MM + M + i) * 32];
#pragma unroll
      for (int i = 0; i < M; ++i) {
        double acc = sv[(i * M) * 32] * ds[0];
#pragma unroll
        for (int l = 1; l < M; ++l) acc = fma(sv[(i * M + l) * 32], ds[l], acc);
        sk[i] = sp[i] + acc;  // :218
      }
      __syncwarp(wmask);  // every lane has consumed its column: the slot may be refilled
      if (issued >= k0) { issue_next(); --issued; }
      slot = (slot + 1 == D) ? 0 : slot + 1;
      phase ^= (slot == 0) ? 1u : 0u;
      state_margins<MODEL>(mc, sk);  // :221
#pragma unroll
      for (int i = 0; i < M; ++i) ss[i] = sk[i];
      emit_inputs(k, in.u + (size_t)k * in.u_ts, in.u_js, ss[M - 1], pre_cur, prec_cur);  // :229
    }
    return;
  }
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
    if (k > k0) {
      const int pp = REV ? (T - k) : (k - 1);
      pre_nxt = group_pre(pp); prec_nxt = group_cost(pp);
      if (pfd > 0 && pp >= P.T_hist) prefetch_rows(pp);
    }
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

// ring depth of the staged recursion for `tiles` one-warp CTAs: as deep as the shared memory of an SM allows when every
// tile is resident, 0 = the batch is too large (the plain kernel is then bound by HBM, not by latency)
static int backward_stages(int tiles) {
  if (const char *e = getenv("EPI_BWD_STAGES")) {
    const int d = atoi(e);
    return d < 0 ? 0 : (d > kBwdMaxStages ? kBwdMaxStages : d);
  }
  int dev = 0, n_sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, dev);
  const int per_sm = (tiles + n_sms - 1) / n_sms;
  if (per_sm > 4) return 0;  // measured: 235 tiles 0.39 against 0.56 ms, 461 tiles 0.61 against 0.63; beyond that HBM binds
  const int d = (int)((size_t)220 * 1024 / ((size_t)per_sm * (48 * 32 * 8 + 1024)));
  return d < 3 ? 0 : (d > 8 ? 8 : d);
}

template <int MODEL, bool TILED>
static void launch_bwd_model(const EkfParams &p, cudaStream_t st, bool want_p) {
  const int block = (model_dim(MODEL) == 6) ? 32 : 64;
  const int grid = (p.B + block - 1) / block;
  if constexpr (MODEL == EPI_MODEL_OPTCTRL && TILED) {
    // the sweep's call shape (per-day scalars for the rollout, schedule to u_fore only, group inputs) on a small batch
    const bool sweep_shape = !want_p && p.bwd_prefetch > 0 && p.T - p.k0 >= 2 && !p.S_SMOOTH.p && !p.u_opt_smooth.p &&
                             p.dot_day.p && p.cost_day.p && p.weights && !p.u_trj.p && !p.init_per_traj;
    const int stages = sweep_shape ? backward_stages(grid) : 0;
    if (stages > 0) {
      EkfParams q = p;
      q.bwd_stages = stages;
      const size_t smem = (size_t)stages * 48 * 32 * sizeof(double);
      auto kern = eks_backward_kernel<MODEL, false, true, true, true>;
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      kern<<<grid, 32, smem, st>>>(q);
      return;
    }
  }
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
