// ekf_pair.cu -- forward EKF pass with TWO WARPS PER 32-TRAJECTORY TILE (sm_100a, FP64, --fmad=false), for small
// batches of the 6-state state+costate model (EPI_MODEL_OPTCTRL) in the call shape of the fused sweep (tiled
// scratch tape, constant Q, no innovation monitor, none of the optional per-day outputs).
//
// Why: the forward pass (GenericExtendedKalmanFilter.m:98-186) walks the T days of a trajectory one after the other.
// A region shard of the strong-scaling sweep (TrainPredictPrescribeNPI.m:93) is a few hundred warps -- one or two per
// SM -- and one warp issues its ~1340 FP64 instructions per day on ONE scheduler's FP64 pipe at one per two cycles:
// ~2 us a day however idle the rest of the GPU is (measured: 1.11 ms for 7500 trajectories x 561 days; narrower warps
// change nothing, six lanes per trajectory lose their instruction-level parallelism, csrc/ekf_rows.cu).  Here the two
// warps of a 64-thread CTA share a tile: warp h owns rows 3h..3h+2 of P for all 32 trajectories (lane = trajectory, so
// every thread keeps the one-thread kernel's parallelism and the tape stores stay 256-byte rows), the scalar model
// callbacks are evaluated by both, and the pieces of the other half that a step needs (three innovation-covariance
// terms, the gains, a 3x3 diagonal block twice, the 3x3 cross block of the two symmetrisations) cross through shared
// memory at five CTA barriers per day: 42 doubles per thread and direction.
//
// Bit-identical to ekf_forward.cu: every matrix element is produced by one thread with the one-thread operation
// sequence (first term a*b, then fma, index ascending; structural zeros skipped); P(l,j) of the exactly symmetric
// pages is read as the thread's own P(j,l) where the row l belongs to the other warp.
#include <cstdlib>

#include "ekf_common.cuh"

#ifndef EPI_PAIR_MAX_BATCH
#define EPI_PAIR_MAX_BATCH 16384
#endif

namespace epi {

namespace {

constexpr int kXSlots = 45;  // exchange slots per direction and day (each written once a day)

// slot bases of the five exchanges
constexpr int kXa = 0;    // 3 innovation terms PCt[0..2] + 6 entries of P-(0..2, 0..2)   (warp 0 -> warp 1 only)
constexpr int kXb = 9;    // 3 gains
constexpr int kXc = 12;   // 9 cross entries of the unsymmetrised P+
constexpr int kXd = 21;   // 6 entries of the diagonal block of P+
constexpr int kXe = 27;   // 9 cross entries of the unsymmetrised P-(k+1)
static_assert(kXe + 9 <= kXSlots, "exchange slots");

// rows 3H..3H+2 of A P and of A P A' + Q (GenericExtendedKalmanFilter.m:158) before the symmetrisation
template <int H>
EPI_DI void time_update_rows(const Mat<6, false> &A, const double (&own)[3][6], const double (&oth)[3][6], const double *__restrict__ Q,
                             double (&qu)[3][6]) {
  constexpr int M = 6;
  // P(k|k)(l, j): rows 3H..3H+2 are `own`, the other three rows `oth`
  auto Pf = [&](int l, int j) { return (l / 3 == H) ? own[l % 3][j] : oth[l % 3][j]; };
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    double AP[M];
#pragma unroll
    for (int j = 0; j < M; ++j) {
      double acc = 0.0;
      bool first = true;
#pragma unroll
      for (int l = 0; l < M; ++l)
        if (a_nz(M, 3 * H + a, l)) {
          acc = first ? A(3 * H + a, l) * Pf(l, j) : fma(A(3 * H + a, l), Pf(l, j), acc);
          first = false;
        }
      AP[j] = acc;
    }
#pragma unroll
    for (int j = 0; j < M; ++j) {
      double acc = 0.0;
      bool first = true;
#pragma unroll
      for (int l = 0; l < M; ++l)
        if (a_nz(M, j, l)) {
          acc = first ? AP[l] * A(j, l) : fma(AP[l], A(j, l), acc);
          first = false;
        }
      qu[a][j] = acc + __ldg(Q + j * M + (3 * H + a));   // constant Q (read-only path): q_elem(i, j) = Q[j*M + i]
    }
  }
}

}  // namespace

template <int MODEL>
__global__ void __launch_bounds__(64) ekf_forward_pair_kernel(const __grid_constant__ EkfParams P) {
  constexpr int M = 6;
  static_assert(model_dim(MODEL) == 6 && !model_legacy(MODEL) && !model_flipped(MODEL), "generic forward 6-state model");
  __shared__ double xbuf[2][kXSlots][32];   // [writer warp][slot][lane]
  const int lane = threadIdx.x & 31, h = threadIdx.x >> 5, o = 1 - h;
  const int r0 = 3 * h;                     // first own row (the other warp's: c0 = 3 - r0)
  int b = blockIdx.x * 32 + lane;
  const bool live = b < P.B;
  if (!live) b = P.B - 1;                    // shadow lanes of a ragged tile: same barriers, no stores
  const TrajIn in = traj_inputs(P, b, M);
  const ModelConsts mc = load_consts(in.prm);
  const int T = P.T, L = P.L, k0 = P.k0;
  const double gamma = P.gamma, v_bar = P.v_bar, eps = in.eps;
  const InvDiv by_gamma = make_invdiv(gamma);
  const Tape<true> tSm = make_tape<true>(P.S_MINUS, M, T - k0, b, k0), tSp = make_tape<true>(P.S_PLUS, M, T - k0, b, k0);
  const Tape<true> tPm = make_tape<true>(P.P_MINUS, 21, T - k0, b, k0), tPp = make_tape<true>(P.P_PLUS, 21, T - k0, b, k0);
  auto send = [&](int slot, double v) { xbuf[h][slot][lane] = v; };
  auto recv = [&](int slot) { return xbuf[o][slot][lane]; };

  // rows r0..r0+2 of a symmetric page as full rows: pr[a][j] = P(r0 + a, j)
  double s[M], pr[3][M];
  {
    const double *si, *pi;
    size_t sst, pst;
    if (P.init_per_traj) { si = P.s_init_t.p + P.s_init_t.off + b; sst = (size_t)P.s_init_t.stride;
                           pi = P.Ps_init_t.p + P.Ps_init_t.off + b; pst = (size_t)P.Ps_init_t.stride; }
    else                 { si = P.s_init_g + in.g * M; sst = 1; pi = P.Ps_init_g + (size_t)in.g * 36; pst = 1; }
#pragma unroll
    for (int i = 0; i < M; ++i) s[i] = si[(size_t)i * sst];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int j = 0; j < M; ++j) {
        // load_mat<M, true> keeps the upper triangle: P(i,j), i <= j, = field j*M + i
        const int i = r0 + a, lo = i < j ? i : j, hi = i < j ? j : i;
        pr[a][j] = pi[(size_t)(hi * M + lo) * pst];
      }
  }
  auto day_x = [&](int kk) { return __ldg(in.x + (size_t)kk * in.x_ts); };
  auto day_R = [&](int kk) { return (P.r_mode == EPI_R_CONST) ? in.R_const : __ldg(in.R + (size_t)kk * in.R_ts); };
  auto day_pre = [&](int kk) { return in.dot_grp ? __ldg(in.dot_grp + kk) : __longlong_as_double(0x7ff8000000000000ll); };
  double x_nxt = day_x(0), R_nxt = day_R(0), pre_nxt = day_pre(0);
  // packed-page field of entry (i, j), i <= j
  auto fld = [](int i, int j) { return Mat<6, true>::idx(i, j); };

#pragma unroll 1
  for (int k = 0; k < T; ++k) {
    const double x_cur = x_nxt, Rk = R_nxt, pre = pre_nxt;
    if (k + 1 < T) { x_nxt = day_x(k + 1); R_nxt = day_R(k + 1); pre_nxt = day_pre(k + 1); }
    // :100-101 the a-priori estimate: own states, own rows of the packed page
    if (live && k >= k0) {
      double *__restrict__ d = tSm.at_day(k);
#pragma unroll
      for (int a = 0; a < 3; ++a) d[tSm.f(r0 + a)] = (h == 0) ? s[a] : s[3 + a];
      double *__restrict__ dp = tPm.at_day(k);
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int j = 0; j < M; ++j)
          if (j >= r0 + a) dp[tPm.f(fld(r0 + a, j))] = pr[a][j];
    }
    double C[3];
    const double xhat = obs_model<MODEL>(mc.obs_type, s, v_bar, C);  // :115-119
    const bool valid = !(x_cur != x_cur);                               // :122
    const bool any_valid = __syncthreads_or(valid ? 1 : 0) != 0;        // CTA-uniform: the barriers below are taken by all or none

    double pp[3][M], sp[M];   // own rows of P(k|k)
    // (:131-134) no observation: the a-priori estimate is kept, :138 symmetrises it
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int j = 0; j < M; ++j) pp[a][j] = (pr[a][j] + pr[a][j]) / 2.0;
#pragma unroll
    for (int i = 0; i < M; ++i) sp[i] = s[i];
    if (any_valid) {
      const double innov = x_cur - xhat;  // :123
      double pct[3];   // PCt of the own rows (== CP: P is exactly symmetric)
#pragma unroll
      for (int a = 0; a < 3; ++a) pct[a] = fma(pr[a][2], C[2], fma(pr[a][1], C[1], pr[a][0] * C[0]));
      // exchange A: warp 0 owns rows 0..2 -- the three terms of the innovation variance and the block P-(0..2,0..2)
      if (h == 0) {
#pragma unroll
        for (int a = 0; a < 3; ++a) send(kXa + a, pct[a]);
        send(kXa + 3, pr[0][0]); send(kXa + 4, pr[0][1]); send(kXa + 5, pr[0][2]);
        send(kXa + 6, pr[1][1]); send(kXa + 7, pr[1][2]); send(kXa + 8, pr[2][2]);
      }
      __syncthreads();
      double cp[3], P0[M], P1[M], P2[M];   // rows 0..2 of P(k|k-1)
      if (h == 0) {
#pragma unroll
        for (int a = 0; a < 3; ++a) cp[a] = pct[a];
#pragma unroll
        for (int j = 0; j < M; ++j) { P0[j] = pr[0][j]; P1[j] = pr[1][j]; P2[j] = pr[2][j]; }
      } else {
#pragma unroll
        for (int a = 0; a < 3; ++a) cp[a] = recv(kXa + a);
        P0[0] = recv(kXa + 3); P0[1] = recv(kXa + 4); P0[2] = recv(kXa + 5);
        P1[0] = P0[1];         P1[1] = recv(kXa + 6); P1[2] = recv(kXa + 7);
        P2[0] = P0[2];         P2[1] = P1[2];         P2[2] = recv(kXa + 8);
#pragma unroll
        for (int a = 0; a < 3; ++a) { P0[3 + a] = pr[a][0]; P1[3 + a] = pr[a][1]; P2[3 + a] = pr[a][2]; }   // P(l, 3+a) = P(3+a, l)
      }
      const double S0 = fma(cp[2], C[2], fma(cp[1], C[1], cp[0] * C[0]));
      const double denom = S0 + gamma * Rk;  // :124
      const InvDiv by_denom = make_invdiv(denom);
      double K[M];
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        const double kk = div_by(pct[a], by_denom);
        if (h == 0) K[a] = kk; else K[3 + a] = kk;
        send(kXb + a, kk);
      }
      __syncthreads();   // exchange B: the gains of the other half
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        const double kk = recv(kXb + a);
        if (h == 0) K[3 + a] = kk; else K[a] = kk;
      }
      double Mx[M][3];  // I - K*C, columns 0..2
#pragma unroll
      for (int i = 0; i < M; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) Mx[i][j] = ((i == j) ? 1.0 : 0.0) - K[i] * C[j];
      // own rows of MP = (I - K C) P and of the Joseph form (:127) before the symmetrisation
      double pu[3][M];
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        double mx0, mx1, mx2, kr;
        if (h == 0) { mx0 = Mx[a][0]; mx1 = Mx[a][1]; mx2 = Mx[a][2]; kr = K[a] * Rk; }
        else        { mx0 = Mx[3 + a][0]; mx1 = Mx[3 + a][1]; mx2 = Mx[3 + a][2]; kr = K[3 + a] * Rk; }
        double MP[M];
#pragma unroll
        for (int j = 0; j < M; ++j) {
          const double acc = fma(mx2, P2[j], fma(mx1, P1[j], mx0 * P0[j]));
          MP[j] = (h == 1) ? (acc + pr[a][j]) : acc;   // rows 3..5: + P(i,j)
        }
#pragma unroll
        for (int j = 0; j < M; ++j) {
          double mij = fma(MP[2], Mx[j][2], fma(MP[1], Mx[j][1], MP[0] * Mx[j][0]));
          if (j >= 3) mij = mij + MP[j];
          const double nij = mij + kr * K[j];
          pu[a][j] = div_by(nij, by_gamma);
        }
      }
      // exchange C: the cross block (own rows x the other warp's columns)
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int cidx = 0; cidx < 3; ++cidx) send(kXc + 3 * a + cidx, h == 0 ? pu[a][3 + cidx] : pu[a][cidx]);
      __syncthreads();
      if (valid) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
#pragma unroll
          for (int j = 0; j < M; ++j) {
            // p_ji: own block from the own registers, cross block = entry (row j - c0 of the other warp, column a)
            double pji;
            if (h == 0) pji = (j < 3) ? pu[j][a] : recv(kXc + 3 * (j - 3) + a);
            else        pji = (j >= 3) ? pu[j - 3][3 + a] : recv(kXc + 3 * j + a);
            pp[a][j] = (pu[a][j] + pji) / 2.0;   // :138 (p_ij + p_ji == p_ji + p_ij)
          }
        }
#pragma unroll
        for (int i = 0; i < M; ++i) sp[i] = s[i] + K[i] * innov;  // :129
      }
    }
    state_margins<MODEL>(mc, sp);  // :141
    // :167-169 the a-posteriori estimate
    if (live && k >= k0) {
      double *__restrict__ d = tSp.at_day(k);
#pragma unroll
      for (int a = 0; a < 3; ++a) d[tSp.f(r0 + a)] = (h == 0) ? sp[a] : sp[3 + a];
      double *__restrict__ dp = tPp.at_day(k);
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int j = 0; j < M; ++j)
          if (j >= r0 + a) dp[tPp.f(fld(r0 + a, j))] = pp[a][j];
    }
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
    // exchange D: the diagonal block of P(k|k) of the other warp (its rows x its columns)
    send(kXd + 0, h == 0 ? pp[0][0] : pp[0][3]); send(kXd + 1, h == 0 ? pp[0][1] : pp[0][4]); send(kXd + 2, h == 0 ? pp[0][2] : pp[0][5]);
    send(kXd + 3, h == 0 ? pp[1][1] : pp[1][4]); send(kXd + 4, h == 0 ? pp[1][2] : pp[1][5]); send(kXd + 5, h == 0 ? pp[2][2] : pp[2][5]);
    __syncthreads();
    // the other warp's rows of P(k|k): po[a][j] = P(c0 + a, j)
    double po[3][M];
    {
      const double d00 = recv(kXd + 0), d01 = recv(kXd + 1), d02 = recv(kXd + 2), d11 = recv(kXd + 3), d12 = recv(kXd + 4),
                   d22 = recv(kXd + 5);
      const double blk[3][3] = {{d00, d01, d02}, {d01, d11, d12}, {d02, d12, d22}};
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int cidx = 0; cidx < 3; ++cidx) {
          // its diagonal block as received; its cross block = the transpose of ours: P(c0 + a, r0 + c) = P(r0 + c, c0 + a)
          if (h == 0) { po[a][3 + cidx] = blk[a][cidx]; po[a][cidx] = pp[cidx][3 + a]; }
          else        { po[a][cidx] = blk[a][cidx];     po[a][3 + cidx] = pp[cidx][a]; }
        }
    }
    // own rows of A P and of A P A' + Q (:158) before the symmetrisation (h is warp-uniform: two straight-line paths)
    double qu[3][M];
    if (h == 0) time_update_rows<0>(A, pp, po, in.Q, qu); else time_update_rows<1>(A, pp, po, in.Q, qu);
    // exchange E: the cross block of the unsymmetrised P(k+1|k)
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int cidx = 0; cidx < 3; ++cidx) send(kXe + 3 * a + cidx, h == 0 ? qu[a][3 + cidx] : qu[a][cidx]);
    __syncthreads();
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int j = 0; j < M; ++j) {
        double pji;
        if (h == 0) pji = (j < 3) ? qu[j][a] : recv(kXe + 3 * (j - 3) + a);
        else        pji = (j >= 3) ? qu[j - 3][3 + a] : recv(kXe + 3 * j + a);
        pr[a][j] = (qu[a][j] + pji) / 2.0;   // :161
      }
    state_margins<MODEL>(mc, sn);  // :164
#pragma unroll
    for (int i = 0; i < M; ++i) s[i] = sn[i];
  }
}

// the call shape this kernel implements, and when it pays: few enough tiles that the one-warp-per-tile launch leaves
// most schedulers idle (measured crossover in DESIGN.md 4); EPI_PAIR=0/1 forces it off/on
bool pair_forward_wanted(const EkfParams &p) {
  const bool monitor = (p.rho.p != nullptr) || (p.beta != 1.0);
  const bool ok = p.model == EPI_MODEL_OPTCTRL && p.tiled && !monitor && p.q_mode == EPI_Q_CONST && !p.u_opt.p && !p.K_GAIN.p &&
                  !p.innov.p && p.B > 0 && p.T > 0;
  if (!ok) return false;
  // opt-in: measured slower than the one-warp kernel at every shard size ("two warps per tile")
  if (const char *e = getenv("EPI_PAIR")) return atoi(e) != 0 && p.B <= EPI_PAIR_MAX_BATCH;
  return false;
}

void launch_ekf_forward_pair(const EkfParams &p, cudaStream_t st) {
  const unsigned grid = (unsigned)((p.B + 31) / 32);
  if (p.model == EPI_MODEL_OPTCTRL) ekf_forward_pair_kernel<EPI_MODEL_OPTCTRL><<<grid, 64, 0, st>>>(p);
}

}  // namespace epi
