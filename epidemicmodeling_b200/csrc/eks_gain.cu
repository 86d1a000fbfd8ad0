// eks_gain.cu -- smoother-gain pass (sm_100a, FP64, --fmad=false).
//
// The fixed-interval smoother gain  J_k = (P+_k A_k') pinv(P-_{k+1})
// (Tools/GenericExtendedKalmanFilter.m:206-217; legacy: mrdivide,
// Tools/NewCaseEKFEstimatorWithOptimalNPI.m:131-132) depends only on the forward tape,
// NOT on the backward recursion, so it is computed for every (day, trajectory) pair in
// parallel: (T-1)*B threads.  This is where most of the FP64 work is (the 6x6 Jacobi pinv)
// and it has all the parallelism a B200 wants even for a single-region sweep.
#include "ekf_common.cuh"

namespace epi {

#ifndef EPI_GAIN_MIN_BLOCKS
#define EPI_GAIN_MIN_BLOCKS 3
#endif

template <int MODEL, bool TILED>
__global__ void __launch_bounds__(128, EPI_GAIN_MIN_BLOCKS) eks_gain_kernel(const __grid_constant__ EkfParams P) {
  constexpr int M = model_dim(MODEL);
  constexpr bool LEG = model_legacy(MODEL);
  constexpr bool SYM = !LEG;
  constexpr bool REV = model_flipped(MODEL);
  constexpr int MM = M * M;
  constexpr int PF = (SYM && TILED) ? M * (M + 1) / 2 : MM;
  constexpr bool PACKED = SYM && TILED;
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t Bpad = (size_t)((P.B + 31) / 32) * 32;  // every warp = one 32-trajectory tile
  const int T = P.T, L = P.L;
  const int k0 = P.k0;
  if (tid >= (size_t)(T - 1 - k0) * Bpad) return;
  const int k = k0 + (int)(tid / Bpad);
  const int b = (int)(tid % Bpad);
  if (b >= P.B) return;
  const int pos = REV ? (T - 1 - k) : k;
  const int posn = REV ? (T - 2 - k) : (k + 1);
  const TrajIn in = traj_inputs(P, b, M);
  const ModelConsts mc = load_consts(in.prm);

  const Tape<TILED> tSp = make_tape<TILED>(P.S_PLUS, M, T - k0, b, k0);
  const Tape<TILED> tPm = make_tape<TILED>(P.P_MINUS, PF, T - k0, b, k0), tPp = make_tape<TILED>(P.P_PLUS, PF, T - k0, b, k0);
  const Tape<true> tJ = make_tape<true>(P.J, MM, T - 1 - k0, b, k0);

  double sp[M];
  {
    const double *__restrict__ d = tSp.at_day(pos);
#pragma unroll
    for (int i = 0; i < M; ++i) sp[i] = d[tSp.f(i)];
  }
  double a25 = 0.0;
  if (M == 6) {
    const double pre = in.dot_grp ? __ldg(in.dot_grp + pos) : __longlong_as_double(0x7ff8000000000000ll);
    if (!(pre == pre)) {
      const InputPass ip = input_pass<MODEL, true, false>(mc, in.eps, sp[M - 1], in.u + (size_t)pos * in.u_ts,
                                                          in.u_js, L, nullptr, 0, nullptr);
      a25 = ip.a25;
    }
  }
  Mat<M, false> A;
  state_jacobian<MODEL>(mc, in.eps, sp, a25, A);  // :206

  Mat<M, false> Jm;
  int rank = M;
  bool bad = false;
  if (!LEG) {
    Mat<M, true> Pn, X;
    tape_load_cov<M, true, TILED, PACKED>(Pn, tPm, tPm.at_day(posn));
#pragma unroll
    for (int q = 0; q < Mat<M, true>::N; ++q) bad |= !(fabs(Pn.v[q]) <= 1.79769313486231570815e308);  // :211
    if (bad) {
#pragma unroll
      for (int q = 0; q < MM; ++q) Jm.v[q] = 0.0;  // :213
    } else {
      VRegs<M> v;  // eigenvectors in registers (a shared-memory V was measured slower: 9.1 vs 8.0 ms)
      rank = pinv_sym<M>(Pn, X, v);
      Mat<M, true> Pp;
      tape_load_cov<M, true, TILED, PACKED>(Pp, tPp, tPp.at_day(pos));
      Mat<M, false> PAt;
#pragma unroll
      for (int i = 0; i < M; ++i)
#pragma unroll
        for (int j = 0; j < M; ++j) PAt.at(i, j) = mul_X_At_ij<M>(Pp, A, i, j);
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
    tape_load_cov<M, false, TILED, false>(Pp, tPp, tPp.at_day(pos));
    {
      Mat<M, false> Pn;
      tape_load_cov<M, false, TILED, false>(Pn, tPm, tPm.at_day(posn));
#pragma unroll
      for (int i = 0; i < M; ++i)
#pragma unroll
        for (int j = 0; j < M; ++j) lu.at(i, j) = Pn(j, i);
    }
#pragma unroll
    for (int i = 0; i < M; ++i)
#pragma unroll
      for (int j = 0; j < M; ++j) rhs.at(j, i) = mul_X_At_ij<M>(Pp, A, i, j);  // rhs = (P+ A')'
    lu_solve_inplace<M>(lu, rhs);
#pragma unroll
    for (int i = 0; i < M; ++i)
#pragma unroll
      for (int j = 0; j < M; ++j) Jm.at(i, j) = rhs(j, i);
  }
  {
    double *__restrict__ d = tJ.at_day(k);
#pragma unroll
    for (int q = 0; q < MM; ++q) d[tJ.f(q)] = Jm.v[q];  // row-major J(i,l) = field i*M + l
  }
  if (P.status && (bad || rank < M)) {
    // max over days of ((M - rank) << 8 | guard-hit); 0 == clean, full rank everywhere
    atomicMax(P.status + b, ((M - rank) << 8) | (bad ? 1 : 0));
  }
}

template <int MODEL>
static void launch_gain_model(const EkfParams &p, cudaStream_t st) {
  if (p.T - p.k0 < 2) return;
  const size_t Bpad = (size_t)((p.B + 31) / 32) * 32;
  const size_t total = (size_t)(p.T - 1 - p.k0) * Bpad;
  const int block = 128;
  const unsigned grid = (unsigned)((total + block - 1) / block);
  if (p.tiled) eks_gain_kernel<MODEL, true><<<grid, block, 0, st>>>(p);
  else eks_gain_kernel<MODEL, false><<<grid, block, 0, st>>>(p);
}

void launch_eks_gain(const EkfParams &p, cudaStream_t st) {
#define CALL(MDL) launch_gain_model<MDL>(p, st)
  EPI_DISPATCH_MODEL(p.model, CALL)
#undef CALL
}

}  // namespace epi
