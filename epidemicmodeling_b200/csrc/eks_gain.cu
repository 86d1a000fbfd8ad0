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

#ifndef EPI_GAIN_BLOCK
#define EPI_GAIN_BLOCK 128
#endif
#ifndef EPI_GAIN_MIN_BLOCKS
#define EPI_GAIN_MIN_BLOCKS (512 / EPI_GAIN_BLOCK)
#endif

// A_k = df/ds at S_PLUS(:,k)  (GenericExtendedKalmanFilter.m:206)
template <int MODEL, bool TILED>
EPI_DI void gain_jacobian(const TrajIn &in, const Tape<TILED> &tSp, int pos, int L, Mat<model_dim(MODEL), false> &A) {
  constexpr int M = model_dim(MODEL);
  const ModelConsts mc = load_consts(in.prm);
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
  state_jacobian<MODEL>(mc, in.eps, sp, a25, A);
}

template <int MODEL, bool TILED>
__global__ void __launch_bounds__(EPI_GAIN_BLOCK, EPI_GAIN_MIN_BLOCKS) eks_gain_kernel(const __grid_constant__ EkfParams P) {
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
  // days of this launch: all of k0 .. T-2, or the [gk_lo, gk_hi) of one time chunk of the piped schedule
  const int klo = (P.gk_hi > P.gk_lo) ? P.gk_lo : k0, khi = (P.gk_hi > P.gk_lo) ? P.gk_hi : T - 1;
  if (tid >= (size_t)(khi - klo) * Bpad) return;
#ifndef EPI_GAIN_TILE_MAJOR
  // consecutive warps (and CTAs) = consecutive days of ONE tile: their tape pages are contiguous in the
  // [tile][day][field][32] scratch layout (6.38 -> 6.20 ms against the tile-major order below)
  const size_t wid = tid >> 5, nd = (size_t)(khi - klo);
  const int k = klo + (int)(wid % nd);
  const int b = (int)((wid / nd) * 32 + (tid & 31));
#else
  const int k = klo + (int)(tid / Bpad);
  const int b = (int)(tid % Bpad);
#endif
  // lanes of this warp that hold a trajectory (the pair-at-a-time pinv is warp-synchronous)
  const unsigned wmask = __ballot_sync(0xffffffffu, b < P.B);
  if (b >= P.B) return;
  const int pos = REV ? (T - 1 - k) : k;
  const int posn = REV ? (T - 2 - k) : (k + 1);
  const TrajIn in = traj_inputs(P, b, M);

  const Tape<TILED> tSp = make_tape<TILED>(P.S_PLUS, M, T - k0, b, k0);
  const Tape<TILED> tPm = make_tape<TILED>(P.P_MINUS, PF, T - k0, b, k0), tPp = make_tape<TILED>(P.P_PLUS, PF, T - k0, b, k0);
  const Tape<true> tJ = make_tape<true>(P.J, MM, T - 1 - k0, b, k0);

  // rotation stack of pinv_sym (generic models only)
  extern __shared__ double rot_stack[];  // generic models: [DS][EPI_GAIN_BLOCK] words (dynamic: 48 KB and more)
  Mat<M, false> A, Jm;
  int rank = M;
  bool bad = false, stored = false;
  if (!LEG) {
    Mat<M, true> Pn, X;
    tape_load_cov<M, true, TILED, PACKED>(Pn, tPm, tPm.at_day(posn));
    {
      // P+ and S+ are read only after the eigen-iteration: pull their lines into L2 now
      const double *__restrict__ dpp = tPp.at_day(pos);
#pragma unroll
      for (int q = 0; q < PF; ++q) asm volatile("prefetch.global.L2 [%0];" ::"l"(dpp + tPp.f(q)));
      const double *__restrict__ dsp = tSp.at_day(pos);
#pragma unroll
      for (int q = 0; q < M; ++q) asm volatile("prefetch.global.L2 [%0];" ::"l"(dsp + tSp.f(q)));
    }
#pragma unroll
    for (int q = 0; q < Mat<M, true>::N; ++q) bad |= !(fabs(Pn.v[q]) <= 1.79769313486231570815e308);  // :211
#if EPI_PINV_MODE == 4 && !defined(EPI_GAIN_SKIP_PINV)
    if constexpr (M == 6) {
      // the matrix goes to this thread's shared-memory column; every lane of the warp enters, a guarded
      // (non-finite) matrix takes no rotation
      double *col = rot_stack + threadIdx.x;
#pragma unroll
      for (int q = 0; q < 21; ++q) col[q * EPI_GAIN_BLOCK] = Pn.v[q];
      rank = pinv_sym6_smem<EPI_GAIN_BLOCK>(col, wmask, bad);
#pragma unroll
      for (int q = 0; q < 21; ++q) X.v[q] = col[q * EPI_GAIN_BLOCK];
      if (bad) rank = M;
    }
#endif
#if (EPI_PINV_MODE == 3 || EPI_PINV_MODE == 5) && !defined(EPI_GAIN_SKIP_PINV)
    if constexpr (M == 6) {
      // every lane of the warp enters; a guarded (non-finite) matrix takes no rotation
      rank = pinv_sym6_pairs<EPI_GAIN_BLOCK, EPI_PINV_MODE == 5>(Pn, X, rot_stack + threadIdx.x, wmask, bad);
      if (bad) rank = M;
    }
#endif
    if (bad) {
#pragma unroll
      for (int q = 0; q < MM; ++q) Jm.v[q] = 0.0;  // :213
    } else {
#ifdef EPI_GAIN_SKIP_PINV  // experiment: cost of everything but the eigen-iteration
      X = Pn;
#else
      if (!(EPI_PINV_MODE >= 3 && M == 6)) rank = pinv_sym<M, EPI_GAIN_BLOCK>(Pn, X, rot_stack + threadIdx.x);
#endif
#ifdef EPI_GAIN_PHASE_SYNC
      __syncthreads();
#endif
      // the state Jacobian is evaluated AFTER the eigen-iteration so that neither its entries nor
      // the model constants are live (or spilled) across it
      gain_jacobian<MODEL, TILED>(in, tSp, pos, L, A);  // :206
      // J = (P+ A') X row by row (:215): only one row of P+ and of P+ A' is live at a time, and
      // the finished row goes straight to the tape (row-major J(i,l) = field i*M + l)
      const double *__restrict__ dp = tPp.at_day(pos);
      double *__restrict__ dj = tJ.at_day(k);
      // row i of the symmetric P+ page for run-time i: element (i,l) is field l*M + i of a full
      // page; of a packed page it is field  i*M - i(i-1)/2 - i + l  for l >= i  and
      // l*M - l(l-1)/2 - l + i  for l < i  -- one run-time base each, the rest immediates
      auto load_row = [&](int i, double(&row)[M]) {
        const double *__restrict__ up = dp + tPp.f(PACKED ? (i * M - (i * (i - 1)) / 2 - i) : i);
        const double *__restrict__ lw = dp + tPp.f(i);
#pragma unroll
        for (int l = 0; l < M; ++l) {
          if (PACKED) row[l] = (l < i) ? lw[tPp.f(l * M - (l * (l - 1)) / 2 - l)] : up[tPp.f(l)];
          else row[l] = up[tPp.f(l * M)];
        }
      };
      double prow[M];
      load_row(0, prow);
#ifndef EPI_GAIN_ROW_UNROLL
#define EPI_GAIN_ROW_UNROLL 2  // row body ~100 instructions; measured kernel ms for 1 / 2 / 3 / 6: 6.49 / 6.39 / 6.59 / 6.38
#endif
      constexpr int kRowUnroll = EPI_GAIN_ROW_UNROLL;
#pragma unroll kRowUnroll
      for (int i = 0; i < M; ++i) {
        double pnext[M], pat[M];
        load_row(i + 1 < M ? i + 1 : i, pnext);  // next row's loads in flight during this row's products
#pragma unroll
        for (int j = 0; j < M; ++j) {
          double acc = 0.0;
          bool first = true;
#pragma unroll
          for (int l = 0; l < M; ++l)
            if (a_nz(M, j, l)) {
              acc = first ? prow[l] * A(j, l) : fma(prow[l], A(j, l), acc);
              first = false;
            }
          pat[j] = acc;
        }
#pragma unroll
        for (int j = 0; j < M; ++j) {
          double acc = pat[0] * X(0, j);
#pragma unroll
          for (int l = 1; l < M; ++l) acc = fma(pat[l], X(l, j), acc);
          dj[tJ.f(i * M + j)] = acc;
        }
#pragma unroll
        for (int l = 0; l < M; ++l) prow[l] = pnext[l];
      }
      stored = true;
    }
  } else {
    // legacy :132   J = (P+ A') / P-   via LU with partial pivoting of (P-)'
    gain_jacobian<MODEL, TILED>(in, tSp, pos, L, A);  // :206
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
  if (!stored) {
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
  const size_t total = (size_t)((p.gk_hi > p.gk_lo) ? (p.gk_hi - p.gk_lo) : (p.T - 1 - p.k0)) * Bpad;
  if (total == 0) return;
  const int block = EPI_GAIN_BLOCK;
  const unsigned grid = (unsigned)((total + block - 1) / block);
  constexpr int M = model_dim(MODEL);
#if EPI_PINV_MODE == 4
  constexpr size_t words6 = SmemPinv6<EPI_GAIN_BLOCK>::SMEM_WORDS;
#elif EPI_PINV_MODE == 5
  constexpr size_t words6 = RotStackU<EPI_GAIN_BLOCK>::SMEM_WORDS + 21 * EPI_GAIN_BLOCK;
#elif EPI_PINV_MODE == 3
  constexpr size_t words6 = RotStackU<EPI_GAIN_BLOCK>::SMEM_WORDS;
#elif EPI_PINV_MODE == 0
  constexpr size_t words6 = RotStack<6, EPI_GAIN_BLOCK>::SMEM_WORDS;
#else
  constexpr size_t words6 = RotStackSel<6, EPI_GAIN_BLOCK>::SMEM_WORDS;
#endif
  const size_t smem = model_legacy(MODEL) ? 0 : sizeof(double) * (M == 6 ? words6 : (size_t)RotStack<M, EPI_GAIN_BLOCK>::SMEM_WORDS);
  if (p.tiled) {
    cudaFuncSetAttribute(eks_gain_kernel<MODEL, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    eks_gain_kernel<MODEL, true><<<grid, block, smem, st>>>(p);
  } else {
    cudaFuncSetAttribute(eks_gain_kernel<MODEL, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    eks_gain_kernel<MODEL, false><<<grid, block, smem, st>>>(p);
  }
}

void launch_eks_gain(const EkfParams &p, cudaStream_t st) {
#define CALL(MDL) launch_gain_model<MDL>(p, st)
  EPI_DISPATCH_MODEL(p.model, CALL)
#undef CALL
}

}  // namespace epi
