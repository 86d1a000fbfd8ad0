// epi_device.cuh -- device-side building blocks shared by the sm_100a kernels.
//
// Arithmetic contract (DESIGN.md "Arithmetic contract"): IEEE binary64, compiled
// with --fmad=false so that no mul+add is contracted implicitly; fma() appears
// only where the contract says a matrix/dot product accumulates
// (first term a*b, further terms fma(a,b,acc), index ascending).  Scalar
// expressions are written in the order MATLAB parses the reference's source
// (left to right, unary minus binds before '*').
//
// The per-step callbacks of the reference (NlinStateUpdate, StateJacobians,
// ObsJacobian, NlinObsUpdate, StateHardMargins, ObsHardMargins; the Hessian
// terms are identically zero) are force-inlined templates on the model variant
// so state, covariance and Jacobian live in registers.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/epi_b200.h"

#define EPI_DI __device__ __forceinline__

namespace epi {

// MATLAB min/max (non-NaN operand wins), written as the same select the oracle uses.
EPI_DI double mmax(double a, double b) { return (b > a || a != a) ? b : a; }
EPI_DI double mmin(double a, double b) { return (b < a || a != a) ? b : a; }

constexpr double kEps = 2.220446049250313e-16;  // MATLAB eps

// ---------------------------------------------------------------------------
// IEEE-correct division by a divisor that is reused (the fading factor gamma: 42 quotients
// per step; the innovation variance: m quotients per step).  With y = RN(1/b) computed once
// by a true division:   q0 = RN(a*y);  r = a - q0*b (exact, one FMA);  q = RN(q0 + r*y)
// is the correctly rounded a/b (Markstein's theorem: y correctly rounded, q0 faithful),
// provided nothing over/underflows -- guarded by exponent-range checks, everything else
// (zeros aside) takes the true division out of line.  3 FP64 instructions instead of the
// ~12-instruction Newton sequence + slow-path call the compiler emits for every a/b.
// Verified against a/b on 6e8 random and near-halfway operands.
// ---------------------------------------------------------------------------
struct InvDiv {
  double b, y;
  bool ok;
};
static __device__ __noinline__ double slow_div(double a, double b) { return a / b; }
EPI_DI InvDiv make_invdiv(double b) {
  InvDiv d;
  d.b = b;
  d.y = 1.0 / b;
  const unsigned e = ((unsigned)__double2hiint(b) >> 20) & 0x7ffu;
  d.ok = (e - 923u) <= 200u;  // 2^-100 <= |b| <= 2^100 (excludes 0, denormals, inf, NaN)
  return d;
}
// Batch form: run many quotients by the same divisor on the fast path unconditionally while
// tracking the exponent range of the numerators; the caller re-does the whole block with true
// divisions in the (rare) case the range check fails.  No per-quotient branch.
struct ExpRange {
  unsigned lo = 2047u, hi = 0u;
  EPI_DI bool safe() const { return lo >= 123u && hi <= 1923u; }  // every |a| in [2^-900, 2^900]
};
EPI_DI double div_fast(double a, const InvDiv &d, ExpRange &r) {
  const unsigned e = ((unsigned)__double2hiint(a) >> 20) & 0x7ffu;
  r.lo = min(r.lo, e);
  r.hi = max(r.hi, e);
  const double q0 = a * d.y;
  const double rem = fma(-q0, d.b, a);
  return fma(rem, d.y, q0);
}
EPI_DI double div_by(double a, const InvDiv &d) {
  const double q0 = a * d.y;
  const double r = fma(-q0, d.b, a);
  const double q = fma(r, d.y, q0);
  const unsigned e = ((unsigned)__double2hiint(a) >> 20) & 0x7ffu;
  if (d.ok && (e - 123u) <= 1800u) return q;  // 2^-900 <= |a| <= 2^900
  if (d.ok && a == 0.0) return q0;            // signed zero of a/b
  return slow_div(a, d.b);
}

__host__ __device__ constexpr int model_dim(int model) { return model >= EPI_MODEL_OPTCTRL ? 6 : 3; }
__host__ __device__ constexpr bool model_flipped(int model) {
  return model == EPI_MODEL_SIALPHA_FLIPPED || model == EPI_MODEL_OPTCTRL_FLIPPED;
}
__host__ __device__ constexpr bool model_legacy(int model) { return model >= EPI_MODEL_LEGACY_TOOLS; }

// structural non-zero pattern of the state Jacobian A
// (SIAlphaModelEKF.m:64-73; SIAlphaModelEKFOptControlled.m:90-132)
__host__ __device__ constexpr bool a_nz(int M, int i, int l) {
  if (M == 3) return (i < 2) ? true : (l == 2);
  return i <= 1   ? (l < 3)
         : i == 2 ? (l == 2 || l == 5)
         : i == 3 ? (l >= 1 && l <= 4)
         : i == 4 ? (l == 0 || (l >= 2 && l <= 4))
                  : (l != 2);
}

// ---------------------------------------------------------------------------
// small dense matrix in registers; SYM = packed upper triangle (exactly
// symmetric matrices only).  All indices fold at compile time after unrolling.
// ---------------------------------------------------------------------------
template <int M, bool SYM>
struct Mat {
  static constexpr int N = SYM ? M * (M + 1) / 2 : M * M;
  double v[N];
  static __host__ __device__ constexpr int idx(int i, int j) {
    return SYM ? (i <= j ? (i * M - i * (i - 1) / 2 + (j - i)) : (j * M - j * (j - 1) / 2 + (i - j)))
               : i * M + j;
  }
  EPI_DI double operator()(int i, int j) const { return v[idx(i, j)]; }
  EPI_DI double &at(int i, int j) { return v[idx(i, j)]; }
};

// the same interface backed by shared memory ([element][thread], conflict-free): used for the
// 6x6 product temporaries of the forward pass to relieve register pressure
template <int M, int STRIDE>
struct SMat {
  double *base;  // this thread's column of a [M*M][STRIDE] buffer
  EPI_DI double operator()(int i, int j) const { return base[(i * M + j) * STRIDE]; }
  EPI_DI double &at(int i, int j) { return base[(i * M + j) * STRIDE]; }
};

// trajectory-minor SoA addressing: X[t][f][b]
struct Soa {
  size_t B;
  int F;
  EPI_DI size_t operator()(int t, int f, size_t b) const { return ((size_t)t * F + f) * B + b; }
};

template <int M, bool SYM>
EPI_DI void load_mat(Mat<M, SYM> &P, const double *__restrict__ base, size_t B) {
  // base points at element [t][0][b]; entry (i,j) of the column-major m x m page is field j*M+i
#pragma unroll
  for (int i = 0; i < M; ++i)
#pragma unroll
    for (int j = (SYM ? i : 0); j < M; ++j) P.at(i, j) = base[(size_t)(j * M + i) * B];
}
template <int M, bool SYM>
EPI_DI void store_mat(const Mat<M, SYM> &P, double *__restrict__ base, size_t B) {
#pragma unroll
  for (int i = 0; i < M; ++i)
#pragma unroll
    for (int j = 0; j < M; ++j) base[(size_t)(j * M + i) * B] = P(i, j);
}

// ---------------------------------------------------------------------------
// model callbacks
// ---------------------------------------------------------------------------

// Loop-invariant scalars of the reference's `params` struct, read once per thread
// (the compiler cannot hoist them itself: tape stores may alias the struct).
struct ModelConsts {
  double dt, beta, gamma, b, alpha_min, alpha_max, s_min, i_min, sigma;
  int obs_type;
  const epi_model_params *p;  // per-NPI tables a, u_min, u_max, w (read-only path)
};
EPI_DI ModelConsts load_consts(const epi_model_params *__restrict__ p) {
  ModelConsts c;
  c.dt = __ldg(&p->dt); c.beta = __ldg(&p->beta); c.gamma = __ldg(&p->gamma); c.b = __ldg(&p->b);
  c.alpha_min = __ldg(&p->alpha_min); c.alpha_max = __ldg(&p->alpha_max);
  c.s_min = __ldg(&p->s_min); c.i_min = __ldg(&p->i_min); c.sigma = __ldg(&p->sigma);
  c.obs_type = __ldg(&p->obs_type);
  c.p = p;
  return c;
}

// StateHardMargins: SIAlphaModelEKF.m:27-31 (s_min/i_min floors); every other
// model clamps s,i to [0,1] (SIAlphaModelBackwardEKF.m:48-52,
// SIAlphaModelEKFOptControlled.m:27-31, NewCaseEKFEstimatorWithOptimalNPI.m:150-154).
template <int MODEL>
EPI_DI void state_margins(const ModelConsts &c, double *s) {
  const double lo_s = (MODEL == EPI_MODEL_SIALPHA) ? c.s_min : 0.0;
  const double lo_i = (MODEL == EPI_MODEL_SIALPHA) ? c.i_min : 0.0;
  s[0] = mmin(1.0, mmax(lo_s, s[0]));
  s[1] = mmin(1.0, mmax(lo_i, s[1]));
  s[2] = mmin(c.alpha_max, mmax(c.alpha_min, s[2]));
}

// One pass over the L NPI inputs of a day.  Fuses what the reference does in
// NlinStateUpdate (bang-bang substitution of NaN inputs,
// SIAlphaModelEKFOptControlled.m:49-58, and the gamma*a'*(u_max-u) term, :67)
// and in StateJacobians (the A(3,6) slope, :107-114): both evaluate the same
// phi at the same state, so one evaluation serves both.
struct InputPass {
  double dot;   // gamma * a' * (u_max - u_opt)
  double a25;   // A(3,6)
  double cost;  // sum_j w_day[j] * u_opt[j]   (NPICost's weights.*inputs for this day)
};
// PRELOAD: the day's inputs and weights are read into registers before the first schedule entry is stored (the stores
// of u_out otherwise fence the loads into batches: three exposed latencies per day in the small-batch recursion, ncu)
template <int MODEL, bool WANT_A25, bool WANT_COST, bool PRELOAD = false>
EPI_DI InputPass input_pass(const ModelConsts &c, double eps, double s5,
                            const double *__restrict__ u, size_t u_stride, int L,
                            double *__restrict__ u_out, size_t uo_stride,
                            const double *__restrict__ w_day) {
  constexpr bool SIX = model_dim(MODEL) == 6;
  constexpr bool FLIP = model_flipped(MODEL);
  const epi_model_params *__restrict__ p = c.p;
  InputPass r;
  r.dot = 0.0; r.a25 = 0.0; r.cost = 0.0;
  const double gamma = c.gamma;
  const double gs = SIX ? gamma * s5 : 0.0;
  const double lo = SIX ? (-1.0) / c.sigma : 0.0, hi = SIX ? 1.0 / c.sigma : 0.0;
  const double slope = (SIX && WANT_A25) ? (gamma * c.dt) * (c.sigma / 2.0) : 0.0;
  // Branch-free per NPI: the constants of the bang-bang rule are loaded (read-only path) whether or
  // not the input turns out to be missing, so no load waits behind the NaN test of the previous NPI --
  // fetched inside that branch they were a chain of twelve exposed memory latencies on every day to
  // optimise (ncu r01) -- and the compiler is free to batch them as far as registers allow.
  double u_in[PRELOAD ? EPI_LMAX : 1], w_in[(PRELOAD && WANT_COST) ? EPI_LMAX : 1];
  if (PRELOAD) {
#pragma unroll
    for (int j = 0; j < EPI_LMAX; ++j)
      if (j < L) {
        u_in[j] = __ldg(u + (size_t)j * u_stride);
        if (WANT_COST) w_in[j] = __ldg(w_day + j);
      }
  }
#pragma unroll
  for (int j = 0; j < EPI_LMAX; ++j) {
    if (j < L) {
      double uj = PRELOAD ? u_in[j] : u[(size_t)j * u_stride];
      const double aj = __ldg(&p->a[j]);
      const double umax = __ldg(&p->u_max[j]);
      if (SIX) {
        const double wj = __ldg(&p->w[j]);
        const double umin = __ldg(&p->u_min[j]);
        const bool opt = (uj != uj);
        const double phi = eps * wj - gs * aj;
        if (WANT_A25) {
          const double term = (slope * aj) * (umax - umin);
          const double a25n = FLIP ? (r.a25 + term) : (r.a25 - term);
          r.a25 = (opt && phi > lo && phi < hi) ? a25n : r.a25;
        }
        const bool to_min = model_legacy(MODEL) ? (phi >= 0.0) : (phi > 0.0);
        uj = opt ? (to_min ? umin : umax) : uj;
      }
      const double g = gamma * aj;
      const double d = umax - uj;
      r.dot = (j == 0) ? g * d : fma(g, d, r.dot);
      if (WANT_COST) {
        const double wu = ((PRELOAD && WANT_COST) ? w_in[j] : w_day[j]) * uj;
        r.cost = (j == 0) ? wu : (r.cost + wu);
      }
      if (u_out) u_out[(size_t)j * uo_stride] = uj;
    }
  }
  return r;
}

// state equations of NlinStateUpdate given the input term `dot`
// (SIAlphaModelEKF.m:44-46, SIAlphaModelBackwardEKF.m:65-67,
//  SIAlphaModelEKFOptControlled.m:60-72, ...BackwardEKFOptControlled.m:81-93)
template <int MODEL>
EPI_DI void state_eqs(const ModelConsts &c, double eps, const double *s, double dot, double *sn) {
  constexpr bool SIX = model_dim(MODEL) == 6;
  constexpr bool FLIP = model_flipped(MODEL);
  const double dt = c.dt, beta = c.beta, gamma = c.gamma;
  const double f2 = (((-gamma) * s[2]) + gamma * c.b) + dot;
  double x0, x1, x2;
  if (!FLIP) {
    x0 = s[0] - ((dt * s[2]) * s[0]) * s[1];
    x1 = s[1] + dt * (((s[2] * s[0]) * s[1]) - beta * s[1]);
    x2 = s[2] + dt * f2;
  } else {
    x0 = s[0] + ((dt * s[2]) * s[0]) * s[1];
    x1 = s[1] - dt * (((s[2] * s[0]) * s[1]) - beta * s[1]);
    x2 = s[2] - dt * f2;
  }
  const double lo_s = (MODEL == EPI_MODEL_SIALPHA) ? c.s_min : 0.0;
  const double lo_i = (MODEL == EPI_MODEL_SIALPHA) ? c.i_min : 0.0;
  sn[0] = mmax(lo_s, mmin(1.0, x0));
  sn[1] = mmax(lo_i, mmin(1.0, x1));
  sn[2] = mmax(c.alpha_min, mmin(c.alpha_max, x2));
  if (SIX) {
    const double rho = (s[3] - s[4]) - (1.0 - eps);
    if (!FLIP) {
      sn[3] = s[3] + ((dt * rho) * s[2]) * s[1];
      sn[4] = s[4] + dt * (((rho * s[2]) * s[0]) + beta * s[4]);
      sn[5] = s[5] + dt * (((rho * s[0]) * s[1]) + gamma * s[5]);
    } else {
      sn[3] = s[3] - ((dt * rho) * s[2]) * s[1];
      sn[4] = s[4] - dt * (((rho * s[2]) * s[0]) + beta * s[4]);
      sn[5] = s[5] - dt * (((rho * s[0]) * s[1]) + gamma * s[5]);
    }
  }
}

// StateJacobians (SIAlphaModelEKF.m:62-76, SIAlphaModelBackwardEKF.m:83-97,
// SIAlphaModelEKFOptControlled.m:88-135, ...Backward...:109-156,
// NewCaseEKFEstimatorWithOptimalNPI.m:211-257).  Only the structural non-zeros
// (a_nz) are written; B = I is structural.
template <int MODEL>
EPI_DI void state_jacobian(const ModelConsts &c, double eps, const double *s, double a25,
                           Mat<model_dim(MODEL), false> &A) {
  constexpr bool SIX = model_dim(MODEL) == 6;
  constexpr bool FLIP = model_flipped(MODEL);
  const double dt = c.dt, beta = c.beta, gamma = c.gamma;
  if (!FLIP) {
    A.at(0, 0) = 1.0 - (dt * s[2]) * s[1];
    A.at(0, 1) = ((-dt) * s[2]) * s[0];
    A.at(0, 2) = ((-dt) * s[0]) * s[1];
    A.at(1, 0) = (dt * s[1]) * s[2];
    A.at(1, 1) = 1.0 + dt * (s[0] * s[2] - beta);
    A.at(1, 2) = (dt * s[0]) * s[1];
    A.at(2, 2) = 1.0 - dt * gamma;
  } else {
    A.at(0, 0) = 1.0 + (dt * s[2]) * s[1];
    A.at(0, 1) = (dt * s[2]) * s[0];
    A.at(0, 2) = (dt * s[0]) * s[1];
    A.at(1, 0) = ((-dt) * s[1]) * s[2];
    A.at(1, 1) = 1.0 - dt * (s[0] * s[2] - beta);
    A.at(1, 2) = ((-dt) * s[0]) * s[1];
    A.at(2, 2) = 1.0 + dt * gamma;
  }
  if (SIX) {
    A.at(2, 5) = a25;
    const double rho = (s[3] - s[4]) - (1.0 - eps);
    if (!FLIP) {
      A.at(3, 1) = (dt * s[2]) * rho;
      A.at(3, 2) = (dt * s[1]) * rho;
      A.at(3, 3) = 1.0 + (dt * s[1]) * s[2];
      A.at(3, 4) = ((-dt) * s[1]) * s[2];
      A.at(4, 0) = (dt * s[2]) * rho;
      A.at(4, 2) = (dt * s[0]) * rho;
      A.at(4, 3) = (dt * s[0]) * s[2];
      A.at(4, 4) = 1.0 - dt * (s[0] * s[2] - beta);
      A.at(5, 0) = (dt * s[1]) * rho;
      A.at(5, 1) = (dt * s[0]) * rho;
      A.at(5, 3) = (dt * s[0]) * s[1];
      A.at(5, 4) = ((-dt) * s[0]) * s[1];
      A.at(5, 5) = 1.0 + dt * gamma;
    } else {
      A.at(3, 1) = ((-dt) * s[2]) * rho;
      A.at(3, 2) = ((-dt) * s[1]) * rho;
      A.at(3, 3) = 1.0 - (dt * s[1]) * s[2];
      A.at(3, 4) = (dt * s[1]) * s[2];
      A.at(4, 0) = ((-dt) * s[2]) * rho;
      A.at(4, 2) = ((-dt) * s[0]) * rho;
      A.at(4, 3) = ((-dt) * s[0]) * s[2];
      A.at(4, 4) = 1.0 + dt * (s[0] * s[2] - beta);
      A.at(5, 0) = ((-dt) * s[1]) * rho;
      A.at(5, 1) = ((-dt) * s[0]) * rho;
      A.at(5, 3) = ((-dt) * s[0]) * s[1];
      A.at(5, 4) = (dt * s[0]) * s[1];
      A.at(5, 5) = 1.0 - dt * gamma;
    }
  }
}

// ObsJacobian + NlinObsUpdate + ObsHardMargins (SIAlphaModelEKF.m:34-36,51-59,79-89;
// MatlabCodeGenerator/ObsHardMargins.m:2-4 is the identity and
// MatlabCodeGenerator/NlinObsUpdate.m:3 knows NEWCASES only).
template <int MODEL>
EPI_DI double obs_model(int obs_type, const double *s, double v_bar, double *C) {
  double xh;
  if (MODEL == EPI_MODEL_LEGACY_CODEGEN || obs_type == EPI_OBS_NEWCASES) {
    C[0] = s[1] * s[2];
    C[1] = s[0] * s[2];
    C[2] = s[0] * s[1];
    xh = (s[0] * s[1]) * s[2] + v_bar;
  } else {
    C[0] = -1.0; C[1] = 0.0; C[2] = 0.0;
    xh = (1.0 - s[0]) + v_bar;
  }
  if (MODEL != EPI_MODEL_LEGACY_CODEGEN) xh = mmax(0.0, xh);
  return xh;
}

// ---------------------------------------------------------------------------
// structured products (the summation order is part of the arithmetic contract)
// ---------------------------------------------------------------------------
// R = A * P, A with pattern a_nz
template <int M, bool SYMP, class RMat>
EPI_DI void mul_A_P(const Mat<M, false> &A, const Mat<M, SYMP> &P, RMat &R) {
#pragma unroll
  for (int i = 0; i < M; ++i)
#pragma unroll
    for (int j = 0; j < M; ++j) {
      double acc = 0.0;
      bool first = true;
#pragma unroll
      for (int l = 0; l < M; ++l)
        if (a_nz(M, i, l)) {
          acc = first ? A(i, l) * P(l, j) : fma(A(i, l), P(l, j), acc);
          first = false;
        }
      R.at(i, j) = acc;
    }
}
// (X * A')(i,j) = sum_{l in nz(A row j)} X(i,l) A(j,l)
template <int M, class XMat>
EPI_DI double mul_X_At_ij(const XMat &X, const Mat<M, false> &A, int i, int j) {
  double acc = 0.0;
  bool first = true;
#pragma unroll
  for (int l = 0; l < M; ++l)
    if (a_nz(M, j, l)) {
      acc = first ? X(i, l) * A(j, l) : fma(X(i, l), A(j, l), acc);
      first = false;
    }
  return acc;
}

// Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3",
// SC'11): ten rounds of two 32x32->64 multiplies and a key schedule.
struct Philox4 { unsigned v[4]; };
EPI_DI Philox4 philox4x32_10(unsigned c0, unsigned c1, unsigned c2, unsigned c3, unsigned k0, unsigned k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const unsigned hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const unsigned n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  Philox4 o;
  o.v[0] = c0; o.v[1] = c1; o.v[2] = c2; o.v[3] = c3;
  return o;
}
// the random-schedule rule of include/epi_b200.h (EPI_U_PHILOX): levels of NPIs 4q..4q+3 on `day`
// for 0-based scenario `sc` of `region`; held scenarios (2*(sc+1) < G) use day 0
EPI_DI bool schedule_held(long long sc, int G) { return 2 * (sc + 1) < (long long)G; }
EPI_DI unsigned level_from_word(unsigned word, int lo, int hi) {
  return (unsigned)lo + __umulhi(word, (unsigned)(hi - lo + 1));
}

// spacing of doubles at |x| (MATLAB eps(x)), normal x
EPI_DI double eps_of(double x) {
  const unsigned long long u = (unsigned long long)__double_as_longlong(fabs(x));
  const unsigned long long e = (u >> 52) & 0x7ffull;
  if (e == 0x7ffull) return __longlong_as_double(0x7ff8000000000000ll);
  if (e <= 52) return 4.9406564584124654e-324;
  return __longlong_as_double((long long)((e - 52) << 52));
}

}  // namespace epi
