// rt_expfit.cu -- Tools/Rt_ExpFitEKF.m (sm_100a, FP64, --fmad=false): the 2-state
// exponential-fit EKF + fixed-interval smoother with second-order (Hessian) terms, one thread
// per trajectory (call site testScripts/test04FullFeatureExtMLpipeline.m:217-219).
// State s = [new cases; growth exponent lambda]; f(s) = [s1 exp(ts s2) + w1; sigma tanh((alpha s2 + w2)/sigma)];
// observation x = s1 + v.  Unlike the generic filter this function keeps the simple covariance
// update, adapts a scalar R, and uses mrdivide in the smoother (Rt_ExpFitEKF.m:59,100,110).
// Arithmetic = oracle orc_rt_expfit_ekf operation for operation; exp/tanh are the CUDA math
// library's (not correctly rounded on either side) => tolerance parity, not bitwise.
#include "ekf_common.cuh"

namespace epi {

struct M2 { double v[4]; };  // row-major {m11, m12, m21, m22}
EPI_DI M2 mm2(const M2 &a, const M2 &b) {
  M2 c;
  c.v[0] = fma(a.v[1], b.v[2], a.v[0] * b.v[0]); c.v[1] = fma(a.v[1], b.v[3], a.v[0] * b.v[1]);
  c.v[2] = fma(a.v[3], b.v[2], a.v[2] * b.v[0]); c.v[3] = fma(a.v[3], b.v[3], a.v[2] * b.v[1]);
  return c;
}
EPI_DI M2 tr2(const M2 &a) { M2 t; t.v[0] = a.v[0]; t.v[1] = a.v[2]; t.v[2] = a.v[1]; t.v[3] = a.v[3]; return t; }
EPI_DI double half_trace2(const M2 &P, const M2 &F) { const M2 a = mm2(P, F); return (a.v[0] + a.v[3]) / 2.0; }
EPI_DI double half_trace4(const M2 &P, const M2 &Fi, const M2 &Fj) {
  const M2 c = mm2(mm2(mm2(P, Fi), P), Fj);
  return (c.v[0] + c.v[3]) / 2.0;
}

// element (t, f) of a [T][F][B] array
EPI_DI double *at(const TArr &a, int F, int t, int f, int b) {
  return a.p + ((size_t)t * F + f) * (size_t)a.stride + (size_t)a.off + b;
}

__global__ void __launch_bounds__(64) rt_expfit_kernel(const __grid_constant__ RtParams P) {
  extern __shared__ double win[];  // [3][W][blockDim.x] ring buffers, newest at `head`
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= P.B) return;
  const long long g = (P.b0 + b) / P.G;
  const int T = P.T, W = P.W;
  const double ts = P.params[3 * g + 0], alpha = P.params[3 * g + 1], sigma = P.params[3 * g + 2];
  const double w1 = P.w_bar[2 * g + 0], w2 = P.w_bar[2 * g + 1];
  const double gamma = P.gamma, beta = P.beta, v_bar = P.v_bar;
  double R = P.R[g];
  // pages are column-major in the ABI (field j*2 + i), row-major in registers
  M2 Q, Pm;
  Q.v[0] = P.Q[4 * g + 0]; Q.v[1] = P.Q[4 * g + 2]; Q.v[2] = P.Q[4 * g + 1]; Q.v[3] = P.Q[4 * g + 3];
  Pm.v[0] = P.Ps_init[4 * g + 0]; Pm.v[1] = P.Ps_init[4 * g + 2]; Pm.v[2] = P.Ps_init[4 * g + 1]; Pm.v[3] = P.Ps_init[4 * g + 3];
  double sm0 = P.s_init.p[P.s_init.off + b], sm1 = P.s_init.p[(size_t)P.s_init.stride + P.s_init.off + b];
  const size_t bs = blockDim.x;
  double *wm = win + threadIdx.x;
  for (int j = 0; j < 3 * W; ++j) wm[(size_t)j * bs] = 0.0;
  int head = 0;

  for (int k = 0; k < T; ++k) {
    *at(P.S_MINUS, 2, k, 0, b) = sm0; *at(P.S_MINUS, 2, k, 1, b) = sm1;  // :37-38
    *at(P.P_MINUS, 4, k, 0, b) = Pm.v[0]; *at(P.P_MINUS, 4, k, 1, b) = Pm.v[2];
    *at(P.P_MINUS, 4, k, 2, b) = Pm.v[1]; *at(P.P_MINUS, 4, k, 3, b) = Pm.v[3];
    const double xhat = (sm0 + v_bar) + 0.0 + 0.0;  // :53 (observation Hessian terms are identically zero)
    const double xk = P.x.p[(size_t)k * P.x.stride + P.x.off + b];
    const bool valid = !(xk != xk);  // :56
    double innov, K0, K1, sp0, sp1;
    M2 Pp;
    if (valid) {
      innov = xk - xhat;
      const double denom = ((Pm.v[0] + gamma * ((1.0 * R) * 1.0)) + 0.0) + 0.0;  // :58
      K0 = Pm.v[0] / denom; K1 = Pm.v[2] / denom;
      M2 Mx;
      Mx.v[0] = 1.0 - K0; Mx.v[1] = 0.0 - 0.0; Mx.v[2] = 0.0 - K1; Mx.v[3] = 1.0 - 0.0;
      const M2 MP = mm2(Mx, Pm);
#pragma unroll
      for (int q = 0; q < 4; ++q) Pp.v[q] = MP.v[q] / gamma;  // :59
      sp0 = sm0 + K0 * innov; sp1 = sm1 + K1 * innov;        // :60
    } else {  // :62-65
      innov = 0.0; K0 = 0.0; K1 = 0.0; Pp = Pm; sp0 = sm0; sp1 = sm1;
    }
    const double E = exp(ts * sp1);
    const double tnh = tanh((alpha * sp1 + w2) / sigma);
    double fs0 = 0.0, fs1 = 0.0, fw0 = 0.0, fw1 = 0.0;
    M2 Fsp, Fwp;
#pragma unroll
    for (int q = 0; q < 4; ++q) { Fsp.v[q] = 0.0; Fwp.v[q] = 0.0; }
    if (P.order == 2) {  // :153-196
      const double f12 = ts * E;
      M2 Fs[2], Fw[2];
      Fs[0].v[0] = 0.0; Fs[0].v[1] = f12; Fs[0].v[2] = f12; Fs[0].v[3] = ((ts * ts) * sp0) * E;
      Fs[1].v[0] = 0.0; Fs[1].v[1] = 0.0; Fs[1].v[2] = 0.0;
      Fs[1].v[3] = ((((-2.0) * (alpha * alpha)) / sigma) * tnh) * (1.0 - tnh * tnh);
      Fw[0].v[0] = 0.0; Fw[0].v[1] = 0.0; Fw[0].v[2] = 0.0; Fw[0].v[3] = 0.0;
      Fw[1].v[0] = 0.0; Fw[1].v[1] = 0.0; Fw[1].v[2] = 0.0; Fw[1].v[3] = (((-2.0) / sigma) * tnh) * (1.0 - tnh * tnh);
      fs0 = half_trace2(Pp, Fs[0]); fs1 = half_trace2(Pp, Fs[1]);
      fw0 = half_trace2(Q, Fw[0]); fw1 = half_trace2(Q, Fw[1]);
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          Fsp.v[2 * i + j] = half_trace4(Pp, Fs[i], Fs[j]);
          Fwp.v[2 * i + j] = half_trace4(Q, Fw[i], Fw[j]);
        }
    }
    sm0 = ((sp0 * E + w1) + fs0) + fw0;  // :80
    sm1 = ((sigma * tnh) + fs1) + fw1;
    const double omt = 1.0 - tnh * tnh;
    M2 A, Bm;
    A.v[0] = E; A.v[1] = (ts * sp0) * E; A.v[2] = 0.0; A.v[3] = alpha * omt;  // :139-150
    Bm.v[0] = 1.0; Bm.v[1] = 0.0; Bm.v[2] = 0.0; Bm.v[3] = omt;
    const M2 APA = mm2(mm2(A, Pp), tr2(A)), BQB = mm2(mm2(Bm, Q), tr2(Bm));
#pragma unroll
    for (int q = 0; q < 4; ++q) Pm.v[q] = ((APA.v[q] + BQB.v[q]) + Fsp.v[q]) + Fwp.v[q];  // :82
    *at(P.S_PLUS, 2, k, 0, b) = sp0; *at(P.S_PLUS, 2, k, 1, b) = sp1;  // :85-87
    *at(P.P_PLUS, 4, k, 0, b) = Pp.v[0]; *at(P.P_PLUS, 4, k, 1, b) = Pp.v[2];
    *at(P.P_PLUS, 4, k, 2, b) = Pp.v[1]; *at(P.P_PLUS, 4, k, 3, b) = Pp.v[3];
    if (P.K_GAIN.p) { *at(P.K_GAIN, 2, k, 0, b) = K0; *at(P.K_GAIN, 2, k, 1, b) = K1; }
    if (P.innov.p) *at(P.innov, 1, k, 0, b) = innov;
    // :90-101 innovation monitor (ring buffers summed newest -> oldest over all W slots)
    const int cnt = (k + 1 < W) ? (k + 1) : W;
    head = (head == 0) ? (W - 1) : (head - 1);
    wm[(size_t)(0 * W + head) * bs] = innov;
    double sM = 0.0;
    for (int j = 0, q = head; j < W; ++j) { sM += wm[(size_t)(0 * W + q) * bs]; q = (q + 1 == W) ? 0 : q + 1; }
    const double mu = sM / (double)cnt;
    const double cc = (innov - mu) * (innov - mu);
    wm[(size_t)(1 * W + head) * bs] = cc;
    wm[(size_t)(2 * W + head) * bs] = cc / R;  // :97
    double sN = 0.0;
    for (int j = 0, q = head; j < W; ++j) { sN += wm[(size_t)(2 * W + q) * bs]; q = (q + 1 == W) ? 0 : q + 1; }
    if (P.rho.p) *at(P.rho, 1, k, 0, b) = sN / (double)cnt;
    if (beta != 1.0 && valid) {  // :99-101
      double sC = 0.0;
      for (int j = 0, q = head; j < W; ++j) { sC += wm[(size_t)(1 * W + q) * bs]; q = (q + 1 == W) ? 0 : q + 1; }
      R = beta * R + ((1.0 - beta) * sC) / (double)cnt;
    }
  }

  // :104-115 smoother over this thread's own tape
  if (!P.S_SMOOTH.p || T < 1) return;
  double ss0 = *at(P.S_PLUS, 2, T - 1, 0, b), ss1 = *at(P.S_PLUS, 2, T - 1, 1, b);
  M2 Ps;
  Ps.v[0] = *at(P.P_PLUS, 4, T - 1, 0, b); Ps.v[2] = *at(P.P_PLUS, 4, T - 1, 1, b);
  Ps.v[1] = *at(P.P_PLUS, 4, T - 1, 2, b); Ps.v[3] = *at(P.P_PLUS, 4, T - 1, 3, b);
  *at(P.S_SMOOTH, 2, T - 1, 0, b) = ss0; *at(P.S_SMOOTH, 2, T - 1, 1, b) = ss1;
  if (P.P_SMOOTH.p) {
    *at(P.P_SMOOTH, 4, T - 1, 0, b) = Ps.v[0]; *at(P.P_SMOOTH, 4, T - 1, 1, b) = Ps.v[2];
    *at(P.P_SMOOTH, 4, T - 1, 2, b) = Ps.v[1]; *at(P.P_SMOOTH, 4, T - 1, 3, b) = Ps.v[3];
  }
  for (int k = T - 2; k >= 0; --k) {
    const double sp0 = *at(P.S_PLUS, 2, k, 0, b), sp1 = *at(P.S_PLUS, 2, k, 1, b);
    const double sn0 = *at(P.S_MINUS, 2, k + 1, 0, b), sn1 = *at(P.S_MINUS, 2, k + 1, 1, b);
    M2 Pp, Pn;
    Pp.v[0] = *at(P.P_PLUS, 4, k, 0, b); Pp.v[2] = *at(P.P_PLUS, 4, k, 1, b);
    Pp.v[1] = *at(P.P_PLUS, 4, k, 2, b); Pp.v[3] = *at(P.P_PLUS, 4, k, 3, b);
    Pn.v[0] = *at(P.P_MINUS, 4, k + 1, 0, b); Pn.v[2] = *at(P.P_MINUS, 4, k + 1, 1, b);
    Pn.v[1] = *at(P.P_MINUS, 4, k + 1, 2, b); Pn.v[3] = *at(P.P_MINUS, 4, k + 1, 3, b);
    const double E = exp(ts * sp1);
    const double tnh = tanh((alpha * sp1 + w2) / sigma);
    M2 A;
    A.v[0] = E; A.v[1] = (ts * sp0) * E; A.v[2] = 0.0; A.v[3] = alpha * (1.0 - tnh * tnh);
    const M2 PAt = mm2(Pp, tr2(A));
    // :110  J = (P+ A') / P-  as defined in DESIGN.md: LU with partial pivoting of (P-)'
    Mat<2, false> lu, rhs;
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j) { lu.at(i, j) = Pn.v[2 * j + i]; rhs.at(i, j) = PAt.v[2 * j + i]; }
    lu_solve_inplace<2>(lu, rhs);
    M2 J;
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j) J.v[2 * i + j] = rhs(j, i);
    const double d0 = ss0 - sn0, d1 = ss1 - sn1;
    const double n0 = sp0 + fma(J.v[1], d1, J.v[0] * d0), n1 = sp1 + fma(J.v[3], d1, J.v[2] * d0);  // :111
    M2 D;
#pragma unroll
    for (int q = 0; q < 4; ++q) D.v[q] = Pn.v[q] - Ps.v[q];
    const M2 JDJ = mm2(mm2(J, D), tr2(J));
#pragma unroll
    for (int q = 0; q < 4; ++q) Ps.v[q] = Pp.v[q] - JDJ.v[q];  // :112
    ss0 = n0; ss1 = n1;
    *at(P.S_SMOOTH, 2, k, 0, b) = ss0; *at(P.S_SMOOTH, 2, k, 1, b) = ss1;
    if (P.P_SMOOTH.p) {
      *at(P.P_SMOOTH, 4, k, 0, b) = Ps.v[0]; *at(P.P_SMOOTH, 4, k, 1, b) = Ps.v[2];
      *at(P.P_SMOOTH, 4, k, 2, b) = Ps.v[1]; *at(P.P_SMOOTH, 4, k, 3, b) = Ps.v[3];
    }
  }
}

void launch_rt_expfit(const RtParams &p, cudaStream_t st) {
  if (p.B <= 0 || p.T <= 0) return;
  const int block = 64;
  const size_t smem = (size_t)3 * p.W * block * sizeof(double);
  cudaFuncSetAttribute(rt_expfit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  rt_expfit_kernel<<<(p.B + block - 1) / block, block, smem, st>>>(p);
}

}  // namespace epi
