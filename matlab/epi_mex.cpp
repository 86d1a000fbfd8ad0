// epi_mex.cpp -- MATLAB/Octave MEX gateway of libepi_b200.so.
//
// One gateway, dispatched by a leading command string; the .m shims in this
// directory carry the reference's exact function names and signatures and call
//     [outs...] = epi_mex('<command>', args...)
// This file is a pure marshalling layer over the tested C ABI
// (include/epi_b200.h): no arithmetic happens here.  It CANNOT be compiled in
// the build image (no mex.h / mkoctfile); build it on a MATLAB/Octave host with
//     mex  -I../include epi_mex.cpp -L../epidemicmodeling_b200 -lepi_b200
//     mkoctfile --mex -I../include epi_mex.cpp -L../epidemicmodeling_b200 -lepi_b200
//
// Ownership: prhs is read-only and never retained; every plhs is allocated with
// mxCreate* (MATLAB frees).  The library context is created on first use and
// destroyed by mexAtExit.  A non-zero status becomes mexErrMsgIdAndTxt("epi:...").
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "mex.h"

#include "epi_b200.h"

static epi_ctx *g_ctx = nullptr;

// contexts on the other GPUs of the box (g_ctx is GPU 0): created on first use by the multi-GPU sweep
static std::vector<epi_ctx *> g_peers;
static void at_exit() {
  for (epi_ctx *p : g_peers) epi_destroy(p);
  g_peers.clear();
  if (g_ctx) { epi_destroy(g_ctx); g_ctx = nullptr; }
}
static epi_ctx *ctx() {
  if (!g_ctx) {
    int rc = epi_create(0, &g_ctx);
    if (rc != EPI_OK) mexErrMsgIdAndTxt("epi:create", "%s", epi_last_error(nullptr));
    mexAtExit(at_exit);
  }
  return g_ctx;
}
// [ctx 0, ctx 1, ...] for up to n_gpus GPUs (n_gpus <= 0: every GPU epi_create accepts)
static std::vector<epi_ctx *> ctxs(int n_gpus) {
  std::vector<epi_ctx *> v(1, ctx());
  for (int d = 1; n_gpus <= 0 || d < n_gpus; ++d) {
    if ((size_t)d > g_peers.size()) {
      epi_ctx *c = nullptr;
      if (epi_create(d, &c) != EPI_OK) {
        if (n_gpus > 0) mexErrMsgIdAndTxt("epi:create", "GPU %d: %s", d, epi_last_error(nullptr));
        break;  // no more devices
      }
      g_peers.push_back(c);
    }
    v.push_back(g_peers[(size_t)d - 1]);
  }
  return v;
}
static void check(int rc) {
  if (rc == EPI_OK) return;
  const char *id = rc == EPI_ERR_ORDER ? "epi:order" : rc == EPI_ERR_OBS_TYPE ? "epi:obsType"
                 : rc == EPI_ERR_QR_SHAPE ? "epi:covShape" : rc == EPI_ERR_ARG ? "epi:arg" : "epi:cuda";
  mexErrMsgIdAndTxt(id, "%s", epi_last_error(g_ctx));
}
static double scalar(const mxArray *a) { return mxGetScalar(a); }
static const double *dbl(const mxArray *a) {
  if (!mxIsDouble(a) || mxIsComplex(a)) mexErrMsgIdAndTxt("epi:arg", "real double arrays expected");
  return mxGetPr(a);
}
static double field_scalar(const mxArray *s, const char *name, double dflt) {
  const mxArray *f = mxGetField(s, 0, name);
  return (f && !mxIsEmpty(f)) ? mxGetScalar(f) : dflt;
}
static void field_vec(const mxArray *s, const char *name, double *dst, int L) {
  for (int j = 0; j < EPI_LMAX; ++j) dst[j] = mxGetNaN();
  const mxArray *f = mxGetField(s, 0, name);
  if (!f || mxIsEmpty(f)) return;
  const double *p = dbl(f);
  const size_t n = mxGetNumberOfElements(f);
  // a 12 x T matrix keeps column 1 only: phi(kk) linear-indexes it
  // (SIAlphaModelEKFOptControlled.m:49-52); column-major => the first L entries
  for (int j = 0; j < L; ++j) dst[j] = (n == 1) ? p[0] : p[j];
}
// the reference's `params` struct -> epi_model_params
static epi_model_params to_params(const mxArray *s, int L) {
  if (!mxIsStruct(s)) mexErrMsgIdAndTxt("epi:arg", "params must be a struct");
  epi_model_params p;
  std::memset(&p, 0, sizeof p);
  const double nan = mxGetNaN();
  p.dt = field_scalar(s, "dt", nan); p.beta = field_scalar(s, "beta", nan);
  p.gamma = field_scalar(s, "gamma", nan); p.b = field_scalar(s, "b", 0.0);
  p.alpha_min = field_scalar(s, "alpha_min", nan); p.alpha_max = field_scalar(s, "alpha_max", nan);
  p.s_min = field_scalar(s, "s_min", 0.0); p.i_min = field_scalar(s, "i_min", 0.0);
  p.epsilon = field_scalar(s, "epsilon", nan); p.sigma = field_scalar(s, "sigma", nan);
  field_vec(s, "a", p.a, L); field_vec(s, "u_min", p.u_min, L);
  field_vec(s, "u_max", p.u_max, L); field_vec(s, "w", p.w, L);
  p.L = L;
  p.obs_type = EPI_OBS_NEWCASES;
  const mxArray *ot = mxGetField(s, 0, "obs_type");
  if (ot && mxIsChar(ot)) {
    char buf[32];
    mxGetString(ot, buf, sizeof buf);
    if (!std::strcmp(buf, "NEWCASES")) p.obs_type = EPI_OBS_NEWCASES;
    else if (!std::strcmp(buf, "TOTALCASES")) p.obs_type = EPI_OBS_TOTALCASES;
    else mexErrMsgIdAndTxt("epi:obsType", "unknown observation type");  // SIAlphaModelEKF.m:57
  }
  return p;
}

// [s,e,i,r,p] = epi_mex('seirp', saturated, rates(7xK), ic(5x1), K, dt, sat(6x1))
static void cmd_seirp(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]) {
  if (nrhs < 6) mexErrMsgIdAndTxt("epi:arg", "seirp: 6 arguments expected");
  epi_seirp_args a;
  std::memset(&a, 0, sizeof a);
  a.mem = EPI_MEM_HOST; a.B = 1; a.K = (int)scalar(prhs[3]); a.dt = scalar(prhs[4]);
  a.saturated = (int)scalar(prhs[0]);
  a.rate_mode = EPI_RATES_SHARED_SERIES;
  // rates arrive as K x 7 (column-major) == [7][K] row-major, the ABI's SHARED_SERIES layout
  a.rates = dbl(prhs[1]); a.ic = dbl(prhs[2]);
  const double *sat = dbl(prhs[5]);
  a.beta_0 = sat[0]; a.beta_s = sat[1]; a.mu_0 = sat[2]; a.mu_s = sat[3]; a.sigma = sat[4]; a.i_0 = sat[5];
  a.out_mode = EPI_SEIRP_OUT_FULL;
  std::vector<double> out((size_t)5 * a.K);
  a.out = out.data();
  check(epi_seirp_batch(ctx(), &a));
  for (int f = 0; f < 5 && f < (nlhs > 0 ? nlhs : 1); ++f) {
    plhs[f] = mxCreateDoubleMatrix(1, a.K, mxREAL);
    std::memcpy(mxGetPr(plhs[f]), out.data() + (size_t)f * a.K, sizeof(double) * a.K);
  }
}

// [s,i,alpha] = epi_mex('sialpha_controlled', u(LxK), x0(3), params, noise_std(3), K, noise(3xK | []))
static void cmd_rollout(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]) {
  if (nrhs < 6) mexErrMsgIdAndTxt("epi:arg", "sialpha_controlled: 6 arguments expected");
  const int L = (int)mxGetM(prhs[0]);
  epi_model_params p = to_params(prhs[2], L);
  epi_rollout_args a;
  std::memset(&a, 0, sizeof a);
  a.mem = EPI_MEM_HOST; a.B = 1; a.K = (int)scalar(prhs[4]); a.L = L; a.G = 1;
  a.prm = &p; a.x0 = dbl(prhs[1]); a.noise_std = dbl(prhs[3]);
  a.u_kind = EPI_U_F64; a.u = dbl(prhs[0]);  // L x K column-major == [K][L][1]
  a.noise = mxIsEmpty(prhs[5]) ? nullptr : dbl(prhs[5]);
  mxArray *o[3];
  for (int f = 0; f < 3; ++f) o[f] = mxCreateDoubleMatrix(1, a.K, mxREAL);
  a.s = mxGetPr(o[0]); a.i = mxGetPr(o[1]); a.alpha = mxGetPr(o[2]);
  check(epi_rollout_cost_batch(ctx(), &a));
  for (int f = 0; f < 3; ++f) {
    if (f < (nlhs > 0 ? nlhs : 1)) plhs[f] = o[f]; else mxDestroyArray(o[f]);
  }
}

// [s,i] = epi_mex('si_controlled', alpha(1xK), beta, s0, i0, K, dt)
static void cmd_si(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]) {
  if (nrhs < 6) mexErrMsgIdAndTxt("epi:arg", "si_controlled: 6 arguments expected");
  epi_si_args a;
  std::memset(&a, 0, sizeof a);
  const double beta = scalar(prhs[1]), s0 = scalar(prhs[2]), i0 = scalar(prhs[3]);
  a.mem = EPI_MEM_HOST; a.B = 1; a.K = (int)scalar(prhs[4]); a.dt = scalar(prhs[5]);
  a.alpha = dbl(prhs[0]); a.beta = &beta; a.s0 = &s0; a.i0 = &i0;
  plhs[0] = mxCreateDoubleMatrix(1, a.K, mxREAL);
  mxArray *oi = mxCreateDoubleMatrix(1, a.K, mxREAL);
  a.s = mxGetPr(plhs[0]); a.i = mxGetPr(oi);
  check(epi_si_controlled_batch(ctx(), &a));
  if (nlhs > 1) plhs[1] = oi; else mxDestroyArray(oi);
}

// [S_MINUS, S_PLUS, P_MINUS, P_PLUS, K_GAIN, S_SMOOTH, P_SMOOTH, innovations, rho] =
//   epi_mex('rt_expfit', x(1xT), s_init(2x1), params(3x1), w_bar(2x1), v_bar, Ps_init(2x2), Q_w(2x2), R_v, beta,
//           gamma, inv_monitor_len, order)                                   Tools/Rt_ExpFitEKF.m:1
static void cmd_rt_expfit(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]) {
  if (nrhs < 12) mexErrMsgIdAndTxt("epi:arg", "rt_expfit: 12 arguments expected");
  epi_rt_expfit_args a;
  std::memset(&a, 0, sizeof a);
  const double R = scalar(prhs[7]);
  a.mem = EPI_MEM_HOST; a.B = 1; a.G = 1; a.T = (int)mxGetNumberOfElements(prhs[0]);
  a.x = dbl(prhs[0]); a.s_init = dbl(prhs[1]); a.params = dbl(prhs[2]); a.w_bar = dbl(prhs[3]);
  a.v_bar = scalar(prhs[4]); a.Ps_init = dbl(prhs[5]); a.Q = dbl(prhs[6]); a.R = &R;   // 2x2 pages are column-major on both sides
  a.beta = scalar(prhs[8]); a.gamma = scalar(prhs[9]); a.W = (int)scalar(prhs[10]); a.order = (int)scalar(prhs[11]);
  const mwSize T = (mwSize)a.T, d3[3] = {2, 2, T}, k3[3] = {2, 1, T};
  mxArray *o[9];
  o[0] = mxCreateDoubleMatrix(2, T, mxREAL); o[1] = mxCreateDoubleMatrix(2, T, mxREAL);
  o[2] = mxCreateNumericArray(3, d3, mxDOUBLE_CLASS, mxREAL); o[3] = mxCreateNumericArray(3, d3, mxDOUBLE_CLASS, mxREAL);
  o[4] = mxCreateNumericArray(3, k3, mxDOUBLE_CLASS, mxREAL);
  o[5] = mxCreateDoubleMatrix(2, T, mxREAL); o[6] = mxCreateNumericArray(3, d3, mxDOUBLE_CLASS, mxREAL);
  o[7] = mxCreateDoubleMatrix(1, T, mxREAL); o[8] = mxCreateDoubleMatrix(T, 1, mxREAL);
  a.S_MINUS = mxGetPr(o[0]); a.S_PLUS = mxGetPr(o[1]); a.P_MINUS = mxGetPr(o[2]); a.P_PLUS = mxGetPr(o[3]);
  a.K_GAIN = mxGetPr(o[4]); a.S_SMOOTH = mxGetPr(o[5]); a.P_SMOOTH = mxGetPr(o[6]);
  a.innovations = mxGetPr(o[7]); a.rho = mxGetPr(o[8]);
  check(epi_rt_expfit_batch(ctx(), &a));   // B = 1: [T][F][1] is MATLAB's column-major F x T
  const int want = nlhs > 0 ? nlhs : 1;
  for (int f = 0; f < 9; ++f) {
    if (f < want) plhs[f] = o[f]; else mxDestroyArray(o[f]);
  }
}

// [J0,J1] = epi_mex('npicost', newcases(1xT), inputs(LxT), weights(LxT))
static void cmd_npicost(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]) {
  if (nrhs < 3) mexErrMsgIdAndTxt("epi:arg", "npicost: 3 arguments expected");
  epi_npicost_args a;
  std::memset(&a, 0, sizeof a);
  a.mem = EPI_MEM_HOST; a.B = 1; a.L = (int)mxGetM(prhs[1]); a.T = (int)mxGetN(prhs[1]); a.G = 1;
  a.newcases = dbl(prhs[0]); a.inputs = dbl(prhs[1]); a.weights = dbl(prhs[2]);
  double J0 = 0, J1 = 0;
  a.J0 = &J0; a.J1 = &J1;
  check(epi_npicost_batch(ctx(), &a));
  plhs[0] = mxCreateDoubleScalar(J0);
  if (nlhs > 1) plhs[1] = mxCreateDoubleScalar(J1);
}

// [on_front, I_opt] = epi_mex('pareto', J0(1xn), J1(1xn))      (I_opt 1-based)
static void cmd_pareto(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]) {
  if (nrhs < 2) mexErrMsgIdAndTxt("epi:arg", "pareto: 2 arguments expected");
  epi_pareto_args a;
  std::memset(&a, 0, sizeof a);
  a.mem = EPI_MEM_HOST; a.n_sets = 1; a.n = (int)mxGetNumberOfElements(prhs[0]);
  a.J0 = dbl(prhs[0]); a.J1 = dbl(prhs[1]);
  plhs[0] = mxCreateLogicalMatrix(1, a.n);
  int iopt = 0;
  a.on_front = (unsigned char *)mxGetLogicals(plhs[0]); a.I_opt = &iopt;
  check(epi_pareto_batch(ctx(), &a));
  if (nlhs > 1) plhs[1] = mxCreateDoubleScalar((double)(iopt + 1));
}

// [u_opt, u_opt_smooth, S_MINUS, S_PLUS, S_SMOOTH, P_MINUS, P_PLUS, P_SMOOTH, K_GAIN, innovations, rho] =
//   epi_mex('ekf_eks', model, u, x, params, s_init, Ps_init, s_final, Ps_final, w_bar, v_bar, Q_w, R_v,
//           beta, gamma, inv_monitor_len, order)
// Q/R shape dispatch follows GenericExtendedKalmanFilter.m:64-91.
// arguments 0..15 of 'ekf_eks' -> epi_ekf_args (one trajectory; Q/R shape dispatch of :64-91)
struct EkfCall {
  epi_ekf_args a;
  epi_model_params p;
  std::vector<double> qbuf;
  int m, L, T;
  bool legacy;
};
static void ekf_fill(EkfCall &c, int nrhs, const mxArray *prhs[]) {
  if (nrhs < 16) mexErrMsgIdAndTxt("epi:arg", "ekf_eks: 16 arguments expected");
  epi_ekf_args &a = c.a;
  std::vector<double> &qbuf = c.qbuf;
  const int model = (int)scalar(prhs[0]);
  const int m = model >= EPI_MODEL_OPTCTRL ? 6 : 3;
  const int L = (int)mxGetM(prhs[1]), T = (int)mxGetN(prhs[1]);
  if ((int)mxGetNumberOfElements(prhs[2]) != T) mexErrMsgIdAndTxt("epi:arg", "x must be 1 x T");
  c.p = to_params(prhs[3], L);
  std::memset(&a, 0, sizeof a);
  a.mem = EPI_MEM_HOST; a.model = model; a.B = 1; a.T = T; a.L = L; a.G = 1; a.prm = &c.p;
  a.u = dbl(prhs[1]); a.x = dbl(prhs[2]);
  a.s_init = dbl(prhs[4]); a.Ps_init = dbl(prhs[5]); a.s_final = dbl(prhs[6]); a.Ps_final = dbl(prhs[7]);
  a.v_bar = scalar(prhs[9]);
  // Q_w
  const mxArray *Q = prhs[10];
  const size_t qr = mxGetM(Q), qn = mxGetNumberOfElements(Q);
  const size_t qc = mxGetNumberOfDimensions(Q) > 2 ? mxGetDimensions(Q)[1] : mxGetN(Q);
  const bool legacy = model >= EPI_MODEL_LEGACY_TOOLS;
  if (qr == qc) {
    if (qn == 1) {  // scalar: B*q*B' with B = I
      qbuf.assign((size_t)m * m, 0.0);
      for (int i = 0; i < m; ++i) qbuf[(size_t)i * m + i] = scalar(Q);
      a.q_mode = EPI_Q_CONST; a.Q = qbuf.data();
    } else if (qn == (size_t)m * m) { a.q_mode = EPI_Q_CONST; a.Q = dbl(Q); }
    else if (!legacy && qn == (size_t)m * m * T) { a.q_mode = EPI_Q_PERDAY_FULL; a.Q = dbl(Q); }
    else mexErrMsgIdAndTxt("epi:covShape", "Process noise covariance noise mismatch");
  } else if (!legacy && (qr == 1 || qc == 1) && qn == (size_t)T) { a.q_mode = EPI_Q_PERDAY_SCALAR; a.Q = dbl(Q); }
  else mexErrMsgIdAndTxt("epi:covShape", "Process noise covariance noise mismatch");
  // R_v
  const mxArray *R = prhs[11];
  const size_t rr = mxGetM(R), rn = mxGetNumberOfElements(R);
  const size_t rc = mxGetNumberOfDimensions(R) > 2 ? mxGetDimensions(R)[1] : mxGetN(R);
  if (rr == rc && rn == 1) { a.r_mode = EPI_R_CONST; a.fixed_R = 1; }
  else if (!legacy && rr == rc && rn == (size_t)T) { a.r_mode = EPI_R_PERDAY; a.fixed_R = 1; }
  else if (!legacy && (rr == 1 || rc == 1) && rn == (size_t)T) { a.r_mode = EPI_R_PERDAY; a.fixed_R = 0; }
  else mexErrMsgIdAndTxt("epi:covShape", "Observation noise covariance noise mismatch");
  a.R = dbl(R);
  a.beta = scalar(prhs[12]); a.gamma = scalar(prhs[13]); a.W = (int)scalar(prhs[14]); a.order = (int)scalar(prhs[15]);
  c.m = m; c.L = L; c.T = T; c.legacy = legacy;
}

static void cmd_ekf(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]) {
  EkfCall c;
  ekf_fill(c, nrhs, prhs);
  epi_ekf_args &a = c.a;
  const int m = c.m, L = c.L, T = c.T;
  const bool legacy = c.legacy;
  const mwSize d3[3] = {(mwSize)m, (mwSize)m, (mwSize)T}, dk[3] = {(mwSize)m, 1, (mwSize)T};
  mxArray *o[11];
  o[0] = mxCreateDoubleMatrix(L, T, mxREAL); o[1] = mxCreateDoubleMatrix(L, T, mxREAL);
  o[2] = mxCreateDoubleMatrix(m, T, mxREAL); o[3] = mxCreateDoubleMatrix(m, T, mxREAL);
  o[4] = mxCreateDoubleMatrix(m, T, mxREAL);
  o[5] = mxCreateNumericArray(3, d3, mxDOUBLE_CLASS, mxREAL); o[6] = mxCreateNumericArray(3, d3, mxDOUBLE_CLASS, mxREAL);
  o[7] = mxCreateNumericArray(3, d3, mxDOUBLE_CLASS, mxREAL); o[8] = mxCreateNumericArray(3, dk, mxDOUBLE_CLASS, mxREAL);
  o[9] = mxCreateDoubleMatrix(1, T, mxREAL); o[10] = mxCreateDoubleMatrix(T, 1, mxREAL);
  a.u_opt = mxGetPr(o[0]); a.u_opt_smooth = legacy ? nullptr : mxGetPr(o[1]);
  a.S_MINUS = mxGetPr(o[2]); a.S_PLUS = mxGetPr(o[3]); a.S_SMOOTH = mxGetPr(o[4]);
  a.P_MINUS = mxGetPr(o[5]); a.P_PLUS = mxGetPr(o[6]); a.P_SMOOTH = mxGetPr(o[7]);
  a.K_GAIN = mxGetPr(o[8]); a.innovations = mxGetPr(o[9]); a.rho = mxGetPr(o[10]);
  check(epi_ekf_eks_batch(ctx(), &a));
  const int want = nlhs > 0 ? nlhs : 1;
  for (int f = 0; f < 11; ++f) {
    if (f < want) plhs[f] = o[f]; else mxDestroyArray(o[f]);
  }
}

// [S_PLUS, S_SMOOTH] = epi_mex('ekf_eks_masked', model, u, x, params, s_init, Ps_init, s_final, Ps_final, w_bar,
//                              v_bar, Q_w, R_v, beta, gamma, inv_monitor_len, order, num_forecast_days)
// The masked-horizon re-runs of Tools/ForecastQualityAssessment.m:383-386 as ONE batch: trajectory `start`
// (1..num_forecast_days) is the same call with observations(T-start+1 : T) = NaN.  S_PLUS, S_SMOOTH come back
// as m x T x num_forecast_days (page `start` = what the reference's loop iteration computes).
static void cmd_ekf_masked(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]) {
  if (nrhs < 17) mexErrMsgIdAndTxt("epi:arg", "ekf_eks_masked: 17 arguments expected");
  EkfCall c;
  ekf_fill(c, nrhs, prhs);
  epi_ekf_args &a = c.a;
  const int m = c.m, T = c.T, nf = (int)scalar(prhs[16]);
  if (nf < 1 || nf > T) mexErrMsgIdAndTxt("epi:arg", "ekf_eks_masked: num_forecast_days must be in 1..T");
  const double *x = dbl(prhs[2]);
  std::vector<double> xb((size_t)T * nf);            // [T][B], B = nf
  for (int t = 0; t < T; ++t)
    for (int b = 0; b < nf; ++b) xb[(size_t)t * nf + b] = (t >= T - (b + 1)) ? mxGetNaN() : x[t];
  a.B = nf; a.G = nf; a.x_per_traj = 1; a.x = xb.data();
  std::vector<double> sp((size_t)T * m * nf), ss((size_t)T * m * nf);   // ABI layout [T][m][B]
  a.S_PLUS = sp.data(); a.S_SMOOTH = ss.data();
  check(epi_ekf_eks_batch(ctx(), &a));
  const mwSize d3[3] = {(mwSize)m, (mwSize)T, (mwSize)nf};
  mxArray *o[2] = {mxCreateNumericArray(3, d3, mxDOUBLE_CLASS, mxREAL), mxCreateNumericArray(3, d3, mxDOUBLE_CLASS, mxREAL)};
  const std::vector<double> *src[2] = {&sp, &ss};
  for (int k = 0; k < 2; ++k) {
    double *dst = mxGetPr(o[k]);
    for (int b = 0; b < nf; ++b)
      for (int t = 0; t < T; ++t)
        for (int i = 0; i < m; ++i) dst[(size_t)i + (size_t)m * ((size_t)t + (size_t)T * b)] = (*src[k])[((size_t)t * m + i) * nf + b];
  }
  plhs[0] = o[0];
  if (nlhs > 1) plhs[1] = o[1]; else mxDestroyArray(o[1]);
}

// [J0, J1, on_front, I_opt, u_knee] = epi_mex('sweep', params(1xnR struct array), eps(1xnE), u(LxTxnR),
//     x(TxnR), R(TxnR), s_init(6xnR), Ps_init(6x6xnR), s_final(6xnR), Ps_final(6x6xnR), Q(6x6xnR),
//     beta_ekf, gamma_ekf, W, x0(3xnR), newcases_hist(T_histxnR), weights(LxTxnR), lean [, n_gpus])
// n_gpus (optional, default 1): shard the regions over that many GPUs of the box in this one blocking call
// (epi_sweep_multi; 0 = all GPUs); the outputs are the same arrays, bit for bit.
// The batched form of the loop body of Tools/TrainPredictPrescribeNPI.m:421-495 + :624-633 for all
// regions x all epsilon.  MATLAB's column-major (L x T x nR) is the ABI's per-region [T][L] layout.
// Outputs: J0, J1, on_front are nE x nR; I_opt 1 x nR (1-based); u_knee L x T_fore x nR.
static void cmd_sweep(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]) {
  if (nrhs < 17) mexErrMsgIdAndTxt("epi:arg", "sweep: 17 arguments expected");
  const mwSize *ud = mxGetDimensions(prhs[2]);
  const int L = (int)ud[0], T = (int)ud[1];
  const int nR = mxGetNumberOfDimensions(prhs[2]) > 2 ? (int)ud[2] : 1;
  const int nE = (int)mxGetNumberOfElements(prhs[1]);
  const int T_hist = (int)mxGetM(prhs[14]);
  if ((int)mxGetNumberOfElements(prhs[0]) != nR) mexErrMsgIdAndTxt("epi:arg", "sweep: one params struct per region");
  std::vector<epi_model_params> prm((size_t)nR);
  for (int r = 0; r < nR; ++r) {
    // to_params reads element 0 of a struct array: copy element r into a 1x1 view
    mxArray *one = mxCreateStructMatrix(1, 1, 0, nullptr);
    const int nf = mxGetNumberOfFields(prhs[0]);
    for (int f = 0; f < nf; ++f) {
      const char *name = mxGetFieldNameByNumber(prhs[0], f);
      mxAddField(one, name);
      const mxArray *v = mxGetFieldByNumber(prhs[0], r, f);
      if (v) mxSetField(one, 0, name, mxDuplicateArray(v));
    }
    prm[(size_t)r] = to_params(one, L);
    mxDestroyArray(one);
  }
  epi_sweep_args a;
  std::memset(&a, 0, sizeof a);
  a.mem = EPI_MEM_HOST; a.n_regions = nR; a.n_eps = nE; a.T = T; a.T_hist = T_hist; a.L = L;
  a.prm = prm.data(); a.eps = dbl(prhs[1]); a.u = dbl(prhs[2]); a.x = dbl(prhs[3]); a.R = dbl(prhs[4]);
  a.s_init = dbl(prhs[5]); a.Ps_init = dbl(prhs[6]); a.s_final = dbl(prhs[7]); a.Ps_final = dbl(prhs[8]);
  a.Q = dbl(prhs[9]); a.beta_ekf = scalar(prhs[10]); a.gamma_ekf = scalar(prhs[11]); a.W = (int)scalar(prhs[12]);
  a.x0 = dbl(prhs[13]); a.newcases_hist = T_hist > 0 ? dbl(prhs[14]) : nullptr; a.weights = dbl(prhs[15]);
  a.lean = (int)scalar(prhs[16]);
  const int Tf = T - T_hist;
  mxArray *oJ0 = mxCreateDoubleMatrix(nE, nR, mxREAL), *oJ1 = mxCreateDoubleMatrix(nE, nR, mxREAL);
  mxArray *oF = mxCreateLogicalMatrix(nE, nR);
  std::vector<int> iopt((size_t)nR);
  const mwSize kd[3] = {(mwSize)L, (mwSize)Tf, (mwSize)nR};
  mxArray *oK = mxCreateNumericArray(3, kd, mxDOUBLE_CLASS, mxREAL);
  a.J0 = mxGetPr(oJ0); a.J1 = mxGetPr(oJ1); a.on_front = (unsigned char *)mxGetLogicals(oF);
  a.I_opt = iopt.data(); a.u_knee = mxGetPr(oK);
  const int n_gpus = nrhs > 17 ? (int)scalar(prhs[17]) : 1;
  if (n_gpus == 1) {
    check(epi_sweep(ctx(), &a));
  } else {
    std::vector<epi_ctx *> cs = ctxs(n_gpus);
    check(epi_sweep_multi(cs.data(), (int)cs.size(), &a));
  }
  mxArray *oI = mxCreateDoubleMatrix(1, nR, mxREAL);
  for (int r = 0; r < nR; ++r) mxGetPr(oI)[r] = (double)(iopt[(size_t)r] + 1);
  mxArray *o[5] = {oJ0, oJ1, oF, oI, oK};
  const int want = nlhs > 0 ? nlhs : 1;
  for (int f = 0; f < 5; ++f) {
    if (f < want) plhs[f] = o[f]; else mxDestroyArray(o[f]);
  }
}

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]) {
  if (nrhs < 1 || !mxIsChar(prhs[0])) mexErrMsgIdAndTxt("epi:arg", "first argument must be a command string");
  char cmd[64];
  mxGetString(prhs[0], cmd, sizeof cmd);
  const std::string c(cmd);
  if (c == "seirp") cmd_seirp(nlhs, plhs, nrhs - 1, prhs + 1);
  else if (c == "sialpha_controlled") cmd_rollout(nlhs, plhs, nrhs - 1, prhs + 1);
  else if (c == "si_controlled") cmd_si(nlhs, plhs, nrhs - 1, prhs + 1);
  else if (c == "npicost") cmd_npicost(nlhs, plhs, nrhs - 1, prhs + 1);
  else if (c == "rt_expfit") cmd_rt_expfit(nlhs, plhs, nrhs - 1, prhs + 1);
  else if (c == "pareto") cmd_pareto(nlhs, plhs, nrhs - 1, prhs + 1);
  else if (c == "ekf_eks") cmd_ekf(nlhs, plhs, nrhs - 1, prhs + 1);
  else if (c == "ekf_eks_masked") cmd_ekf_masked(nlhs, plhs, nrhs - 1, prhs + 1);
  else if (c == "sweep") cmd_sweep(nlhs, plhs, nrhs - 1, prhs + 1);
  else mexErrMsgIdAndTxt("epi:arg", "unknown command '%s'", cmd);
}
