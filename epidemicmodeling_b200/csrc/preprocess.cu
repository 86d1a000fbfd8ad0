// preprocess.cu -- per-region preprocessing that feeds the filters (sm_100a, FP64, --fmad=false):
// Tools/TrainPredictPrescribeNPI.m:121-128 (NPI forward fill), :162-187 (new-case series: diff, clip,
// NaN fill, causal 7-day moving average, zero-phase 4-tap moving average, normalisation, cumulative
// series), :200-201 (I0), :240 (observation-noise estimate R_v).  One thread per region, sequential
// in time; arithmetic = oracle orc_preprocess_region operation for operation (MATLAB's filter is
// direct form II transposed, filtfilt = odd reflection of 3(nb-1) samples + steady-state initial
// conditions, forward and reversed pass).
#include "epi_device.cuh"
#include "epi_internal.h"

namespace epi {

constexpr int kMaxTaps = 32;

__global__ void __launch_bounds__(64) preprocess_kernel(const __grid_constant__ PreprocParams P) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= P.B) return;
  const int T = P.T, L = P.L, W = P.W;
  const size_t B = (size_t)P.B;
  const double N = P.population[b];
  // :121-128 forward fill, then "no NPI" for what is still missing
  for (int j = 0; j < L; ++j) {
    double prev = __longlong_as_double(0x7ff8000000000000ll);
    for (int i = 0; i < T; ++i) {
      double v = P.ip_in[((size_t)i * L + j) * B + b];
      if (i > 0 && v != v && !(prev != prev)) v = prev;
      prev = v;  // the filled value propagates through a run of NaNs, a leading NaN stays NaN ...
      P.ip_out[((size_t)i * L + j) * B + b] = (v != v) ? 0.0 : v;  // ... and becomes 0 (:128)
    }
  }
  // :162-172
  int last = -1;
  double prev_cc = P.cc[b];
  for (int t = 0; t < T; ++t) {
    const double c = P.cc[(size_t)t * B + b];
    double d = c - prev_cc;
    prev_cc = c;
    if (d < 0.0) d = 0.0;
    if (!(d != d)) last = t;
    P.refined[(size_t)t * B + b] = d;
  }
  if (last >= 0 && last != T - 1) P.refined[(size_t)(T - 1) * B + b] = P.refined[(size_t)last * B + b];
  // NaN -> 0, causal moving average (:173, DF2T), cumulative and normalised series
  double bt[kMaxTaps], z[kMaxTaps];
  for (int i = 0; i < W; ++i) { bt[i] = 1.0 / (double)W; z[i] = 0.0; }
  double cs = 0.0, acc0 = 0.0;
  int cnt0 = 0;
  for (int t = 0; t < T; ++t) {
    double x = P.refined[(size_t)t * B + b];
    if (x != x) x = 0.0;
    P.refined[(size_t)t * B + b] = x;
    const double y = (W > 1) ? (bt[0] * x + z[0]) : (bt[0] * x);
    for (int i = 0; i + 2 < W; ++i) z[i] = bt[i + 1] * x + z[i + 1];
    if (W > 1) z[W - 2] = bt[W - 1] * x;
    P.smoothed[(size_t)t * B + b] = y;
    P.normalized[(size_t)t * B + b] = y / N;
    cs += y;
    P.confirmed_norm[(size_t)t * B + b] = cs / N;
    if (y > 0.0 && cnt0 < P.n_first) { acc0 += y; ++cnt0; }  // :200
  }
  const double mean0 = cnt0 ? acc0 / (double)cnt0 : __longlong_as_double(0x7ff8000000000000ll);
  P.I0[b] = mmax(P.min_cases, mean0);  // :201
  // :174 zero-phase moving average of Wh = round(W/2) taps
  const int Wh = (W + 1) / 2;  // round(W/2), half away from zero
  const int nfact = (3 * (Wh - 1) > 1) ? 3 * (Wh - 1) : 1;
  const int ne = T + 2 * nfact;
  double *xe = P.scratch + b, *ye = P.scratch + (size_t)ne * B + b;  // [ne][B] each
  double zi[kMaxTaps];
  for (int i = 0; i < Wh; ++i) bt[i] = 1.0 / (double)Wh;
  for (int i = Wh - 2; i >= 0; --i) zi[i] = bt[i + 1] + ((i + 1 < Wh - 1) ? zi[i + 1] : 0.0);
  const double x0 = P.refined[b], xl = P.refined[(size_t)(T - 1) * B + b];
  for (int i = 0; i < nfact; ++i) xe[(size_t)i * B] = 2.0 * x0 - P.refined[(size_t)(nfact - i) * B + b];
  for (int i = 0; i < T; ++i) xe[(size_t)(nfact + i) * B] = P.refined[(size_t)i * B + b];
  for (int i = 0; i < nfact; ++i) xe[(size_t)(nfact + T + i) * B] = 2.0 * xl - P.refined[(size_t)(T - 2 - i) * B + b];
  auto fir = [&](const double *src, bool reversed, double *dst) {
    const double first = src[(size_t)(reversed ? ne - 1 : 0) * B];
    for (int i = 0; i + 1 < Wh; ++i) z[i] = zi[i] * first;
    for (int t = 0; t < ne; ++t) {
      const double x = src[(size_t)(reversed ? ne - 1 - t : t) * B];
      const double y = (Wh > 1) ? (bt[0] * x + z[0]) : (bt[0] * x);
      for (int i = 0; i + 2 < Wh; ++i) z[i] = bt[i + 1] * x + z[i + 1];
      if (Wh > 1) z[Wh - 2] = bt[Wh - 1] * x;
      dst[(size_t)t * B] = y;
    }
  };
  fir(xe, false, ye);   // forward pass
  fir(ye, true, xe);    // pass over the reversed sequence: xe[t] = output for reversed index t
  for (int i = 0; i < T; ++i) {
    const double zl = xe[(size_t)(ne - 1 - nfact - i) * B];
    P.zerolag[(size_t)i * B + b] = zl;
    const double e = (zl - P.refined[(size_t)i * B + b]) / N;
    P.R_v[(size_t)i * B + b] = 0.1 * (e * e);  // :240
  }
}

void launch_preprocess(const PreprocParams &p, cudaStream_t st) {
  if (p.B <= 0) return;
  preprocess_kernel<<<(p.B + 63) / 64, 64, 0, st>>>(p);
}

}  // namespace epi
