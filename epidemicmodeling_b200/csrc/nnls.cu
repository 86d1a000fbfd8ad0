// nnls.cu -- the non-negative regression between the EKF rounds (sm_100a, FP64, --fmad=false):
// Tools/TrainPredictPrescribeNPI.m:264-278 and :326-339 -- reg_coef_a = lsqnonneg(X, y) followed by
// the alternating intercept loop (up to 100 rounds, both the intercept and the error of a round use
// the PREVIOUS coefficients, as the .m is written).  One thread per region: 12 unknowns, a few hundred
// rows, so the Gram moments are formed once and the Lawson-Hanson active-set iteration runs on them
// (w = X'(d - Xa) = X'd - G a).  The passive-set least squares is DEFINED through the normal equations
// (Cholesky in index order, non-positive pivot => variable dropped); MATLAB's backslash uses QR.
// Arithmetic = oracle orc_nnls_affine operation for operation.
#include "epi_device.cuh"
#include "epi_internal.h"

namespace epi {

constexpr int kP = EPI_LMAX;

struct NnlsWork {
  double G[kP][kP];
};

__device__ static void nnls_solve_passive(const NnlsWork &W, const double *c, const int *P, int p, double *z) {
  int idx[kP], dead[kP], m = 0;
  double Lc[kP][kP], yv[kP];
  for (int j = 0; j < p; ++j) { z[j] = 0.0; if (P[j]) idx[m++] = j; }
  for (int i = 0; i < m; ++i) {
    for (int j = 0; j <= i; ++j) {
      double acc = W.G[idx[i]][idx[j]];
      for (int k = 0; k < j; ++k) acc = fma(-Lc[i][k], Lc[j][k], acc);
      if (j < i) {
        Lc[i][j] = dead[j] ? 0.0 : acc / Lc[j][j];
      } else {
        dead[i] = !(acc > 0.0);
        Lc[i][i] = dead[i] ? 1.0 : sqrt(acc);
      }
    }
    if (dead[i]) for (int j = 0; j < i; ++j) Lc[i][j] = 0.0;
  }
  for (int i = 0; i < m; ++i) {
    double acc = c[idx[i]];
    for (int k = 0; k < i; ++k) acc = fma(-Lc[i][k], yv[k], acc);
    yv[i] = dead[i] ? 0.0 : acc / Lc[i][i];
  }
  for (int i = m - 1; i >= 0; --i) {
    double acc = yv[i];
    for (int k = i + 1; k < m; ++k) acc = fma(-Lc[k][i], z[idx[k]], acc);
    z[idx[i]] = dead[i] ? 0.0 : acc / Lc[i][i];
  }
}

// lsqnonneg on the moments (Lawson & Hanson ch. 23 as in lsqnonneg.m: first maximum of w over Z enters,
// alpha = min x/(x - z) over the non-positive passive entries, |x| < tol leaves, itmax = 3p inner steps)
__device__ static void nnls_moments(const NnlsWork &W, const double *c, double tol, int p, double *x) {
  int P[kP], Z[kP];
  double w[kP], z[kP];
  for (int j = 0; j < p; ++j) { P[j] = 0; Z[j] = 1; x[j] = 0.0; w[j] = c[j]; }
  const int itmax = 3 * p;
  int iter = 0;
  for (;;) {
    bool anyZ = false, go = false;
    for (int j = 0; j < p; ++j)
      if (Z[j]) { anyZ = true; if (w[j] > tol) go = true; }
    if (!anyZ || !go) break;
    int t = -1;
    double best = 0.0;
    for (int j = 0; j < p; ++j)
      if (Z[j] && (t < 0 || w[j] > best)) { best = w[j]; t = j; }
    P[t] = 1; Z[t] = 0;
    nnls_solve_passive(W, c, P, p, z);
    bool stop = false;
    for (;;) {
      bool neg = false;
      for (int j = 0; j < p; ++j) if (P[j] && z[j] <= 0.0) neg = true;
      if (!neg) break;
      if (++iter > itmax) { stop = true; break; }
      double alpha = __longlong_as_double(0x7ff0000000000000ll);
      for (int j = 0; j < p; ++j)
        if (P[j] && z[j] <= 0.0) { const double a = x[j] / (x[j] - z[j]); if (a < alpha) alpha = a; }
      for (int j = 0; j < p; ++j) x[j] = x[j] + alpha * (z[j] - x[j]);
      for (int j = 0; j < p; ++j) { Z[j] = ((fabs(x[j]) < tol && P[j]) || Z[j]) ? 1 : 0; P[j] = !Z[j]; }
      nnls_solve_passive(W, c, P, p, z);
    }
    for (int j = 0; j < p; ++j) x[j] = z[j];
    if (stop) break;
    for (int j = 0; j < p; ++j) {
      double acc = c[j];
      for (int l = 0; l < p; ++l) acc = fma(-W.G[j][l], x[l], acc);
      w[j] = acc;
    }
  }
}

__global__ void __launch_bounds__(32) nnls_affine_kernel(const __grid_constant__ NnlsParams Q) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= Q.B) return;
  const int n = Q.n, p = Q.p;
  const size_t B = (size_t)Q.B;
  const double *__restrict__ X = Q.X + b;  // X(i, j) = X[(i*p + j) * B]
  const double *__restrict__ y = Q.y + b;  // y(i)    = y[i * B]
  NnlsWork W;
  double Xty[kP], Xt1[kP], cs[kP], c[kP], a[kP], at[kP];
  for (int j = 0; j < p; ++j) {
    double sy = 0.0, s1 = 0.0, sa = 0.0;
    for (int i = 0; i < n; ++i) {
      const double x = X[((size_t)i * p + j) * B];
      sy = fma(x, y[(size_t)i * B], sy); s1 += x; sa += fabs(x);
    }
    Xty[j] = sy; Xt1[j] = s1; cs[j] = sa;
    for (int l = 0; l <= j; ++l) {
      double g = 0.0;
      for (int i = 0; i < n; ++i) g = fma(X[((size_t)i * p + j) * B], X[((size_t)i * p + l) * B], g);
      W.G[j][l] = g; W.G[l][j] = g;
    }
  }
  double norm1 = 0.0;
  for (int j = 0; j < p; ++j) norm1 = mmax(norm1, cs[j]);
  const double tol = ((10.0 * 2.220446049250313e-16) * norm1) * (double)(n > p ? n : p);
  auto row_xa = [&](int i) {
    double xa = X[((size_t)i * p) * B] * a[0];
    for (int j = 1; j < p; ++j) xa = fma(X[((size_t)i * p + j) * B], a[j], xa);
    return xa;
  };
  double bb = 0.0;
  for (int j = 0; j < p; ++j) c[j] = Xty[j];
  nnls_moments(W, c, tol, p, a);  // :264
  double min_err = 0.0;
  for (int i = 0; i < n; ++i) { const double r = y[(size_t)i * B] - row_xa(i); min_err += r * r; }  // :266
  int k = 0;
  for (; k < Q.max_alt; ++k) {  // :267-277
    for (int j = 0; j < p; ++j) c[j] = Xty[j] - bb * Xt1[j];
    nnls_moments(W, c, tol, p, at);
    double sm = 0.0;
    for (int i = 0; i < n; ++i) sm += y[(size_t)i * B] - row_xa(i);
    const double b_t = sm / (double)n;
    double err = 0.0;
    for (int i = 0; i < n; ++i) { const double r = (y[(size_t)i * B] - row_xa(i)) - b_t; err += r * r; }
    if (err < min_err) {
      for (int j = 0; j < p; ++j) a[j] = at[j];
      bb = b_t; min_err = err;
    } else {
      break;
    }
  }
  for (int j = 0; j < p; ++j) Q.a[(size_t)j * B + b] = a[j];
  Q.b[b] = bb;
  if (Q.n_alt) Q.n_alt[b] = k;
}

void launch_nnls_affine(const NnlsParams &q, cudaStream_t st) {
  if (q.B <= 0) return;
  nnls_affine_kernel<<<(q.B + 31) / 32, 32, 0, st>>>(q);
}

}  // namespace epi
