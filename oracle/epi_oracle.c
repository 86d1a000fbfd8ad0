/*
 * epi_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 * See epi_oracle.h for the arithmetic contract and the "parity unpinned" note.
 * Build: make -C oracle   (gcc -O2 -ffp-contract=off -mfma -fopenmp)
 *
 * Every block cites the reference file:line (relative to the reference root)
 * whose behaviour it restates.  Nothing here is copied from the reference: the
 * reference is MATLAB, this is a from-scratch C restatement of its formulas.
 */
#include "epi_oracle.h"

#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MM 6 /* max state dimension */

/* MATLAB min/max: the non-NaN operand wins.  The exact select form below is
 * mirrored by the CUDA kernels so that ties / signed zeros agree bitwise. */
static inline double mmax(double a, double b) { return (b > a || a != a) ? b : a; }
static inline double mmin(double a, double b) { return (b < a || a != a) ? b : a; }

static const double ORC_EPS = 2.220446049250313e-16; /* MATLAB eps */

/* ------------------------------------------------------------------------- */
/* SEIRP  (Tools/SEIRP.m:13-32)                                              */
/* ------------------------------------------------------------------------- */
void orc_seirp(const double *alpha_e, const double *alpha_i, const double *kappa,
               const double *rho, const double *beta, const double *mu,
               const double *gamma, int rs, double s0, double e0, double i0, double r0,
               double p0, int K, double dt, double *s, double *e, double *i, double *r,
               double *p) {
  if (K <= 0) return;
  s[0] = s0; e[0] = e0; i[0] = i0; r[0] = r0; p[0] = p0; /* SEIRP.m:20-24 */
  for (int t = 0; t + 1 < K; ++t) {                       /* SEIRP.m:26 */
    double ae = alpha_e[t * rs], ai = alpha_i[t * rs], ka = kappa[t * rs], ro = rho[t * rs];
    double be = beta[t * rs], m_ = mu[t * rs], ga = gamma[t * rs];
    double S = s[t], E = e[t], I = i[t], R = r[t], P = p[t];
    /* SEIRP.m:27-31, parsed as MATLAB does: unary minus first, then left-to-right */
    s[t + 1] = ((((-ae) * S) * E - (ai * S) * I) + ga * R) * dt + S;
    e[t + 1] = (((((ae * S) * E) + ((ai * S) * I)) - ka * E) - ro * E) * dt + E;
    i[t + 1] = ((ka * E - be * I) - m_ * I) * dt + I;
    r[t + 1] = ((be * I + ro * E) - ga * R) * dt + R;
    p[t + 1] = (m_ * I) * dt + P;
  }
}

/* SEIRPSaturatedResource  (Tools/SEIRPSaturatedResource.m:13-36) */
void orc_seirp_saturated(const double *alpha_e, const double *alpha_i, const double *kappa,
                         const double *rho, const double *gamma, int rs, double s0, double e0,
                         double i0, double r0, double p0, int K, double dt, double beta_0,
                         double beta_s, double mu_0, double mu_s, double sigma, double i_0,
                         double *s, double *e, double *i, double *r, double *p) {
  if (K <= 0) return;
  s[0] = s0; e[0] = e0; i[0] = i0; r[0] = r0; p[0] = p0;
  for (int t = 0; t + 1 < K; ++t) {
    double ae = alpha_e[t * rs], ai = alpha_i[t * rs], ka = kappa[t * rs], ro = rho[t * rs];
    double ga = gamma[t * rs];
    double S = s[t], E = e[t], I = i[t], R = r[t], P = p[t];
    double h = (tanh((I - i_0) / sigma) + 1.0) / 2.0;      /* :27 */
    double be = (beta_s - beta_0) * h + beta_0;             /* :28 */
    double m_ = (mu_s - mu_0) * h + mu_0;                   /* :29 */
    s[t + 1] = ((((-ae) * S) * E - (ai * S) * I) + ga * R) * dt + S;
    e[t + 1] = (((((ae * S) * E) + ((ai * S) * I)) - ka * E) - ro * E) * dt + E;
    i[t + 1] = ((ka * E - be * I) - m_ * I) * dt + I;
    r[t + 1] = ((be * I + ro * E) - ga * R) * dt + R;
    p[t + 1] = (m_ * I) * dt + P;
  }
}

/* ------------------------------------------------------------------------- */
/* SIalpha_Controlled (Tools/SIalpha_Controlled.m:15-32), SI_Controlled,     */
/* NPICost, Pareto                                                           */
/* ------------------------------------------------------------------------- */

/* gamma * a' * (u_max - u): (gamma*a') element-wise, then the 12-term dot
 * product, DEFINED as first term a*b then fma, index ascending. */
static double input_dot(double gamma, const double *a, const double *u_max, const double *u,
                        int L) {
  double acc = 0.0;
  for (int j = 0; j < L; ++j) {
    double g = gamma * a[j];
    double d = u_max[j] - u[j];
    acc = (j == 0) ? g * d : fma(g, d, acc);
  }
  return acc;
}

void orc_sialpha_controlled(const double *u, int L, double s0, double i0, double alpha0,
                            const double *u_max, double alpha_min, double alpha_max,
                            double gamma, const double *a, double b, double beta,
                            double s_std, double i_std, double a_std, int K, double dt,
                            const double *noise, double *s, double *i, double *alpha) {
  double S = s0, I = i0, A = alpha0;
  for (int t = 0; t < K; ++t) { /* :24-28 ; randn order s, i, alpha */
    double ns = noise ? noise[3 * t + 0] : 0.0;
    double ni = noise ? noise[3 * t + 1] : 0.0;
    double na = noise ? noise[3 * t + 2] : 0.0;
    double asi = (A * S) * I;
    double Sn = mmax(0.0, mmin(1.0, S - dt * (asi + ns * s_std)));
    double In = mmax(0.0, mmin(1.0, I + dt * ((asi - beta * I) + ni * i_std)));
    double dot = input_dot(gamma, a, u_max, u + (size_t)L * t, L);
    double An = mmax(alpha_min,
                     mmin(alpha_max, A + dt * (((((-gamma) * A) + gamma * b) + dot) + na * a_std)));
    S = Sn; I = In; A = An;
    s[t] = S; i[t] = I; alpha[t] = A; /* :30-32 initial condition dropped */
  }
}

void orc_si_controlled(const double *alpha, double beta, double s0, double i0, int K, double dt,
                       double *s, double *i) {
  if (K <= 0) return;
  s[0] = s0; i[0] = i0; /* SI_Controlled.m:15-16 */
  for (int t = 0; t + 1 < K; ++t) { /* :19-22 */
    double S = s[t], I = i[t], A = alpha[t];
    s[t + 1] = mmax(0.0, mmin(1.0, S - ((dt * A) * S) * I));
    i[t + 1] = mmax(0.0, mmin(1.0, I + dt * (((A * S) * I) - beta * I)));
  }
}

void orc_npicost(const double *newcases, int T, const double *inputs, const double *weights,
                 int L, double *J0, double *J1) {
  double a0 = 0.0;
  for (int t = 0; t < T; ++t) a0 += newcases[t];
  *J0 = a0 / (double)T; /* NPICost.m:6 */
  /* NPICost.m:9-10: mean(weighted_inputs(:)).  MATLAB's sum order is
   * unspecified; DEFINED here as column (= day) sums in row order, then the
   * day sums in day order. */
  double a1 = 0.0;
  for (int t = 0; t < T; ++t) {
    double c = weights[(size_t)t * L] * inputs[(size_t)t * L];
    for (int j = 1; j < L; ++j) c = c + weights[(size_t)t * L + j] * inputs[(size_t)t * L + j];
    a1 += c;
  }
  *J1 = a1 / (double)((size_t)L * T);
}

void orc_pareto(const double *J0, const double *J1, int n, unsigned char *on_front,
                int *I_opt) {
  /* TrainPredictPrescribeNPI.m:624-627 : strict dominance in both coordinates */
  for (int i = 0; i < n; ++i) {
    int cnt = 0;
    for (int j = 0; j < n; ++j) cnt += (J0[j] < J0[i]) && (J1[j] < J1[i]);
    on_front[i] = (cnt == 0);
  }
  /* :633 knee point; MATLAB max/min skip NaN, min returns the first minimum */
  double m0 = NAN, m1 = NAN;
  for (int i = 0; i < n; ++i) { m0 = mmax(m0, J0[i]); m1 = mmax(m1, J1[i]); }
  int best = 0; double bv = NAN;
  for (int i = 0; i < n; ++i) {
    double q0 = J0[i] / m0, q1 = J1[i] / m1;
    double v = q0 * q0 + q1 * q1;
    if (v != v) continue;
    if (bv != bv || v < bv) { bv = v; best = i; }
  }
  if (I_opt) *I_opt = best;
}

/* ------------------------------------------------------------------------- */
/* random NPI schedules: TrainPredictPrescribeNPI.m:499-510 on a counter-based stream */
/* ------------------------------------------------------------------------- */
/* Philox4x32-10 (Salmon et al., SC'11).  Third-party algorithm, restated from the paper:
 * round: (hi0,lo0) = M0*c0, (hi1,lo1) = M1*c2; c = {hi1^c1^k0, lo1, hi0^c3^k1, lo0};
 * key += (W0, W1) after every round; 10 rounds. */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    c0 = n0; c1 = (uint32_t)p1; c2 = n2; c3 = (uint32_t)p0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
/* Schedule of Monte-Carlo scenario `scenario` (0-based) of `region`, n_scenarios per region:
 * scenarios with (scenario+1) < n_scenarios/2 hold one randi([u_min(j), u_max(j)]) per NPI over
 * all K days (:502-503), the others draw per NPI per day (:505-507).  Draw (j, day) =
 * word (j mod 4) of Philox(counter {day, j/4, scenario, region}, key seed), day = 0 when held;
 * level = lo + floor(word*(hi-lo+1)/2^32).  u [K][L] uint8. */
void orc_random_schedule(uint64_t seed, uint32_t region, uint32_t scenario, int n_scenarios, int L, int K,
                         const double *u_min, const double *u_max, unsigned char *u) {
  const int held = 2 * ((long long)scenario + 1) < (long long)n_scenarios;
  const uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  for (int t = 0; t < K; ++t)
    for (int j = 0; j < L; ++j) {
      const uint32_t ctr[4] = {held ? 0u : (uint32_t)t, (uint32_t)(j / 4), scenario, region};
      uint32_t w[4];
      orc_philox4x32_10(ctr, key, w);
      const int lo = (int)u_min[j], hi = (int)u_max[j];
      u[t * L + j] = (unsigned char)(lo + (int)(((uint64_t)w[j & 3] * (uint64_t)(hi - lo + 1)) >> 32));
    }
}

/* ------------------------------------------------------------------------- */
/* pinv / mrdivide as DEFINED by the oracle                                  */
/* ------------------------------------------------------------------------- */

/* spacing of doubles at |x| (MATLAB eps(x)) for normal x */
static double eps_of(double x) {
  union { double d; uint64_t u; } v;
  v.d = fabs(x);
  uint64_t e = (v.u >> 52) & 0x7ffu;
  if (e == 0x7ffu) return NAN;
  if (e <= 52) return 4.9406564584124654e-324; /* subnormal spacing */
  v.u = (e - 52) << 52;
  return v.d;
}

#define ORC_JACOBI_MAXSWEEP 16
static const double ORC_JACOBI_REL = 2.168404344971009e-19; /* 2^-62 */

/* Pair order of one Jacobi sweep: "sets" of index-disjoint pairs (round-robin
 * tournament).  Rotations of one set commute exactly (their angles read
 * a_pp, a_qq, a_pq of disjoint index pairs), which is what lets the CUDA kernel
 * compute the three angles of a 6x6 set side by side; the APPLICATION order is
 * still the listed order, pair by pair, and a pair (p,q) is rotated as listed
 * (p > q occurs: the table is the circle method -- pairs at positions (0,1),
 * (2,3), (4,5), then positions 1..5 rotate -- which the kernel executes as one
 * rolled loop).  m = 2 and m = 3 have one pair per set (m = 3: the classical
 * row-cyclic order). */
static const int ORC_JSETS6[5][3][2] = {{{0, 1}, {2, 3}, {4, 5}}, {{0, 3}, {1, 5}, {2, 4}},
                                        {{0, 5}, {3, 4}, {1, 2}}, {{0, 4}, {5, 2}, {3, 1}},
                                        {{0, 2}, {4, 1}, {5, 3}}};
static const int ORC_JSETS3[3][1][2] = {{{0, 1}}, {{0, 2}}, {{1, 2}}};
static const int ORC_JSETS2[1][1][2] = {{{0, 1}}};

/* Rotation annihilating a_pq (apq != 0), tan(2 phi) = apq / d with d = (aqq - app)/2:
 *   r = sqrt(d^2 + apq^2),  t = sgn(d) apq / (|d| + r)  (sgn(0) = +1),
 *   c = sqrt((|d| + r) / (2 r)),  s = t c
 * -- the same angle as the textbook theta = d/apq, t = sgn(theta)/(|theta| + sqrt(theta^2+1)),
 * c = 1/sqrt(t^2+1), with 2 divisions + 2 square roots on a dependency chain of 3 instead of
 * 3 + 2 on a chain of 5.  When d^2 + apq^2 leaves [2^-900, 2^900] (squares about to
 * under/overflow) the textbook form, which cannot overflow, is used instead. */
static void jacobi_angle(double app, double aqq, double apq, double *t_, double *c_, double *s_) {
  double d = 0.5 * (aqq - app);
  double r2 = fma(d, d, apq * apq);
  double t, c;
  if (r2 > 1.1830521861667747e-271 && r2 < 8.452712498170644e+270) { /* 2^-900 .. 2^900 */
    double ad = fabs(d);
    double r = sqrt(r2);
    double den = ad + r;
    t = apq / den;
    if (d < 0.0) t = -t;
    c = sqrt(den / (r + r));
  } else {
    double theta = d / apq;
    double at = fabs(theta);
    t = 1.0 / (at + sqrt(at * at + 1.0));
    if (theta < 0.0) t = -t;
    c = 1.0 / sqrt(t * t + 1.0);
  }
  *t_ = t; *c_ = c; *s_ = t * c;
}

/* Threshold Jacobi on the symmetric matrix whose upper triangle is A
 * (column-major mxm), pairs in the set order above.  A pair (p,q) is ACTIVE iff
 * |a_pq| > 2^-62 * max_p|a_pp| (threshold recomputed at each sweep start); the
 * iteration stops when no pair is active at a sweep start (or after 16 sweeps).
 * A set with no active pair is skipped; in an executed set every pair is
 * rotated, the idle ones by the identity (t = 0, c = 1, s = 0) -- that changes
 * no value (at most the sign of a zero) and makes the set branch-free in the
 * kernel.  lambda_i = diagonal after convergence; the eigenvector matrix
 * V = R_1 R_2 ... R_n (R_k: plane rotation with R_pp = R_qq = c, R_pq = s,
 * R_qp = -s) is never formed:
 *   pinv = V W V' = R_1 ( ... (R_n W R_n') ... ) R_1',   W = diag([|lambda_i| > tol] / lambda_i),
 *   tol = m * eps(max|lambda|)   (MATLAB pinv's default tolerance)
 * is evaluated by REPLAYING the recorded rotations, last first, as two-sided
 * updates of the symmetric X (30 flops per rotation instead of 24 for V plus
 * the m^3 products; and the kernel needs no m x m eigenvector registers). */
int orc_pinv_sym(const double *Ain, int m, double *X, int *rank) {
  double a[MM][MM];
  struct { int p, q; double c, s; } rec[ORC_JACOBI_MAXSWEEP * MM * (MM - 1) / 2];
  int n_rec = 0;
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < m; ++j) a[i][j] = (i <= j) ? Ain[j * m + i] : Ain[i * m + j];
  const int (*sets)[2] = m == 6 ? &ORC_JSETS6[0][0] : m == 3 ? &ORC_JSETS3[0][0] : m == 2 ? &ORC_JSETS2[0][0] : NULL;
  const int n_set = m == 6 ? 3 : 1, n_sets = m == 6 ? 5 : m == 3 ? 3 : 1;
  int sweep = 0;
  for (; sets && sweep < ORC_JACOBI_MAXSWEEP; ++sweep) {
    double dmax = 0.0, offmax = 0.0;
    for (int p = 0; p < m; ++p) dmax = mmax(dmax, fabs(a[p][p]));
    for (int p = 0; p < m; ++p)
      for (int q = p + 1; q < m; ++q) offmax = mmax(offmax, fabs(a[p][q]));
    double thr = dmax * ORC_JACOBI_REL;
    if (!(offmax > thr)) break;
    for (int st = 0; st < n_sets; ++st) {
      int act[3], any_act = 0;
      double tt[3], cc[3], ss[3];
      for (int i = 0; i < n_set; ++i) { /* the set's angles read disjoint entries */
        const int p = sets[st * n_set + i][0], q = sets[st * n_set + i][1];
        act[i] = fabs(a[p][q]) > thr;
        any_act |= act[i];
        tt[i] = 0.0; cc[i] = 1.0; ss[i] = 0.0;
        if (act[i]) jacobi_angle(a[p][p], a[q][q], a[p][q], &tt[i], &cc[i], &ss[i]);
      }
      if (!any_act) continue;
      for (int i = 0; i < n_set; ++i) {
        const int p = sets[st * n_set + i][0], q = sets[st * n_set + i][1];
        const double t = tt[i], c = cc[i], s = ss[i];
        double app = a[p][p], aqq = a[q][q], apq = a[p][q];
        a[p][p] = app - t * apq;
        a[q][q] = aqq + t * apq;
        if (act[i]) { a[p][q] = 0.0; a[q][p] = 0.0; }
        for (int r = 0; r < m; ++r) {
          if (r != p && r != q) {
            double g = a[r][p], h = a[r][q];
            double gp = fma(c, g, -(s * h));
            double hp = fma(s, g, c * h);
            a[r][p] = gp; a[p][r] = gp;
            a[r][q] = hp; a[q][r] = hp;
          }
        }
        rec[n_rec].p = p; rec[n_rec].q = q; rec[n_rec].c = c; rec[n_rec].s = s;
        ++n_rec;
      }
    }
  }
  double lmax = 0.0;
  for (int i = 0; i < m; ++i) lmax = mmax(lmax, fabs(a[i][i]));
  double tol = (double)m * eps_of(lmax);
  double x[MM][MM];
  int rk = 0;
  for (int i = 0; i < m; ++i) {
    int keep = fabs(a[i][i]) > tol;
    for (int j = 0; j < m; ++j) x[i][j] = 0.0;
    x[i][i] = keep ? 1.0 / a[i][i] : 0.0;
    rk += keep;
  }
  for (int k = n_rec - 1; k >= 0; --k) { /* X <- R_k X R_k' */
    const int p = rec[k].p, q = rec[k].q;
    const double c = rec[k].c, s = rec[k].s;
    for (int r = 0; r < m; ++r) {
      if (r != p && r != q) {
        double g = x[r][p], h = x[r][q];
        double gp = fma(c, g, s * h);
        double hp = fma(c, h, -(s * g));
        x[r][p] = gp; x[p][r] = gp;
        x[r][q] = hp; x[q][r] = hp;
      }
    }
    double xpp = x[p][p], xpq = x[p][q], xqq = x[q][q];
    double u1 = fma(c, xpp, s * xpq), u2 = fma(c, xpq, s * xqq);       /* (R X)(p,p), (R X)(p,q) */
    double w1 = fma(c, xpq, -(s * xpp)), w2 = fma(c, xqq, -(s * xpq)); /* (R X)(q,p), (R X)(q,q) */
    x[p][p] = fma(c, u1, s * u2);
    x[p][q] = fma(c, u2, -(s * u1)); x[q][p] = x[p][q];
    x[q][q] = fma(c, w2, -(s * w1));
  }
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < m; ++j) X[j * m + i] = x[i][j];
  if (rank) *rank = rk;
  return sweep;
}

/* X = B / A  :=  (A' \ B')'  by Gaussian elimination with partial pivoting
 * (first maximal |pivot| wins), multipliers l = a_ik / a_kk, updates
 * fma(-l, a_kj, a_ij), unit-lower forward substitution, back substitution
 * with one division per unknown.  No singularity test (the reference only
 * warns, NewCaseEKFEstimatorWithOptimalNPI.m:132). */
void orc_mrdivide(const double *B, const double *A, int m, double *X) {
  double lu[MM][MM], rhs[MM][MM];
  /* lu = A' ; rhs = B' (row r of rhs = column r of B' = row r ... ) */
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < m; ++j) {
      lu[i][j] = A[i * m + j];  /* (A')[i][j] = A[j][i] = A_colmajor[i*m + j] */
      rhs[i][j] = B[i * m + j]; /* (B')[i][j] = B[j][i] */
    }
  for (int k = 0; k < m; ++k) {
    int piv = k; double best = fabs(lu[k][k]);
    for (int r = k + 1; r < m; ++r) {
      double c = fabs(lu[r][k]);
      if (c > best) { best = c; piv = r; }
    }
    if (piv != k)
      for (int j = 0; j < m; ++j) {
        double t = lu[k][j]; lu[k][j] = lu[piv][j]; lu[piv][j] = t;
        t = rhs[k][j]; rhs[k][j] = rhs[piv][j]; rhs[piv][j] = t;
      }
    for (int r = k + 1; r < m; ++r) {
      double l = lu[r][k] / lu[k][k];
      for (int j = k + 1; j < m; ++j) lu[r][j] = fma(-l, lu[k][j], lu[r][j]);
      for (int j = 0; j < m; ++j) rhs[r][j] = fma(-l, rhs[k][j], rhs[r][j]);
    }
  }
  /* back substitution: Y (m x m) with lu_upper * Y = rhs */
  for (int j = 0; j < m; ++j)
    for (int r = m - 1; r >= 0; --r) {
      double acc = rhs[r][j];
      for (int c = r + 1; c < m; ++c) acc = fma(-lu[r][c], rhs[c][j], acc);
      rhs[r][j] = acc / lu[r][r];
    }
  /* X = Y' : X[i][j] = Y[j][i]; column-major X[j*m+i] */
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < m; ++j) X[j * m + i] = rhs[j][i];
}

/* ------------------------------------------------------------------------- */
/* model callbacks                                                           */
/* ------------------------------------------------------------------------- */
static inline int model_dim(int model) { return model >= ORC_OPTCTRL ? 6 : 3; }
static inline int model_flipped(int model) {
  return model == ORC_SIALPHA_FLIPPED || model == ORC_OPTCTRL_FLIPPED;
}
static inline int model_legacy(int model) { return model >= ORC_LEGACY_TOOLS; }

/* structural non-zero pattern of the state Jacobian A (row -> columns),
 * SIAlphaModelEKF.m:64-73 and SIAlphaModelEKFOptControlled.m:90-132 */
static const int A3_nnz[3] = {3, 3, 1};
static const int A3_col[3][3] = {{0, 1, 2}, {0, 1, 2}, {2, 0, 0}};
static const int A6_nnz[6] = {3, 3, 2, 4, 4, 5};
static const int A6_col[6][5] = {{0, 1, 2, 0, 0}, {0, 1, 2, 0, 0}, {2, 5, 0, 0, 0},
                                 {1, 2, 3, 4, 0}, {0, 2, 3, 4, 0}, {0, 1, 3, 4, 5}};

/* StateHardMargins: SIAlphaModelEKF.m:27-31 (s_min/i_min floors); all other
 * models clamp s,i to [0,1] (SIAlphaModelBackwardEKF.m:48-52,
 * SIAlphaModelEKFOptControlled.m:27-31, NewCaseEKF...m:150-154). */
static void state_margins(int model, const orc_params *p, double *s) {
  double lo_s = (model == ORC_SIALPHA) ? p->s_min : 0.0;
  double lo_i = (model == ORC_SIALPHA) ? p->i_min : 0.0;
  s[0] = mmin(1.0, mmax(lo_s, s[0]));
  s[1] = mmin(1.0, mmax(lo_i, s[1]));
  s[2] = mmin(p->alpha_max, mmax(p->alpha_min, s[2]));
}

/* NlinStateUpdate: SIAlphaModelEKF.m:39-48, SIAlphaModelBackwardEKF.m:60-69,
 * SIAlphaModelEKFOptControlled.m:39-74, ...BackwardEKFOptControlled.m:60-95,
 * NewCaseEKF...m:162-197 (tie-break >= at :175). */
static void nlin_state_update(int model, const orc_params *p, const double *u_in,
                              const double *s, double *u_out, double *sn) {
  const int L = p->L, six = model_dim(model) == 6, flip = model_flipped(model);
  double uu[ORC_LMAX];
  for (int j = 0; j < L; ++j) uu[j] = u_in[j];
  if (six) {
    double gs = p->gamma * s[5];
    for (int j = 0; j < L; ++j) {
      double phi = p->epsilon * p->w[j] - gs * p->a[j];
      if (uu[j] != uu[j]) {
        int to_min = model_legacy(model) ? (phi >= 0.0) : (phi > 0.0);
        uu[j] = to_min ? p->u_min[j] : p->u_max[j];
      }
    }
  }
  double dot = input_dot(p->gamma, p->a, p->u_max, uu, L);
  double dt = p->dt;
  double x0, x1, x2;
  double f2 = (((-p->gamma) * s[2]) + p->gamma * p->b) + dot;
  if (!flip) {
    x0 = s[0] - ((dt * s[2]) * s[0]) * s[1];
    x1 = s[1] + dt * (((s[2] * s[0]) * s[1]) - p->beta * s[1]);
    x2 = s[2] + dt * f2;
  } else {
    x0 = s[0] + ((dt * s[2]) * s[0]) * s[1];
    x1 = s[1] - dt * (((s[2] * s[0]) * s[1]) - p->beta * s[1]);
    x2 = s[2] - dt * f2;
  }
  double lo_s = (model == ORC_SIALPHA) ? p->s_min : 0.0;
  double lo_i = (model == ORC_SIALPHA) ? p->i_min : 0.0;
  sn[0] = mmax(lo_s, mmin(1.0, x0));
  sn[1] = mmax(lo_i, mmin(1.0, x1));
  sn[2] = mmax(p->alpha_min, mmin(p->alpha_max, x2));
  if (six) {
    double rho = (s[3] - s[4]) - (1.0 - p->epsilon);
    if (!flip) {
      sn[3] = s[3] + ((dt * rho) * s[2]) * s[1];
      sn[4] = s[4] + dt * (((rho * s[2]) * s[0]) + p->beta * s[4]);
      sn[5] = s[5] + dt * (((rho * s[0]) * s[1]) + p->gamma * s[5]);
    } else {
      sn[3] = s[3] - ((dt * rho) * s[2]) * s[1];
      sn[4] = s[4] - dt * (((rho * s[2]) * s[0]) + p->beta * s[4]);
      sn[5] = s[5] - dt * (((rho * s[0]) * s[1]) + p->gamma * s[5]);
    }
  }
  if (u_out) for (int j = 0; j < L; ++j) u_out[j] = uu[j];
}

/* StateJacobians: SIAlphaModelEKF.m:62-76, SIAlphaModelBackwardEKF.m:83-97,
 * SIAlphaModelEKFOptControlled.m:88-135, ...Backward...:109-156,
 * NewCaseEKF...m:211-257.  A is row-major A[i][j]; B = I is structural. */
static void state_jacobian(int model, const orc_params *p, const double *u, const double *s,
                           double A[MM][MM]) {
  const int six = model_dim(model) == 6, flip = model_flipped(model), L = p->L;
  const double dt = p->dt;
  memset(A, 0, sizeof(double) * MM * MM);
  if (!flip) {
    A[0][0] = 1.0 - (dt * s[2]) * s[1];
    A[0][1] = ((-dt) * s[2]) * s[0];
    A[0][2] = ((-dt) * s[0]) * s[1];
    A[1][0] = (dt * s[1]) * s[2];
    A[1][1] = 1.0 + dt * (s[0] * s[2] - p->beta);
    A[1][2] = (dt * s[0]) * s[1];
    A[2][2] = 1.0 - dt * p->gamma;
  } else {
    A[0][0] = 1.0 + (dt * s[2]) * s[1];
    A[0][1] = (dt * s[2]) * s[0];
    A[0][2] = (dt * s[0]) * s[1];
    A[1][0] = ((-dt) * s[1]) * s[2];
    A[1][1] = 1.0 - dt * (s[0] * s[2] - p->beta);
    A[1][2] = ((-dt) * s[0]) * s[1];
    A[2][2] = 1.0 + dt * p->gamma;
  }
  if (!six) return;
  double gs = p->gamma * s[5];
  double lo = (-1.0) / p->sigma, hi = 1.0 / p->sigma;
  double a25 = 0.0;
  for (int j = 0; j < L; ++j) {
    double phi = p->epsilon * p->w[j] - gs * p->a[j];
    if (u[j] != u[j]) {
      if (phi > lo && phi < hi) {
        double term = (((p->gamma * dt) * (p->sigma / 2.0)) * p->a[j]) * (p->u_max[j] - p->u_min[j]);
        a25 = flip ? (a25 + term) : (a25 - term);
      }
    }
  }
  A[2][5] = a25;
  double rho = (s[3] - s[4]) - (1.0 - p->epsilon);
  if (!flip) {
    A[3][1] = (dt * s[2]) * rho;
    A[3][2] = (dt * s[1]) * rho;
    A[3][3] = 1.0 + (dt * s[1]) * s[2];
    A[3][4] = ((-dt) * s[1]) * s[2];
    A[4][0] = (dt * s[2]) * rho;
    A[4][2] = (dt * s[0]) * rho;
    A[4][3] = (dt * s[0]) * s[2];
    A[4][4] = 1.0 - dt * (s[0] * s[2] - p->beta);
    A[5][0] = (dt * s[1]) * rho;
    A[5][1] = (dt * s[0]) * rho;
    A[5][3] = (dt * s[0]) * s[1];
    A[5][4] = ((-dt) * s[0]) * s[1];
    A[5][5] = 1.0 + dt * p->gamma;
  } else {
    A[3][1] = ((-dt) * s[2]) * rho;
    A[3][2] = ((-dt) * s[1]) * rho;
    A[3][3] = 1.0 - (dt * s[1]) * s[2];
    A[3][4] = (dt * s[1]) * s[2];
    A[4][0] = ((-dt) * s[2]) * rho;
    A[4][2] = ((-dt) * s[0]) * rho;
    A[4][3] = ((-dt) * s[0]) * s[2];
    A[4][4] = 1.0 + dt * (s[0] * s[2] - p->beta);
    A[5][0] = ((-dt) * s[1]) * rho;
    A[5][1] = ((-dt) * s[0]) * rho;
    A[5][3] = ((-dt) * s[0]) * s[1];
    A[5][4] = (dt * s[0]) * s[1];
    A[5][5] = 1.0 - dt * p->gamma;
  }
}

/* ObsJacobian + NlinObsUpdate + ObsHardMargins:
 * SIAlphaModelEKF.m:34-36,51-59,79-89 (identical in every model file);
 * MatlabCodeGenerator/ObsHardMargins.m:2-4 is the identity and
 * MatlabCodeGenerator/NlinObsUpdate.m:3 knows NEWCASES only. */
static int obs_model(int model, const orc_params *p, const double *s, double v_bar, double C[3],
                     double *xhat) {
  int ot = (model == ORC_LEGACY_CODEGEN) ? ORC_OBS_NEWCASES : p->obs_type;
  double xh;
  if (ot == ORC_OBS_NEWCASES) {
    C[0] = s[1] * s[2]; C[1] = s[0] * s[2]; C[2] = s[0] * s[1];
    xh = (s[0] * s[1]) * s[2] + v_bar;
  } else if (ot == ORC_OBS_TOTALCASES) {
    C[0] = -1.0; C[1] = 0.0; C[2] = 0.0;
    xh = (1.0 - s[0]) + v_bar;
  } else {
    return -1;
  }
  if (model != ORC_LEGACY_CODEGEN) xh = mmax(0.0, xh);
  *xhat = xh;
  return 0;
}

/* R = A * P with A's row pattern */
static void mul_A_P(int m, const double A[MM][MM], const double P[MM][MM], double R[MM][MM]) {
  for (int i = 0; i < m; ++i) {
    int nn = (m == 3) ? A3_nnz[i] : A6_nnz[i];
    const int *cl = (m == 3) ? A3_col[i] : A6_col[i];
    for (int j = 0; j < m; ++j) {
      double acc = A[i][cl[0]] * P[cl[0]][j];
      for (int t = 1; t < nn; ++t) acc = fma(A[i][cl[t]], P[cl[t]][j], acc);
      R[i][j] = acc;
    }
  }
}
/* R = X * A' with A's row pattern: R[i][j] = sum_{l in nz(A row j)} X[i][l] A[j][l] */
static void mul_X_At(int m, const double X[MM][MM], const double A[MM][MM], double R[MM][MM]) {
  for (int j = 0; j < m; ++j) {
    int nn = (m == 3) ? A3_nnz[j] : A6_nnz[j];
    const int *cl = (m == 3) ? A3_col[j] : A6_col[j];
    for (int i = 0; i < m; ++i) {
      double acc = X[i][cl[0]] * A[j][cl[0]];
      for (int t = 1; t < nn; ++t) acc = fma(X[i][cl[t]], A[j][cl[t]], acc);
      R[i][j] = acc;
    }
  }
}
/* dense R = X * Y */
static void mul_dense(int m, const double X[MM][MM], const double Y[MM][MM], double R[MM][MM]) {
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < m; ++j) {
      double acc = X[i][0] * Y[0][j];
      for (int l = 1; l < m; ++l) acc = fma(X[i][l], Y[l][j], acc);
      R[i][j] = acc;
    }
}
/* dense R = X * Y' */
static void mul_dense_T(int m, const double X[MM][MM], const double Y[MM][MM], double R[MM][MM]) {
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < m; ++j) {
      double acc = X[i][0] * Y[j][0];
      for (int l = 1; l < m; ++l) acc = fma(X[i][l], Y[j][l], acc);
      R[i][j] = acc;
    }
}

static void get_Q(int m, int q_mode, const double *Q, int k, double Qk[MM][MM]) {
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < m; ++j) {
      if (q_mode == ORC_Q_CONST) Qk[i][j] = Q[j * m + i];
      else if (q_mode == ORC_Q_PERDAY_FULL) Qk[i][j] = Q[(size_t)k * m * m + j * m + i];
      else Qk[i][j] = (i == j) ? Q[k] : 0.0; /* B*q*B' with B = I */
    }
}

static void load_cm(int m, const double *src, double M[MM][MM]) {
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < m; ++j) M[i][j] = src[j * m + i];
}
static void store_cm(int m, const double M[MM][MM], double *dst) {
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < m; ++j) dst[j * m + i] = M[i][j];
}

/* measurement update shared by generic and legacy variants:
 * gain (GenericExtendedKalmanFilter.m:124 / NewCaseEKF...m:63) and M = I - K*C */
static void gain_and_M(int m, const double P[MM][MM], const double C[3], double gammaR,
                       double K[MM], double M[MM][MM]) {
  double PCt[MM], CP[3];
  for (int i = 0; i < m; ++i)
    PCt[i] = fma(P[i][2], C[2], fma(P[i][1], C[1], P[i][0] * C[0]));
  for (int j = 0; j < 3; ++j)
    CP[j] = fma(C[2], P[2][j], fma(C[1], P[1][j], C[0] * P[0][j]));
  double S0 = fma(CP[2], C[2], fma(CP[1], C[1], CP[0] * C[0])); /* (C*P)*C' */
  double denom = S0 + gammaR;                                      /* + Gsp + Gvp (zeros) */
  for (int i = 0; i < m; ++i) K[i] = PCt[i] / denom;
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < m; ++j) {
      double e = (i == j) ? 1.0 : 0.0;
      M[i][j] = (j < 3) ? (e - K[i] * C[j]) : e; /* eye(m) - Kgain*Ck, C(4:6) = 0 */
    }
}
/* R = M * P with M = [dense cols 0..2 | identity cols 3..m-1] */
static void mul_M_P(int m, const double M[MM][MM], const double P[MM][MM], double R[MM][MM]) {
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < m; ++j) {
      double acc = fma(M[i][2], P[2][j], fma(M[i][1], P[1][j], M[i][0] * P[0][j]));
      if (i >= 3) acc = acc + P[i][j];
      R[i][j] = acc;
    }
}
/* R = X * M' */
static void mul_X_Mt(int m, const double X[MM][MM], const double M[MM][MM], double R[MM][MM]) {
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < m; ++j) {
      double acc = fma(X[i][2], M[j][2], fma(X[i][1], M[j][1], X[i][0] * M[j][0]));
      if (j >= 3) acc = acc + X[i][j];
      R[i][j] = acc;
    }
}

/* ------------------------------------------------------------------------- */
/* generic EKF + smoother core (no time flip)                                */
/* Tools/GenericExtendedKalmanFilter.m:41-233                                */
/* ------------------------------------------------------------------------- */
static int ekf_generic_core(int model, const orc_params *p, int T, const double *u,
                            const double *x, const double *s_init, const double *Ps_init,
                            const double *s_final, const double *Ps_final, double v_bar,
                            int q_mode, const double *Q, int r_mode, int fixed_R,
                            const double *Rin, double beta, double gamma, int W,
                            double *u_opt, double *u_opt_smooth, double *S_MINUS,
                            double *S_PLUS, double *S_SMOOTH, double *P_MINUS, double *P_PLUS,
                            double *P_SMOOTH, double *K_GAIN, double *innov_out, double *rho) {
  const int m = model_dim(model), L = p->L;
  const size_t mm = (size_t)m * m;
  double *R = (double *)malloc(sizeof(double) * (size_t)T);
  double *win = (double *)calloc((size_t)3 * W, sizeof(double));
  double *imean = win, *icov = win + W, *icovn = win + 2 * W; /* :55-57 */
  for (int k = 0; k < T; ++k) R[k] = (r_mode == ORC_R_CONST) ? Rin[0] : Rin[k]; /* :79-88 */

  double s[MM], P[MM][MM];
  for (int i = 0; i < m; ++i) s[i] = s_init[i];
  load_cm(m, Ps_init, P);
  if (u_opt_smooth) memset(u_opt_smooth, 0, sizeof(double) * (size_t)L * T); /* :95 */

  for (int k = 0; k < T; ++k) { /* :98 */
    for (int i = 0; i < m; ++i) S_MINUS[(size_t)k * m + i] = s[i]; /* :100 */
    store_cm(m, P, P_MINUS + k * mm);                               /* :101 */
    double C[3], xhat;
    if (obs_model(model, p, s, v_bar, C, &xhat)) { free(R); free(win); return -3; } /* :115-119 */
    double K[MM], sp[MM], Pp[MM][MM], innov;
    const double xk = x[k];
    if (!(xk != xk)) { /* :122 */
      innov = xk - xhat; /* :123 */
      double M[MM][MM], MP[MM][MM], MPM[MM][MM];
      gain_and_M(m, P, C, gamma * R[k], K, M); /* :124 */
      mul_M_P(m, M, P, MP);
      mul_X_Mt(m, MP, M, MPM);
      for (int i = 0; i < m; ++i)
        for (int j = 0; j < m; ++j)
          Pp[i][j] = (MPM[i][j] + (K[i] * R[k]) * K[j]) / gamma; /* :127 Joseph form */
      for (int i = 0; i < m; ++i) sp[i] = s[i] + K[i] * innov; /* :129 */
    } else { /* :131-134 */
      innov = 0.0;
      for (int i = 0; i < m; ++i) { K[i] = 0.0; sp[i] = s[i]; }
      memcpy(Pp, P, sizeof(P));
    }
    { /* :138 symmetrise */
      double t[MM][MM];
      for (int i = 0; i < m; ++i)
        for (int j = 0; j < m; ++j) t[i][j] = (Pp[i][j] + Pp[j][i]) / 2.0;
      memcpy(Pp, t, sizeof(t));
    }
    state_margins(model, p, sp); /* :141 */
    double sn[MM], A[MM][MM], AP[MM][MM], APA[MM][MM], Qk[MM][MM];
    nlin_state_update(model, p, u + (size_t)L * k, sp, u_opt ? u_opt + (size_t)L * k : NULL, sn); /* :155 */
    state_jacobian(model, p, u + (size_t)L * k, sp, A); /* :157 */
    mul_A_P(m, A, Pp, AP);
    mul_X_At(m, AP, A, APA);
    get_Q(m, q_mode, Q, k, Qk);
    for (int i = 0; i < m; ++i)
      for (int j = 0; j < m; ++j) P[i][j] = APA[i][j] + Qk[i][j]; /* :158 */
    {
      double t[MM][MM];
      for (int i = 0; i < m; ++i)
        for (int j = 0; j < m; ++j) t[i][j] = (P[i][j] + P[j][i]) / 2.0; /* :161 */
      memcpy(P, t, sizeof(t));
    }
    state_margins(model, p, sn); /* :164 */
    for (int i = 0; i < m; ++i) s[i] = sn[i];
    for (int i = 0; i < m; ++i) S_PLUS[(size_t)k * m + i] = sp[i]; /* :167 */
    store_cm(m, Pp, P_PLUS + k * mm);
    if (K_GAIN) for (int i = 0; i < m; ++i) K_GAIN[(size_t)k * m + i] = K[i];
    if (innov_out) innov_out[k] = innov;
    /* :172-185 innovation monitor; window slot 0 = newest, sums run 0..W-1 */
    int cnt = (k + 1 < W) ? k + 1 : W;
    for (int j = W - 1; j > 0; --j) imean[j] = imean[j - 1];
    imean[0] = innov;
    double sm = 0.0;
    for (int j = 0; j < W; ++j) sm += imean[j];
    double mu = sm / (double)cnt;
    double cc = (innov - mu) * (innov - mu);
    for (int j = W - 1; j > 0; --j) { icov[j] = icov[j - 1]; icovn[j] = icovn[j - 1]; }
    icov[0] = cc;
    icovn[0] = cc / (R[k] + ORC_EPS);
    double sn_ = 0.0;
    for (int j = 0; j < W; ++j) sn_ += icovn[j];
    if (rho) rho[k] = sn_ / (double)cnt;
    if (beta != 1.0 && !(xk != xk) && fixed_R && k + 1 < T) { /* :180 */
      double sc = 0.0;
      for (int j = 0; j < W; ++j) sc += icov[j];
      double R_estim = sc / (double)cnt;
      R[k + 1] = beta * R[k] + (1.0 - beta) * R_estim; /* :184 */
    }
  }

  /* smoother :189-230 */
  double ss[MM], Ps[MM][MM];
  for (int i = 0; i < m; ++i) ss[i] = S_PLUS[(size_t)(T - 1) * m + i];
  load_cm(m, P_PLUS + (size_t)(T - 1) * mm, Ps);
  for (int i = 0; i < m; ++i)
    if (!(s_final[i] != s_final[i])) ss[i] = s_final[i]; /* :195-196 */
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < m; ++j) {
      double f = Ps_final[j * m + i];
      if (!(f != f)) Ps[i][j] = f; /* :198-202 element-wise */
    }
  for (int i = 0; i < m; ++i) S_SMOOTH[(size_t)(T - 1) * m + i] = ss[i];
  store_cm(m, Ps, P_SMOOTH + (size_t)(T - 1) * mm);

  for (int k = T - 2; k >= 0; --k) { /* :204 */
    double spk[MM], A[MM][MM], Pp[MM][MM], Pm[MM][MM], J[MM][MM];
    for (int i = 0; i < m; ++i) spk[i] = S_PLUS[(size_t)k * m + i];
    state_jacobian(model, p, u + (size_t)L * k, spk, A); /* :206 */
    load_cm(m, P_PLUS + k * mm, Pp);
    load_cm(m, P_MINUS + (k + 1) * mm, Pm);
    int bad = 0;
    for (int i = 0; i < m; ++i)
      for (int j = 0; j < m; ++j) bad |= !(fabs(Pm[i][j]) <= 1.79769313486231570815e308); /* :211 */
    if (bad) {
      memset(J, 0, sizeof(J)); /* :213 */
    } else {
      double PAt[MM][MM], X[MM][MM], Xc[MM * MM];
      mul_X_At(m, Pp, A, PAt);
      orc_pinv_sym(P_MINUS + (k + 1) * mm, m, Xc, NULL);
      load_cm(m, Xc, X);
      mul_dense(m, PAt, X, J); /* :215 */
    }
    double ds[MM], sk[MM];
    for (int l = 0; l < m; ++l) ds[l] = ss[l] - S_MINUS[(size_t)(k + 1) * m + l];
    for (int i = 0; i < m; ++i) {
      double acc = J[i][0] * ds[0];
      for (int l = 1; l < m; ++l) acc = fma(J[i][l], ds[l], acc);
      sk[i] = spk[i] + acc; /* :218 */
    }
    state_margins(model, p, sk); /* :221 */
    double D[MM][MM], JD[MM][MM], JDJ[MM][MM], Pn[MM][MM];
    for (int i = 0; i < m; ++i)
      for (int j = 0; j < m; ++j) D[i][j] = Pm[i][j] - Ps[i][j];
    mul_dense(m, J, D, JD);
    mul_dense_T(m, JD, J, JDJ);
    for (int i = 0; i < m; ++i)
      for (int j = 0; j < m; ++j) Pn[i][j] = Pp[i][j] - JDJ[i][j]; /* :223 */
    for (int i = 0; i < m; ++i)
      for (int j = 0; j < m; ++j) Ps[i][j] = (Pn[i][j] + Pn[j][i]) / 2.0; /* :226 */
    for (int i = 0; i < m; ++i) ss[i] = sk[i];
    for (int i = 0; i < m; ++i) S_SMOOTH[(size_t)k * m + i] = ss[i];
    store_cm(m, Ps, P_SMOOTH + k * mm);
    if (u_opt_smooth) {
      double dummy[MM];
      nlin_state_update(model, p, u + (size_t)L * k, ss, u_opt_smooth + (size_t)L * k, dummy); /* :229 */
    }
  }
  free(R); free(win);
  return 0;
}

/* ------------------------------------------------------------------------- */
/* legacy monolith core: Tools/NewCaseEKFEstimatorWithOptimalNPI.m:9-143     */
/* ------------------------------------------------------------------------- */
static int ekf_legacy_core(int model, const orc_params *p, int T, const double *u,
                           const double *x, const double *s_init, const double *Ps_init,
                           const double *s_final, const double *Ps_final, double v_bar,
                           const double *Q, double R0, double beta, double gamma, int W,
                           double *u_opt, double *S_MINUS, double *S_PLUS, double *S_SMOOTH,
                           double *P_MINUS, double *P_PLUS, double *P_SMOOTH, double *K_GAIN,
                           double *innov_out, double *rho) {
  const int m = 6, L = p->L;
  const size_t mm = 36;
  double *win = (double *)calloc((size_t)3 * W, sizeof(double));
  double *imean = win, *icov = win + W, *icovn = win + 2 * W;
  double R = R0; /* :31 scalar, adapted in place */
  double s[MM], P[MM][MM], Qm[MM][MM];
  for (int i = 0; i < m; ++i) s[i] = s_init[i];
  load_cm(m, Ps_init, P);
  load_cm(m, Q, Qm);
  for (int k = 0; k < T; ++k) {
    for (int i = 0; i < m; ++i) S_MINUS[(size_t)k * m + i] = s[i];
    store_cm(m, P, P_MINUS + k * mm);
    double C[3], xhat;
    if (obs_model(model, p, s, v_bar, C, &xhat)) { free(win); return -3; }
    double K[MM], sp[MM], Pp[MM][MM], innov;
    const double xk = x[k];
    if (!(xk != xk)) {
      innov = xk - xhat;
      double M[MM][MM], MP[MM][MM];
      gain_and_M(m, P, C, gamma * R, K, M); /* :63 */
      mul_M_P(m, M, P, MP);
      for (int i = 0; i < m; ++i)
        for (int j = 0; j < m; ++j) Pp[i][j] = MP[i][j] / gamma; /* :64 */
      for (int i = 0; i < m; ++i) sp[i] = s[i] + K[i] * innov;
    } else {
      innov = 0.0;
      for (int i = 0; i < m; ++i) { K[i] = 0.0; sp[i] = s[i]; }
      memcpy(Pp, P, sizeof(P));
    }
    state_margins(model, p, sp); /* :74 (no symmetrisation in the legacy file) */
    double sn[MM], A[MM][MM], AP[MM][MM], APA[MM][MM];
    nlin_state_update(model, p, u + (size_t)L * k, sp, u_opt ? u_opt + (size_t)L * k : NULL, sn);
    state_jacobian(model, p, u + (size_t)L * k, sp, A);
    mul_A_P(m, A, Pp, AP);
    mul_X_At(m, AP, A, APA);
    for (int i = 0; i < m; ++i)
      for (int j = 0; j < m; ++j) P[i][j] = APA[i][j] + Qm[i][j]; /* :91 */
    state_margins(model, p, sn); /* :94 */
    for (int i = 0; i < m; ++i) s[i] = sn[i];
    for (int i = 0; i < m; ++i) S_PLUS[(size_t)k * m + i] = sp[i];
    store_cm(m, Pp, P_PLUS + k * mm);
    if (K_GAIN) for (int i = 0; i < m; ++i) K_GAIN[(size_t)k * m + i] = K[i];
    if (innov_out) innov_out[k] = innov;
    int cnt = (k + 1 < W) ? k + 1 : W; /* :102-112 */
    for (int j = W - 1; j > 0; --j) imean[j] = imean[j - 1];
    imean[0] = innov;
    double sm = 0.0;
    for (int j = 0; j < W; ++j) sm += imean[j];
    double mu = sm / (double)cnt;
    double cc = (innov - mu) * (innov - mu);
    for (int j = W - 1; j > 0; --j) { icov[j] = icov[j - 1]; icovn[j] = icovn[j - 1]; }
    icov[0] = cc;
    icovn[0] = cc / R; /* :108 no +eps */
    double sn_ = 0.0;
    for (int j = 0; j < W; ++j) sn_ += icovn[j];
    if (rho) rho[k] = sn_ / (double)cnt;
    if (beta != 1.0 && !(xk != xk)) { /* :110 */
      double sc = 0.0;
      for (int j = 0; j < W; ++j) sc += icov[j];
      R = beta * R + ((1.0 - beta) * sc) / (double)cnt; /* :111 */
    }
  }
  double ss[MM], Ps[MM][MM];
  for (int i = 0; i < m; ++i) ss[i] = S_PLUS[(size_t)(T - 1) * m + i];
  load_cm(m, P_PLUS + (size_t)(T - 1) * mm, Ps);
  for (int i = 0; i < m; ++i)
    if (!(s_final[i] != s_final[i])) ss[i] = s_final[i]; /* :122-123 */
  { /* :125-127 sub-matrix assignment P_SMOOTH(row, col, T) = Ps_final(row, col) */
    int rowset[MM] = {0}, colset[MM] = {0};
    for (int i = 0; i < m; ++i)
      for (int j = 0; j < m; ++j) {
        double f = Ps_final[j * m + i];
        if (!(f != f)) { rowset[i] = 1; colset[j] = 1; }
      }
    for (int i = 0; i < m; ++i)
      for (int j = 0; j < m; ++j)
        if (rowset[i] && colset[j]) Ps[i][j] = Ps_final[j * m + i];
  }
  for (int i = 0; i < m; ++i) S_SMOOTH[(size_t)(T - 1) * m + i] = ss[i];
  store_cm(m, Ps, P_SMOOTH + (size_t)(T - 1) * mm);
  for (int k = T - 2; k >= 0; --k) { /* :129-139 */
    double spk[MM], A[MM][MM], Pp[MM][MM], Pm[MM][MM], J[MM][MM], PAt[MM][MM];
    double Nc[MM * MM], Xc[MM * MM];
    for (int i = 0; i < m; ++i) spk[i] = S_PLUS[(size_t)k * m + i];
    state_jacobian(model, p, u + (size_t)L * k, spk, A);
    load_cm(m, P_PLUS + k * mm, Pp);
    load_cm(m, P_MINUS + (k + 1) * mm, Pm);
    mul_X_At(m, Pp, A, PAt);
    store_cm(m, PAt, Nc);
    orc_mrdivide(Nc, P_MINUS + (k + 1) * mm, m, Xc); /* :132 */
    load_cm(m, Xc, J);
    double ds[MM], sk[MM];
    for (int l = 0; l < m; ++l) ds[l] = ss[l] - S_MINUS[(size_t)(k + 1) * m + l];
    for (int i = 0; i < m; ++i) {
      double acc = J[i][0] * ds[0];
      for (int l = 1; l < m; ++l) acc = fma(J[i][l], ds[l], acc);
      sk[i] = spk[i] + acc;
    }
    state_margins(model, p, sk);
    double D[MM][MM], JD[MM][MM], JDJ[MM][MM];
    for (int i = 0; i < m; ++i)
      for (int j = 0; j < m; ++j) D[i][j] = Pm[i][j] - Ps[i][j];
    mul_dense(m, J, D, JD);
    mul_dense_T(m, JD, J, JDJ);
    for (int i = 0; i < m; ++i)
      for (int j = 0; j < m; ++j) Ps[i][j] = Pp[i][j] - JDJ[i][j]; /* :138 no symmetrisation */
    for (int i = 0; i < m; ++i) ss[i] = sk[i];
    for (int i = 0; i < m; ++i) S_SMOOTH[(size_t)k * m + i] = ss[i];
    store_cm(m, Ps, P_SMOOTH + k * mm);
  }
  free(win);
  return 0;
}

static void flip_cols(double *a, size_t col, int T) {
  if (!a) return;
  for (int k = 0; k < T / 2; ++k)
    for (size_t i = 0; i < col; ++i) {
      double t = a[(size_t)k * col + i];
      a[(size_t)k * col + i] = a[(size_t)(T - 1 - k) * col + i];
      a[(size_t)(T - 1 - k) * col + i] = t;
    }
}

int orc_ekf_eks(int model, const orc_params *prm, int T, const double *u, const double *x,
                const double *s_init, const double *Ps_init, const double *s_final,
                const double *Ps_final, double v_bar, int q_mode, const double *Q, int r_mode,
                int fixed_R, const double *R, double beta, double gamma, int W, int order,
                double *u_opt, double *u_opt_smooth, double *S_MINUS, double *S_PLUS,
                double *S_SMOOTH, double *P_MINUS, double *P_PLUS, double *P_SMOOTH,
                double *K_GAIN, double *innovations, double *rho) {
  if (order == 2) order = 1; /* every model's Hessian terms are identically zero
                                (SIAlphaModelEKF.m:92-109 etc.): order 2 == order 1 */
  if (order != 1) return -2; /* GenericExtendedKalmanFilter.m:111 'Undefined order' */
  if (model < 0 || model > ORC_LEGACY_CODEGEN || T < 1 || W < 1) return -1;
  const int m = model_dim(model), L = prm->L;
  const size_t mm = (size_t)m * m;
  /* the tape the smoother needs is always materialised */
  double *tS_M = S_MINUS ? S_MINUS : (double *)malloc(sizeof(double) * m * T);
  double *tS_P = S_PLUS ? S_PLUS : (double *)malloc(sizeof(double) * m * T);
  double *tS_S = S_SMOOTH ? S_SMOOTH : (double *)malloc(sizeof(double) * m * T);
  double *tP_M = P_MINUS ? P_MINUS : (double *)malloc(sizeof(double) * mm * T);
  double *tP_P = P_PLUS ? P_PLUS : (double *)malloc(sizeof(double) * mm * T);
  double *tP_S = P_SMOOTH ? P_SMOOTH : (double *)malloc(sizeof(double) * mm * T);
  int rc;
  if (model_legacy(model)) {
    rc = ekf_legacy_core(model, prm, T, u, x, s_init, Ps_init, s_final, Ps_final, v_bar, Q,
                         R[0], beta, gamma, W, u_opt, tS_M, tS_P, tS_S, tP_M, tP_P, tP_S,
                         K_GAIN, innovations, rho);
  } else if (!model_flipped(model)) {
    rc = ekf_generic_core(model, prm, T, u, x, s_init, Ps_init, s_final, Ps_final, v_bar,
                          q_mode, Q, r_mode, fixed_R, R, beta, gamma, W, u_opt, u_opt_smooth,
                          tS_M, tS_P, tS_S, tP_M, tP_P, tP_S, K_GAIN, innovations, rho);
  } else {
    /* SIAlphaModelBackwardEKF.m:19-40: flip u and x, swap init/final; Q_w and
     * R_v are NOT flipped (:27); flip every output back except rho (:40 is a
     * no-op on the squeezed T x 1 rho). */
    double *uf = (double *)malloc(sizeof(double) * (size_t)L * T);
    double *xf = (double *)malloc(sizeof(double) * (size_t)T);
    for (int k = 0; k < T; ++k) {
      memcpy(uf + (size_t)L * k, u + (size_t)L * (T - 1 - k), sizeof(double) * L);
      xf[k] = x[T - 1 - k];
    }
    rc = ekf_generic_core(model, prm, T, uf, xf, s_final, Ps_final, s_init, Ps_init, v_bar,
                          q_mode, Q, r_mode, fixed_R, R, beta, gamma, W, u_opt, u_opt_smooth,
                          tS_M, tS_P, tS_S, tP_M, tP_P, tP_S, K_GAIN, innovations, rho);
    flip_cols(u_opt, L, T); flip_cols(u_opt_smooth, L, T);
    flip_cols(S_MINUS, m, T); flip_cols(S_PLUS, m, T); flip_cols(S_SMOOTH, m, T);
    flip_cols(P_MINUS, mm, T); flip_cols(P_PLUS, mm, T); flip_cols(P_SMOOTH, mm, T);
    flip_cols(K_GAIN, m, T); flip_cols(innovations, 1, T);
    free(uf); free(xf);
  }
  if (!S_MINUS) free(tS_M);
  if (!S_PLUS) free(tS_P);
  if (!S_SMOOTH) free(tS_S);
  if (!P_MINUS) free(tP_M);
  if (!P_PLUS) free(tP_P);
  if (!P_SMOOTH) free(tP_S);
  return rc;
}

/* ------------------------------------------------------------------------- */
/* one region of the Pareto sweep: TrainPredictPrescribeNPI.m:421-495,624-633 */
/* ------------------------------------------------------------------------- */
static void sweep_one(const orc_sweep_region *rg, double eps, const double *noise, double *J0,
                      double *J1, double *u_fore) {
  const int L = rg->prm.L, T = rg->T, Th = rg->T_hist, Tf = T - Th;
  orc_params p = rg->prm;
  p.epsilon = eps; /* :424 */
  double *u = (double *)malloc(sizeof(double) * (size_t)L * T);
  double *uos = (double *)malloc(sizeof(double) * (size_t)L * T);
  double *nc = (double *)malloc(sizeof(double) * (size_t)T);
  double *rs = (double *)malloc(sizeof(double) * (size_t)3 * (Tf > 0 ? Tf : 1));
  memcpy(u, rg->u_hist, sizeof(double) * (size_t)L * Th);
  for (size_t k = (size_t)L * Th; k < (size_t)L * T; ++k) u[k] = NAN; /* :458 */
  orc_ekf_eks(ORC_OPTCTRL, &p, T, u, rg->x, rg->s_init, rg->Ps_init, rg->s_final, rg->Ps_final,
              0.0, ORC_Q_CONST, rg->Q, ORC_R_PERDAY, 0, rg->R, rg->beta_ekf, rg->gamma_ekf,
              rg->W, 1, NULL, uos, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL); /* :460 */
  orc_sialpha_controlled(uos + (size_t)L * Th, L, rg->s_h, rg->i_h, rg->alpha_h, p.u_max,
                         p.alpha_min, p.alpha_max, p.gamma, p.a, p.b, p.beta, rg->noise_std[0],
                         rg->noise_std[1], rg->noise_std[2], Tf, p.dt, noise, rs, rs + Tf,
                         rs + 2 * Tf); /* :481 */
  memcpy(nc, rg->newcases_hist, sizeof(double) * (size_t)Th);
  for (int t = 0; t < Tf; ++t) nc[Th + t] = (rs[t] * rs[Tf + t]) * rs[2 * Tf + t]; /* :493 s.*i.*alpha */
  orc_npicost(nc, T, uos, rg->weights, L, J0, J1);
  if (u_fore) memcpy(u_fore, uos + (size_t)L * Th, sizeof(double) * (size_t)L * Tf);
  free(u); free(uos); free(nc); free(rs);
}

void orc_sweep_region_run(const orc_sweep_region *rg, const double *eps, int n_eps, double *J0,
                          double *J1, double *u_fore, unsigned char *on_front, int *I_opt) {
  const int L = rg->prm.L, Tf = rg->T - rg->T_hist;
  for (int e = 0; e < n_eps; ++e)
    sweep_one(rg, eps[e], rg->noise ? rg->noise + (size_t)e * 3 * Tf : NULL, J0 + e, J1 + e,
              u_fore ? u_fore + (size_t)e * L * Tf : NULL);
  if (on_front) orc_pareto(J0, J1, n_eps, on_front, I_opt);
}

void orc_sweep_batch(const orc_sweep_region *rg, int n_regions, const double *eps, int n_eps,
                     double *J0, double *J1, unsigned char *on_front, int *I_opt,
                     int n_threads) {
  const long total = (long)n_regions * n_eps;
#ifdef _OPENMP
  if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel for schedule(dynamic, 4)
  for (long q = 0; q < total; ++q) {
    int r = (int)(q / n_eps), e = (int)(q % n_eps);
    const int Tf = rg[r].T - rg[r].T_hist;
    sweep_one(&rg[r], eps[e], rg[r].noise ? rg[r].noise + (size_t)e * 3 * Tf : NULL,
              J0 + q, J1 + q, NULL);
  }
  if (on_front)
    for (int r = 0; r < n_regions; ++r)
      orc_pareto(J0 + (size_t)r * n_eps, J1 + (size_t)r * n_eps, n_eps,
                 on_front + (size_t)r * n_eps, I_opt ? I_opt + r : NULL);
}

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* ------------------------------------------------------------------------- */
/* Rt_ExpFitEKF: 2-state exponential-fit EKF/EKS with second-order terms     */
/* (Tools/Rt_ExpFitEKF.m:1-227) -- SURVEY 8f-3                               */
/* ------------------------------------------------------------------------- */
/* 2x2 matrices are row-major double[4] = {m11, m12, m21, m22}.  Products are DEFINED as
 * c_ij = fma(a_i2, b_2j, a_i1 * b_1j) with every term kept (structural zeros included:
 * exact for finite operands), as the arithmetic contract of DESIGN.md states. */
static void mm2(const double *a, const double *b, double *c) {
  double t[4];
  t[0] = fma(a[1], b[2], a[0] * b[0]); t[1] = fma(a[1], b[3], a[0] * b[1]);
  t[2] = fma(a[3], b[2], a[2] * b[0]); t[3] = fma(a[3], b[3], a[2] * b[1]);
  c[0] = t[0]; c[1] = t[1]; c[2] = t[2]; c[3] = t[3];
}
static void tr2(const double *a, double *t) { t[0] = a[0]; t[1] = a[2]; t[2] = a[1]; t[3] = a[3]; }
/* trace(P*Fi*P*Fj)/2 evaluated left to right (:178) */
static double half_trace4(const double *P, const double *Fi, const double *Fj) {
  double a[4], b[4], c[4];
  mm2(P, Fi, a); mm2(a, P, b); mm2(b, Fj, c);
  return (c[0] + c[3]) / 2.0;
}
static double half_trace2(const double *P, const double *F) {
  double a[4];
  mm2(P, F, a);
  return (a[0] + a[3]) / 2.0;
}

/* x[T] (NaN = missing), s_init[2], params[3] = {time_scale, alpha, sigma} (:121-123), w_bar[2],
 * Ps_init / Q row-major 2x2, scalar R.  Outputs (any may be NULL except the four tape arrays):
 * S_* [T][2], P_* [T][4] row-major, K_GAIN [T][2], innovations [T], rho [T].
 * Returns 0, or -2 for order not in {1,2} (:47,:75 'Undefined order'). */
int orc_rt_expfit_ekf(const double *x, int T, const double *s_init, const double *params, const double *w_bar,
                      double v_bar, const double *Ps_init, const double *Q, double R, double beta, double gamma,
                      int W, int order, double *S_MINUS, double *S_PLUS, double *P_MINUS, double *P_PLUS,
                      double *K_GAIN, double *S_SMOOTH, double *P_SMOOTH, double *innovations, double *rho) {
  if (order != 1 && order != 2) return -2;
  const double ts = params[0], alpha = params[1], sigma = params[2];
  double sm[2] = {s_init[0], s_init[1]}, Pm[4] = {Ps_init[0], Ps_init[1], Ps_init[2], Ps_init[3]};
  double *winM = (double *)calloc((size_t)3 * W, sizeof(double)), *winC = winM + W, *winN = winM + 2 * W;
  for (int k = 0; k < T; ++k) {
    S_MINUS[2 * k] = sm[0]; S_MINUS[2 * k + 1] = sm[1];                 /* :37-38 */
    for (int q = 0; q < 4; ++q) P_MINUS[4 * k + q] = Pm[q];
    /* :40-49 observation Hessian terms: Gs = Gv = {0} => gs = Gsp = gv = Gvp = 0 for either order */
    const double xhat = (sm[0] + v_bar) + 0.0 + 0.0;                       /* :53, :131 */
    double innov, Kg[2], sp[2], Pp[4];
    if (!(x[k] != x[k])) {                                                 /* :56 */
      innov = x[k] - xhat;                                                 /* :57 */
      const double denom = ((Pm[0] + gamma * ((1.0 * R) * 1.0)) + 0.0) + 0.0; /* :58, C = [1 0], D = 1 */
      Kg[0] = Pm[0] / denom; Kg[1] = Pm[2] / denom;
      const double M[4] = {1.0 - Kg[0], 0.0 - 0.0, 0.0 - Kg[1], 1.0 - 0.0}; /* eye(m) - Kgain*C */
      double MP[4];
      mm2(M, Pm, MP);
      for (int q = 0; q < 4; ++q) Pp[q] = MP[q] / gamma;                   /* :59 */
      sp[0] = sm[0] + Kg[0] * innov; sp[1] = sm[1] + Kg[1] * innov;        /* :60 */
    } else {                                                               /* :62-65 */
      innov = 0.0; Kg[0] = Kg[1] = 0.0;
      for (int q = 0; q < 4; ++q) Pp[q] = Pm[q];
      sp[0] = sm[0]; sp[1] = sm[1];
    }
    const double E = exp(ts * sp[1]);
    const double tnh = tanh((alpha * sp[1] + w_bar[1]) / sigma);
    double fs[2] = {0, 0}, fw[2] = {0, 0}, Fsp[4] = {0, 0, 0, 0}, Fwp[4] = {0, 0, 0, 0};
    if (order == 2) {                                                      /* :153-196 */
      const double f12 = ts * E;
      const double Fs1[4] = {0.0, f12, f12, ((ts * ts) * sp[0]) * E};
      const double Fs2[4] = {0.0, 0.0, 0.0, ((((-2.0) * (alpha * alpha)) / sigma) * tnh) * (1.0 - tnh * tnh)};
      const double Fw1[4] = {0.0, 0.0, 0.0, 0.0};
      const double Fw2[4] = {0.0, 0.0, 0.0, (((-2.0) / sigma) * tnh) * (1.0 - tnh * tnh)};
      const double *Fs[2] = {Fs1, Fs2}, *Fw[2] = {Fw1, Fw2};
      for (int i = 0; i < 2; ++i) {
        fs[i] = half_trace2(Pp, Fs[i]);
        fw[i] = half_trace2(Q, Fw[i]);
        for (int j = 0; j < 2; ++j) {
          Fsp[2 * i + j] = half_trace4(Pp, Fs[i], Fs[j]);
          Fwp[2 * i + j] = half_trace4(Q, Fw[i], Fw[j]);
        }
      }
    }
    /* :80 state update (:120-127) + second-order means */
    sm[0] = ((sp[0] * E + w_bar[0]) + fs[0]) + fw[0];
    sm[1] = ((sigma * tnh) + fs[1]) + fw[1];
    /* :81-82 */
    const double omt = 1.0 - tnh * tnh;
    const double A[4] = {E, (ts * sp[0]) * E, 0.0, alpha * omt};
    const double Bm[4] = {1.0, 0.0, 0.0, omt};
    double At[4], Bt[4], AP[4], APA[4], BQ[4], BQB[4];
    tr2(A, At); tr2(Bm, Bt);
    mm2(A, Pp, AP); mm2(AP, At, APA);
    mm2(Bm, Q, BQ); mm2(BQ, Bt, BQB);
    for (int q = 0; q < 4; ++q) Pm[q] = ((APA[q] + BQB[q]) + Fsp[q]) + Fwp[q];
    S_PLUS[2 * k] = sp[0]; S_PLUS[2 * k + 1] = sp[1];                      /* :85-87 */
    for (int q = 0; q < 4; ++q) P_PLUS[4 * k + q] = Pp[q];
    if (K_GAIN) { K_GAIN[2 * k] = Kg[0]; K_GAIN[2 * k + 1] = Kg[1]; }
    if (innovations) innovations[k] = innov;
    /* :90-101 innovation monitor, windows newest first */
    const int cnt = (k + 1 < W) ? (k + 1) : W;
    for (int j = W - 1; j > 0; --j) { winM[j] = winM[j - 1]; winC[j] = winC[j - 1]; winN[j] = winN[j - 1]; }
    winM[0] = innov;
    double sM = 0.0;
    for (int j = 0; j < W; ++j) sM += winM[j];
    const double mu = sM / (double)cnt;
    const double cc = (innov - mu) * (innov - mu);
    winC[0] = cc;
    winN[0] = cc / R;                                                      /* :97 (no eps) */
    double sN = 0.0;
    for (int j = 0; j < W; ++j) sN += winN[j];
    if (rho) rho[k] = sN / (double)cnt;
    if (beta != 1.0 && !(x[k] != x[k])) {                                  /* :99-101 */
      double sC = 0.0;
      for (int j = 0; j < W; ++j) sC += winC[j];
      R = beta * R + ((1.0 - beta) * sC) / (double)cnt;
    }
  }
  free(winM);
  /* :104-115 smoother */
  if (S_SMOOTH && T > 0) {
    double ss[2] = {S_PLUS[2 * (T - 1)], S_PLUS[2 * (T - 1) + 1]}, Ps[4];
    for (int q = 0; q < 4; ++q) Ps[q] = P_PLUS[4 * (T - 1) + q];
    S_SMOOTH[2 * (T - 1)] = ss[0]; S_SMOOTH[2 * (T - 1) + 1] = ss[1];
    if (P_SMOOTH) for (int q = 0; q < 4; ++q) P_SMOOTH[4 * (T - 1) + q] = Ps[q];
    for (int k = T - 2; k >= 0; --k) {
      const double *sp = S_PLUS + 2 * k, *Pp = P_PLUS + 4 * k, *Pn = P_MINUS + 4 * (k + 1), *sn = S_MINUS + 2 * (k + 1);
      const double E = exp(ts * sp[1]);
      const double tnh = tanh((alpha * sp[1] + w_bar[1]) / sigma);
      const double A[4] = {E, (ts * sp[0]) * E, 0.0, alpha * (1.0 - tnh * tnh)};
      double At[4], PAt[4], Bc[4], Ac[4], Jc[4], J[4];
      tr2(A, At); mm2(Pp, At, PAt);
      /* orc_mrdivide takes column-major m x m */
      Bc[0] = PAt[0]; Bc[1] = PAt[2]; Bc[2] = PAt[1]; Bc[3] = PAt[3];
      Ac[0] = Pn[0]; Ac[1] = Pn[2]; Ac[2] = Pn[1]; Ac[3] = Pn[3];
      orc_mrdivide(Bc, Ac, 2, Jc);                                         /* :110 */
      J[0] = Jc[0]; J[1] = Jc[2]; J[2] = Jc[1]; J[3] = Jc[3];
      const double d0 = ss[0] - sn[0], d1 = ss[1] - sn[1];
      const double n0 = sp[0] + fma(J[1], d1, J[0] * d0), n1 = sp[1] + fma(J[3], d1, J[2] * d0); /* :111 */
      double D[4], Jt[4], JD[4], JDJ[4];
      for (int q = 0; q < 4; ++q) D[q] = Pn[q] - Ps[q];
      tr2(J, Jt); mm2(J, D, JD); mm2(JD, Jt, JDJ);
      for (int q = 0; q < 4; ++q) Ps[q] = Pp[q] - JDJ[q];                  /* :112 */
      ss[0] = n0; ss[1] = n1;
      S_SMOOTH[2 * k] = ss[0]; S_SMOOTH[2 * k + 1] = ss[1];
      if (P_SMOOTH) for (int q = 0; q < 4; ++q) P_SMOOTH[4 * k + q] = Ps[q];
    }
  }
  return 0;
}

/* ------------------------------------------------------------------------- */
/* Per-region preprocessing that feeds the filters              */
/* Tools/TrainPredictPrescribeNPI.m:121-128 (NPI fill), :165-187 (case series), */
/* :200-201 (I0), :240 (R_v)                                                   */
/* ------------------------------------------------------------------------- */
/* y = filter(b, 1, x, zi) for an FIR b[0..nb-1] in MATLAB's direct form II transposed:
 *   y(n) = b1 x(n) + z1;  z_i = b_{i+1} x(n) + z_{i+1}  (i < nb-1);  z_{nb-1} = b_nb x(n).
 * z (nb-1 values) is updated in place. */
static void fir_df2t(const double *b, int nb, const double *x, int n, double *z, double *y) {
  for (int t = 0; t < n; ++t) {
    const double xt = x[t];
    const double yt = (nb > 1) ? (b[0] * xt + z[0]) : (b[0] * xt);
    for (int i = 0; i + 2 < nb; ++i) z[i] = b[i + 1] * xt + z[i + 1];
    if (nb > 1) z[nb - 2] = b[nb - 1] * xt;
    if (y) y[t] = yt;
  }
}
/* MATLAB filtfilt(b, a, x) for an FIR with scalar a (normalised first, as filter does):
 * odd reflection of nfact = 3(nb-1) samples at both ends, initial states zi * (first sample)
 * with zi the steady state of a unit input (zi_i = sum_{j>i} b_j), forward pass, reversed pass. */
static int fir_filtfilt(const double *b, int nb, const double *x, int n, double *y) {
  const int nfact = (3 * (nb - 1) > 1) ? 3 * (nb - 1) : 1;
  if (n <= nfact) return -1; /* "Data length must be larger than 3 times the filter order" */
  const int ne = n + 2 * nfact;
  double *xe = (double *)malloc(sizeof(double) * (size_t)ne * 2), *ye = xe + ne;
  double zi[32], z[32];
  for (int i = 0; i + 1 < nb; ++i) { /* zi(i) = b(i+1) + zi(i+1) solved from the last one up */
    zi[i] = 0.0;
  }
  for (int i = nb - 2; i >= 0; --i) zi[i] = b[i + 1] + ((i + 1 < nb - 1) ? zi[i + 1] : 0.0);
  for (int i = 0; i < nfact; ++i) xe[i] = 2.0 * x[0] - x[nfact - i];
  for (int i = 0; i < n; ++i) xe[nfact + i] = x[i];
  for (int i = 0; i < nfact; ++i) xe[nfact + n + i] = 2.0 * x[n - 1] - x[n - 2 - i];
  for (int i = 0; i + 1 < nb; ++i) z[i] = zi[i] * xe[0];
  fir_df2t(b, nb, xe, ne, z, ye);
  for (int i = 0; i < ne / 2; ++i) { double t = ye[i]; ye[i] = ye[ne - 1 - i]; ye[ne - 1 - i] = t; }
  for (int i = 0; i + 1 < nb; ++i) z[i] = zi[i] * ye[0];
  fir_df2t(b, nb, ye, ne, z, xe);
  for (int i = 0; i < n; ++i) y[i] = xe[ne - 1 - nfact - i];
  free(xe);
  return 0;
}

/* One region.  cc[T] cumulative confirmed cases (NaN allowed), ip[T][L] NPI levels (NaN allowed,
 * filled in place), population N, W = SmoothingWinLen.  Outputs [T] each: refined, smoothed,
 * zerolag, normalized (= smoothed/N), confirmed_norm (= cumsum(smoothed)/N), R_v; *I0.
 * Returns 0, -1 if T < 2 (:166 "Insufficient data"), -2 if T <= 3*(round(W/2)-1) (filtfilt). */
int orc_preprocess_region(const double *cc, int T, double N, int W, int n_first, double min_cases, double *ip,
                          int L, double *refined, double *smoothed, double *zerolag, double *normalized,
                          double *confirmed_norm, double *R_v, double *I0) {
  if (T < 2) return -1;
  /* :121-128 */
  for (int j = 0; j < L; ++j)
    for (int i = 1; i < T; ++i)
      if (ip[i * L + j] != ip[i * L + j] && !(ip[(i - 1) * L + j] != ip[(i - 1) * L + j])) ip[i * L + j] = ip[(i - 1) * L + j];
  for (int q = 0; q < T * L; ++q)
    if (ip[q] != ip[q]) ip[q] = 0.0;
  /* :162-172 */
  for (int t = 0; t < T; ++t) {
    double d = cc[t] - cc[t > 0 ? t - 1 : 0]; /* diff([cc(1); cc]) */
    if (d < 0.0) d = 0.0;
    refined[t] = d;
  }
  if (refined[T - 1] != refined[T - 1]) {
    int last = -1;
    for (int t = T - 1; t >= 0; --t)
      if (!(refined[t] != refined[t])) { last = t; break; }
    if (last >= 0) refined[T - 1] = refined[last];
  }
  for (int t = 0; t < T; ++t)
    if (refined[t] != refined[t]) refined[t] = 0.0;
  /* :173 causal moving average */
  double b[32], z[32];
  for (int i = 0; i < W; ++i) { b[i] = 1.0 / (double)W; z[i] = 0.0; }
  fir_df2t(b, W, refined, T, z, smoothed);
  /* :174 zero-phase moving average of round(W/2) taps */
  const int Wh = (int)floor((double)W / 2.0 + 0.5);
  for (int i = 0; i < Wh; ++i) b[i] = 1.0 / (double)Wh;
  if (fir_filtfilt(b, Wh, refined, T, zerolag)) return -2;
  /* :175-180, :240 */
  double cs = 0.0;
  for (int t = 0; t < T; ++t) {
    normalized[t] = smoothed[t] / N;
    cs += smoothed[t];
    confirmed_norm[t] = cs / N;
    const double e = (zerolag[t] - refined[t]) / N;
    R_v[t] = 0.1 * (e * e);
  }
  /* :200-201 */
  double acc = 0.0;
  int cnt = 0;
  for (int t = 0; t < T && cnt < n_first; ++t)
    if (smoothed[t] > 0.0) { acc += smoothed[t]; ++cnt; }
  const double mean = cnt ? acc / (double)cnt : NAN; /* mean([]) = NaN */
  *I0 = mmax(min_cases, mean);
  return 0;
}

/* ------------------------------------------------------------------------- */
/* Non-negative regression between the EKF rounds               */
/* Tools/TrainPredictPrescribeNPI.m:264-278 (and :326-339): lsqnonneg +       */
/* alternating intercept                                                       */
/* ------------------------------------------------------------------------- */
#define NNLS_PMAX 12
/* Least squares on the passive set through the normal equations G_PP z_P = c_P (Cholesky in
 * ascending index order; a non-positive pivot drops that variable: z = 0).  MATLAB solves
 * C(:,P)\d by QR; the normal equations are this oracle's DEFINITION. */
static void nnls_solve_passive(const double G[NNLS_PMAX][NNLS_PMAX], const double *c, const int *P, int p, double *z) {
  int idx[NNLS_PMAX], m = 0;
  double Lc[NNLS_PMAX][NNLS_PMAX], yv[NNLS_PMAX];
  int dead[NNLS_PMAX];
  for (int j = 0; j < p; ++j) { z[j] = 0.0; if (P[j]) idx[m++] = j; }
  for (int i = 0; i < m; ++i) {
    for (int j = 0; j <= i; ++j) {
      double acc = G[idx[i]][idx[j]];
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
  for (int i = 0; i < m; ++i) { /* L y = c */
    double acc = c[idx[i]];
    for (int k = 0; k < i; ++k) acc = fma(-Lc[i][k], yv[k], acc);
    yv[i] = dead[i] ? 0.0 : acc / Lc[i][i];
  }
  for (int i = m - 1; i >= 0; --i) { /* L' z = y */
    double acc = yv[i];
    for (int k = i + 1; k < m; ++k) acc = fma(-Lc[k][i], z[idx[k]], acc);
    z[idx[i]] = dead[i] ? 0.0 : acc / Lc[i][i];
  }
}
/* lsqnonneg (Lawson & Hanson 1974, ch. 23, as MATLAB's lsqnonneg.m implements it) on the moments
 * G = C'C and c = C'd:  w = C'(d - Cx) = c - Gx. */
static void nnls_moments(const double G[NNLS_PMAX][NNLS_PMAX], const double *c, double tol, int p, double *x) {
  int P[NNLS_PMAX], Z[NNLS_PMAX];
  double w[NNLS_PMAX], z[NNLS_PMAX];
  for (int j = 0; j < p; ++j) { P[j] = 0; Z[j] = 1; x[j] = 0.0; w[j] = c[j]; }
  const int itmax = 3 * p;
  int iter = 0;
  for (;;) {
    int anyZ = 0, t = -1;
    double best = 0.0;
    for (int j = 0; j < p; ++j)
      if (Z[j]) {
        anyZ = 1;
        if (w[j] > tol && (t < 0 || w[j] > best)) { best = w[j]; t = j; } /* first maximum among w(Z) */
      }
    if (!anyZ || t < 0) break;
    /* t = argmax over ALL of Z (MATLAB: [~,t] = max(wz)); the test above is any(w(Z) > tol) */
    t = -1;
    for (int j = 0; j < p; ++j)
      if (Z[j] && (t < 0 || w[j] > best)) { best = w[j]; t = j; }
    P[t] = 1; Z[t] = 0;
    nnls_solve_passive(G, c, P, p, z);
    int stop = 0;
    for (;;) {
      int neg = 0;
      for (int j = 0; j < p; ++j) if (P[j] && z[j] <= 0.0) neg = 1;
      if (!neg) break;
      if (++iter > itmax) { stop = 1; break; } /* lsqnonneg.m: exitflag 0, x = z */
      double alpha = INFINITY;
      for (int j = 0; j < p; ++j)
        if (P[j] && z[j] <= 0.0) { const double a = x[j] / (x[j] - z[j]); if (a < alpha) alpha = a; }
      for (int j = 0; j < p; ++j) x[j] = x[j] + alpha * (z[j] - x[j]);
      for (int j = 0; j < p; ++j) { Z[j] = (fabs(x[j]) < tol && P[j]) || Z[j]; P[j] = !Z[j]; }
      nnls_solve_passive(G, c, P, p, z);
    }
    for (int j = 0; j < p; ++j) x[j] = z[j];
    if (stop) break;
    for (int j = 0; j < p; ++j) { /* w = c - G x */
      double acc = c[j];
      for (int l = 0; l < p; ++l) acc = fma(-G[j][l], x[l], acc);
      w[j] = acc;
    }
  }
}
/* X [n][p] row-major, y [n]  ->  a [p] >= 0, b, number of accepted alternations.
 * :264 a = lsqnonneg(X, y), b = 0; then up to 100 rounds of  a' = lsqnonneg(X, y - b),
 * b' = mean(y - X a) and err' = sum((y - X a - b').^2) -- both with the OLD a, as written --
 * accepted while err' < err (:267-277). */
int orc_nnls_affine(const double *X, const double *y, int n, int p, int max_alt, double *a, double *b_out) {
  double G[NNLS_PMAX][NNLS_PMAX], Xty[NNLS_PMAX], Xt1[NNLS_PMAX], cs[NNLS_PMAX], c[NNLS_PMAX] = {0}, at[NNLS_PMAX];
  for (int j = 0; j < p; ++j) {
    double sy = 0.0, s1 = 0.0, sa = 0.0;
    for (int i = 0; i < n; ++i) { sy = fma(X[i * p + j], y[i], sy); s1 += X[i * p + j]; sa += fabs(X[i * p + j]); }
    Xty[j] = sy; Xt1[j] = s1; cs[j] = sa;
    for (int l = 0; l <= j; ++l) {
      double g = 0.0;
      for (int i = 0; i < n; ++i) g = fma(X[i * p + j], X[i * p + l], g);
      G[j][l] = g; G[l][j] = g;
    }
  }
  double norm1 = 0.0;
  for (int j = 0; j < p; ++j) norm1 = mmax(norm1, cs[j]);
  const double tol = ((10.0 * 2.220446049250313e-16) * norm1) * (double)(n > p ? n : p); /* lsqnonneg.m default TolX */
  double b = 0.0;
  for (int j = 0; j < p; ++j) c[j] = Xty[j];
  nnls_moments(G, c, tol, p, a);
  double min_err = 0.0;
  for (int i = 0; i < n; ++i) {
    double xa = X[i * p] * a[0];
    for (int j = 1; j < p; ++j) xa = fma(X[i * p + j], a[j], xa);
    const double r = y[i] - xa;
    min_err += r * r;
  }
  int k = 0;
  for (; k < max_alt; ++k) {
    for (int j = 0; j < p; ++j) c[j] = Xty[j] - b * Xt1[j];   /* X'(y - b) */
    nnls_moments(G, c, tol, p, at);
    double sm = 0.0;
    for (int i = 0; i < n; ++i) {
      double xa = X[i * p] * a[0];
      for (int j = 1; j < p; ++j) xa = fma(X[i * p + j], a[j], xa);
      sm += y[i] - xa;
    }
    const double b_t = sm / (double)n;
    double err = 0.0;
    for (int i = 0; i < n; ++i) {
      double xa = X[i * p] * a[0];
      for (int j = 1; j < p; ++j) xa = fma(X[i * p + j], a[j], xa);
      const double r = (y[i] - xa) - b_t;
      err += r * r;
    }
    if (err < min_err) {
      for (int j = 0; j < p; ++j) a[j] = at[j];
      b = b_t; min_err = err;
    } else {
      break;
    }
  }
  *b_out = b;
  return k;
}
/* plain lsqnonneg for the tests */
void orc_lsqnonneg(const double *X, const double *y, int n, int p, double *a) {
  double dummy;
  /* max_alt = 0: only the first call of orc_nnls_affine */
  orc_nnls_affine(X, y, n, p, 0, a, &dummy);
}
