// epi_linalg_experiments.cuh -- alternative formulations of the 6x6 Jacobi pinv (same iteration, same bits
// as pinv_sym6 / oracle orc_pinv_sym), built only with -DEPI_PINV_MODE=1..5 (tools/build_variants.py).
// None is the shipped path: each was measured SLOWER than the set-wise rolled form on the sweep workload
// (DESIGN.md 4 has the numbers and the ncu evidence); they are kept so that the measurements can be repeated.
//   1  rolled set loop, every pair behind its own branch (idle pairs skipped), 1-/3-wide angles     6.8 ms
//   2  the same with the sets unrolled on fixed positions                                            9.8 ms
//   3  pair-at-a-time, warp-synchronous, one-hot dispatch on fixed register positions                7.0 ms
//   4  matrix in the thread's shared-memory column, one generic rotation and replay body             7.65 ms
//   5  mode 3 forward + mode 4 replay                                                                 7.3 ms
//   (set-wise rolled form, EPI_PINV_MODE 0: 6.2 ms)
#pragma once

static __device__ __noinline__ JacobiRot3 jacobi_rotation3_call(double app0, double app1, double app2, double aqq0,
                                                                double aqq1, double aqq2, double apq0, double apq1,
                                                                double apq2) {
  const double app[3] = {app0, app1, app2}, aqq[3] = {aqq0, aqq1, aqq2}, apq[3] = {apq0, apq1, apq2};
  return jacobi_rotation3(app, aqq, apq);  // one out-of-line copy for the unrolled-set variant
}

// Per-thread stack of the recorded rotations: the first DS words live in shared memory
// ([word][thread], conflict-free), the rest -- matrices that need unusually many rotations --
// in local memory.  Words are raw 64-bit patterns (c, s, or a sweep's set mask).
template <int M, int NT>
struct RotStackSel {
  static constexpr int NPAIR = M * (M - 1) / 2;
#ifndef EPI_ROT_DS
#define EPI_ROT_DS 48
#endif
  static constexpr int DS = (M == 6) ? EPI_ROT_DS : 24;
  static constexpr int CAP = kJacobiMaxSweep * (2 * NPAIR + 1);
  static constexpr int SMEM_WORDS = DS * NT;
  double *sm;  // this thread's column of the CTA's [DS][NT] buffer
  double loc[CAP - DS];
  int sp;
  // one code path for both homes of a word: the slot address is selected (generic pointer), not the
  // access duplicated -- every push/pop site is a single store/load
  EPI_DI double *slot(int i) { return i < DS ? sm + i * NT : loc + (i - DS); }
  EPI_DI void push(double x) { *slot(sp) = x; ++sp; }
  EPI_DI double pop() { --sp; return *slot(sp); }
  template <int NW>
  EPI_DI void push_n(const double (&x)[NW]) {
#pragma unroll
    for (int i = 0; i < NW; ++i) *slot(sp + i) = x[i];
    sp += NW;
  }
  template <int NW>
  EPI_DI void pop_n(double (&x)[NW]) {  // x[i] = the word pushed as x[i]
    sp -= NW;
#pragma unroll
    for (int i = 0; i < NW; ++i) x[i] = *slot(sp + i);
  }
};

// One rotation angle without branches: the 1-wide form of jacobi_rotation3 (same operation
// sequence per operand, so the same bits); used when a set has a single active pair.
#ifdef EPI_ROT1_INLINE
EPI_DI
#else
static __device__ __noinline__
#endif
JacobiRot jacobi_rotation1(double app, double aqq, double apq) {
  const double d = 0.5 * (aqq - app);
  const double r2v[1] = {fma(d, d, apq * apq)};
  bool ok = jacobi_in_range(r2v[0]);
  double r[1];
  ok = sqrt_n<1>(r2v, r) && ok;
  const double den0 = fabs(d) + r[0];
  const double num[2] = {apq, den0}, den[2] = {den0, r[0] + r[0]};
  double quo[2];
  ok = div_n<2>(num, den, quo) && ok;
  const double carg[1] = {quo[1]};
  double c[1];
  ok = sqrt_n<1>(carg, c) && ok;
  if (!ok) return jacobi_rotation_cold(app, aqq, apq);
  JacobiRot o;
  o.t = (d < 0.0) ? -quo[0] : quo[0];
  o.c = c[0];
  o.s = o.t * o.c;
  return o;
}

EPI_DI JacobiRot jacobi_rotation1_inline(double app, double aqq, double apq) {
  const double d = 0.5 * (aqq - app);
  const double r2v[1] = {fma(d, d, apq * apq)};
  bool ok = jacobi_in_range(r2v[0]);
  double r[1];
  ok = sqrt_n<1>(r2v, r) && ok;
  const double den0 = fabs(d) + r[0];
  const double num[2] = {apq, den0}, den[2] = {den0, r[0] + r[0]};
  double quo[2];
  ok = div_n<2>(num, den, quo) && ok;
  const double carg[1] = {quo[1]};
  double c[1];
  ok = sqrt_n<1>(carg, c) && ok;
  if (!ok) return jacobi_rotation_cold(app, aqq, apq);
  JacobiRot o;
  o.t = (d < 0.0) ? -quo[0] : quo[0];
  o.c = c[0];
  o.s = o.t * o.c;
  return o;
}

// 1 / x for six operands side by side (the eigenvalue reciprocals of the pinv); bit-identical
// to the operator wherever div_n's validity test holds, the operator itself elsewhere.
EPI_DI void recip6(const double (&x)[6], double (&y)[6]) {
  const double one[6] = {1.0, 1.0, 1.0, 1.0, 1.0, 1.0};
  if (!div_n<6>(one, x, y)) {
#pragma unroll 1
    for (int i = 0; i < 6; ++i) y[i] = 1.0 / x[i];
  }
}


// M = 6.  Oracle orc_pinv_sym: threshold Jacobi, pairs in round-robin sets (circle method), a pair is
// rotated iff it is ACTIVE (|a_pq| > thr), idle pairs are skipped.  On the sweep workload the lanes of a
// warp (32 epsilon of one region on one day) agree on which pairs are active almost always (measured on
// the CPU: 7.8 pair rotations per warp-day against 7.7 per lane), and past the first ~100 days a set has
// ONE active pair -- so every pair sits behind its own (in practice warp-uniform) branch instead of
// rotating idle pairs by the identity: 7.8 instead of 12.7 rotations + replays per matrix.  A set with
// two or three active pairs evaluates its angles side by side (jacobi_rotation3), a set with one takes
// the 1-wide form.
//
// Rolled set loop (the unrolled form of 15 forward rotation sites overflows the instruction cache):
// the pairs of a set always sit at POSITIONS (0,1), (2,3), (4,5); after each set positions 1..5 rotate
// (new position i holds old position PI[i]), which returns to the identity after the 5 sets of a sweep
// and generates exactly the oracle's ORC_JSETS6 order and orientation.  The rotation stack holds (c, s)
// of the executed rotations only, and one 15-bit pair mask per sweep.
// One set of three index-disjoint pairs (P[i], Q[i]) of the 6x6 iteration; returns the 3-bit mask of
// the pairs it rotated.  The indices are compile-time constants after inlining/unrolling.
template <class Stack>
EPI_DI unsigned jacobi_set6(Mat<6, true> &a, double thr, Stack &stk, int p0, int q0, int p1, int q1, int p2, int q2) {
  constexpr int M = 6;
  const int P[3] = {p0, p1, p2}, Q[3] = {q0, q1, q2};
  bool act[3];
  double app[3], aqq[3], apq[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    app[i] = a(P[i], P[i]); aqq[i] = a(Q[i], Q[i]); apq[i] = a(P[i], Q[i]);
    act[i] = fabs(apq[i]) > thr;
  }
  const int n_act = (act[0] ? 1 : 0) + (act[1] ? 1 : 0) + (act[2] ? 1 : 0);
  unsigned mask = 0;
  if (n_act) {
    double rt[3], rc[3], rs[3];
    if (n_act > 1) {
      double sapp[3], saqq[3], sapq[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) {  // an idle pair gets a benign triple so that the shared fast path is taken
        sapp[i] = act[i] ? app[i] : 0.0; saqq[i] = act[i] ? aqq[i] : 0.0; sapq[i] = act[i] ? apq[i] : 1.0;
      }
#if EPI_PINV_MODE == 2
      const JacobiRot3 rot = jacobi_rotation3_call(sapp[0], sapp[1], sapp[2], saqq[0], saqq[1], saqq[2], sapq[0], sapq[1], sapq[2]);
#else
      const JacobiRot3 rot = jacobi_rotation3(sapp, saqq, sapq);
#endif
#pragma unroll
      for (int i = 0; i < 3; ++i) { rt[i] = rot.t[i]; rc[i] = rot.c[i]; rs[i] = rot.s[i]; }
    } else {
      const double p1v = act[0] ? app[0] : (act[1] ? app[1] : app[2]);
      const double q1v = act[0] ? aqq[0] : (act[1] ? aqq[1] : aqq[2]);
      const double o1v = act[0] ? apq[0] : (act[1] ? apq[1] : apq[2]);
      const JacobiRot rot = jacobi_rotation1(p1v, q1v, o1v);
#pragma unroll
      for (int i = 0; i < 3; ++i) { rt[i] = rot.t; rc[i] = rot.c; rs[i] = rot.s; }
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      if (act[i]) {
        const int p = P[i], q = Q[i];
        const double t = rt[i], c = rc[i], s = rs[i];
        a.at(p, p) = app[i] - t * apq[i];
        a.at(q, q) = aqq[i] + t * apq[i];
        a.at(p, q) = 0.0;
#pragma unroll
        for (int r = 0; r < M; ++r)
          if (r != p && r != q) {
            const double g = a(r, p), h = a(r, q);
            a.at(r, p) = fma(c, g, -(s * h));
            a.at(r, q) = fma(s, g, c * h);
          }
        const double cs[2] = {c, s};
        stk.template push_n<2>(cs);
        mask |= 1u << i;
      }
    }
  }
  return mask;
}

template <int NT>
EPI_DI int pinv_sym6_perpair(Mat<6, true> &a, Mat<6, true> &X, double *stack_smem) {
  constexpr int M = 6;
  RotStackSel<M, NT> stk;
  stk.sm = stack_smem;
  stk.sp = 0;
  int nsw = 0;

  for (int sweep = 0; sweep < kJacobiMaxSweep; ++sweep) {
    double dmax = 0.0;
#pragma unroll
    for (int p = 0; p < M; ++p) dmax = mmax(dmax, fabs(a(p, p)));
    const double thr = dmax * kJacobiRel;
    if (!any_offdiag_gt<M>(a, thr)) break;
    unsigned mask = 0;
#if EPI_PINV_MODE == 2
    // sets unrolled on the oracle's fixed (p, q) table: no position bookkeeping
    using JSF = JacobiSets<6>;
#pragma unroll
    for (int st = 0; st < 5; ++st)
      mask |= jacobi_set6(a, thr, stk, JSF::p(st, 0), JSF::q(st, 0), JSF::p(st, 1), JSF::q(st, 1), JSF::p(st, 2),
                          JSF::q(st, 2)) << (3 * st);
#else
#pragma unroll 1
    for (int st = 0; st < 5; ++st) {
      mask |= jacobi_set6(a, thr, stk, 0, 1, 2, 3, 4, 5) << (3 * st);
      {
        constexpr int PI[6] = {0, 3, 1, 5, 2, 4};
        Mat<M, true> b;
#pragma unroll
        for (int i = 0; i < M; ++i)
#pragma unroll
          for (int j = i; j < M; ++j) b.at(i, j) = a(PI[i], PI[j]);
        a = b;
      }
    }
#endif
    stk.push(__longlong_as_double((long long)mask));
    ++nsw;
  }
  double lmax = 0.0;
#pragma unroll
  for (int i = 0; i < M; ++i) lmax = mmax(lmax, fabs(a(i, i)));
  const double tol = (double)M * eps_of(lmax);
  int rank = 0;
#pragma unroll
  for (int i = 0; i < M; ++i)
#pragma unroll
    for (int j = i; j < M; ++j) X.at(i, j) = 0.0;
  {
    double lam[6], inv[6];
#pragma unroll
    for (int i = 0; i < M; ++i) lam[i] = a(i, i);
    recip6(lam, inv);
#pragma unroll
    for (int i = 0; i < M; ++i) {
      const bool keep = fabs(lam[i]) > tol;
      X.at(i, i) = keep ? inv[i] : 0.0;
      rank += keep ? 1 : 0;
    }
  }
  // X <- R_k X R_k', last rotation first
  for (; nsw > 0; --nsw) {
    const unsigned mask = (unsigned)__double_as_longlong(stk.pop());
    // unrolled over the pairs with the oracle's (p,q) table: no position bookkeeping on X
    using JS = JacobiSets<6>;
#pragma unroll
    for (int st = 4; st >= 0; --st) {
      if (((mask >> (3 * st)) & 7u) == 0) continue;
#pragma unroll
      for (int i = 2; i >= 0; --i) {
        if (mask & (1u << (3 * st + i))) {
          const int p = JS::p(st, i), q = JS::q(st, i);
          double cs[2];
          stk.template pop_n<2>(cs);
          const double c = cs[0], s = cs[1];
#pragma unroll
          for (int r = 0; r < M; ++r)
            if (r != p && r != q) {
              const double g = X(r, p), h = X(r, q);
              X.at(r, p) = fma(c, g, s * h);
              X.at(r, q) = fma(c, h, -(s * g));
            }
          const double xpp = X(p, p), xpq = X(p, q), xqq = X(q, q);
          const double u1 = fma(c, xpp, s * xpq), u2 = fma(c, xpq, s * xqq);
          const double w1 = fma(c, xpq, -(s * xpp)), w2 = fma(c, xqq, -(s * xpq));
          X.at(p, p) = fma(c, u1, s * u2);
          X.at(p, q) = fma(c, u2, -(s * u1));
          X.at(q, q) = fma(c, w2, -(s * w1));
        }
      }
    }
  }
  return rank;
}

// ---- M = 6, pair-at-a-time warp-synchronous form (EPI_PINV_MODE 3) --------------------------------
// The same iteration as pinv_sym6 (oracle orc_pinv_sym: pairs visited in the circle-method order, a pair
// rotated iff |a_pq| > thr when it is visited), organised around what the sweep workload looks like: the 32
// lanes of a warp (32 epsilon of one region and day) agree on the active pairs almost always and only ~8 of
// the 15 pairs of a sweep are active.  The warp walks the UNION of its lanes' active pairs: a 15-bit activity
// mask per lane, the next pair = lowest set bit of the warp-wide OR, ONE inline copy of the angle code, and two
// warp-uniform 15-way switches (gather app/aqq/apq; apply the rotation on the fixed register positions).  Idle
// pairs cost nothing and the matrix never moves between registers (the rolled set loop spent a quarter of the
// kernel's instructions on position permutations and idle-pair selects).  Lanes whose pair is idle are
// predicated off, so every lane performs exactly its own oracle sequence.  The rotation stack pointer is
// warp-uniform: every lane pushes (c, s) of every rotation of the warp plus its own 15-bit "rotated" mask per
// sweep, and the replay walks the same union backwards.
#define EPI_FOR_PAIRS15(X_) X_(0) X_(1) X_(2) X_(3) X_(4) X_(5) X_(6) X_(7) X_(8) X_(9) X_(10) X_(11) X_(12) X_(13) X_(14)

// position of the unordered pair {i, j} in the visiting order of a sweep
__host__ __device__ constexpr int jacobi_pair_bit6(int i, int j) {
  for (int n = 0; n < 15; ++n) {
    const int p = JacobiSets<6>::p(n / 3, n % 3), q = JacobiSets<6>::q(n / 3, n % 3);
    if ((p == i && q == j) || (p == j && q == i)) return n;
  }
  return -1;
}

EPI_DI unsigned jacobi_active_mask6(const Mat<6, true> &a, double thr) {
  using JS = JacobiSets<6>;
  unsigned m = 0;
#pragma unroll
  for (int n = 0; n < 15; ++n)
    if (fabs(a(JS::p(n / 3, n % 3), JS::q(n / 3, n % 3))) > thr) m |= 1u << n;
  return m;
}

// entry numbers / visiting-order positions per pair, for the generic (shared-memory) rotation bodies
struct JacobiTab6 {
  // per pair n of the visiting order: byte offsets (entry * stride) are formed at run time from entry numbers
  unsigned char pp[15], qq[15], pq[15];
  unsigned char rp[15][4], rq[15][4];   // entries (r, p), (r, q) for the four r != p, q (ascending r)
  unsigned char bp[15][4], bq[15][4];   // position of the pairs {r, p}, {r, q} in the visiting order
};
__host__ __device__ constexpr JacobiTab6 make_jacobi_tab6() {
  JacobiTab6 t{};
  for (int n = 0; n < 15; ++n) {
    const int p = JacobiSets<6>::p(n / 3, n % 3), q = JacobiSets<6>::q(n / 3, n % 3);
    t.pp[n] = (unsigned char)Mat<6, true>::idx(p, p);
    t.qq[n] = (unsigned char)Mat<6, true>::idx(q, q);
    t.pq[n] = (unsigned char)Mat<6, true>::idx(p, q);
    int k = 0;
    for (int r = 0; r < 6; ++r)
      if (r != p && r != q) {
        t.rp[n][k] = (unsigned char)Mat<6, true>::idx(r, p);
        t.rq[n][k] = (unsigned char)Mat<6, true>::idx(r, q);
        t.bp[n][k] = (unsigned char)jacobi_pair_bit6(r, p);
        t.bq[n][k] = (unsigned char)jacobi_pair_bit6(r, q);
        ++k;
      }
  }
  return t;
}
static __constant__ JacobiTab6 c_jtab6 = make_jacobi_tab6();

// rows r != p, q of one rotation, and the activity bits of the eight entries it changes
template <int P, int Q, int R>
EPI_DI void jacobi_rotate_row6(Mat<6, true> &a, double c, double s, double thr, unsigned &m) {
  if constexpr (R != P && R != Q) {
    const double g = a(R, P), h = a(R, Q);
    const double gn = fma(c, g, -(s * h)), hn = fma(s, g, c * h);
    a.at(R, P) = gn;
    a.at(R, Q) = hn;
    constexpr int b1 = jacobi_pair_bit6(R, P), b2 = jacobi_pair_bit6(R, Q);
    m = (fabs(gn) > thr) ? (m | (1u << b1)) : (m & ~(1u << b1));
    m = (fabs(hn) > thr) ? (m | (1u << b2)) : (m & ~(1u << b2));
  }
}
template <int P, int Q, int... R>
EPI_DI void jacobi_rotate_rows6(Mat<6, true> &a, double c, double s, double thr, unsigned &m, std::integer_sequence<int, R...>) {
  (jacobi_rotate_row6<P, Q, R>(a, c, s, thr, m), ...);
}

template <int NT>
struct RotStackU {  // warp-uniform stack pointer; [word][thread] shared-memory words, local-memory overflow
#ifndef EPI_ROT_DSU
#define EPI_ROT_DSU 27
#endif
  static constexpr int DS = EPI_ROT_DSU;
  static constexpr int CAP = kJacobiMaxSweep * (2 * 15 + 1);
  static constexpr int SMEM_WORDS = DS * NT;
};

// keeps the compiler from correlating two dispatches on the same pair index (jump threading would fuse
// them into one 15-way region whose exits carry a copy of the whole matrix each)
EPI_DI unsigned opaque_u32(unsigned x) {
  asm volatile("" : "+r"(x));
  return x;
}

template <int NT, bool SMEM_REPLAY>
EPI_DI int pinv_sym6_pairs(Mat<6, true> &a, Mat<6, true> &X, double *stack_smem_, unsigned wmask, bool skip) {
  constexpr int M = 6;
  using JS = JacobiSets<6>;
  using ST = RotStackU<NT>;
  // SMEM_REPLAY: the first 21 words of the column hold X during the replay, the stack follows
  double *xcol = stack_smem_;
  double *stack_smem = stack_smem_ + (SMEM_REPLAY ? 21 * NT : 0);
  double loc[ST::CAP - ST::DS];
  int sp = 0;  // warp-uniform
  auto push = [&](double x) {
    if (sp < ST::DS) stack_smem[sp * NT] = x; else loc[sp - ST::DS] = x;
    ++sp;
  };
  auto pop = [&]() -> double {
    --sp;
    return sp < ST::DS ? stack_smem[sp * NT] : loc[sp - ST::DS];
  };
  auto push2 = [&](double x, double y) {  // one (warp-uniform) capacity test per rotation
    if (sp + 2 <= ST::DS) { stack_smem[sp * NT] = x; stack_smem[(sp + 1) * NT] = y; sp += 2; }
    else { push(x); push(y); }
  };
  auto pop2 = [&](double &x, double &y) {
    if (sp <= ST::DS) { sp -= 2; x = stack_smem[sp * NT]; y = stack_smem[(sp + 1) * NT]; }
    else { y = pop(); x = pop(); }
  };
  int nsw = 0;
  bool done = skip;
  for (int sweep = 0; sweep < kJacobiMaxSweep; ++sweep) {
    double dmax = 0.0;
#pragma unroll
    for (int p = 0; p < M; ++p) dmax = mmax(dmax, fabs(a(p, p)));
    const double thr = dmax * kJacobiRel;
    unsigned m = done ? 0u : jacobi_active_mask6(a, thr);
    done = (m == 0u);  // the oracle leaves its sweep loop for good
    unsigned rem = __reduce_or_sync(wmask, m);
    if (rem == 0u) break;
    unsigned rec = 0;
    while (rem != 0u) {
      const int n = __ffs(rem) - 1;
      // one-hot dispatch on the fixed register positions (nested bit tests: 5 set groups x 3 pairs); a lane
      // whose pair is idle has no bit set and skips the rotation.  Each site reads its pair straight from the
      // matrix registers, calls the one out-of-line copy of the angle code, rotates in place and re-tests the
      // eight entries the rotation changed (the activity mask of the pairs still to visit).
      const unsigned hot = opaque_u32(((m >> n) & 1u) << n);
      double c = 1.0, s = 0.0;
#define EPI_ROTATE(N_) if (hot & (1u << (N_))) { constexpr int p = JS::p((N_) / 3, (N_) % 3), q = JS::q((N_) / 3, (N_) % 3); \
          asm volatile(""); const double app = a(p, p), aqq = a(q, q), apq = a(p, q); \
          const JacobiRot rot = jacobi_rotation1(app, aqq, apq); \
          const double t = rot.t; c = rot.c; s = rot.s; \
          a.at(p, p) = app - t * apq; a.at(q, q) = aqq + t * apq; a.at(p, q) = 0.0; \
          jacobi_rotate_rows6<p, q>(a, c, s, thr, m, std::make_integer_sequence<int, 6>{}); \
          m &= ~(1u << (N_)); }
#define EPI_ROTATE3(S_) if (hot & (7u << (3 * (S_)))) { EPI_ROTATE(3 * (S_)) EPI_ROTATE(3 * (S_) + 1) EPI_ROTATE(3 * (S_) + 2) }
      EPI_ROTATE3(0) EPI_ROTATE3(1) EPI_ROTATE3(2) EPI_ROTATE3(3) EPI_ROTATE3(4)
#undef EPI_ROTATE3
#undef EPI_ROTATE
      rec |= hot;
      push2(c, s);
      rem = __reduce_or_sync(wmask, m) & (0xfffffffeu << n);
    }
    push(__longlong_as_double((long long)rec));
    ++nsw;
  }
  double lmax = 0.0;
#pragma unroll
  for (int i = 0; i < M; ++i) lmax = mmax(lmax, fabs(a(i, i)));
  const double tol = (double)M * eps_of(lmax);
  int rank = 0;
#pragma unroll
  for (int i = 0; i < M; ++i)
#pragma unroll
    for (int j = i; j < M; ++j) X.at(i, j) = 0.0;
  {
    double lam[6], inv[6];
#pragma unroll
    for (int i = 0; i < M; ++i) lam[i] = a(i, i);
    recip6(lam, inv);
#pragma unroll
    for (int i = 0; i < M; ++i) {
      const bool keep = fabs(lam[i]) > tol;
      X.at(i, i) = keep ? inv[i] : 0.0;
      rank += keep ? 1 : 0;
    }
  }
  // X <- R_k X R_k', last rotation first
  if constexpr (SMEM_REPLAY) {
    // one generic replay body on the shared-memory column (entry offsets of the warp-uniform pair from the
    // constant table): 15 register-position replay sites would be 9 KB of hot code
    auto XS = [&](int e) -> double & { return xcol[e * NT]; };
#pragma unroll
    for (int e = 0; e < 21; ++e) XS(e) = X.v[e];
    for (; nsw > 0; --nsw) {
      const unsigned rec = (unsigned)__double_as_longlong(pop());
      unsigned rem = __reduce_or_sync(wmask, rec);
      while (rem != 0u) {
        const int n = 31 - __clz(rem);  // warp-uniform
        rem &= ~(1u << n);
        double c, s;
        pop2(c, s);
        const bool act = (rec >> n) & 1u;
        const int ipp = c_jtab6.pp[n], iqq = c_jtab6.qq[n], ipq = c_jtab6.pq[n];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int irp = c_jtab6.rp[n][k], irq = c_jtab6.rq[n][k];
          const double g = XS(irp), h = XS(irq);
          const double gn = fma(c, g, s * h), hn = fma(c, h, -(s * g));
          if (act) {
            XS(irp) = gn;
            XS(irq) = hn;
          }
        }
        const double xpp = XS(ipp), xpq = XS(ipq), xqq = XS(iqq);
        const double u1 = fma(c, xpp, s * xpq), u2 = fma(c, xpq, s * xqq);
        const double w1 = fma(c, xpq, -(s * xpp)), w2 = fma(c, xqq, -(s * xpq));
        if (act) {
          XS(ipp) = fma(c, u1, s * u2);
          XS(ipq) = fma(c, u2, -(s * u1));
          XS(iqq) = fma(c, w2, -(s * w1));
        }
      }
    }
#pragma unroll
    for (int e = 0; e < 21; ++e) X.v[e] = XS(e);
  } else {
  for (; nsw > 0; --nsw) {
    const unsigned rec = (unsigned)__double_as_longlong(pop());
    unsigned rem = __reduce_or_sync(wmask, rec);
    while (rem != 0u) {
      const int n = 31 - __clz(rem);
      rem &= ~(1u << n);
      double c, s;
      pop2(c, s);
      const unsigned hot = opaque_u32(rec & (1u << n));
#define EPI_REPLAY(N_) if (hot & (1u << (N_))) { constexpr int p = JS::p((N_) / 3, (N_) % 3), q = JS::q((N_) / 3, (N_) % 3); \
          _Pragma("unroll") for (int r = 0; r < M; ++r) if (r != p && r != q) { \
            const double g = X(r, p), h = X(r, q); \
            X.at(r, p) = fma(c, g, s * h); X.at(r, q) = fma(c, h, -(s * g)); } \
          const double xpp = X(p, p), xpq = X(p, q), xqq = X(q, q); \
          const double u1 = fma(c, xpp, s * xpq), u2 = fma(c, xpq, s * xqq); \
          const double w1 = fma(c, xpq, -(s * xpp)), w2 = fma(c, xqq, -(s * xpq)); \
          X.at(p, p) = fma(c, u1, s * u2); X.at(p, q) = fma(c, u2, -(s * u1)); X.at(q, q) = fma(c, w2, -(s * w1)); }
#define EPI_REPLAY3(S_) if (hot & (7u << (3 * (S_)))) { EPI_REPLAY(3 * (S_)) EPI_REPLAY(3 * (S_) + 1) EPI_REPLAY(3 * (S_) + 2) }
      EPI_REPLAY3(0) EPI_REPLAY3(1) EPI_REPLAY3(2) EPI_REPLAY3(3) EPI_REPLAY3(4)
#undef EPI_REPLAY3
#undef EPI_REPLAY
    }
  }
  }
  return rank;
}

// ---- M = 6, shared-memory-resident form (EPI_PINV_MODE 4) -------------------------------------------
// ncu on the register-resident forms: the kernel is bound by INSTRUCTION FETCH (gcc instruction requests at
// 91 % of peak, SM i-cache hit rate 93 %): 15 forward and 15 replay rotation sites on fixed register
// positions are ~25 KB of hot code for 16 warps in different phases.  Here the matrix lives in the thread's
// shared-memory column ([entry][thread], conflict-free for any per-lane entry) and ONE rotation body and ONE
// replay body serve all 15 pairs: the pair walked by the warp (lowest set bit of the OR of its lanes'
// activity masks) is warp-uniform, so the entry offsets come from a constant table through the uniform
// datapath and the hot loop is a few hundred instructions.  Same iteration, same per-lane operation sequence
// as pinv_sym6_pairs / oracle orc_pinv_sym.

template <int NT>
struct SmemPinv6 {
#ifndef EPI_ROT_DS4
#define EPI_ROT_DS4 27
#endif
  static constexpr int NM = 21;           // packed symmetric matrix
  static constexpr int DS = EPI_ROT_DS4;  // rotation-stack words in shared memory
  static constexpr int CAP = kJacobiMaxSweep * (2 * 15 + 1);
  static constexpr int SMEM_WORDS = (NM + DS) * NT;
};

// col: this thread's column of the CTA's [21 + DS][NT] buffer; on entry the packed matrix is in its first
// 21 words, on return they hold the packed pinv.
template <int NT>
EPI_DI int pinv_sym6_smem(double *col, unsigned wmask, bool skip) {
  constexpr int M = 6;
  using ST = SmemPinv6<NT>;
  double *stack_smem = col + ST::NM * NT;
  double loc[ST::CAP - ST::DS];
  int sp = 0;  // warp-uniform
  auto A = [&](int e) -> double & { return col[e * NT]; };
  auto push = [&](double x) {
    if (sp < ST::DS) stack_smem[sp * NT] = x; else loc[sp - ST::DS] = x;
    ++sp;
  };
  auto pop = [&]() -> double {
    --sp;
    return sp < ST::DS ? stack_smem[sp * NT] : loc[sp - ST::DS];
  };
  auto push2 = [&](double x, double y) {  // one (warp-uniform) capacity test per rotation
    if (sp + 2 <= ST::DS) { stack_smem[sp * NT] = x; stack_smem[(sp + 1) * NT] = y; sp += 2; }
    else { push(x); push(y); }
  };
  auto pop2 = [&](double &x, double &y) {
    if (sp <= ST::DS) { sp -= 2; x = stack_smem[sp * NT]; y = stack_smem[(sp + 1) * NT]; }
    else { y = pop(); x = pop(); }
  };
  int nsw = 0;
  bool done = skip;
  for (int sweep = 0; sweep < kJacobiMaxSweep; ++sweep) {
    double dmax = 0.0;
#pragma unroll
    for (int p = 0; p < M; ++p) dmax = mmax(dmax, fabs(A(Mat<6, true>::idx(p, p))));
    const double thr = dmax * kJacobiRel;
    unsigned m = 0;
#pragma unroll
    for (int n = 0; n < 15; ++n)
      if (fabs(A(Mat<6, true>::idx(JacobiSets<6>::p(n / 3, n % 3), JacobiSets<6>::q(n / 3, n % 3)))) > thr) m |= 1u << n;
    if (done) m = 0u;
    done = (m == 0u);  // the oracle leaves its sweep loop for good
    unsigned rem = __reduce_or_sync(wmask, m);
    if (rem == 0u) break;
    unsigned rec = 0;
    while (rem != 0u) {
      const int n = __ffs(rem) - 1;  // warp-uniform
      const bool act = (m >> n) & 1u;
      const int ipp = c_jtab6.pp[n], iqq = c_jtab6.qq[n], ipq = c_jtab6.pq[n];
      const double app = A(ipp), aqq = A(iqq), apq = A(ipq);
      // a lane whose pair is idle gets a benign triple so that the warp stays on the branch-free fast path
      const JacobiRot rot = jacobi_rotation1_inline(act ? app : 0.0, act ? aqq : 0.0, act ? apq : 1.0);
      const double t = rot.t, c = rot.c, s = rot.s;
      if (act) {
        A(ipp) = app - t * apq;
        A(iqq) = aqq + t * apq;
        A(ipq) = 0.0;
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int irp = c_jtab6.rp[n][k], irq = c_jtab6.rq[n][k];
        const unsigned mp = 1u << c_jtab6.bp[n][k], mq = 1u << c_jtab6.bq[n][k];
        const double g = A(irp), h = A(irq);
        const double gn = fma(c, g, -(s * h)), hn = fma(s, g, c * h);
        if (act) {
          A(irp) = gn;
          A(irq) = hn;
          m = (fabs(gn) > thr) ? (m | mp) : (m & ~mp);
          m = (fabs(hn) > thr) ? (m | mq) : (m & ~mq);
        }
      }
      if (act) rec |= 1u << n;
      m &= ~(1u << n);
      push2(c, s);
      rem = __reduce_or_sync(wmask, m) & (0xfffffffeu << n);
    }
    push(__longlong_as_double((long long)rec));
    ++nsw;
  }
  int rank = 0;
  {
    double lam[6], inv[6];
    double lmax = 0.0;
#pragma unroll
    for (int i = 0; i < M; ++i) {
      lam[i] = A(Mat<6, true>::idx(i, i));
      lmax = mmax(lmax, fabs(lam[i]));
    }
    const double tol = (double)M * eps_of(lmax);
    recip6(lam, inv);
#pragma unroll
    for (int i = 0; i < M; ++i)
#pragma unroll
      for (int j = i + 1; j < M; ++j) A(Mat<6, true>::idx(i, j)) = 0.0;
#pragma unroll
    for (int i = 0; i < M; ++i) {
      const bool keep = fabs(lam[i]) > tol;
      A(Mat<6, true>::idx(i, i)) = keep ? inv[i] : 0.0;
      rank += keep ? 1 : 0;
    }
  }
  // X <- R_k X R_k', last rotation first
  for (; nsw > 0; --nsw) {
    const unsigned rec = (unsigned)__double_as_longlong(pop());
    unsigned rem = __reduce_or_sync(wmask, rec);
    while (rem != 0u) {
      const int n = 31 - __clz(rem);  // warp-uniform
      rem &= ~(1u << n);
      double c, s;
      pop2(c, s);
      const bool act = (rec >> n) & 1u;
      const int ipp = c_jtab6.pp[n], iqq = c_jtab6.qq[n], ipq = c_jtab6.pq[n];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int irp = c_jtab6.rp[n][k], irq = c_jtab6.rq[n][k];
        const double g = A(irp), h = A(irq);
        const double gn = fma(c, g, s * h), hn = fma(c, h, -(s * g));
        if (act) {
          A(irp) = gn;
          A(irq) = hn;
        }
      }
      const double xpp = A(ipp), xpq = A(ipq), xqq = A(iqq);
      const double u1 = fma(c, xpp, s * xpq), u2 = fma(c, xpq, s * xqq);
      const double w1 = fma(c, xpq, -(s * xpp)), w2 = fma(c, xqq, -(s * xpq));
      if (act) {
        A(ipp) = fma(c, u1, s * u2);
        A(ipq) = fma(c, u2, -(s * u1));
        A(iqq) = fma(c, w2, -(s * w1));
      }
    }
  }
  return rank;
}

