"""NumPy/LAPACK twin of the CPU oracle (test infrastructure, NOT product code).

An independently written, readable restatement of the same reference files the
C oracle (epi_oracle.c) follows, using numpy matrix products (BLAS order),
numpy.linalg.pinv (LAPACK SVD, MATLAB's default tolerance) and
numpy.linalg.solve.  Purpose:
  * cross-check the C oracle where the problem is well conditioned (SEIRP,
    rollouts, forward EKF, the 3-state smoother) -- agreement ~1e-10;
  * quantify how much the ill-conditioned 6-state smoother depends on the pinv
    algorithm: oracle(Jacobi pinv) vs twin(SVD pinv);
  * stand in for the interpreter-regime baseline.
Each function cites the reference file:line it restates.  Only tests/ and
bench.py's cpu_baseline leg may import it.
"""
import numpy as np

EPS = np.finfo(np.float64).eps


def mmax(a, b):  # MATLAB max: the non-NaN operand wins
    return b if (b > a or a != a) else a


def mmin(a, b):
    return b if (b < a or a != a) else a


# --- Tools/SEIRP.m:13-32 -------------------------------------------------------
def seirp(alpha_e, alpha_i, kappa, rho, beta, mu, gamma, s0, e0, i0, r0, p0, T, dt):
    K = int(np.floor(T / dt + 0.5))
    vec = lambda v: np.broadcast_to(np.asarray(v, dtype=float).ravel(), (K,)) if np.size(v) == 1 \
        else np.asarray(v, dtype=float).ravel()
    ae, ai, ka, ro, be, m_, ga = (vec(v) for v in (alpha_e, alpha_i, kappa, rho, beta, mu, gamma))
    s, e, i, r, p = (np.zeros(K) for _ in range(5))
    s[0], e[0], i[0], r[0], p[0] = s0, e0, i0, r0, p0
    for t in range(K - 1):
        s[t + 1] = (-ae[t] * s[t] * e[t] - ai[t] * s[t] * i[t] + ga[t] * r[t]) * dt + s[t]
        e[t + 1] = (ae[t] * s[t] * e[t] + ai[t] * s[t] * i[t] - ka[t] * e[t] - ro[t] * e[t]) * dt + e[t]
        i[t + 1] = (ka[t] * e[t] - be[t] * i[t] - m_[t] * i[t]) * dt + i[t]
        r[t + 1] = (be[t] * i[t] + ro[t] * e[t] - ga[t] * r[t]) * dt + r[t]
        p[t + 1] = (m_[t] * i[t]) * dt + p[t]
    return s, e, i, r, p


# --- Tools/SEIRPSaturatedResource.m:13-36 ------------------------------------------
def seirp_saturated(alpha_e, alpha_i, kappa, rho, gamma, s0, e0, i0, r0, p0, T, dt, beta_0, beta_s,
                    mu_0, mu_s, sigma, i_0):
    K = int(np.floor(T / dt + 0.5))
    vec = lambda v: np.broadcast_to(np.asarray(v, dtype=float).ravel(), (K,)) if np.size(v) == 1 \
        else np.asarray(v, dtype=float).ravel()
    ae, ai, ka, ro, ga = (vec(v) for v in (alpha_e, alpha_i, kappa, rho, gamma))
    s, e, i, r, p = (np.zeros(K) for _ in range(5))
    s[0], e[0], i[0], r[0], p[0] = s0, e0, i0, r0, p0
    for t in range(K - 1):
        h = (np.tanh((i[t] - i_0) / sigma) + 1) / 2
        be = (beta_s - beta_0) * h + beta_0
        m_ = (mu_s - mu_0) * h + mu_0
        s[t + 1] = (-ae[t] * s[t] * e[t] - ai[t] * s[t] * i[t] + ga[t] * r[t]) * dt + s[t]
        e[t + 1] = (ae[t] * s[t] * e[t] + ai[t] * s[t] * i[t] - ka[t] * e[t] - ro[t] * e[t]) * dt + e[t]
        i[t + 1] = (ka[t] * e[t] - be * i[t] - m_ * i[t]) * dt + i[t]
        r[t + 1] = (be * i[t] + ro[t] * e[t] - ga[t] * r[t]) * dt + r[t]
        p[t + 1] = (m_ * i[t]) * dt + p[t]
    return s, e, i, r, p


# --- Tools/SIalpha_Controlled.m:15-32 ------------------------------------------------
def sialpha_controlled(u, s0, i0, alpha0, u_max, alpha_min, alpha_max, gamma, a, b, beta, s_std,
                       i_std, a_std, K, dt, noise=None):
    u = np.asarray(u, dtype=float)
    a = np.asarray(a, dtype=float).ravel()
    u_max = np.asarray(u_max, dtype=float).ravel()
    s, i, al = np.zeros(K + 1), np.zeros(K + 1), np.zeros(K + 1)
    s[0], i[0], al[0] = s0, i0, alpha0
    nz = np.zeros((3, K)) if noise is None else np.asarray(noise, dtype=float)
    for t in range(K):
        s[t + 1] = min(max(s[t] - dt * (al[t] * s[t] * i[t] + nz[0, t] * s_std), 0), 1)
        i[t + 1] = min(max(i[t] + dt * (al[t] * s[t] * i[t] - beta * i[t] + nz[1, t] * i_std), 0), 1)
        al[t + 1] = min(max(al[t] + dt * (-gamma * al[t] + gamma * b + gamma * a @ (u_max - u[:, t])
                                         + nz[2, t] * a_std), alpha_min), alpha_max)
    return s[1:], i[1:], al[1:]


# --- Tools/SI_Controlled.m:12-22 -------------------------------------------------------
def si_controlled(alpha, beta, s0, i0, K, dt):
    alpha = np.asarray(alpha, dtype=float).ravel()
    s, i = np.zeros(K), np.zeros(K)
    s[0], i[0] = s0, i0
    for t in range(K - 1):
        s[t + 1] = min(max(s[t] - dt * alpha[t] * s[t] * i[t], 0), 1)
        i[t + 1] = min(max(i[t] + dt * (alpha[t] * s[t] * i[t] - beta * i[t]), 0), 1)
    return s, i


# --- Tools/NPICost.m:6-10 -------------------------------------------------------------------
def npicost(newcases, inputs, weights):
    J0 = np.mean(np.asarray(newcases, dtype=float))
    wi = np.asarray(weights, dtype=float) * np.asarray(inputs, dtype=float)
    return float(J0), float(np.mean(wi.ravel()))


# --- Tools/TrainPredictPrescribeNPI.m:624-633 --------------------------------------------------
def pareto(J0, J1):
    J0, J1 = np.asarray(J0, dtype=float), np.asarray(J1, dtype=float)
    n = J0.size
    mask = np.zeros(n, dtype=bool)
    for k in range(n):
        mask[k] = np.sum((J0 < J0[k]) & (J1 < J1[k])) == 0
    v = (J0 / np.nanmax(J0)) ** 2 + (J1 / np.nanmax(J1)) ** 2
    return mask, int(np.nanargmin(v))


# --- model callbacks ---------------------------------------------------------------------------
class Model:
    """kind: 'sialpha' | 'optctrl' | 'legacy_tools' | 'legacy_codegen'; flipped = Backward wrappers."""

    def __init__(self, kind, params, flipped=False):
        self.kind, self.p, self.flip = kind, params, flipped
        self.m = 3 if kind == "sialpha" else 6
        g = lambda k: np.asarray(params[k], dtype=float).ravel()
        self.a, self.u_max = g("a"), g("u_max")
        if self.m == 6:
            self.u_min = g("u_min")
            w = np.asarray(params["w"], dtype=float)
            self.w = w[:, 0] if (w.ndim == 2 and w.shape[1] > 1) else w.ravel()  # phi(kk) linear index
            if self.w.size == 1:
                self.w = np.full(self.a.size, self.w[0])

    def margins(self, s):
        p = self.p
        lo_s, lo_i = (p["s_min"], p["i_min"]) if (self.kind == "sialpha" and not self.flip) else (0.0, 0.0)
        s = s.copy()
        s[0] = mmin(1.0, mmax(lo_s, s[0]))
        s[1] = mmin(1.0, mmax(lo_i, s[1]))
        s[2] = mmin(p["alpha_max"], mmax(p["alpha_min"], s[2]))
        return s

    def obs(self, s, v_bar):
        ot = "NEWCASES" if self.kind == "legacy_codegen" else self.p.get("obs_type", "NEWCASES")
        C = np.zeros((1, self.m))
        if ot == "NEWCASES":
            C[0, :3] = [s[1] * s[2], s[0] * s[2], s[0] * s[1]]
            xh = s[0] * s[1] * s[2] + v_bar
        elif ot == "TOTALCASES":
            C[0, 0] = -1.0
            xh = 1 - s[0] + v_bar
        else:
            raise ValueError("unknown observation type")
        if self.kind != "legacy_codegen":
            xh = mmax(0.0, xh)
        return C, xh

    def phi(self, s):
        return self.p["epsilon"] * self.w - self.p["gamma"] * s[5] * self.a

    def update(self, u, s):
        p, sg = self.p, (-1.0 if self.flip else 1.0)
        dt, beta, gamma = p["dt"], p["beta"], p["gamma"]
        u = np.array(u, dtype=float)
        if self.m == 6:
            phi = self.phi(s)
            for k in range(u.size):
                if np.isnan(u[k]):
                    to_min = (phi[k] >= 0) if self.kind.startswith("legacy") else (phi[k] > 0)
                    u[k] = self.u_min[k] if to_min else self.u_max[k]
        lo_s, lo_i = (p["s_min"], p["i_min"]) if (self.kind == "sialpha" and not self.flip) else (0.0, 0.0)
        sn = np.zeros(self.m)
        sn[0] = mmax(lo_s, mmin(1.0, s[0] - sg * dt * s[2] * s[0] * s[1]))
        sn[1] = mmax(lo_i, mmin(1.0, s[1] + sg * dt * (s[2] * s[0] * s[1] - beta * s[1])))
        sn[2] = mmax(p["alpha_min"], mmin(p["alpha_max"], s[2] + sg * dt * (
            -gamma * s[2] + gamma * p["b"] + gamma * self.a @ (self.u_max - u))))
        if self.m == 6:
            rho = s[3] - s[4] - (1 - p["epsilon"])
            sn[3] = s[3] + sg * dt * rho * s[2] * s[1]
            sn[4] = s[4] + sg * dt * (rho * s[2] * s[0] + beta * s[4])
            sn[5] = s[5] + sg * dt * (rho * s[0] * s[1] + gamma * s[5])
        return u, sn

    def jac(self, u, s):
        p, sg = self.p, (-1.0 if self.flip else 1.0)
        dt, beta, gamma = p["dt"], p["beta"], p["gamma"]
        A = np.zeros((self.m, self.m))
        A[0, 0] = 1 - sg * dt * s[2] * s[1]
        A[0, 1] = -sg * dt * s[2] * s[0]
        A[0, 2] = -sg * dt * s[0] * s[1]
        A[1, 0] = sg * dt * s[1] * s[2]
        A[1, 1] = 1 + sg * dt * (s[0] * s[2] - beta)
        A[1, 2] = sg * dt * s[0] * s[1]
        A[2, 2] = 1 - sg * dt * gamma
        if self.m == 6:
            phi = self.phi(s)
            for k in range(len(u)):
                if np.isnan(u[k]) and -1.0 / p["sigma"] < phi[k] < 1.0 / p["sigma"]:
                    A[2, 5] -= sg * gamma * dt * (p["sigma"] / 2) * self.a[k] * (self.u_max[k] - self.u_min[k])
            rho = s[3] - s[4] - (1 - p["epsilon"])
            A[3, 1] = sg * dt * s[2] * rho
            A[3, 2] = sg * dt * s[1] * rho
            A[3, 3] = 1 + sg * dt * s[1] * s[2]
            A[3, 4] = -sg * dt * s[1] * s[2]
            A[4, 0] = sg * dt * s[2] * rho
            A[4, 2] = sg * dt * s[0] * rho
            A[4, 3] = sg * dt * s[0] * s[2]
            A[4, 4] = 1 - sg * dt * (s[0] * s[2] - beta)
            A[5, 0] = sg * dt * s[1] * rho
            A[5, 1] = sg * dt * s[0] * rho
            A[5, 3] = sg * dt * s[0] * s[1]
            A[5, 4] = -sg * dt * s[0] * s[1]
            A[5, 5] = 1 + sg * dt * gamma
        return A


def _expand_QR(Q_w, R_v, T, m):
    """GenericExtendedKalmanFilter.m:64-91."""
    Q = np.asarray(Q_w, dtype=float)
    Q2 = np.atleast_2d(Q)
    if Q2.shape[0] == Q2.shape[1]:
        Qp = [Q[:, :, k] for k in range(T)] if Q.ndim == 3 else \
            [Q2 if Q2.shape[0] == m else np.eye(m) * Q2[0, 0]] * T
    elif min(Q2.shape) == 1 and Q.size == T:
        Qp = [np.eye(m) * q for q in Q.ravel()]
    else:
        raise ValueError("Process noise covariance noise mismatch")
    R = np.asarray(R_v, dtype=float)
    R2 = np.atleast_2d(R)
    if R2.shape[0] == R2.shape[1]:
        Rp, fixed = (R.ravel().copy() if R.ndim == 3 else np.full(T, R2[0, 0])), True
    elif min(R2.shape) == 1 and R.size == T:
        Rp, fixed = R.ravel().copy(), False
    else:
        raise ValueError("Observation noise covariance noise mismatch")
    return Qp, Rp, fixed


def generic_ekf(model, u, x, s_init, Ps_init, s_final, Ps_final, v_bar, Q_w, R_v, beta, gamma, W,
                order=1, pinv=None):
    """Tools/GenericExtendedKalmanFilter.m:41-233 (no time flip; see ekf_eks)."""
    if order not in (1, 2):
        raise ValueError("Undefined order")
    pinv = pinv or np.linalg.pinv
    u = np.asarray(u, dtype=float)
    x = np.asarray(x, dtype=float).ravel()
    L, T = u.shape
    m = model.m
    Qp, R, fixed_R = _expand_QR(Q_w, R_v, T, m)
    S_M, S_P = np.zeros((m, T)), np.zeros((m, T))
    P_M, P_P = np.zeros((m, m, T)), np.zeros((m, m, T))
    Kg, innov, rho = np.zeros((m, 1, T)), np.zeros((1, T)), np.zeros((T, 1))
    u_opt, u_opt_s = np.zeros((L, T)), np.zeros((L, T))
    imean, icov, icovn = np.zeros(W), np.zeros(W), np.zeros(W)
    sk, Pk = np.asarray(s_init, dtype=float).ravel().copy(), np.asarray(Ps_init, dtype=float).copy()
    I = np.eye(m)
    for k in range(T):
        S_M[:, k], P_M[:, :, k] = sk, Pk
        C, xh = model.obs(sk, v_bar)
        if not np.isnan(x[k]):
            innov[0, k] = x[k] - xh
            Kgain = Pk @ C.T / (C @ Pk @ C.T + gamma * R[k])
            Pp = ((I - Kgain @ C) @ Pk @ (I - Kgain @ C).T + Kgain * R[k] @ Kgain.T) / gamma
            sp = sk + (Kgain * innov[0, k]).ravel()
        else:
            Kgain, Pp, sp = np.zeros((m, 1)), Pk.copy(), sk.copy()
        Pp = (Pp + Pp.T) / 2.0
        sp = model.margins(sp)
        u_opt[:, k], sk = model.update(u[:, k], sp)
        A = model.jac(u[:, k], sp)
        Pk = A @ Pp @ A.T + Qp[k]
        Pk = (Pk + Pk.T) / 2.0
        sk = model.margins(sk)
        S_P[:, k], P_P[:, :, k], Kg[:, :, k] = sp, Pp, Kgain
        cnt = min(k + 1, W)
        imean = np.concatenate([[innov[0, k]], imean[:-1]])
        mu = imean.sum() / cnt
        cc = (innov[0, k] - mu) ** 2
        icov = np.concatenate([[cc], icov[:-1]])
        icovn = np.concatenate([[cc / (R[k] + EPS)], icovn[:-1]])
        rho[k, 0] = icovn.sum() / cnt
        if beta != 1 and not np.isnan(x[k]) and fixed_R and k + 1 < T:
            R[k + 1] = beta * R[k] + (1 - beta) * (icov.sum() / cnt)
    S_S, P_S = np.zeros((m, T)), np.zeros((m, m, T))
    S_S[:, -1], P_S[:, :, -1] = S_P[:, -1], P_P[:, :, -1]
    sf, Pf = np.asarray(s_final, dtype=float).ravel(), np.asarray(Ps_final, dtype=float)
    S_S[~np.isnan(sf), -1] = sf[~np.isnan(sf)]
    P_S[:, :, -1][~np.isnan(Pf)] = Pf[~np.isnan(Pf)]
    for k in range(T - 2, -1, -1):
        A = model.jac(u[:, k], S_P[:, k])
        pm = P_M[:, :, k + 1]
        J = np.zeros((m, m)) if not np.all(np.isfinite(pm)) else (P_P[:, :, k] @ A.T) @ pinv(pm)
        S_S[:, k] = model.margins(S_P[:, k] + J @ (S_S[:, k + 1] - S_M[:, k + 1]))
        Pn = P_P[:, :, k] - J @ (pm - P_S[:, :, k + 1]) @ J.T
        P_S[:, :, k] = (Pn + Pn.T) / 2.0
        u_opt_s[:, k], _ = model.update(u[:, k], S_S[:, k])
    return dict(u_opt=u_opt, u_opt_smooth=u_opt_s, S_MINUS=S_M, S_PLUS=S_P, S_SMOOTH=S_S,
                P_MINUS=P_M, P_PLUS=P_P, P_SMOOTH=P_S, K_GAIN=Kg, innovations=innov, rho=rho)


def legacy_ekf(model, u, x, s_init, Ps_init, s_final, Ps_final, v_bar, Q_w, R_v, beta, gamma, W):
    """Tools/NewCaseEKFEstimatorWithOptimalNPI.m:9-143."""
    u = np.asarray(u, dtype=float)
    x = np.asarray(x, dtype=float).ravel()
    L, T = u.shape
    m = 6
    Q = np.asarray(Q_w, dtype=float)
    Q = np.eye(m) * Q.ravel()[0] if Q.size == 1 else Q
    R = float(np.asarray(R_v, dtype=float).ravel()[0])
    S_M, S_P = np.zeros((m, T)), np.zeros((m, T))
    P_M, P_P = np.zeros((m, m, T)), np.zeros((m, m, T))
    Kg, innov, rho = np.zeros((m, 1, T)), np.zeros((1, T)), np.zeros((T, 1))
    u_opt = np.zeros((L, T))
    imean, icov, icovn = np.zeros(W), np.zeros(W), np.zeros(W)
    sk, Pk = np.asarray(s_init, dtype=float).ravel().copy(), np.asarray(Ps_init, dtype=float).copy()
    I = np.eye(m)
    for k in range(T):
        S_M[:, k], P_M[:, :, k] = sk, Pk
        C, xh = model.obs(sk, v_bar)
        if not np.isnan(x[k]):
            innov[0, k] = x[k] - xh
            Kgain = Pk @ C.T / (C @ Pk @ C.T + gamma * R)
            Pp = (I - Kgain @ C) @ Pk / gamma
            sp = sk + (Kgain * innov[0, k]).ravel()
        else:
            Kgain, Pp, sp = np.zeros((m, 1)), Pk.copy(), sk.copy()
        sp = model.margins(sp)
        u_opt[:, k], sk = model.update(u[:, k], sp)
        A = model.jac(u[:, k], sp)
        Pk = A @ Pp @ A.T + Q
        sk = model.margins(sk)
        S_P[:, k], P_P[:, :, k], Kg[:, :, k] = sp, Pp, Kgain
        cnt = min(k + 1, W)
        imean = np.concatenate([[innov[0, k]], imean[:-1]])
        mu = imean.sum() / cnt
        cc = (innov[0, k] - mu) ** 2
        icov = np.concatenate([[cc], icov[:-1]])
        icovn = np.concatenate([[cc / R], icovn[:-1]])
        rho[k, 0] = icovn.sum() / cnt
        if beta != 1 and not np.isnan(x[k]):
            R = beta * R + (1 - beta) * icov.sum() / cnt
    S_S, P_S = np.zeros((m, T)), np.zeros((m, m, T))
    S_S[:, -1], P_S[:, :, -1] = S_P[:, -1], P_P[:, :, -1]
    sf, Pf = np.asarray(s_final, dtype=float).ravel(), np.asarray(Ps_final, dtype=float)
    S_S[~np.isnan(sf), -1] = sf[~np.isnan(sf)]
    rows, cols = np.nonzero(~np.isnan(Pf))
    if rows.size:
        P_S[np.ix_(np.unique(rows), np.unique(cols), [T - 1])] = \
            Pf[np.ix_(np.unique(rows), np.unique(cols))][:, :, None]
    for k in range(T - 2, -1, -1):
        A = model.jac(u[:, k], S_P[:, k])
        pm = P_M[:, :, k + 1]
        J = np.linalg.solve(pm.T, (P_P[:, :, k] @ A.T).T).T  # mrdivide
        S_S[:, k] = model.margins(S_P[:, k] + J @ (S_S[:, k + 1] - S_M[:, k + 1]))
        P_S[:, :, k] = P_P[:, :, k] - J @ (pm - P_S[:, :, k + 1]) @ J.T
    return dict(u_opt=u_opt, S_MINUS=S_M, S_PLUS=S_P, S_SMOOTH=S_S, P_MINUS=P_M, P_PLUS=P_P,
                P_SMOOTH=P_S, K_GAIN=Kg, innovations=innov, rho=rho)


def ekf_eks(kind, flipped, u, x, params, s_init, Ps_init, s_final, Ps_final, v_bar, Q_w, R_v, beta,
            gamma, W, order=1, pinv=None):
    """The Tools wrappers: SIAlphaModelEKF / ...BackwardEKF / ...EKFOptControlled /
    ...BackwardEKFOptControlled (flip per SIAlphaModelBackwardEKF.m:19-40) and the legacy monolith."""
    model = Model(kind, params, flipped)
    if kind.startswith("legacy"):
        return legacy_ekf(model, u, x, s_init, Ps_init, s_final, Ps_final, v_bar, Q_w, R_v, beta, gamma, W)
    if not flipped:
        return generic_ekf(model, u, x, s_init, Ps_init, s_final, Ps_final, v_bar, Q_w, R_v, beta,
                           gamma, W, order, pinv)
    u = np.asarray(u, dtype=float)
    x = np.asarray(x, dtype=float).ravel()
    r = generic_ekf(model, u[:, ::-1], x[::-1], s_final, Ps_final, s_init, Ps_init, v_bar, Q_w, R_v,
                    beta, gamma, W, order, pinv)
    out = {}
    for k, v in r.items():
        out[k] = v if k == "rho" else v[..., ::-1].copy()  # rho is returned un-flipped (:40)
    return out
