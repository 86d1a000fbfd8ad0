"""Synthetic workloads shaped like the reference's bundled xprize-sample-data.

The big real input (OxCGRT_latest.csv) is absent from the reference checkout, so
every benchmark/parity workload is synthetic, built the way
testScripts/testPrescribeXPRIZE01.m:95-118 builds its scenario (a noiseless
SIalpha rollout as "history") with the EKF setup of
Tools/TrainPredictPrescribeNPI.m:200-247,423-458.  Per-region (N, a, b) and NPI
cost weights come from the committed fixture data/regions_nonnegls.npz
(extracted from the reference's .mat/.csv by tools/make_region_fixture.py).

Pure numpy input construction -- no oracle, no GPU code.  SURVEY.md section 8(d)
"Configs restated as concrete synthetic inputs" is the specification.
"""
import os

import numpy as np

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "regions_nonnegls.npz")

L_NPI = 12
DT = 1.0
GAMMA = 1.0 / 7.0                      # TrainPredictPrescribeNPI.m:16,216
BETA = -np.log(0.01) / 21.0            # :18-19,221
ALPHA0 = BETA + np.log(2.5) / DT       # :20,222
I0 = 10.0
MIN_CASES = 1.0                        # :14


def load_regions(n_regions=236):
    """235 fixture rows (+ duplicates of the first rows up to n_regions)."""
    d = np.load(_DATA)
    idx = np.arange(n_regions) % d["N"].size
    return dict(names=d["names"][idx], N=d["N"][idx], a=d["a2"][idx], b=d["b2"][idx],
                cost_weights=d["cost_weights"][idx], npi_max=d["npi_max"].copy())


def epsilon_grid_xprize02(n=250):
    """testScripts/testPrescribeXPRIZE02.m:52-53 (incl. the logspace(-12,-eps) quirk)."""
    eps = np.finfo(np.float64).eps
    h = n // 2
    return np.concatenate([np.logspace(-12.0, -eps, h), np.linspace(eps, 1.0 - eps, n - h)])


def _rollout_clean(u, s0, i0, alpha0, u_max, alpha_min, alpha_max, a, b):
    """Noise-free SIalpha rollout used only to synthesise observations."""
    K = u.shape[1]
    s, i, al = np.zeros(K), np.zeros(K), np.zeros(K)
    S, I, A = s0, i0, alpha0
    for t in range(K):
        asi = A * S * I
        Sn = max(0.0, min(1.0, S - DT * asi))
        In = max(0.0, min(1.0, I + DT * (asi - BETA * I)))
        An = max(alpha_min, min(alpha_max, A + DT * (-GAMMA * A + GAMMA * b
                                                     + GAMMA * float(a @ (u_max - u[:, t])))))
        S, I, A = Sn, In, An
        s[t], i[t], al[t] = S, I, A
    return s, i, al


def region_history(reg, r, T_hist, seed_u=2, seed_x=3, noise_rel=0.05):
    """History NPIs (piecewise-constant integer levels changing every 30 days)
    and a noisy new-case series for region r."""
    u_max = reg["npi_max"]
    rng_u = np.random.default_rng(seed_u + 7919 * r)
    n_seg = (T_hist + 29) // 30
    lev = np.stack([rng_u.integers(0, int(u_max[j]) + 1, size=n_seg) for j in range(L_NPI)])
    u_hist = np.repeat(lev, 30, axis=1)[:, :T_hist].astype(np.float64)
    N = reg["N"][r]
    s0, i0 = (N - I0) / N, I0 / N
    s, i, al = _rollout_clean(u_hist, s0, i0, ALPHA0, u_max, 1e-8, 100.0, reg["a"][r], reg["b"][r])
    clean = s * i * al
    rng_x = np.random.default_rng(seed_x + r)
    x = np.maximum(0.0, clean * (1.0 + noise_rel * rng_x.standard_normal(T_hist)))
    R_v = 0.1 * (noise_rel * clean) ** 2 + 1e-30
    return u_hist, x, R_v, clean


def ekf3_setup(reg, r):
    """3-state EKF setup of TrainPredictPrescribeNPI.m:200-239."""
    N = reg["N"][r]
    params = dict(dt=DT, w=np.nan, a=reg["a"][r], b=reg["b"][r], u_min=np.zeros(L_NPI),
                  u_max=reg["npi_max"], s_min=MIN_CASES / N, i_min=MIN_CASES / N,
                  alpha_min=1e-8, alpha_max=100.0, epsilon=np.nan, gamma=GAMMA,
                  obs_type="NEWCASES", beta=BETA, sigma=1000000.0)
    s_std, i_std, a_std = 10.0 * I0 / N, 30.0 * I0 / N, 1e-2
    Q_w = DT ** 2 * np.diag(np.array([s_std, i_std, a_std]) ** 2)
    s_init = np.array([(N - I0) / N, I0 / N, ALPHA0])
    Ps_init = DT ** 2 * np.diag(np.array([10 * s_std, 10 * i_std, 10 * a_std]) ** 2)
    return dict(params=params, Q_w=Q_w, s_init=s_init, Ps_init=Ps_init,
                s_final=np.full(3, np.nan), Ps_final=np.full((3, 3), np.nan),
                w_bar=np.zeros(3), v_bar=0.0, beta_ekf=1.0, gamma_ekf=0.995, W=21, order=1,
                noise_std=(s_std, i_std, a_std))


def ekf6_setup(reg, r, setup3):
    """6-state sweep setup of TrainPredictPrescribeNPI.m:423-457."""
    q_lambda = 0.0001
    params = dict(setup3["params"])
    params["w"] = reg["cost_weights"][r]
    s_init = np.concatenate([setup3["s_init"], np.zeros(3)])
    Q = np.zeros((6, 6)); Q[:3, :3] = setup3["Q_w"]; Q[3:, 3:] = DT ** 2 * np.eye(3) * q_lambda ** 2
    P0 = np.zeros((6, 6)); P0[:3, :3] = setup3["Ps_init"]
    P0[3:, 3:] = 10.0 * DT ** 2 * np.eye(3) * q_lambda ** 2
    s_final = np.array([np.nan, np.nan, np.nan, 0.0, 0.0, 0.0])
    Ps_final = np.zeros((6, 6)); Ps_final[:3, :3] = np.nan
    Ps_final[3, 3] = Ps_final[4, 4] = Ps_final[5, 5] = 1e-8
    return dict(params=params, Q_w=Q, s_init=s_init, Ps_init=P0, s_final=s_final,
                Ps_final=Ps_final, w_bar=np.zeros(6), v_bar=0.0, beta_ekf=1.0, gamma_ekf=0.995,
                W=21, order=1)


def sweep_inputs(n_regions=236, T_hist=441, T_fore=120, seed_u=2, seed_x=3):
    """Per-region inputs of the optimal-NPI Pareto sweep (BASELINE config 4).
    Returns a list of dicts; the 3-state 'fixed input' EKF/EKS that supplies the
    rollout start state and the historic new-case estimate
    (TrainPredictPrescribeNPI.m:373-392) is run by the caller (GPU path in the
    product, oracle in the tests)."""
    reg = load_regions(n_regions)
    out = []
    T = T_hist + T_fore
    for r in range(n_regions):
        u_hist, x_h, R_h, _ = region_history(reg, r, T_hist, seed_u, seed_x)
        s3 = ekf3_setup(reg, r)
        s6 = ekf6_setup(reg, r, s3)
        x = np.concatenate([x_h, np.full(T_fore, np.nan)])            # :364
        R_v = np.concatenate([R_h, np.full(T_fore, R_h.mean())])      # :360
        u_fixed = np.concatenate([u_hist[:, :-1],
                                  np.repeat(u_hist[:, -1:], T_fore + 1, axis=1)], axis=1)  # :375-376
        w = reg["cost_weights"][r]
        weights = np.repeat(w[:, None], T, axis=1)                    # :390
        out.append(dict(region=r, name=str(reg["names"][r]), T=T, T_hist=T_hist, u_hist=u_hist,
                        u_fixed=u_fixed, x=x, R_v=R_v, setup3=s3, setup6=s6, weights=weights))
    return out


def seirp_ensemble(B=1_000_000, seed=1234):
    """BASELINE config 2: per-trajectory constant rates."""
    rng = np.random.default_rng(seed)
    lo = np.array([.3, .001, .02, .04, .05, .005, 0.0])
    hi = np.array([.9, .01, .1, .12, .2, .04, .003])
    rates = lo[:, None] + (hi - lo)[:, None] * rng.random((7, B))  # alpha_e, alpha_i, kappa, rho, beta, mu, gamma
    e0 = 1e-6
    ic = np.zeros((5, B)); ic[0] = 1.0 - e0; ic[1] = e0
    return rates, ic


def random_schedules(n, T_fore, npi_max, rng):
    """Random NPI schedules of TrainPredictPrescribeNPI.m:500-510: the first half
    constant in time, the second half redrawn per day.  Returns [n, L, T_fore]."""
    L = len(npi_max)
    u = np.zeros((n, L, T_fore))
    for sc in range(n):
        for j in range(L):
            if sc + 1 < n / 2:
                u[sc, j, :] = rng.integers(0, int(npi_max[j]) + 1)
            else:
                u[sc, j, :] = rng.integers(0, int(npi_max[j]) + 1, size=T_fore)
    return u
