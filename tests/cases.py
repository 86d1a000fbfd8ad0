"""Seeded parity cases shared by the CPU tests, the GPU tests and the golden-vector
generator (tools/make_golden.py).  Parameter sets are the reference's own:
testScripts/testSEIRP01.m:18-74 (scenarios A-E), testSEIRP02.m:22-41,
testSEIRP03.m:21-35, and the EKF set-ups of Tools/TrainPredictPrescribeNPI.m:200-247,
423-458 on synthetic regions built from the reference's trained parameters
(epidemicmodeling_b200/synthetic.py)."""
import numpy as np

from epidemicmodeling_b200 import synthetic as syn

N_POP = 84.0e6


def seirp_scenarios(short=True):
    """name -> kwargs of SEIRP(...).  `short` trims the 4000-day scenarios to 400 days."""
    dt = 0.1
    out = {}

    def K_of(T):
        return int(np.floor(T / dt + 0.5))

    def mk(T, ae, ai, ka, ro, be, mu, ga):
        K = K_of(T)
        full = lambda v: v * np.ones(K) if np.isscalar(v) else v
        return dict(alpha_e=full(ae), alpha_i=full(ai), kappa=full(ka), rho=full(ro), beta=full(be),
                    mu=full(mu), gamma=full(ga), s0=1 - 1 / N_POP, e0=1 / N_POP, i0=0.0, r0=0.0,
                    p0=0.0, T=T, dt=dt)

    TL = 400 if short else 4000
    out["A"] = mk(50, 0.65, 0.005, 0.05, 0.08, 0.1, 0.02, 0.0)
    out["B"] = mk(TL, 0.65, 0.005, 0.05, 0.08, 0.1, 0.02, 0.001)
    Kc = K_of(120)
    out["C"] = mk(120, 0.65 * np.linspace(1, 0.01, Kc), 0.005 * np.linspace(1, 0.01, Kc), 0.05, 0.08,
                  0.1, 0.02, 0.001)
    out["D"] = mk(TL, 0.65, 0.005, 0.005, 0.08, 0.1, 0.02, 0.001)
    out["E"] = mk(TL, 0.65, 0.005, 0.05, 0.08, 0.1, 0.02, 1 / 365)
    # testSEIRP02.m parameter set 2 (quarantine window)
    K2 = K_of(150)
    qs, qe = int(np.floor(30 / dt + 0.5)), int(np.floor(90 / dt + 0.5))
    ae = np.concatenate([0.6 * np.ones(qs), 0.1 * np.ones(qe - qs), 0.4 * np.ones(K2 - qe)])
    ai = np.concatenate([0.005 * np.ones(qs), 0.001 * np.ones(qe - qs), 0.001 * np.ones(K2 - qe)])
    out["Q"] = mk(150, ae, ai, 0.05, 0.08, 0.1, 0.02, 0.001)
    # BASELINE config 1 "~1 year daily"
    out["Y"] = dict(mk(50, 0.65, 0.005, 0.05, 0.08, 0.1, 0.02, 0.0), T=365, dt=1.0)
    for k in ("alpha_e", "alpha_i", "kappa", "rho", "beta", "mu", "gamma"):
        out["Y"][k] = out["Y"][k][:1] * np.ones(365)
    return out


def seirp_saturated_case():
    """testScripts/testSEIRP03.m:21-35."""
    dt, T = 0.1, 150
    K = int(np.floor(T / dt + 0.5))
    one = np.ones(K)
    return dict(alpha_e=0.6 * one, alpha_i=0.005 * one, kappa=0.05 * one, rho=0.08 * one,
                gamma=0.001 * one, s0=(N_POP - 1) / N_POP, e0=1 / N_POP, i0=0.0, r0=0.0, p0=0.0, T=T,
                dt=dt, beta_0=0.1, beta_s=0.01, mu_0=0.02, mu_s=0.2, sigma=1.0, i_0=0.1)


def ekf3_case(region=0, T_hist=90, T_fore=30, variant="perday"):
    """3-state EKF/EKS call of TrainPredictPrescribeNPI.m:377 (fixed-input round).
    variant: 'perday' (R_v 1xT, beta_ekf=1) | 'adaptive' (scalar R, beta_ekf=0.9,
    TrainNPIPrescriptor.m:163-190) | 'totalcases' | 'endpoint' (finite s_final)."""
    inp = syn.sweep_inputs(n_regions=region + 1, T_hist=T_hist, T_fore=T_fore)[region]
    s3 = inp["setup3"]
    args = dict(u=inp["u_fixed"], x=inp["x"], params=dict(s3["params"]), s_init=s3["s_init"],
                Ps_init=s3["Ps_init"], s_final=s3["s_final"], Ps_final=s3["Ps_final"],
                w_bar=s3["w_bar"], v_bar=s3["v_bar"], Q_w=s3["Q_w"], R_v=inp["R_v"],
                beta=s3["beta_ekf"], gamma=s3["gamma_ekf"], inv_monitor_len=s3["W"], order=1)
    if variant == "adaptive":
        args["R_v"] = np.array([[float(np.mean(inp["R_v"]))]])
        args["beta"] = 0.9
    elif variant == "totalcases":
        args["params"]["obs_type"] = "TOTALCASES"
        N = 1.0 / s3["params"]["s_min"]
        xs = np.cumsum(np.nan_to_num(inp["x"]))
        xs[T_hist:] = np.nan
        args["x"] = xs
        args["R_v"] = np.array([[(3.0 / N) ** 2]])
    elif variant == "endpoint":
        args["s_final"] = np.array([0.9, np.nan, 0.05])
        Pf = np.full((3, 3), np.nan)
        Pf[0, 0], Pf[2, 2] = 1e-6, 1e-5
        args["Ps_final"] = Pf
    elif variant == "backward":
        # the Backward wrappers start the (time-reversed) filter from s_final / Ps_final
        # (SIAlphaModelBackwardEKF.m:21-24), so both must be finite
        args["s_final"] = np.array([0.999, 2e-4, 0.12])
        args["Ps_final"] = np.diag(np.array([1e-3, 1e-4, 1e-2]) ** 2)
        args["s_init"] = np.array([np.nan, np.nan, syn.ALPHA0])
        Pi = np.full((3, 3), np.nan)
        Pi[2, 2] = 1e-4
        args["Ps_init"] = Pi
    return args


def ekf6_case(region=0, T_hist=60, T_fore=30, epsilon=0.3, backward=False):
    """6-state call of TrainPredictPrescribeNPI.m:460 (backward=True: the commented
    SIAlphaModelBackwardEKFOptControlled experiment of :466-472 with finite end conditions)."""
    inp = syn.sweep_inputs(n_regions=region + 1, T_hist=T_hist, T_fore=T_fore)[region]
    s6 = inp["setup6"]
    prm = dict(s6["params"])
    prm["epsilon"] = epsilon
    u = np.concatenate([inp["u_hist"], np.full((syn.L_NPI, T_fore), np.nan)], axis=1)
    c = dict(u=u, x=inp["x"], params=prm, s_init=s6["s_init"], Ps_init=s6["Ps_init"],
             s_final=s6["s_final"], Ps_final=s6["Ps_final"], w_bar=s6["w_bar"], v_bar=0.0,
             Q_w=s6["Q_w"], R_v=inp["R_v"], beta=1.0, gamma=0.995, inv_monitor_len=21, order=1)
    if backward:
        c["s_final"] = np.array([0.999, 2e-4, 0.12, 0.0, 0.0, 0.0])
        c["Ps_final"] = np.diag(np.array([1e-3, 1e-4, 1e-2, 1e-4, 1e-4, 1e-4]) ** 2)
    return c


def legacy_case(region=0, T_hist=60, T_fore=20):
    """Legacy monolith call shaped like Tools/PrescribeNPI.m:126-153,287
    (scalar R = 1e-6 scaled to the data, beta = 0.9, s_init(4:6) = 1)."""
    c = ekf6_case(region, T_hist, T_fore, epsilon=0.2)
    c["Q_w"] = np.diag(np.array([0.01, 0.01, 0.1, 10, 10, 10]) ** 2) * 1e-8
    c["R_v"] = np.array([[float(np.nanmean(c["x"]) ** 2 * 1e-2 + 1e-30)]])
    c["beta"] = 0.9
    c["s_init"] = np.concatenate([c["s_init"][:3], np.ones(3)])
    c["s_final"] = np.full(6, np.nan)
    c["Ps_final"] = np.full((6, 6), np.nan)
    return c


def rollout_case(region=0, K=60, seed=11, noisy=True):
    reg = syn.load_regions(region + 1)
    rng = np.random.default_rng(seed)
    u = syn.random_schedules(2, K, reg["npi_max"], rng)[1]  # per-day random schedule, L x K
    N = reg["N"][region]
    return dict(u=u, s0=(N - 10.0) / N, i0=10.0 / N, alpha0=syn.ALPHA0, u_max=reg["npi_max"],
                alpha_min=1e-8, alpha_max=100.0, gamma=syn.GAMMA, a=reg["a"][region],
                b=reg["b"][region], beta=syn.BETA,
                s_noise_std=100.0 / N if noisy else 0.0, i_noise_std=300.0 / N if noisy else 0.0,
                alpha_noise_std=1e-2 if noisy else 0.0, K=K, dt=1.0,
                noise=rng.standard_normal((3, K)) if noisy else None)


def sweep_case(n_regions=3, n_eps=12, T_hist=60, T_fore=30):
    """Inputs of the fused sweep for a few regions (the 3-state fixed-input
    smoother that feeds it is run by the caller)."""
    inp = syn.sweep_inputs(n_regions=n_regions, T_hist=T_hist, T_fore=T_fore)
    eps = syn.epsilon_grid_xprize02(250)[:: max(1, 250 // n_eps)][:n_eps].copy()
    return inp, eps


def rt_expfit_case(T=220, seed=4, order=1, forecast_days=14):
    """Rt_ExpFitEKF call of testScripts/test04FullFeatureExtMLpipeline.m:198-219 on a synthetic
    smoothed new-case series (exponential growth/decay with a slowly varying exponent)."""
    rng = np.random.default_rng(seed)
    t = np.arange(T)
    lam = 0.06 * np.sin(t / 35.0) + 0.01
    x = 800.0 * np.exp(np.cumsum(lam)) * (1.0 + 0.04 * rng.standard_normal(T))
    x[T - forecast_days:] = np.nan                               # :201 masked forecast horizon
    x[40:43] = np.nan                                            # a reporting gap
    Q_w = np.diag([250.0 ** 2, 3.0e-3 ** 2])
    return dict(x=x.reshape(1, T), s_init=np.array([x[0], lam[0]]), params=np.array([1.0, 0.9, 0.1]),
                w_bar=np.zeros(2), v_bar=0.0, Ps_init=100.0 * Q_w, Q_w=Q_w, R_v=10.0 ** 2, beta=0.9,
                gamma=0.995, inv_monitor_len=21, order=order)
