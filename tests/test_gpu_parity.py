"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the
same seeded inputs and against the committed golden vectors.

Bar (BASELINE.md 4 / north_star): FP64 rel <= 1e-9 for SEIRP/EKF states, <= 1e-6 for
converged optimal-control costs and schedules.  Because the kernels and the oracle
execute the same IEEE-754 operation sequence (DESIGN.md "Arithmetic contract"), the
tests assert the much stronger BIT-EXACT agreement wherever no libm call is involved
(everything except SEIRPSaturatedResource's tanh, held to 1e-12).
"""
import os

import numpy as np
import pytest

import cases
from epidemicmodeling_b200 import _capi as K
from epidemicmodeling_b200 import api, synthetic as syn, workloads as wl
from epidemicmodeling_b200.engine import pack_params

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
EKF_KEYS = ("u_opt", "u_opt_smooth", "S_MINUS", "S_PLUS", "S_SMOOTH", "P_MINUS", "P_PLUS", "P_SMOOTH",
            "K_GAIN", "innovations", "rho")
TOL_STATE = 1e-9   # north_star tolerance for SEIRP / EKF states
TOL_COST = 1e-6    # north_star tolerance for optimal-control costs and schedules


def orc():
    from oracle import oracle
    return oracle


def ekf_args(c):
    return (c["u"], c["x"], c["params"], c["s_init"], c["Ps_init"], c["s_final"], c["Ps_final"],
            c["w_bar"], c["v_bar"], c["Q_w"], c["R_v"], c["beta"], c["gamma"], c["inv_monitor_len"],
            c["order"])


def assert_bits(a, b, what=""):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    if not np.array_equal(a, b, equal_nan=True):
        bad = np.flatnonzero(~((a == b) | (np.isnan(a) & np.isnan(b))).ravel())
        k = bad[0]
        raise AssertionError(f"{what}: {bad.size}/{a.size} entries differ; first at flat {k}: "
                             f"gpu={a.ravel()[k]!r} oracle={b.ravel()[k]!r}")


# ------------------------------------------------------------------------------ SEIRP
@pytest.mark.parametrize("name", ["A", "B", "C", "D", "E", "Q", "Y"])
def test_seirp_signature_bit_exact(engine, name):
    kw = cases.seirp_scenarios()[name]
    got = api.SEIRP(**kw)
    want = orc().SEIRP(**kw)
    g = np.load(os.path.join(GOLD, "seirp.npz"))
    for a, b in zip(got, want):
        assert_bits(a, b, f"SEIRP {name}")
        assert np.max(np.abs(a - b) / (np.abs(b) + 1e-300)) <= TOL_STATE
    assert_bits(np.array([o[0, -1] for o in got]), g[f"{name}_last"], f"golden {name}")
    if name in ("A", "Y"):
        assert_bits(np.concatenate(got), g[f"{name}_full"], f"golden full {name}")


def test_seirp_saturated_tolerance(engine):
    kw = cases.seirp_saturated_case()
    got = api.SEIRPSaturatedResource(**kw)
    want = orc().SEIRPSaturatedResource(**kw)
    for a, b in zip(got, want):
        # tanh is not correctly rounded on either side: tolerance parity
        assert np.max(np.abs(a - b) / (np.abs(b) + 1e-30)) < 1e-12
    g = np.load(os.path.join(GOLD, "seirp.npz"))
    assert np.max(np.abs(np.concatenate(got)[:, ::100] - g["SAT_every100"]) / (np.abs(g["SAT_every100"]) + 1e-30)) < 1e-12


@pytest.mark.parametrize("rate_mode", ["const", "series"])
def test_seirp_ensemble_bit_exact(engine, rate_mode):
    B, Kn = 3000, 97  # ragged: not a multiple of the block size
    rates, ic = syn.seirp_ensemble(B, seed=7)
    if rate_mode == "const":
        out = engine.seirp(rates, ic, Kn, 1.0)
        fin = engine.seirp(rates, ic, Kn, 1.0, out_mode=K.SEIRP_OUT_FINAL)
        rser = None
    else:
        rng = np.random.default_rng(8)
        rser = rates[None] * (1.0 + 0.1 * rng.random((Kn, 7, B)))
        out = engine.seirp(rser, ic, Kn, 1.0, rate_mode=K.RATES_SERIES)
        fin = engine.seirp(rser, ic, Kn, 1.0, rate_mode=K.RATES_SERIES, out_mode=K.SEIRP_OUT_FINAL)
    assert_bits(fin, out[:, -1, :], "final-only mode")
    o = orc()
    for b in list(range(0, B, 211)) + [B - 1]:
        r = [rates[f, b] if rser is None else rser[:, f, b] for f in range(7)]
        want = o.SEIRP(*r, *ic[:, b], Kn * 1.0, 1.0)
        for f in range(5):
            assert_bits(out[f, :, b], want[f][0], f"ensemble b={b} f={f}")


def test_seirp_edge_cases(engine):
    rates, ic = syn.seirp_ensemble(5, seed=1)
    out = engine.seirp(rates, ic, 1, 1.0)                   # K = 1: only the initial condition
    assert_bits(out[:, 0, :], ic)
    assert engine.seirp(rates[:, :0], ic[:, :0], 10, 1.0).shape == (5, 10, 0)   # empty batch


# ------------------------------------------------------------------------------ rollout / cost / Pareto
@pytest.mark.parametrize("noisy", [False, True])
def test_sialpha_controlled_signature(engine, noisy):
    rc = cases.rollout_case(noisy=noisy)
    got = api.SIalpha_Controlled(**{("K_" if k == "K" else k): v for k, v in rc.items()})
    want = orc().SIalpha_Controlled(**rc)
    for a, b in zip(got, want):
        assert_bits(a, b, "SIalpha_Controlled")
    if noisy:
        g = np.load(os.path.join(GOLD, "rollout.npz"))
        assert_bits(got[0], g["s"]); assert_bits(got[1], g["i"]); assert_bits(got[2], g["alpha"])


@pytest.mark.parametrize("u_kind,nS", [("f64", 257), ("u8", 257), ("f64", 1024), ("u8", 1024), ("u8", 1000),
                                          ("u8", 1040), ("f64", 1040)])
def test_rollout_cost_batch(engine, u_kind, nS):
    """BASELINE config 5 shape, small: regions x random schedules x 45 days with NPICost fused.
    nS = 1024 takes the TMA-staged kernel (aligned rows), 257 / 1000(u8: rows not 16-byte
    multiples) the plain one; 45 days = ragged last time tile."""
    nR, Kn, L = 3, 45, 12
    reg = syn.load_regions(nR)
    rng = np.random.default_rng(21)
    B = nR * nS
    u = np.stack([syn.random_schedules(nS, Kn, reg["npi_max"], rng) for _ in range(nR)])  # [nR,nS,L,K]
    u_kb = np.ascontiguousarray(np.transpose(u, (3, 2, 0, 1)).reshape(Kn, L, B))
    noise = rng.standard_normal((Kn, 3, B))
    prm_d = [dict(dt=1.0, beta=syn.BETA, gamma=syn.GAMMA, b=reg["b"][r], a=reg["a"][r], u_max=reg["npi_max"],
                  alpha_min=1e-8, alpha_max=100.0) for r in range(nR)]
    x0 = np.array([[(reg["N"][r] - 10) / reg["N"][r], 10 / reg["N"][r], syn.ALPHA0] for r in range(nR)])
    nstd = np.array([[100 / reg["N"][r], 300 / reg["N"][r], 1e-2] for r in range(nR)])
    w = np.stack([np.repeat(reg["cost_weights"][r][None, :], Kn, axis=0) for r in range(nR)])  # [nR,K,L]
    Th = 30
    j0p, j1p = rng.random(nR), rng.random(nR) * 10
    uin = u_kb.astype(np.uint8) if u_kind == "u8" else u_kb
    res = engine.rollout_cost(pack_params(prm_d, L), x0, uin, Kn, L, G=nS, noise_std=nstd, noise=noise,
                              want_traj=True, want_cost=True, T_total=Th + Kn, j0_prefix=j0p,
                              j1_prefix=j1p, w=w)
    o = orc()
    for b in list(range(0, B, 97)) + [B - 1]:
        r = b // nS
        s, i, al = o.SIalpha_Controlled(u_kb[:, :, b].T, *x0[r], reg["npi_max"], 1e-8, 100.0, syn.GAMMA,
                                        reg["a"][r], reg["b"][r], syn.BETA, *nstd[r], Kn, 1.0,
                                        noise=noise[:, :, b].T)
        assert_bits(res["s"][:, b], s[0]); assert_bits(res["i"][:, b], i[0]); assert_bits(res["alpha"][:, b], al[0])
        nc = (s * i) * al
        a0, a1 = j0p[r], j1p[r]
        for t in range(Kn):
            a0 += nc[0, t]
            c = w[r, t, 0] * u_kb[t, 0, b]
            for j in range(1, L):
                c = c + w[r, t, j] * u_kb[t, j, b]
            a1 += c
        assert res["J0"][b] == a0 / (Th + Kn) and res["J1"][b] == a1 / (L * (Th + Kn))


def test_si_controlled_and_npicost_signatures(engine):
    al = 0.2 + 0.1 * np.sin(np.arange(50) / 5.0)
    got = api.SI_Controlled(al, 0.2, 0.999, 0.001, 50, 1.0)
    want = orc().SI_Controlled(al, 0.2, 0.999, 0.001, 50, 1.0)
    assert_bits(got[0], want[0]); assert_bits(got[1], want[1])
    rng = np.random.default_rng(0)
    nc, u, w = rng.random(37), rng.integers(0, 4, (12, 37)).astype(float), rng.random((12, 37))
    assert api.NPICost(nc, u, w) == orc().NPICost(nc, u, w)


def test_pareto_front(engine):
    g = np.load(os.path.join(GOLD, "rollout.npz"))
    mask, iopt = api.ParetoFront(g["pj0"], g["pj1"])
    assert np.array_equal(mask, g["mask"]) and iopt == int(g["iopt"]) + 1
    rng = np.random.default_rng(9)
    J0, J1 = rng.random((17, 1000)), rng.random((17, 1000))
    J0[:, 5] = J0[:, 2]; J1[:, 5] = J1[:, 2]           # ties
    J0[3, 0] = np.nan                                   # NaN never dominates, skipped by the knee
    J0[4] = 1.0; J1[4] = 2.0                            # all points identical: everything survives
    m, io = engine.pareto(J0, J1)
    o = orc()
    for r in range(17):
        wm, wi = o.pareto(J0[r], J1[r])
        assert np.array_equal(m[r].astype(bool), wm) and io[r] == wi, r
    # the sort-based path (n > 8192) against the O(n^2) oracle, with ties, NaN and +-0 keys
    n = 9000
    K0, K1 = np.round(rng.random((3, n)), 3), np.round(rng.random((3, n)), 3)   # many exact ties
    K0[0, :5] = np.nan; K1[0, 5:9] = np.nan; K0[1, 10] = -0.0; K0[1, 11] = 0.0; K0[2, 7] = np.inf
    ms, ios = engine.pareto(K0, K1)
    for r in range(3):
        wm, wi = o.pareto(K0[r], K1[r])
        assert np.array_equal(ms[r].astype(bool), wm) and ios[r] == wi, ("sorted path", r)
    # idempotence at a size the O(n^2) oracle would not like: front(front) == front
    J0b, J1b = rng.random((1, 12000)), rng.random((1, 12000))
    mb, _ = engine.pareto(J0b, J1b)
    keep = mb[0].astype(bool)
    m2, _ = engine.pareto(J0b[:, keep], J1b[:, keep])
    assert m2.all() and keep.sum() < 100


# ------------------------------------------------------------------------------ EKF / EKS (signatures)
_SIGS = {
    "ekf3_perday": (api.SIAlphaModelEKF, lambda: cases.ekf3_case(0, variant="perday"), 0),
    "ekf3_adaptive": (api.SIAlphaModelEKF, lambda: cases.ekf3_case(1, variant="adaptive"), 0),
    "ekf3_totalcases": (api.SIAlphaModelEKF, lambda: cases.ekf3_case(2, variant="totalcases"), 0),
    "ekf3_endpoint": (api.SIAlphaModelEKF, lambda: cases.ekf3_case(3, variant="endpoint"), 0),
    "ekf3_flipped": (api.SIAlphaModelBackwardEKF, lambda: cases.ekf3_case(4, variant="backward"), 1),
    "ekf6_optctrl": (api.SIAlphaModelEKFOptControlled, lambda: cases.ekf6_case(0), 2),
    "ekf6_flipped": (api.SIAlphaModelBackwardEKFOptControlled, lambda: cases.ekf6_case(1, backward=True), 3),
}


@pytest.mark.parametrize("name", sorted(_SIGS))
def test_ekf_signatures_bit_exact(engine, name):
    fn, mk, model = _SIGS[name]
    c = mk()
    got = dict(zip(EKF_KEYS, fn(*ekf_args(c))))
    want = orc().ekf_eks(model, *ekf_args(c))
    g = np.load(os.path.join(GOLD, f"{name}.npz"))
    for k in EKF_KEYS:
        assert_bits(got[k], want[k], f"{name}.{k} vs oracle")
        assert_bits(got[k], g[k], f"{name}.{k} vs golden")


@pytest.mark.parametrize("log2_scale", [-960, 490])
@pytest.mark.parametrize("which", ["ekf3", "ekf6"])
def test_extreme_covariance_scales_take_the_slow_paths(engine, which, log2_scale):
    """Scaling Ps_init, Q_w, R_v (and a finite Ps_final) by 2^k scales every covariance exactly and
    leaves gains and states alone -- but at 2^-960 / 2^490 the squares inside the Jacobi angle and the
    quotients of the covariance update leave the exponent window of the fast paths (branch-free
    3-wide div/sqrt, reciprocal division), so the kernels must take their true-division / textbook
    fall-backs.  Same bits as the oracle there too."""
    f = 2.0 ** log2_scale
    c = cases.ekf3_case(0) if which == "ekf3" else cases.ekf6_case(0)
    c = dict(c, Ps_init=np.asarray(c["Ps_init"]) * f, Q_w=np.asarray(c["Q_w"]) * f, R_v=np.asarray(c["R_v"]) * f,
             Ps_final=np.asarray(c["Ps_final"]) * f)
    fn, model = (api.SIAlphaModelEKF, 0) if which == "ekf3" else (api.SIAlphaModelEKFOptControlled, 2)
    got = dict(zip(EKF_KEYS, fn(*ekf_args(c))))
    want = orc().ekf_eks(model, *ekf_args(c))
    for k in EKF_KEYS:
        assert_bits(got[k], want[k], f"{which} x 2^{log2_scale}: {k}")
    assert np.isfinite(got["S_SMOOTH"]).all() and np.isfinite(got["P_SMOOTH"]).all()


@pytest.mark.parametrize("variant", ["tools", "codegen"])
def test_legacy_estimator_bit_exact(engine, variant):
    c = cases.legacy_case(0 if variant == "tools" else 1)
    got = api.NewCaseEKFEstimatorWithOptimalNPI(*ekf_args(c), variant=variant)
    o = orc()
    want = o.ekf_eks(o.LEGACY_TOOLS if variant == "tools" else o.LEGACY_CODEGEN, *ekf_args(c))
    order = ("u_opt", "S_MINUS", "S_PLUS", "S_SMOOTH", "P_MINUS", "P_PLUS", "P_SMOOTH", "K_GAIN",
             "innovations", "rho") if variant == "tools" else \
            ("u_opt", "S_MINUS", "S_PLUS", "P_MINUS", "P_PLUS", "K_GAIN", "S_SMOOTH", "P_SMOOTH",
             "innovations", "rho")
    g = np.load(os.path.join(GOLD, f"legacy_{variant}.npz"))
    assert len(got) == 10
    for k, a in zip(order, got):
        assert_bits(a, want[k], f"legacy {variant}.{k}")
        assert_bits(a, g[k], f"legacy {variant}.{k} golden")


def test_generic_filter_entry_and_errors(engine):
    c = cases.ekf3_case(0)
    a = ekf_args(c)
    got = api.GenericExtendedKalmanFilter(a[0], a[1], api.HANDLES_SIALPHA, *a[2:])
    want = api.SIAlphaModelEKF(*a)
    for x, y in zip(got, want):
        assert_bits(x, y)
    with pytest.raises(NotImplementedError):
        api.GenericExtendedKalmanFilter(a[0], a[1], object(), *a[2:])
    bad = list(a); bad[-1] = 3
    with pytest.raises(ValueError, match="Undefined order"):
        api.SIAlphaModelEKF(*bad)
    bad = list(a); bad[10] = np.ones(7)
    with pytest.raises(ValueError, match="Observation noise"):
        api.SIAlphaModelEKF(*bad)
    bad = list(a); bad[2] = dict(c["params"], obs_type="DEATHS")
    with pytest.raises(ValueError, match="unknown observation type"):
        api.SIAlphaModelEKF(*bad)
    # the C ABI itself reports the reference's errors as status codes
    prm = pack_params([c["params"]], 12)
    with pytest.raises(K.EpiError) as ei:
        engine.ekf_eks(K.MODEL_SIALPHA, prm, c["u"].T.copy(), c["x"], c["R_v"], np.eye(3).ravel(),
                       c["s_init"], np.eye(3).ravel(), c["s_final"], np.full(9, np.nan), B=1,
                       T=c["u"].shape[1], L=12, r_mode=K.R_PERDAY, order=7)
    assert ei.value.code == K.ERR_ORDER
    prm[0].obs_type = 5
    with pytest.raises(K.EpiError) as ei:
        engine.ekf_eks(K.MODEL_SIALPHA, prm, c["u"].T.copy(), c["x"], c["R_v"], np.eye(3).ravel(),
                       c["s_init"], np.eye(3).ravel(), c["s_final"], np.full(9, np.nan), B=1,
                       T=c["u"].shape[1], L=12, r_mode=K.R_PERDAY)
    assert ei.value.code == K.ERR_OBS_TYPE


# ------------------------------------------------------------------------------ EKF / EKS (batches)
def _ekf3_replicate_batch(nR=4, nRep=33, T_hist=70, T_fore=20, seed=31):
    """BASELINE config 3 shape, small: regions x noise replicates with per-trajectory x."""
    inp = syn.sweep_inputs(n_regions=nR, T_hist=T_hist, T_fore=T_fore)
    b = wl.fixed_input_batch(inp)
    T = b["T"]
    rng = np.random.default_rng(seed)
    B = nR * nRep
    x = np.empty((T, B))
    for r in range(nR):
        clean = np.nan_to_num(inp[r]["x"])
        for k in range(nRep):
            xr = np.maximum(0.0, clean * (1.0 + 0.05 * rng.standard_normal(T)))
            xr[np.isnan(inp[r]["x"])] = np.nan
            xr[rng.integers(0, T_hist, 3)] = np.nan        # a few missing observations inside the history
            x[:, r * nRep + k] = xr
    return inp, b, x, nRep


def test_ekf3_replicate_batch_bit_exact_and_waves(engine):
    inp, b, x, nRep = _ekf3_replicate_batch()
    B, T, L = x.shape[1], b["T"], b["L"]
    kw = dict(B=B, T=T, L=L, G=nRep, x_per_traj=True, r_mode=K.R_PERDAY, fixed_R=False,
              beta=b["beta"], gamma=b["gamma"], W=b["W"], want_status=True)
    out = engine.ekf_eks(K.MODEL_SIALPHA, b["prm"], b["u"], x, b["R"], b["Q"], b["s_init"], b["Ps_init"],
                         b["s_final"], b["Ps_final"], **kw)
    o = orc()
    for bb in list(range(0, B, 17)) + [B - 1]:
        r = inp[bb // nRep]
        s3 = r["setup3"]
        want = o.ekf_eks(o.SIALPHA, r["u_fixed"], x[:, bb], s3["params"], s3["s_init"], s3["Ps_init"],
                         s3["s_final"], s3["Ps_final"], s3["w_bar"], 0.0, s3["Q_w"], r["R_v"], 1.0,
                         s3["gamma_ekf"], s3["W"], 1)
        assert_bits(out["S_SMOOTH"][:, :, bb].T, want["S_SMOOTH"], f"b={bb} S_SMOOTH")
        assert_bits(out["P_SMOOTH"][:, :, bb].reshape(T, 3, 3).transpose(2, 1, 0), want["P_SMOOTH"], f"b={bb} P_SMOOTH")
        assert_bits(out["K_GAIN"][:, :, bb].T, want["K_GAIN"][:, 0, :], f"b={bb} K")
        assert_bits(out["rho"][:, bb], want["rho"][:, 0], f"b={bb} rho")
        assert_bits(out["u_opt_smooth"][:, :, bb].T, want["u_opt_smooth"])
    assert not out["status"].any()                           # full rank, no NaN/Inf guard hit
    # lean call (tape in packed scratch) in small waves must give the same bits
    engine.set_scratch_limit(40 * (T * 8 * 40))
    try:
        lean = engine.ekf_eks(K.MODEL_SIALPHA, b["prm"], b["u"], x, b["R"], b["Q"], b["s_init"],
                              b["Ps_init"], b["s_final"], b["Ps_final"], outputs=("S_SMOOTH", "P_SMOOTH"),
                              **kw)
    finally:
        engine.set_scratch_limit(0)
    assert_bits(lean["S_SMOOTH"], out["S_SMOOTH"], "waves/packed S_SMOOTH")
    assert_bits(lean["P_SMOOTH"], out["P_SMOOTH"], "waves/packed P_SMOOTH")


def test_ekf3_device_memory_mode_matches_host_mode(engine):
    import torch
    inp, b, x, nRep = _ekf3_replicate_batch(nR=2, nRep=40)
    B, T, L = x.shape[1], b["T"], b["L"]
    kw = dict(B=B, T=T, L=L, G=nRep, x_per_traj=True, r_mode=K.R_PERDAY, fixed_R=False,
              beta=b["beta"], gamma=b["gamma"], W=b["W"], outputs=("S_SMOOTH", "u_opt_smooth", "rho"))
    host = engine.ekf_eks(K.MODEL_SIALPHA, b["prm"], b["u"], x, b["R"], b["Q"], b["s_init"], b["Ps_init"],
                          b["s_final"], b["Ps_final"], **kw)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    d = engine.ekf_eks(K.MODEL_SIALPHA, b["prm"], dev(b["u"]), dev(x), dev(b["R"]), dev(b["Q"]),
                       dev(b["s_init"]), dev(b["Ps_init"]), dev(b["s_final"]), dev(b["Ps_final"]), **kw)
    engine.sync()
    for k in host:
        assert_bits(d[k].cpu().numpy(), host[k], f"device-mode {k}")


def test_ekf6_per_trajectory_epsilon_batch(engine):
    """Generic batched entry with the 6-state model: one region, epsilon per trajectory."""
    c = cases.ekf6_case(2, T_hist=50, T_fore=25)
    eps = np.array([1e-9, 1e-3, 0.05, 0.3, 0.7, 0.999])
    T, L = c["u"].shape[1], 12
    cm = lambda P: np.ascontiguousarray(np.asarray(P).T).ravel()
    out = engine.ekf_eks(K.MODEL_OPTCTRL, pack_params([c["params"]], L), c["u"].T.copy(), c["x"], c["R_v"],
                         cm(c["Q_w"]), c["s_init"], cm(c["Ps_init"]), c["s_final"], cm(c["Ps_final"]),
                         B=eps.size, T=T, L=L, G=eps.size, epsilon=eps, r_mode=K.R_PERDAY, fixed_R=False,
                         beta=1.0, gamma=0.995, W=21, want_status=True)
    o = orc()
    for e, ev in enumerate(eps):
        cc = dict(c, params=dict(c["params"], epsilon=ev))
        want = o.ekf_eks(o.OPTCTRL, *ekf_args(cc))
        for k in ("S_SMOOTH", "u_opt", "u_opt_smooth", "S_PLUS"):
            assert_bits(out[k][:, :, e].T, want[k], f"eps={ev} {k}")
        assert_bits(out["P_SMOOTH"][:, :, e].reshape(T, 6, 6).transpose(2, 1, 0), want["P_SMOOTH"])
        # north_star tolerance for schedules is implied by bit equality; state it anyway
        assert np.max(np.abs(out["u_opt_smooth"][:, :, e].T - want["u_opt_smooth"])) <= TOL_COST
    assert out["status"].shape == (eps.size,)


@pytest.mark.parametrize("model", ["sialpha", "optctrl"])
def test_forward_instantiations_agree(engine, model):
    """The forward pass has three instantiations per model: strided tape (every output requested), tiled
    scratch tape with optional per-day outputs (run-time dispatch in the day loop) and tiled PLAIN (constant
    Q, no optional outputs fixed at compile time: csrc/ekf_forward.cu).  Same bits from all three, for a
    ragged batch that spans several tiles."""
    if model == "sialpha":
        inp, b, x, nRep = _ekf3_replicate_batch(nR=3, nRep=27)
        B, T, L = x.shape[1], b["T"], b["L"]
        args = (K.MODEL_SIALPHA, b["prm"], b["u"], x, b["R"], b["Q"], b["s_init"], b["Ps_init"], b["s_final"],
                b["Ps_final"])
        kw = dict(B=B, T=T, L=L, G=nRep, x_per_traj=True, r_mode=K.R_PERDAY, fixed_R=False, beta=b["beta"],
                  gamma=b["gamma"], W=b["W"])
    else:
        c = cases.ekf6_case(1, T_hist=40, T_fore=20)
        eps = np.concatenate([np.logspace(-9, -0.001, 40), [0.3, 0.999, 1e-12]])
        T, L = c["u"].shape[1], 12
        cm = lambda P: np.ascontiguousarray(np.asarray(P).T).ravel()
        args = (K.MODEL_OPTCTRL, pack_params([c["params"]], L), c["u"].T.copy(), c["x"], c["R_v"], cm(c["Q_w"]),
                c["s_init"], cm(c["Ps_init"]), c["s_final"], cm(c["Ps_final"]))
        kw = dict(B=eps.size, T=T, L=L, G=eps.size, epsilon=eps, r_mode=K.R_PERDAY, fixed_R=False, beta=1.0,
                  gamma=0.995, W=21)
    full = engine.ekf_eks(*args, **kw)                                                   # strided tape
    opt = engine.ekf_eks(*args, outputs=("S_SMOOTH", "u_opt", "K_GAIN", "innovations", "u_opt_smooth"), **kw)
    plain = engine.ekf_eks(*args, outputs=("S_SMOOTH", "u_opt_smooth"), **kw)            # tiled, PLAIN
    for k in ("S_SMOOTH", "u_opt", "K_GAIN", "innovations", "u_opt_smooth"):
        assert_bits(opt[k], full[k], f"tiled+optional outputs: {k}")
    for k in ("S_SMOOTH", "u_opt_smooth"):
        assert_bits(plain[k], full[k], f"tiled PLAIN: {k}")


@pytest.mark.parametrize("nS", [10, 257])
def test_generated_schedules_match_oracle_and_supplied(engine, nS):
    """EPI_U_PHILOX: the schedules drawn in the kernel are the oracle's (bit for bit), integrating
    them equals integrating the same schedules supplied as uint8, and a batch split at a region
    boundary (`first`) reproduces the unsplit result."""
    nR, Kn, L = 3, 37, 12
    reg = syn.load_regions(nR)
    B = nR * nS
    umin = np.zeros(L)
    prm_d = [dict(dt=1.0, beta=syn.BETA, gamma=syn.GAMMA, b=reg["b"][r], a=reg["a"][r], u_max=reg["npi_max"],
                  u_min=umin, alpha_min=1e-8, alpha_max=100.0) for r in range(nR)]
    prm = pack_params(prm_d, L)
    x0 = np.array([[(reg["N"][r] - 10) / reg["N"][r], 10 / reg["N"][r], syn.ALPHA0] for r in range(nR)])
    w = np.stack([np.repeat(reg["cost_weights"][r][None, :], Kn, axis=0) for r in range(nR)])
    seed = 0x1234_5678_9ABC_DEF0
    u = engine.random_schedules(prm, B, Kn, L, nS, seed)
    assert u.dtype == np.uint8 and u.shape == (Kn, L, B)
    o = orc()
    for b in sorted(set(list(range(0, B, 41)) + [nS // 2 - 2, nS // 2 - 1, nS // 2, B - 1])):
        ref = o.random_schedule(seed, b // nS, b % nS, nS, L, Kn, umin, reg["npi_max"])
        assert np.array_equal(u[:, :, b].T, ref), f"schedule of trajectory {b}"
    kw = dict(G=nS, want_traj=True, want_cost=True, T_total=Kn, w=w)
    gen = engine.rollout_cost(prm, x0, None, Kn, L, B=B, seed=seed, **kw)
    sup = engine.rollout_cost(prm, x0, u, Kn, L, **kw)
    for k in ("s", "i", "alpha", "J0", "J1"):
        assert_bits(gen[k], sup[k], f"generated vs supplied {k}")
    # shard: regions 1.. as their own call
    tail = engine.rollout_cost(pack_params(prm_d[1:], L), x0[1:], None, Kn, L, B=B - nS, seed=seed, first=nS,
                               G=nS, want_traj=False, want_cost=True, T_total=Kn, w=w[1:])
    assert_bits(tail["J0"], gen["J0"][nS:]); assert_bits(tail["J1"], gen["J1"][nS:])
    with pytest.raises(K.EpiError):
        engine.rollout_cost(prm, x0, None, Kn, L, B=B, seed=seed, first=1, **kw)   # not a multiple of G
    bad = pack_params([dict(d, u_min=np.full(L, np.nan)) for d in prm_d], L)
    with pytest.raises(K.EpiError):
        engine.rollout_cost(bad, x0, None, Kn, L, B=B, seed=seed, **kw)            # level bounds are required


# ------------------------------------------------------------------------------ preprocessing + CSV pipeline
def test_preprocess_batch_bit_exact(engine):
    rng = np.random.default_rng(8)
    T, B, L = 97, 70, 12
    nc = np.maximum(0, rng.poisson(40, (T, B)) + np.arange(T)[:, None] * rng.random(B))
    cc = np.cumsum(nc, axis=0)
    cc[rng.integers(1, T, 200), rng.integers(0, B, 200)] = np.nan
    cc[T - 1, ::5] = np.nan
    cc[40, 3] -= 500.0
    cc[:, 7] = 0.0                                       # a region without cases: I0 = min_cases
    cc[:, 9] = np.nan                                    # a region without data at all
    ip = rng.integers(0, 5, (T, L, B)).astype(float)
    ip[rng.integers(0, T, 900), rng.integers(0, L, 900), rng.integers(0, B, 900)] = np.nan
    pop = rng.uniform(1e5, 3e8, B)
    got = engine.preprocess(cc, pop, ip)
    o = orc()
    for b in range(B):
        ref = o.preprocess_region(cc[:, b], ip[:, :, b], pop[b])
        for k in ("refined", "smoothed", "zerolag", "normalized", "confirmed_norm", "R_v"):
            assert_bits(got[k][:, b], ref[k], f"{k} region {b}")
        assert_bits(got["ip_filled"][:, :, b], ref["ip"], f"ip region {b}")
        assert_bits(got["I0"][b:b + 1], np.array([ref["I0"]]), f"I0 region {b}")
    with pytest.raises(K.EpiError):
        engine.preprocess(cc[:9], pop, ip[:9])          # filtfilt's data-length rule


def test_csv_to_prescriptions_pipeline(engine, tmp_path):
    """OxCGRT-format CSV -> preprocess -> 3-state EKF/EKS -> optimal-NPI sweep -> knee -> prescription
    CSV, every numeric stage on the device; J0/J1 against the oracle on the same preprocessed inputs."""
    import importlib.util
    import pandas as pd
    from epidemicmodeling_b200 import pipeline, xprize_io as xio
    spec = importlib.util.spec_from_file_location("pfc", os.path.join(os.path.dirname(HERE), "tools", "prescribe_from_csv.py"))
    pfc = importlib.util.module_from_spec(spec); spec.loader.exec_module(pfc)
    data, out = str(tmp_path / "ox.csv"), str(tmp_path / "presc.csv")
    regions, start, end = pfc.synthetic_oxcgrt(data, 3, 60)
    eps = syn.epsilon_grid_xprize02(10)
    Tf = 20
    res = pipeline.prescribe_from_csv(engine, data, start, end, Tf, eps, regions, out_file=out,
                                      forecast_dates=[f"2020-06-{d + 1:02d}" for d in range(Tf)])
    assert res["J0"].shape == (3, 10) and res["u_knee"].shape == (3, Tf, 12)
    assert np.isfinite(res["J0"]).all() and np.isfinite(res["J1"]).all()
    df = pd.read_csv(out, keep_default_na=False)
    assert list(df.columns)[:4] == ["PrescriptionIndex", "CountryName", "RegionName", "Date"] and len(df) == 3 * Tf
    lv = df[xio.NPI_COLUMNS].to_numpy()
    assert (lv >= 0).all() and (lv <= xio.NPI_MAXES[None, :]).all()
    assert not lv[Tf - 1::Tf].any()                      # u_opt_smooth(:, T) = 0 quirk reaches the file
    # oracle on the same (device-preprocessed) inputs
    o = orc()
    pre, ids = res["pre"], res["ids"]
    inputs = [pipeline.setup_from_preprocessed(pre, b, regions[g]["N"], regions[g]["a"], regions[g]["b"], xio.NPI_MAXES,
                                               regions[g]["weights"], 60, Tf) for b, g in enumerate(ids)]
    S = wl.run_fixed_input(engine, inputs)
    for r, rin in enumerate(inputs):
        s6, Th, Sr = rin["setup6"], rin["T_hist"], S[:, :, r].T
        reg = o.SweepRegion(s6["params"], rin["T"], Th, rin["u_hist"], rin["x"], rin["R_v"], s6["s_init"], s6["Ps_init"],
                            s6["s_final"], s6["Ps_final"], s6["Q_w"], 1.0, 0.995, 21, Sr[0, Th - 1], Sr[1, Th - 1],
                            Sr[2, Th - 1], (Sr[0, :Th] * Sr[1, :Th]) * Sr[2, :Th], rin["weights"])
        j0, j1, m, io, _ = o.sweep_region(reg, eps)
        assert_bits(res["J0"][r], j0, "pipeline J0"); assert_bits(res["J1"][r], j1, "pipeline J1")
        assert res["I_opt"][r] == io


def test_nnls_affine_batch_bit_exact(engine):
    rng = np.random.default_rng(12)
    n, p, B = 90, 12, 40
    umax = np.array([3, 3, 2, 4, 2, 3, 2, 4, 2, 3, 2, 4.0])
    U = np.stack([np.stack([rng.integers(0, int(m) + 1, n // 10).repeat(10) for m in umax], 1) for _ in range(B)], 2)
    X = np.ascontiguousarray(umax[None, :, None] - U.astype(float))          # [n, p, B]
    atrue = np.where(rng.random((p, B)) < 0.5, rng.random((p, B)) * 0.05, 0.0)
    y = np.einsum("npb,pb->nb", X, atrue) + 0.03 + 0.004 * rng.standard_normal((n, B))
    X[:, 4, 3] = 0.0; X[:, 7, 5] = X[:, 2, 5]                                 # zero / collinear columns
    a, b, k = engine.nnls_affine(X, y)
    o = orc()
    for r in range(B):
        ra, rb, rk = o.nnls_affine(X[:, :, r], y[:, r])
        assert_bits(a[:, r], ra, f"a region {r}"); assert_bits(b[r:r + 1], np.array([rb]), f"b region {r}")
        assert k[r] == rk
    a0, b0, k0 = engine.nnls_affine(X, y, max_alt=0)                           # plain lsqnonneg
    assert not b0.any() and not k0.any()
    assert_bits(a0[:, 0], o.lsqnonneg(X[:, :, 0], y[:, 0]))
    with pytest.raises(K.EpiError):
        engine.nnls_affine(np.zeros((5, 13, 2)), np.zeros((5, 2)))


def test_training_rounds_pipeline(engine, tmp_path):
    """preprocess -> EKF round 1 (zero inputs) -> regression -> EKF round 2 (real inputs) -> regression,
    all regions at once; the second-round fit against the oracle on the same smoothed alpha."""
    import importlib.util
    from epidemicmodeling_b200 import pipeline, xprize_io as xio
    spec = importlib.util.spec_from_file_location("pfc", os.path.join(os.path.dirname(HERE), "tools", "prescribe_from_csv.py"))
    pfc = importlib.util.module_from_spec(spec); spec.loader.exec_module(pfc)
    data = str(tmp_path / "ox.csv")
    regions, start, end = pfc.synthetic_oxcgrt(data, 4, 120)
    ids, dates, cc, _, ip = xio.read_oxcgrt(data, start, end)
    pops = np.array([regions[g]["N"] for g in ids])
    pre = engine.preprocess(cc, pops, ip)
    tr = pipeline.train_regions(engine, pre, pops, xio.NPI_MAXES, 120, 90)
    assert tr["a2"].shape == (12, 4) and (tr["a2"] >= 0).all() and (tr["a1"] >= 0).all() and np.isfinite(tr["b2"]).all()
    assert tr["a2"].any()                                                     # the NPIs explain part of alpha


# ------------------------------------------------------------------------------ Rt_ExpFitEKF
TOL_RT = 1e-9   # north_star tolerance for EKF states; exp/tanh are not correctly rounded on either side


@pytest.mark.parametrize("order", [1, 2])
def test_rt_expfit_signature_mirror(order):
    """api.Rt_ExpFitEKF (Tools/Rt_ExpFitEKF.m:1 signature) against the oracle."""
    c = cases.rt_expfit_case(order=order)
    got, ref = api.Rt_ExpFitEKF(**c), orc().Rt_ExpFitEKF(**c)
    names = ("S_MINUS", "S_PLUS", "P_MINUS", "P_PLUS", "K_GAIN", "S_SMOOTH", "P_SMOOTH", "innovations", "rho")
    for n, a, b in zip(names, got, ref):
        assert a.shape == b.shape, n
        scale = np.max(np.abs(b)) or 1.0
        assert np.max(np.abs(a - b)) / scale <= TOL_RT, f"{n}: {np.max(np.abs(a - b)) / scale:.2e}"
    with pytest.raises(ValueError, match="Undefined order"):
        api.Rt_ExpFitEKF(**dict(c, order=0))


def test_rt_expfit_batch(engine):
    """A batch of series with per-group parameters, ragged block (B = 70 > one 64-thread CTA),
    outputs on request only."""
    B, G, T = 70, 35, 150
    cs = [cases.rt_expfit_case(T=T, seed=100 + b, order=2) for b in range(B)]
    x = np.stack([c["x"][0] for c in cs], axis=1)
    s0 = np.stack([c["s_init"] for c in cs], axis=1)
    prm = np.array([[1.0, 0.9, 0.1], [1.0, 0.8, 0.2]])
    Qs = [np.diag([250.0 ** 2, 3.0e-3 ** 2]), np.diag([100.0 ** 2, 5.0e-3 ** 2])]
    res = engine.rt_expfit(x, s0, prm, np.zeros((2, 2)), np.stack([(100 * q).T.ravel() for q in Qs]),
                           np.stack([q.T.ravel() for q in Qs]), np.array([100.0, 400.0]), T=T, G=G, beta=0.9,
                           gamma=0.995, W=21, order=2, outputs=("S_SMOOTH", "rho"))
    assert set(res) == {"S_SMOOTH", "rho"}
    o = orc()
    for b in (0, 34, 35, 69):
        g = b // G
        ref = o.Rt_ExpFitEKF(x[:, b], s0[:, b], prm[g], [0, 0], 0.0, 100 * Qs[g], Qs[g], [100.0, 400.0][g], 0.9, 0.995, 21, 2)
        for a, r in ((res["S_SMOOTH"][:, :, b].T, ref[5]), (res["rho"][:, b], ref[8])):
            assert np.max(np.abs(a - r)) / np.max(np.abs(r)) <= TOL_RT
    with pytest.raises(K.EpiError):
        engine.rt_expfit(x, s0, prm, np.zeros((2, 2)), np.zeros((2, 4)), np.zeros((2, 4)), np.ones(2), T=T, G=G, order=3)


# ------------------------------------------------------------------------------ fused sweep
def _run_sweep(engine, inp, eps, **kw):
    S = wl.run_fixed_input(engine, inp)
    batch = wl.sweep_batch(inp, S)
    return S, wl.run_sweep(engine, batch, eps, **kw)


def test_sweep_matches_oracle_and_golden(engine):
    inp, eps = cases.sweep_case()
    S, res = _run_sweep(engine, inp, eps, want_front=True, want_u_fore=True, want_u_knee=True, want_P_first=True)
    g = np.load(os.path.join(GOLD, "sweep.npz"))
    assert_bits(res["J0"], g["J0"], "J0 golden"); assert_bits(res["J1"], g["J1"], "J1 golden")
    assert np.array_equal(res["on_front"].astype(bool), g["mask"]) and np.array_equal(res["I_opt"], g["iopt"])
    nR, nE = g["J0"].shape
    Tf, L = inp[0]["T"] - inp[0]["T_hist"], 12
    uf = res["u_fore"].reshape(Tf, L, nR, nE)
    assert_bits(np.transpose(uf, (2, 3, 1, 0)), g["u_fore"], "u_fore golden")
    for r in range(nR):
        assert_bits(res["u_knee"][r].T, g["u_fore"][r, g["iopt"][r]], "knee schedule")
    assert not res["u_fore"][Tf - 1].any()                   # u_opt_smooth(:,T) == 0 quirk reaches the rollout
    # live oracle on the same inputs (fixed-input smoother included)
    o = orc()
    for r, rin in enumerate(inp):
        s3 = rin["setup3"]
        o3 = o.ekf_eks(o.SIALPHA, rin["u_fixed"], rin["x"], s3["params"], s3["s_init"], s3["Ps_init"],
                       s3["s_final"], s3["Ps_final"], s3["w_bar"], 0.0, s3["Q_w"], rin["R_v"], 1.0,
                       s3["gamma_ekf"], s3["W"], 1)
        assert_bits(S[:, :, r].T, o3["S_SMOOTH"], "fixed-input smoother")
    assert np.max(np.abs(res["J0"] - g["J0"]) / np.abs(g["J0"])) <= TOL_COST
    # lean mode (smoother on the days to optimise only) returns the same bits
    batch = wl.sweep_batch(inp, S)
    lean = wl.run_sweep(engine, batch, eps, want_front=True, want_u_fore=True, want_u_knee=True, lean=True)
    for k in ("J0", "J1", "on_front", "I_opt", "u_fore", "u_knee"):
        assert_bits(lean[k], res[k], f"lean {k}")
    with pytest.raises(K.EpiError):
        wl.run_sweep(engine, batch, eps, want_P_first=True, lean=True)


@pytest.mark.parametrize("T_hist,T_fore", [(0, 25), (25, 0), (1, 1), (40, 1)])
def test_sweep_lean_edge_shapes(engine, T_hist, T_fore):
    """No history / no forecast / single days: lean == full, and == the oracle.  With no history there is no
    fixed-input smoother to start the rollout from (TrainPredictPrescribeNPI.m:481 indexes the last historic
    day): the caller supplies the start state and an empty new-case history."""
    inp, eps = cases.sweep_case(n_regions=2, n_eps=5, T_hist=T_hist, T_fore=T_fore)
    if T_hist == 0:
        x0 = np.stack([r["setup6"]["s_init"][:3] for r in inp])
        S = np.zeros((T_fore, 3, len(inp)))
        batch = wl.sweep_batch(inp, S, x0=x0)
    else:
        S = wl.run_fixed_input(engine, inp)
        batch = wl.sweep_batch(inp, S)
    full = wl.run_sweep(engine, batch, eps)
    lean = wl.run_sweep(engine, batch, eps, lean=True)
    for k in ("J0", "J1", "on_front", "I_opt"):
        assert_bits(lean[k], full[k], f"lean {k} T_hist={T_hist} T_fore={T_fore}")
    o = orc()
    for r, rin in enumerate(inp):
        s6 = rin["setup6"]
        reg = o.SweepRegion(s6["params"], rin["T"], T_hist, rin["u_hist"], rin["x"], rin["R_v"], s6["s_init"],
                            s6["Ps_init"], s6["s_final"], s6["Ps_final"], s6["Q_w"], 1.0, 0.995, 21,
                            batch["x0"][r, 0], batch["x0"][r, 1], batch["x0"][r, 2], batch["newcases_hist"][r],
                            rin["weights"])
        j0, j1, m, io, _ = o.sweep_region(reg, eps)
        assert_bits(full["J0"][r], j0, f"J0 T_hist={T_hist} T_fore={T_fore}")
        assert_bits(full["J1"][r], j1, f"J1 T_hist={T_hist} T_fore={T_fore}")
        assert np.array_equal(full["on_front"][r].astype(bool), m) and full["I_opt"][r] == io


def _guard_case(name):
    """6-state inputs that drive P_MINUS out of the well-behaved regime (GenericExtendedKalmanFilter.m:209-215)."""
    c = cases.ekf6_case(0, T_hist=30, T_fore=12)
    z = np.zeros((6, 6))
    if name == "overflow":      # covariances overflow to Inf on day 5: the isnan/isinf guard sets J = 0
        c.update(Ps_init=np.eye(6) * 1e306, Q_w=np.eye(6) * 1e307)
    elif name == "infQ":        # an infinite process-noise entry: guard from the first day on
        c.update(Q_w=np.diag([1e-6, 1e-6, np.inf, 1e-6, 1e-6, 1e-6]))
    elif name == "zero":        # P_MINUS == 0: pinv of the zero matrix, rank 0
        c.update(Ps_init=z.copy(), Q_w=z.copy())
    elif name == "rank1":       # rank-1 covariance, no process noise: rank 1 on every day
        e = z.copy(); e[1, 1] = 1e-4
        c.update(Ps_init=e, Q_w=z.copy())
    elif name == "rank2q":      # rank grows 2 -> 4 -> 5 and stays deficient
        c.update(Ps_init=z.copy(), Q_w=np.diag([1e-8, 1e-8, 0, 0, 0, 0.0]))
    return c


@pytest.mark.parametrize("name", ["overflow", "infQ", "zero", "rank1", "rank2q"])
def test_smoother_guard_and_rank_deficit(engine, name):
    """The P_MINUS NaN/Inf guard (J = 0, :209-214) and rank-deficient pinv (:215): same bits as the oracle
    (NaN positions included) and the status word reports both -- (6 - min rank) << 8 | guard-hit."""
    c = _guard_case(name)
    eps = np.array([1e-6, 0.3, 0.9])
    T, L = c["u"].shape[1], 12
    cm = lambda P: np.ascontiguousarray(np.asarray(P).T).ravel()
    out = engine.ekf_eks(K.MODEL_OPTCTRL, pack_params([c["params"]], L), c["u"].T.copy(), c["x"], c["R_v"],
                         cm(c["Q_w"]), c["s_init"], cm(c["Ps_init"]), c["s_final"], cm(c["Ps_final"]),
                         B=eps.size, T=T, L=L, G=eps.size, epsilon=eps, r_mode=K.R_PERDAY, fixed_R=False,
                         beta=1.0, gamma=0.995, W=21, want_status=True)
    o = orc()
    seen_guard = seen_deficit = False
    for e, ev in enumerate(eps):
        cc = dict(c, params=dict(c["params"], epsilon=ev))
        with np.errstate(all="ignore"):
            want = o.ekf_eks(o.OPTCTRL, *ekf_args(cc))
        for k in ("S_PLUS", "S_SMOOTH", "u_opt", "u_opt_smooth"):
            assert_bits(out[k][:, :, e].T, want[k], f"{name} eps={ev} {k}")
        assert_bits(out["P_SMOOTH"][:, :, e].reshape(T, 6, 6).transpose(2, 1, 0), want["P_SMOOTH"], f"{name} P_SMOOTH")
        status = 0
        for k in range(1, T):
            Pm = want["P_MINUS"][:, :, k]
            if not np.all(np.isfinite(Pm)):
                status = max(status, 1)
            else:
                status = max(status, (6 - o.pinv_sym(Pm)[1]) << 8)
        assert int(out["status"][e]) == status, (name, ev, int(out["status"][e]), status)
        seen_guard |= bool(status & 1)
        seen_deficit |= bool(status >> 8)
    assert seen_guard == (name in ("overflow", "infQ")) and seen_deficit == (name in ("zero", "rank1", "rank2q"))


@pytest.mark.parametrize("segments", [2, 5, 64])
def test_sweep_segmented_forward_is_bit_identical(engine, segments, monkeypatch):
    """The time-segmented persistent forward launch (few-wave batches) resumes from the tape pages
    of the previous segment: same bits as the plain launch, full and lean, ragged last tile,
    more segments than forecast days."""
    inp, eps = cases.sweep_case(n_regions=3, n_eps=15, T_hist=33, T_fore=21)   # 45 trajectories: 2 tiles, one ragged
    S = wl.run_fixed_input(engine, inp)
    batch = wl.sweep_batch(inp, S)
    monkeypatch.setenv("EPI_FWD_SEGMENTS", "1")
    ref = wl.run_sweep(engine, batch, eps, want_front=True, want_u_fore=True, want_P_first=True)
    monkeypatch.setenv("EPI_FWD_SEGMENTS", str(segments))
    seg = wl.run_sweep(engine, batch, eps, want_front=True, want_u_fore=True, want_P_first=True)
    for k in ("J0", "J1", "on_front", "I_opt", "u_fore", "P_first"):
        assert_bits(seg[k], ref[k], f"segmented {k} S={segments}")
    lean = wl.run_sweep(engine, batch, eps, want_front=True, want_u_fore=True, lean=True)
    for k in ("J0", "J1", "on_front", "I_opt", "u_fore"):
        assert_bits(lean[k], ref[k], f"segmented lean {k} S={segments}")


def test_sweep_with_noise_waves_and_device_mode(engine):
    import torch
    inp, eps = cases.sweep_case(n_regions=2, n_eps=9, T_hist=40, T_fore=20)
    S = wl.run_fixed_input(engine, inp)
    batch = wl.sweep_batch(inp, S)
    B, Tf = 2 * eps.size, 20
    rng = np.random.default_rng(77)
    noise = rng.standard_normal((Tf, 3, B))
    nstd = np.array([inp[r]["setup3"]["noise_std"] for r in range(2)])
    ref = wl.run_sweep(engine, batch, eps, noise=noise, noise_std=nstd)
    o = orc()
    for r, rin in enumerate(inp):
        s6 = rin["setup6"]
        Th = rin["T_hist"]
        Sr = S[:, :, r].T
        reg = o.SweepRegion(s6["params"], rin["T"], Th, rin["u_hist"], rin["x"], rin["R_v"], s6["s_init"],
                            s6["Ps_init"], s6["s_final"], s6["Ps_final"], s6["Q_w"], 1.0, 0.995, 21,
                            Sr[0, Th - 1], Sr[1, Th - 1], Sr[2, Th - 1], (Sr[0, :Th] * Sr[1, :Th]) * Sr[2, :Th],
                            rin["weights"], noise_std=tuple(nstd[r]),
                            noise=np.ascontiguousarray(noise[:, :, r * eps.size:(r + 1) * eps.size].transpose(2, 0, 1)))
        j0, j1, m, io, _ = o.sweep_region(reg, eps)
        assert_bits(ref["J0"][r], j0, "noisy J0"); assert_bits(ref["J1"][r], j1, "noisy J1")
        assert np.array_equal(ref["on_front"][r].astype(bool), m) and ref["I_opt"][r] == io
    engine.set_scratch_limit(5 * 60 * 8 * 200)              # forces several waves
    try:
        waved = wl.run_sweep(engine, batch, eps, noise=noise, noise_std=nstd)
    finally:
        engine.set_scratch_limit(0)
    assert_bits(waved["J0"], ref["J0"]); assert_bits(waved["J1"], ref["J1"])
    dbatch = wl.sweep_to_device(batch, eps, "cuda:0")
    dres = wl.run_sweep(engine, dbatch, eps, noise=torch.from_numpy(noise).cuda(),
                        noise_std=torch.from_numpy(nstd).cuda())
    engine.sync()
    assert_bits(dres["J0"].cpu().numpy(), ref["J0"]); assert_bits(dres["J1"].cpu().numpy(), ref["J1"])
    assert np.array_equal(dres["I_opt"].cpu().numpy(), ref["I_opt"])


def test_sweep_full_size_properties(engine):
    """BASELINE config 4 at full size (236 regions x 250 epsilon x (441+120) days): properties
    that do not need the oracle at scale + an oracle spot check of a few (region, epsilon)."""
    nR, Th, Tf = 236, 441, 120
    inp = syn.sweep_inputs(n_regions=nR, T_hist=Th, T_fore=Tf)
    eps = syn.epsilon_grid_xprize02(250)
    S, res = _run_sweep(engine, inp, eps, want_front=True, want_u_knee=True)
    lean = wl.run_sweep(engine, wl.sweep_batch(inp, S), eps, want_front=True, want_u_knee=True, lean=True)
    for k in ("J0", "J1", "on_front", "I_opt", "u_knee"):
        assert_bits(lean[k], res[k], f"full-size lean {k}")
    J0, J1, m = res["J0"], res["J1"], res["on_front"].astype(bool)
    assert J0.shape == (nR, 250) and np.isfinite(J0).all() and np.isfinite(J1).all()
    assert (J0 >= 0).all() and (J1 >= 0).all()
    wmax = np.array([np.mean(r["weights"][:, 0] * np.array(r["setup6"]["params"]["u_max"])) for r in inp])
    assert (J1 <= wmax[:, None] * (1 + 1e-12)).all()          # cost cannot exceed "everything at max"
    assert m.any(axis=1).all()                                # every region has a non-empty front
    m2, _ = engine.pareto(np.where(m, J0, np.inf), np.where(m, J1, np.inf))
    assert np.array_equal(m2.astype(bool) & m, m)             # idempotence
    umax = np.asarray(inp[0]["setup6"]["params"]["u_max"])
    uk = res["u_knee"]
    assert (((uk == 0) | (uk == umax[None, None, :]))).all()  # bang-bang knee schedules
    o = orc()
    for r, e in ((0, 0), (17, 124), (123, 125), (235, 249)):
        rin = inp[r]
        s6 = rin["setup6"]
        Sr = S[:, :, r].T
        reg = o.SweepRegion(s6["params"], rin["T"], Th, rin["u_hist"], rin["x"], rin["R_v"], s6["s_init"],
                            s6["Ps_init"], s6["s_final"], s6["Ps_final"], s6["Q_w"], 1.0, 0.995, 21,
                            Sr[0, Th - 1], Sr[1, Th - 1], Sr[2, Th - 1], (Sr[0, :Th] * Sr[1, :Th]) * Sr[2, :Th],
                            rin["weights"])
        j0, j1, _, _, _ = o.sweep_region(reg, eps[e:e + 1])
        assert J0[r, e] == j0[0] and J1[r, e] == j1[0], (r, e)


# ------------------------------------------------------------------------------ next row
def test_forecast_quality_loop(engine):
    """Tools/ForecastQualityAssessment.m:383-394: masked-horizon re-runs as one batch vs the oracle loop."""
    inp = syn.sweep_inputs(n_regions=2, T_hist=70, T_fore=0)
    nf, look = 9, 5
    res = wl.forecast_quality(engine, inp, nf, look)
    T = inp[0]["T"]
    o = orc()
    for r, rin in enumerate(inp):
        s3 = rin["setup3"]
        for start in (1, 4, nf):
            xp = rin["x"].copy()
            xp[T - start:] = np.nan
            want = o.ekf_eks(o.SIALPHA, rin["u_fixed"], xp, s3["params"], s3["s_init"], s3["Ps_init"], s3["s_final"],
                             s3["Ps_final"], s3["w_bar"], 0.0, s3["Q_w"], rin["R_v"], 1.0, s3["gamma_ekf"], s3["W"], 1)
            b = r * nf + start - 1
            assert_bits(res["S_PLUS"][:, :, b].T, want["S_PLUS"], "forecast-quality S_PLUS")
            assert_bits(res["S_SMOOTH"][:, :, b].T, want["S_SMOOTH"], "forecast-quality S_SMOOTH")
            est = (want["S_SMOOTH"][0] * want["S_SMOOTH"][1]) * want["S_SMOOTH"][2]
            e = 100.0 * np.abs(rin["x"] - est) / rin["x"]
            last = min(T, T - start + look)
            assert np.array_equal(res["EstError_SMOOTH"][r, start - 1, :last - T + start], e[T - start:last])


# ------------------------------------------------------------------------------ per-trajectory inputs, Q/R modes
@pytest.mark.parametrize("q_mode", ["const", "perday_scalar", "perday_full"])
def test_ekf3_everything_per_trajectory(engine, q_mode):
    """u, x, R, initial/final conditions per trajectory and the three Q shapes of
    GenericExtendedKalmanFilter.m:64-77 (square => constant, length-T vector, m x m x T)."""
    rng = np.random.default_rng(101)
    B, T, L, m = 37, 55, 12, 3
    reg = syn.load_regions(4)
    base = [cases.ekf3_case(r, T_hist=T - 10, T_fore=10) for r in range(4)]
    u = np.empty((T, L, B)); x = np.empty((T, B)); R = np.empty((T, B))
    s_init = np.empty((m, B)); Ps_init = np.empty((m * m, B)); s_final = np.empty((m, B)); Ps_final = np.empty((m * m, B))
    per = []
    for b in range(B):
        c = base[b % 4]
        ub = c["u"].copy(); ub[:, rng.integers(0, T, 5)] = rng.integers(0, 3, (L, 5))
        xb = c["x"] * (1 + 0.02 * rng.standard_normal(T))
        Rb = c["R_v"] * (1 + 0.5 * rng.random(T))
        si = c["s_init"] * np.array([1.0, 1 + 0.1 * rng.random(), 1 + 0.05 * rng.random()])
        Pi = np.asarray(c["Ps_init"]) * (1 + rng.random())
        sf = np.array([np.nan, np.nan, 0.1 + 0.05 * rng.random()]) if b % 3 == 0 else np.full(3, np.nan)
        Pf = np.full((3, 3), np.nan)
        if b % 3 == 0:
            Pf[2, 2] = 1e-4
        u[:, :, b], x[:, b], R[:, b] = ub.T, xb, Rb
        s_init[:, b], Ps_init[:, b], s_final[:, b], Ps_final[:, b] = si, Pi.T.ravel(), sf, Pf.T.ravel()
        per.append((c, ub, xb, Rb, si, Pi, sf, Pf))
    prm = pack_params([base[g]["params"] for g in range(4)] * 10, L)[:B]  # one group per trajectory (G = 1)
    prm = pack_params([base[b % 4]["params"] for b in range(B)], L)
    Q0 = np.asarray(base[0]["Q_w"])
    if q_mode == "const":
        Qarg, qm, Qm = np.stack([Q0.T.ravel()] * B), K.Q_CONST, Q0
    elif q_mode == "perday_scalar":
        qv = 1e-10 * (1 + rng.random(T))
        Qarg, qm, Qm = np.stack([qv] * B), K.Q_PERDAY_SCALAR, qv
    else:
        Q3 = np.stack([Q0 * (1 + 0.3 * rng.random()) for _ in range(T)], axis=2)   # m x m x T
        Qarg, qm, Qm = np.stack([np.transpose(Q3, (2, 1, 0)).ravel()] * B), K.Q_PERDAY_FULL, Q3
    out = engine.ekf_eks(K.MODEL_SIALPHA, prm, u, x, R, Qarg, s_init, Ps_init, s_final, Ps_final, B=B, T=T, L=L,
                         G=1, u_per_traj=True, x_per_traj=True, r_mode=K.R_PERDAY, fixed_R=False, r_per_traj=True,
                         q_mode=qm, init_per_traj=True, beta=1.0, gamma=0.995, W=21)
    o = orc()
    for b in (0, 1, 2, 3, 17, B - 1):
        c, ub, xb, Rb, si, Pi, sf, Pf = per[b]
        want = o.ekf_eks(o.SIALPHA, ub, xb, c["params"], si, Pi, sf, Pf, c["w_bar"], 0.0, Qm, Rb, 1.0, 0.995, 21, 1)
        for k in ("S_MINUS", "S_PLUS", "S_SMOOTH", "u_opt", "u_opt_smooth"):
            assert_bits(out[k][:, :, b].T, want[k], f"{q_mode} b={b} {k}")
        for k in ("P_MINUS", "P_PLUS", "P_SMOOTH"):
            assert_bits(out[k][:, :, b].reshape(T, 3, 3).transpose(2, 1, 0), want[k], f"{q_mode} b={b} {k}")
        assert_bits(out["rho"][:, b], want["rho"][:, 0], "rho")


def test_ekf_scalar_R_per_trajectory_adaptive_and_legacy_batch(engine):
    """Constant R given per trajectory with adaptation (beta_ekf = 0.9), generic 3-state and the legacy
    6-state monolith as a batch of different epsilon."""
    c = cases.ekf3_case(1, variant="adaptive")
    B, T, L = 5, c["u"].shape[1], 12
    Rs = float(np.asarray(c["R_v"]).ravel()[0]) * np.array([0.5, 1.0, 2.0, 4.0, 8.0])
    cm = lambda P: np.ascontiguousarray(np.asarray(P).T).ravel()
    out = engine.ekf_eks(K.MODEL_SIALPHA, pack_params([c["params"]], L), c["u"].T.copy(), c["x"], Rs, cm(c["Q_w"]),
                         c["s_init"], cm(c["Ps_init"]), c["s_final"], cm(c["Ps_final"]), B=B, T=T, L=L, G=B,
                         r_mode=K.R_CONST, fixed_R=True, r_per_traj=True, beta=0.9, gamma=c["gamma"], W=21,
                         outputs=("S_SMOOTH", "rho", "K_GAIN"))
    o = orc()
    for b in range(B):
        a = list(ekf_args(c)); a[10] = np.array([[Rs[b]]])
        want = o.ekf_eks(o.SIALPHA, *a)
        assert_bits(out["S_SMOOTH"][:, :, b].T, want["S_SMOOTH"], f"adaptive R b={b}")
        assert_bits(out["rho"][:, b], want["rho"][:, 0], f"adaptive rho b={b}")
    lc = cases.legacy_case(0)
    eps = np.array([0.05, 0.2, 0.6])
    T = lc["u"].shape[1]
    out = engine.ekf_eks(K.MODEL_LEGACY_TOOLS, pack_params([lc["params"]], L), lc["u"].T.copy(), lc["x"],
                         np.asarray(lc["R_v"]).ravel(), cm(lc["Q_w"]), lc["s_init"], cm(lc["Ps_init"]), lc["s_final"],
                         cm(lc["Ps_final"]), B=3, T=T, L=L, G=3, epsilon=eps, r_mode=K.R_CONST, beta=lc["beta"],
                         gamma=lc["gamma"], W=21, outputs=("u_opt", "S_SMOOTH", "P_SMOOTH", "rho"))
    for e, ev in enumerate(eps):
        a = list(ekf_args(dict(lc, params=dict(lc["params"], epsilon=float(ev)))))
        want = o.ekf_eks(o.LEGACY_TOOLS, *a)
        assert_bits(out["S_SMOOTH"][:, :, e].T, want["S_SMOOTH"], f"legacy batch eps={ev}")
        assert_bits(out["P_SMOOTH"][:, :, e].reshape(T, 6, 6).transpose(2, 1, 0), want["P_SMOOTH"], "legacy P_SMOOTH")
        assert_bits(out["u_opt"][:, :, e].T, want["u_opt"], "legacy u_opt")


# ------------------------------------------------------------------------------ BASELINE full sizes (device memory)
def test_config2_seirp_ensemble_full_size(engine):
    """BASELINE config 2: 1M parameter sets x 365 days.  Properties: conservation of s+e+i+r+p
    (the right-hand sides of SEIRP.m:27-31 sum to zero), final-only mode == last sample; plus an
    oracle spot check."""
    import torch
    B, Kn = 1_000_000, 365
    rates, ic = syn.seirp_ensemble(B)
    rd, icd = torch.from_numpy(rates).cuda(), torch.from_numpy(ic).cuda()
    out = engine.seirp(rd, icd, Kn, 1.0)
    fin = engine.seirp(rd, icd, Kn, 1.0, out_mode=K.SEIRP_OUT_FINAL)
    engine.sync()
    tot = out.sum(dim=0)
    assert float((tot - 1.0).abs().max()) < 1e-12
    assert torch.equal(fin, out[:, -1, :])
    assert bool(torch.isfinite(out).all()) and float(out.min()) > -1e-12
    o = orc()
    for b in (0, 31, 255, 256, 499_999, B - 1):
        want = o.SEIRP(*rates[:, b], *ic[:, b], float(Kn), 1.0)
        got = out[:, :, b].cpu().numpy()
        for f in range(5):
            assert_bits(got[f], want[f][0], f"config2 b={b} f={f}")


def test_config3_ekf3_replicates_full_width_sample(engine):
    """BASELINE config 3 shape (236 regions x replicates x 400 days), 200 replicates per region in device
    memory (the 10k-replicate run is tools/bench_configs.py): bounds + oracle spot check."""
    import torch
    nR, nRep, T = 236, 200, 400
    inp = syn.sweep_inputs(n_regions=nR, T_hist=T, T_fore=0)
    b = wl.fixed_input_batch(inp)
    rng = np.random.default_rng(3)
    B = nR * nRep
    clean = np.stack([r["x"] for r in inp])                                   # [nR,T]
    x = np.maximum(0.0, np.repeat(clean.T[:, :, None], nRep, axis=2) * (1 + 0.05 * rng.standard_normal((T, nR, nRep))))
    x = np.ascontiguousarray(x.reshape(T, B))
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    from epidemicmodeling_b200.engine import params_to_device
    out = engine.ekf_eks(K.MODEL_SIALPHA, params_to_device(b["prm"], "cuda:0"), dev(b["u"]), dev(x), dev(b["R"]),
                         dev(b["Q"]), dev(b["s_init"]), dev(b["Ps_init"]), dev(b["s_final"]), dev(b["Ps_final"]),
                         B=B, T=T, L=b["L"], G=nRep, x_per_traj=True, r_mode=K.R_PERDAY, fixed_R=False, beta=1.0,
                         gamma=b["gamma"], W=b["W"], outputs=("S_SMOOTH",), want_status=True)
    engine.sync()
    S = out["S_SMOOTH"]
    assert bool(torch.isfinite(S).all())
    assert float(S[:, 0].min()) >= 0 and float(S[:, 0].max()) <= 1 and float(S[:, 1].min()) >= 0 and float(S[:, 1].max()) <= 1
    assert float(S[:, 2].min()) >= 1e-8 and float(S[:, 2].max()) <= 100.0      # StateHardMargins (:27-31)
    assert int(out["status"].max().item()) == 0
    o = orc()
    for bb in (0, 199, 200, 23_456, B - 1):
        r = inp[bb // nRep]
        s3 = r["setup3"]
        want = o.ekf_eks(o.SIALPHA, r["u_fixed"], x[:, bb], s3["params"], s3["s_init"], s3["Ps_init"], s3["s_final"],
                         s3["Ps_final"], s3["w_bar"], 0.0, s3["Q_w"], r["R_v"], 1.0, s3["gamma_ekf"], s3["W"], 1)
        assert_bits(S[:, :, bb].cpu().numpy().T, want["S_SMOOTH"], f"config3 b={bb}")


def test_config5_monte_carlo_scoring_wide(engine):
    """BASELINE config 5 shape at 236 regions x 8192 uint8 schedules x 120 days on the device (TMA-staged
    kernel) + per-region Pareto of the cloud: cost bounds, front idempotence, oracle spot check."""
    import torch
    from epidemicmodeling_b200.engine import params_to_device
    nR, nS, Kn, L = 236, 8192, 120, 12
    reg = syn.load_regions(nR)
    B = nR * nS
    g = torch.Generator(device="cuda").manual_seed(5)
    u8 = torch.empty((Kn, L, B), dtype=torch.uint8, device="cuda")
    for j in range(L):
        u8[:, j, :] = torch.randint(0, int(reg["npi_max"][j]) + 1, (Kn, B), dtype=torch.uint8, device="cuda", generator=g)
    prm = pack_params([dict(dt=1.0, beta=syn.BETA, gamma=syn.GAMMA, b=reg["b"][r], a=reg["a"][r], u_max=reg["npi_max"],
                            alpha_min=1e-8, alpha_max=100.0) for r in range(nR)], L)
    x0 = np.array([[(reg["N"][r] - 10) / reg["N"][r], 10 / reg["N"][r], syn.ALPHA0] for r in range(nR)])
    w = np.stack([np.repeat(reg["cost_weights"][r][None, :], Kn, axis=0) for r in range(nR)])
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    res = engine.rollout_cost(params_to_device(prm, "cuda:0"), dev(x0), u8, Kn, L, G=nS, B=B, want_traj=False,
                              want_cost=True, T_total=Kn, j0_prefix=dev(np.zeros(nR)), j1_prefix=dev(np.zeros(nR)), w=dev(w))
    J0, J1 = res["J0"].view(nR, nS), res["J1"].view(nR, nS)
    mask, iopt = engine.pareto(J0, J1)
    engine.sync()
    assert bool(torch.isfinite(J0).all()) and bool(torch.isfinite(J1).all()) and float(J0.min()) >= 0
    wmax = torch.from_numpy(np.array([np.mean(reg["cost_weights"][r] * reg["npi_max"]) for r in range(nR)])).cuda()
    assert bool((J1 <= wmax[:, None] * (1 + 1e-12)).all())
    m = mask.bool()
    assert bool(m.any(dim=1).all())
    inf = torch.full_like(J0, float("inf"))
    m2, _ = engine.pareto(torch.where(m, J0, inf), torch.where(m, J1, inf))
    engine.sync()
    assert bool(((m2.bool() & m) == m).all())
    o = orc()
    uh = u8.cpu().numpy()
    for bb in (0, 8191, 8192, 1_000_003, B - 1):
        r = bb // nS
        s, i, al = o.SIalpha_Controlled(uh[:, :, bb].T.astype(float), *x0[r], reg["npi_max"], 1e-8, 100.0, syn.GAMMA,
                                        reg["a"][r], reg["b"][r], syn.BETA, 0.0, 0.0, 0.0, Kn, 1.0)
        j0, j1 = o.NPICost((s * i) * al, uh[:, :, bb].T.astype(float), w[r].T)
        assert float(res["J0"][bb]) == j0 and float(res["J1"][bb]) == j1, bb


# ------------------------------------------------------------------------------ multi-GPU through the C ABI
@pytest.mark.parametrize("n_regions", [5, 2, 1])
def test_sweep_multi_gpu_one_host_call(engine, n_regions):
    """epi_sweep_multi: ONE blocking host-memory call shards the regions over the GPUs of the box
    (TrainPredictPrescribeNPI.m:93 is the region loop) and returns the same arrays as the single-GPU call,
    bit for bit -- including the trajectory-minor ones (noise in; u_fore, P_first out) that go through
    per-shard staging, ragged shards and shards with no region at all."""
    import torch
    from epidemicmodeling_b200.api import get_engine
    n_dev = torch.cuda.device_count()
    if n_dev < 2:
        pytest.skip("needs at least 2 GPUs")
    peers = [get_engine(d) for d in range(1, min(n_dev, 4))]
    inp, eps = cases.sweep_case(n_regions=n_regions, n_eps=7, T_hist=35, T_fore=14)
    S = wl.run_fixed_input(engine, inp)
    batch = wl.sweep_batch(inp, S)
    B, Tf = n_regions * eps.size, 14
    rng = np.random.default_rng(5)
    noise = rng.standard_normal((Tf, 3, B))
    nstd = np.array([r["setup3"]["noise_std"] for r in inp])
    kw = dict(noise=noise, noise_std=nstd, want_front=True, want_u_knee=True, want_u_fore=True, want_P_first=True)
    one = wl.run_sweep(engine, batch, eps, **kw)
    multi = wl.run_sweep(engine, batch, eps, peers=peers, **kw)
    for k in ("J0", "J1", "on_front", "I_opt", "u_knee", "u_fore", "P_first"):
        assert_bits(multi[k], one[k], f"multi-GPU {k} (n_regions={n_regions}, {1 + len(peers)} GPUs)")
    # the peers really worked: their contexts launched kernels (unless they had no region to take)
    if n_regions >= 2:
        assert peers[0].launch_count > 0
    # device arrays belong to one GPU: refused, not silently run on one device
    dbatch = wl.sweep_to_device(batch, eps, "cuda:0")
    with pytest.raises(ValueError):
        wl.run_sweep(engine, dbatch, eps, peers=peers)


# ------------------------------------------------------------------------------ round-2 robustness items
def test_device_mode_validates_params_without_draining_the_stream(engine):
    """EPI_MEM_DEVICE calls cannot read epi_model_params on the host: a one-CTA kernel checks them and the
    violation ('unknown observation type', SIAlphaModelEKF.m:57) surfaces at the next sync / call."""
    import torch
    from epidemicmodeling_b200.engine import params_to_device
    inp, b, x, nRep = _ekf3_replicate_batch(nR=2, nRep=8)
    B, T, L = x.shape[1], b["T"], b["L"]
    kw = dict(B=B, T=T, L=L, G=nRep, x_per_traj=True, r_mode=K.R_PERDAY, fixed_R=False, beta=b["beta"], gamma=b["gamma"],
              W=b["W"], outputs=("S_SMOOTH",))
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    args = [dev(b[k]) for k in ("u",)] + [dev(x), dev(b["R"]), dev(b["Q"]), dev(b["s_init"]), dev(b["Ps_init"]),
                                          dev(b["s_final"]), dev(b["Ps_final"])]
    bad = pack_params([dict(r["setup3"]["params"]) for r in inp], L)
    bad[1].obs_type = 7
    engine.ekf_eks(K.MODEL_SIALPHA, params_to_device(bad, "cuda:0"), *args, **kw)      # enqueued, no error yet
    with pytest.raises(K.EpiError) as e:
        engine.sync()
    assert e.value.code == K.ERR_OBS_TYPE and "unknown observation type" in e.value.msg
    engine.sync()                                                                       # reported once
    good = engine.ekf_eks(K.MODEL_SIALPHA, params_to_device(b["prm"], "cuda:0"), *args, **kw)
    engine.sync()
    host = engine.ekf_eks(K.MODEL_SIALPHA, b["prm"], b["u"], x, b["R"], b["Q"], b["s_init"], b["Ps_init"], b["s_final"],
                          b["Ps_final"], **kw)
    assert_bits(good["S_SMOOTH"].cpu().numpy(), host["S_SMOOTH"], "device mode after a reported violation")
    with pytest.raises(K.EpiError) as e:                                                # host mode: immediate
        engine.ekf_eks(K.MODEL_SIALPHA, bad, b["u"], x, b["R"], b["Q"], b["s_init"], b["Ps_init"], b["s_final"],
                       b["Ps_final"], **kw)
    assert e.value.code == K.ERR_OBS_TYPE


def test_monitor_window_limit_is_a_clear_argument_error(engine):
    """The innovation-monitor window lives in shared memory: a window that cannot fit is refused with a message
    that says so (the reference accepts any inv_monitor_len; 151 days for m = 3, 302 for m = 6)."""
    c = cases.ekf3_case(0, T_hist=40, T_fore=10)
    from epidemicmodeling_b200 import api
    a = lambda W: api.SIAlphaModelEKF(c["u"], c["x"], c["params"], c["s_init"], c["Ps_init"], c["s_final"], c["Ps_final"],
                                     c["w_bar"], c["v_bar"], c["Q_w"], c["R_v"], c["beta"], c["gamma"], W, 1)
    ok = a(151)
    o = orc()
    want = o.ekf_eks(o.SIALPHA, *ekf_args(dict(c, inv_monitor_len=151)))
    assert_bits(ok[4], want["S_SMOOTH"], "W = 151 S_SMOOTH"); assert_bits(ok[10], want["rho"], "W = 151 rho")
    with pytest.raises(K.EpiError) as e:
        a(152)
    assert e.value.code == K.ERR_ARG and "inv_monitor_len" in e.value.msg and "shared memory" in e.value.msg


def test_random_schedules_device_mode_on_a_fresh_engine():
    """Engine.random_schedules(device=True) as the FIRST call of a context: the kernel must run on torch's
    current stream (the output tensor and the temporary params tensor belong to it)."""
    import torch
    from epidemicmodeling_b200.engine import Engine
    nR, nS, Kn, L = 2, 33, 21, 12
    reg = syn.load_regions(nR)
    prm = pack_params([dict(dt=1.0, beta=syn.BETA, gamma=syn.GAMMA, b=reg["b"][r], a=reg["a"][r], u_max=reg["npi_max"],
                            u_min=np.zeros(L), alpha_min=1e-8, alpha_max=100.0) for r in range(nR)], L)
    fresh = Engine(0)
    try:
        side = torch.cuda.Stream()
        with torch.cuda.stream(side):
            ud = fresh.random_schedules(prm, nR * nS, Kn, L, nS, 99, device=True)
            got = ud.cpu().numpy()                     # ordered after the kernel on the same stream
        uh = fresh.random_schedules(prm, nR * nS, Kn, L, nS, 99)
        assert np.array_equal(got, uh)
    finally:
        fresh.close()


@pytest.mark.parametrize("shape", [(3, 15, 33, 21), (2, 7, 1, 9), (1, 4, 12, 0)])
def test_lane_group_kernels_are_bit_identical(engine, shape, monkeypatch):
    """csrc/ekf_rows.cu: six lanes per trajectory (one covariance row per lane) for the forward pass and the
    smoother recursion of the sweep -- the same bits as the one-thread kernels, full and lean, with ragged warps
    (5 trajectories per warp), and against the oracle."""
    nR, nE, Th, Tf = shape
    inp, eps = cases.sweep_case(n_regions=nR, n_eps=nE, T_hist=Th, T_fore=Tf)
    S = wl.run_fixed_input(engine, inp)
    batch = wl.sweep_batch(inp, S)
    kw = dict(want_front=True, want_u_fore=True, want_u_knee=True)
    monkeypatch.setenv("EPI_ROWS", "0")
    monkeypatch.setenv("EPI_FWD_SEGMENTS", "1")
    one = wl.run_sweep(engine, batch, eps, **kw)
    monkeypatch.setenv("EPI_ROWS", "1")
    rows = wl.run_sweep(engine, batch, eps, **kw)
    for k in ("J0", "J1", "on_front", "I_opt", "u_fore", "u_knee"):
        assert_bits(rows[k], one[k], f"lane-group {k} {shape}")
    lean = wl.run_sweep(engine, batch, eps, lean=True, **kw)
    for k in ("J0", "J1", "on_front", "I_opt", "u_fore", "u_knee"):
        assert_bits(lean[k], one[k], f"lane-group lean {k} {shape}")
    o = orc()
    for r, rin in enumerate(inp):
        s6 = rin["setup6"]
        Sr = S[:, :, r].T
        reg = o.SweepRegion(s6["params"], rin["T"], Th, rin["u_hist"], rin["x"], rin["R_v"], s6["s_init"],
                            s6["Ps_init"], s6["s_final"], s6["Ps_final"], s6["Q_w"], 1.0, 0.995, 21,
                            Sr[0, Th - 1], Sr[1, Th - 1], Sr[2, Th - 1], (Sr[0, :Th] * Sr[1, :Th]) * Sr[2, :Th], rin["weights"])
        j0, j1, m, io, _ = o.sweep_region(reg, eps)
        assert_bits(rows["J0"][r], j0, f"lane-group J0 vs oracle {shape}"); assert_bits(rows["J1"][r], j1, "J1 vs oracle")
    # the generic batched entry takes the same kernels when its tape is scratch (S_SMOOTH only)
    c = cases.ekf6_case(1, T_hist=30, T_fore=11)
    ee = np.array([1e-9, 1e-3, 0.05, 0.3, 0.7, 0.999, 0.2])
    T, L = c["u"].shape[1], 12
    cm = lambda P: np.ascontiguousarray(np.asarray(P).T).ravel()
    call = lambda: engine.ekf_eks(K.MODEL_OPTCTRL, pack_params([c["params"]], L), c["u"].T.copy(), c["x"], c["R_v"],
                                  cm(c["Q_w"]), c["s_init"], cm(c["Ps_init"]), c["s_final"], cm(c["Ps_final"]),
                                  B=ee.size, T=T, L=L, G=ee.size, epsilon=ee, r_mode=K.R_PERDAY, fixed_R=False,
                                  beta=1.0, gamma=0.995, W=21, outputs=("S_SMOOTH", "u_opt_smooth"))
    g_rows = call()
    monkeypatch.setenv("EPI_ROWS", "0")
    g_one = call()
    for k in ("S_SMOOTH", "u_opt_smooth"):
        assert_bits(g_rows[k], g_one[k], f"lane-group generic entry {k}")


@pytest.mark.parametrize("shape", [(3, 15, 33, 21), (2, 7, 1, 9), (1, 4, 12, 0), (2, 33, 20, 6)])
def test_two_warp_forward_is_bit_identical(engine, shape, monkeypatch):
    """csrc/ekf_pair.cu: the forward pass of small batches with two warps per 32-trajectory tile (each warp owns three
    rows of P, five shared-memory exchanges per day) -- the same bits as the one-warp kernel: full and lean sweeps,
    ragged tiles, days without observation, and the generic entry with a per-trajectory epsilon."""
    nR, nE, Th, Tf = shape
    inp, eps = cases.sweep_case(n_regions=nR, n_eps=nE, T_hist=Th, T_fore=Tf)
    S = wl.run_fixed_input(engine, inp)
    batch = wl.sweep_batch(inp, S)
    kw = dict(want_front=True, want_u_fore=True, want_u_knee=True)
    monkeypatch.setenv("EPI_PAIR", "0")
    monkeypatch.setenv("EPI_FWD_SEGMENTS", "1")
    one = wl.run_sweep(engine, batch, eps, want_P_first=True, **kw)
    monkeypatch.setenv("EPI_PAIR", "1")
    two = wl.run_sweep(engine, batch, eps, want_P_first=True, **kw)
    for k in ("J0", "J1", "on_front", "I_opt", "u_fore", "u_knee", "P_first"):
        assert_bits(two[k], one[k], f"two-warp forward {k} {shape}")
    lean = wl.run_sweep(engine, batch, eps, lean=True, **kw)
    for k in ("J0", "J1", "on_front", "I_opt", "u_fore", "u_knee"):
        assert_bits(lean[k], one[k], f"two-warp forward lean {k} {shape}")
    c = cases.ekf6_case(1, T_hist=30, T_fore=11)
    ee = np.array([1e-9, 1e-3, 0.05, 0.3, 0.7, 0.999, 0.2])
    T, L = c["u"].shape[1], 12
    cm = lambda P: np.ascontiguousarray(np.asarray(P).T).ravel()
    call = lambda: engine.ekf_eks(K.MODEL_OPTCTRL, pack_params([c["params"]], L), c["u"].T.copy(), c["x"], c["R_v"],
                                  cm(c["Q_w"]), c["s_init"], cm(c["Ps_init"]), c["s_final"], cm(c["Ps_final"]),
                                  B=ee.size, T=T, L=L, G=ee.size, epsilon=ee, r_mode=K.R_PERDAY, fixed_R=False,
                                  beta=1.0, gamma=0.995, W=21, outputs=("S_SMOOTH", "P_SMOOTH", "u_opt_smooth"), want_status=True)
    g_two = call()
    monkeypatch.setenv("EPI_PAIR", "0")
    g_one = call()
    for k in ("S_SMOOTH", "P_SMOOTH", "u_opt_smooth", "status"):
        assert_bits(g_two[k], g_one[k], f"two-warp forward, generic entry {k}")


@pytest.mark.parametrize("shape", [(3, 15, 60, 30), (2, 40, 50, 70), (1, 33, 100, 12)])
def test_piped_schedule_is_bit_identical(engine, shape, monkeypatch):
    """Small sweeps run forward || gains (csrc/ekf_forward.cu `ekf_forward_piped_kernel`, capi.cu): the forward pass as
    4-warp CTAs that keep their SM, the gains of a time chunk on a second stream behind a stream wait on the forward
    pass's progress counter.  Same bits as the one-stream schedule: full and lean sweeps, ragged tiles, 1..9 chunks."""
    nR, nE, Th, Tf = shape
    inp, eps = cases.sweep_case(n_regions=nR, n_eps=nE, T_hist=Th, T_fore=Tf)
    S = wl.run_fixed_input(engine, inp)
    batch = wl.sweep_batch(inp, S)
    kw = dict(want_front=True, want_u_fore=True, want_u_knee=True)
    monkeypatch.setenv("EPI_PIPE", "0")
    one = wl.run_sweep(engine, batch, eps, want_P_first=True, **kw)
    for chunks in ("8", "1", "9"):
        monkeypatch.setenv("EPI_PIPE", "1")
        monkeypatch.setenv("EPI_PIPE_CHUNKS", chunks)
        two = wl.run_sweep(engine, batch, eps, want_P_first=True, **kw)
        for k in ("J0", "J1", "on_front", "I_opt", "u_fore", "u_knee", "P_first"):
            assert_bits(two[k], one[k], f"piped schedule ({chunks} chunks) {k} {shape}")
        lean = wl.run_sweep(engine, batch, eps, lean=True, **kw)
        for k in ("J0", "J1", "on_front", "I_opt", "u_fore", "u_knee"):
            assert_bits(lean[k], one[k], f"piped schedule ({chunks} chunks) lean {k} {shape}")


@pytest.mark.parametrize("shape", [(3, 15, 33, 21), (2, 7, 1, 9), (1, 4, 12, 0), (2, 40, 50, 70)])
def test_staged_backward_is_bit_identical(engine, shape, monkeypatch):
    """Small sweeps run the smoother recursion with the tape pages streamed by TMA bulk copies into a shared-memory ring
    (csrc/eks_backward.cu, STAGED): the same bits as the register-prefetch kernel for ring depths 1, 2, 5 and 16 (deeper
    than the run is long), full and lean sweeps, ragged tiles, a run without history."""
    nR, nE, Th, Tf = shape
    inp, eps = cases.sweep_case(n_regions=nR, n_eps=nE, T_hist=Th, T_fore=Tf)
    S = wl.run_fixed_input(engine, inp)
    batch = wl.sweep_batch(inp, S)
    kw = dict(want_front=True, want_u_fore=True, want_u_knee=True)
    monkeypatch.setenv("EPI_BWD_STAGES", "0")
    one = wl.run_sweep(engine, batch, eps, **kw)
    for stages in ("1", "2", "5", "16"):
        monkeypatch.setenv("EPI_BWD_STAGES", stages)
        two = wl.run_sweep(engine, batch, eps, **kw)
        for k in ("J0", "J1", "on_front", "I_opt", "u_fore", "u_knee"):
            assert_bits(two[k], one[k], f"staged backward ({stages} stages) {k} {shape}")
        lean = wl.run_sweep(engine, batch, eps, lean=True, **kw)
        for k in ("J0", "J1", "on_front", "I_opt", "u_fore", "u_knee"):
            assert_bits(lean[k], one[k], f"staged backward ({stages} stages) lean {k} {shape}")


@pytest.mark.parametrize("shape", [(2, 9, 40, 30), (3, 40, 70, 12)])
def test_sweep_with_missing_history_npis(engine, shape, monkeypatch):
    """NPIs missing INSIDE the history (NaN entries of u on given days): the state update replaces them by the bang-bang
    rule (SIAlphaModelEKFOptControlled.m:41-50), so those days are evaluated per trajectory in the forward pass, the
    gains, the backward recursion and the history cost of the rollout (no per-group shortcut).  Against the oracle on the
    same inputs, bit for bit, under every schedule of the sweep (one stream / piped, register-prefetch / staged
    backward, segmented forward) and in device mode."""
    nR, nE, Th, Tf = shape
    inp, eps = cases.sweep_case(n_regions=nR, n_eps=nE, T_hist=Th, T_fore=Tf)
    rng = np.random.default_rng(11)
    for rin in inp:
        uh = np.array(rin["u_hist"], dtype=np.float64, copy=True)
        for _ in range(6):
            uh[rng.integers(0, uh.shape[0]), rng.integers(0, Th)] = np.nan
        uh[:, Th // 2] = np.nan            # a day with every NPI missing
        rin["u_hist"] = uh
    S = wl.run_fixed_input(engine, inp)
    batch = wl.sweep_batch(inp, S)
    o = orc()
    ref = []
    for r, rin in enumerate(inp):
        s6 = rin["setup6"]
        reg = o.SweepRegion(s6["params"], rin["T"], Th, rin["u_hist"], rin["x"], rin["R_v"], s6["s_init"],
                            s6["Ps_init"], s6["s_final"], s6["Ps_final"], s6["Q_w"], 1.0, 0.995, 21,
                            batch["x0"][r, 0], batch["x0"][r, 1], batch["x0"][r, 2], batch["newcases_hist"][r],
                            rin["weights"])
        ref.append(o.sweep_region(reg, eps, want_u=True))
    kw = dict(want_front=True, want_u_fore=True)
    for env in ({}, {"EPI_PIPE": "0", "EPI_BWD_STAGES": "0"}, {"EPI_PIPE": "1", "EPI_BWD_STAGES": "3"},
                {"EPI_PIPE": "0", "EPI_FWD_SEGMENTS": "3"}):
        for k in ("EPI_PIPE", "EPI_BWD_STAGES", "EPI_FWD_SEGMENTS"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        res = wl.run_sweep(engine, batch, eps, **kw)
        uf = res["u_fore"].reshape(Tf, 12, nR, nE)
        for r in range(nR):
            j0, j1, m, io, u = ref[r]
            assert_bits(res["J0"][r], j0, f"J0 region {r} {env}")
            assert_bits(res["J1"][r], j1, f"J1 region {r} {env}")
            assert np.array_equal(res["on_front"][r].astype(bool), m) and res["I_opt"][r] == io
            assert_bits(np.transpose(uf[:, :, r, :], (2, 1, 0)), u, f"u_fore region {r} {env}")
    dbatch = wl.sweep_to_device(batch, eps, "cuda:0")
    dres = wl.run_sweep(engine, dbatch, None, want_front=True)
    assert_bits(dres["J0"].cpu().numpy(), res["J0"], "device mode J0")
    assert_bits(dres["J1"].cpu().numpy(), res["J1"], "device mode J1")


def test_pareto_pruned_path_special_values(engine):
    """Large point sets are pruned by J0 buckets before the sort (csrc/pareto_sorted.cu): the mask must stay the O(n^2)
    predicate's for ties, NaN / +-Inf / +-0 in either coordinate, a degenerate J0 range (every point in one bucket), an
    overflowing range, negative costs and a set whose front is everything."""
    rng = np.random.default_rng(21)
    n = 9500
    o = orc()
    sets0, sets1 = [], []
    a0, a1 = np.round(rng.random(n), 2), np.round(rng.random(n), 2)                     # heavy ties
    a0[:7] = [np.nan, np.inf, -np.inf, 0.0, -0.0, np.inf, np.nan]
    a1[3:12] = [np.nan, 0.0, -0.0, np.inf, -np.inf, np.nan, 5.0, -5.0, 0.5]
    sets0.append(a0); sets1.append(a1)
    sets0.append(np.full(n, 3.25)); sets1.append(rng.random(n))                           # degenerate range
    b0 = rng.standard_normal(n) * 1e307; b0[0], b0[1] = 1.7e308, -1.7e308                 # hi - lo overflows
    sets0.append(b0); sets1.append(rng.standard_normal(n))
    c0 = np.sort(rng.random(n)); sets0.append(c0); sets1.append(1.0 - c0)                 # everything on the front
    sets0.append(-rng.random(n) * 1e-300); sets1.append(-rng.random(n))                   # tiny negative J0, negative J1
    d0 = np.round(10.0 ** rng.uniform(-300, 300, n), 0) * rng.choice([-1.0, 1.0], n)     # 600 decades, both signs, ties at 0
    d0[:6] = [5e-324, -5e-324, 0.0, -0.0, 2.2250738585072014e-308, -2.2250738585072014e-308]
    sets0.append(d0); sets1.append(np.round(rng.standard_normal(n), 1))
    e0 = 10.0 ** rng.uniform(-8, -2, n); sets0.append(e0); sets1.append(3.0 * rng.random(n) ** 2 + 1e-3 / e0 ** 0.25)  # config-5-like
    J0, J1 = np.stack(sets0), np.stack(sets1)
    m, io = engine.pareto(J0, J1)
    for r in range(J0.shape[0]):
        wm, wi = o.pareto(J0[r], J1[r])
        assert np.array_equal(m[r].astype(bool), wm), ("pruned path mask", r, int(m[r].sum()), int(wm.sum()))
        assert io[r] == wi, ("knee", r)
    os.environ["EPI_PARETO_PRUNE"] = "0"        # the whole-set sort must agree as well
    try:
        m0, io0 = engine.pareto(J0, J1)
    finally:
        del os.environ["EPI_PARETO_PRUNE"]
    assert np.array_equal(m0, m) and np.array_equal(io0, io)
    # a realistic cloud at config 5's size: idempotence and front size
    J0b = rng.random((2, 100_000)); J1b = 1.0 / (J0b + 0.05) + 0.1 * rng.random((2, 100_000))
    mb, _ = engine.pareto(J0b, J1b)
    for r in range(2):
        keep = mb[r].astype(bool)
        wm, _ = o.pareto(J0b[r, keep], J1b[r, keep])
        assert wm.all() and 10 < keep.sum() < 5000
        # no dropped front point: every point is dominated by some kept point or kept itself (checked on a sample)
        k0, k1 = J0b[r, keep], J1b[r, keep]
        for i in rng.integers(0, 100_000, 300):
            assert keep[i] or np.any((k0 < J0b[r, i]) & (k1 < J1b[r, i]))
