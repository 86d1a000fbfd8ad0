"""matlab/epi_mex.cpp EXECUTED: the gateway is built against tests/mexhost (an in-memory implementation of the
MEX C API subset it uses) and driven out of process with MATLAB-shaped arrays.  `not gpu`: it links, dispatches
and raises the reference's error identifiers before any device work.  `gpu`: every command returns the same bits
as the ctypes path (epidemicmodeling_b200.api / workloads) on the parity cases -- the marshalling (column-major
pages, struct arrays, 1-based knee index, Q/R shape dispatch) is what is under test."""
import os

import numpy as np
import pytest

import cases
import mexhost

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def mparams(p):
    """The reference's params struct as MATLAB holds it (column vectors)."""
    out = {}
    for k, v in p.items():
        out[k] = v if isinstance(v, str) else (np.asarray(v, dtype=np.float64).reshape(-1, 1) if np.ndim(v) else float(v))
    return out


def ekf_margs(model, c):
    col = lambda v: np.asarray(v, dtype=np.float64).reshape(-1, 1)
    R = np.asarray(c["R_v"], dtype=np.float64)
    return (float(model), c["u"], np.asarray(c["x"]).reshape(1, -1), mparams(c["params"]), col(c["s_init"]), c["Ps_init"],
            col(c["s_final"]), c["Ps_final"], col(c["w_bar"]), float(c["v_bar"]), c["Q_w"],
            R.reshape(1, -1) if R.ndim == 1 else R, float(c["beta"]), float(c["gamma"]), float(c["inv_monitor_len"]),
            float(c["order"]))


def test_gateway_links_dispatches_and_reports_reference_errors():
    mexhost.build()
    with pytest.raises(mexhost.MexError) as e:
        mexhost.call("no_such_command")
    assert e.value.id == "epi:arg" and "unknown command" in e.value.msg
    with pytest.raises(mexhost.MexError) as e:
        mexhost.call("ekf_eks", 0.0)
    assert e.value.id == "epi:arg"
    c = cases.ekf3_case()
    bad = dict(c, params=dict(c["params"], obs_type="CASES"))
    with pytest.raises(mexhost.MexError) as e:
        mexhost.call("ekf_eks", *ekf_margs(0, bad), nlhs=11)
    assert e.value.id == "epi:obsType" and e.value.msg == "unknown observation type"     # SIAlphaModelEKF.m:57
    badQ = dict(c, Q_w=np.ones((2, 5)))
    with pytest.raises(mexhost.MexError) as e:
        mexhost.call("ekf_eks", *ekf_margs(0, badQ), nlhs=11)
    assert e.value.id == "epi:covShape"                                                   # GenericExtendedKalmanFilter.m:75
    with pytest.raises(mexhost.MexError) as e:
        mexhost.call("ekf_eks_masked", *ekf_margs(0, c))                                  # num_forecast_days missing
    assert e.value.id == "epi:arg"


def _bits(a, b, what):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    assert np.array_equal(a, b, equal_nan=True), what


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["ekf3_perday", "ekf3_adaptive", "ekf3_totalcases", "ekf3_flipped", "ekf6_optctrl", "ekf6_flipped",
                                  "legacy_tools"])
def test_gateway_ekf_eks_matches_ctypes_path(name):
    from epidemicmodeling_b200 import api
    table = {"ekf3_perday": (0, api.SIAlphaModelEKF, cases.ekf3_case(variant="perday")),
             "ekf3_adaptive": (0, api.SIAlphaModelEKF, cases.ekf3_case(variant="adaptive")),
             "ekf3_totalcases": (0, api.SIAlphaModelEKF, cases.ekf3_case(variant="totalcases")),
             "ekf3_flipped": (1, api.SIAlphaModelBackwardEKF, cases.ekf3_case(variant="backward")),
             "ekf6_optctrl": (2, api.SIAlphaModelEKFOptControlled, cases.ekf6_case()),
             "ekf6_flipped": (3, api.SIAlphaModelBackwardEKFOptControlled, cases.ekf6_case(backward=True)),
             "legacy_tools": (4, api.NewCaseEKFEstimatorWithOptimalNPI, cases.legacy_case())}
    model, fn, c = table[name]
    want = fn(c["u"], c["x"], c["params"], c["s_init"], c["Ps_init"], c["s_final"], c["Ps_final"], c["w_bar"], c["v_bar"],
              c["Q_w"], c["R_v"], c["beta"], c["gamma"], c["inv_monitor_len"], c["order"])
    got = mexhost.call("ekf_eks", *ekf_margs(model, c), nlhs=11)
    if name.startswith("legacy"):            # the gateway returns the 11-slot generic order; the legacy shim drops slot 2
        got = [got[0]] + got[2:]
    assert len(got) == len(want)
    for k, (g, w) in enumerate(zip(got, want)):
        w = np.asarray(w)
        _bits(np.asarray(g).reshape(w.shape) if np.asarray(g).size == w.size else g, w, f"{name} output {k}")


@pytest.mark.gpu
def test_gateway_sweep_and_masked_batch_match_ctypes_path():
    import torch
    from epidemicmodeling_b200 import api, workloads as wl
    eng = api.get_engine(0)
    inp, eps = cases.sweep_case(n_regions=3, n_eps=9, T_hist=40, T_fore=15)
    S = wl.run_fixed_input(eng, inp)
    batch = wl.sweep_batch(inp, S)
    want = wl.run_sweep(eng, batch, eps, want_front=True, want_u_knee=True)
    nR, T, Th, L = 3, batch["T"], batch["T_hist"], batch["L"]
    F = lambda a, axes: np.transpose(np.asarray(a), axes)            # per-region row-major pages -> MATLAB dims
    margs = ([mparams(r["setup6"]["params"]) for r in inp], eps.reshape(1, -1),
             F(batch["u"], (2, 1, 0)),                                 # [nR,T,L] -> L x T x nR
             batch["x"].T, batch["R"].T, batch["s_init"].T,
             F(batch["Ps_init"].reshape(nR, 6, 6), (2, 1, 0)), batch["s_final"].T,
             F(batch["Ps_final"].reshape(nR, 6, 6), (2, 1, 0)), F(batch["Q"].reshape(nR, 6, 6), (2, 1, 0)),
             float(batch["beta_ekf"]), float(batch["gamma_ekf"]), float(batch["W"]), batch["x0"].T,
             batch["newcases_hist"].T, F(batch["weights"], (2, 1, 0)), 0.0)
    J0, J1, front, iopt, uknee = mexhost.call("sweep", *margs, nlhs=5)
    _bits(J0.T, want["J0"], "sweep J0"); _bits(J1.T, want["J1"], "sweep J1")
    assert np.array_equal(front.T, want["on_front"].astype(bool))
    assert np.array_equal(iopt.ravel().astype(int) - 1, want["I_opt"])                    # 1-based in MATLAB
    _bits(np.transpose(uknee, (2, 1, 0)), want["u_knee"], "sweep u_knee")
    # lean flag and the multi-GPU path of the same command (all GPUs of the box; 1 GPU: same call)
    for extra in ((1.0,), (0.0, 0.0)):
        a = list(margs)
        a[16:] = extra
        J0b, J1b = mexhost.call("sweep", *a, nlhs=2)
        _bits(J0b, J0, f"sweep variant {extra} J0"); _bits(J1b, J1, f"sweep variant {extra} J1")
    if torch.cuda.device_count() >= 2:
        a = list(margs) + [2.0]
        J0c, J1c, frontc, ioptc, ukc = mexhost.call("sweep", *a, nlhs=5)
        _bits(J0c, J0, "sweep n_gpus=2 J0"); _bits(ukc, uknee, "sweep n_gpus=2 u_knee")
        assert np.array_equal(ioptc, iopt) and np.array_equal(frontc, front)

    # masked-horizon batch == workloads.forecast_quality == the reference's loop of single calls
    rin, nf = inp[1], 6
    c = cases.ekf3_case(1, T_hist=40, T_fore=15)
    xfull = np.nan_to_num(c["x"], nan=float(np.nanmean(c["x"])))                          # a fully observed series
    c = dict(c, x=xfull)
    SP, SS = mexhost.call("ekf_eks_masked", *ekf_margs(0, c), float(nf), nlhs=2)
    assert SP.shape == (3, c["u"].shape[1], nf)
    for start in range(1, nf + 1):
        xp = xfull.copy()
        xp[len(xp) - start:] = np.nan                                                      # ForecastQualityAssessment.m:385
        one = api.SIAlphaModelEKF(c["u"], xp, c["params"], c["s_init"], c["Ps_init"], c["s_final"], c["Ps_final"],
                                  c["w_bar"], c["v_bar"], c["Q_w"], c["R_v"], c["beta"], c["gamma"], c["inv_monitor_len"], 1)
        _bits(SP[:, :, start - 1], one[3], f"masked S_PLUS start={start}")
        _bits(SS[:, :, start - 1], one[4], f"masked S_SMOOTH start={start}")


@pytest.mark.gpu
def test_gateway_small_commands_match_ctypes_path():
    from epidemicmodeling_b200 import api
    kw = cases.seirp_scenarios()["A"]
    K = int(np.floor(kw["T"] / kw["dt"] + 0.5))
    rates = np.stack([kw[k] for k in ("alpha_e", "alpha_i", "kappa", "rho", "beta", "mu", "gamma")])   # 7 x K
    ic = np.array([kw[k] for k in ("s0", "e0", "i0", "r0", "p0")]).reshape(5, 1)
    got = mexhost.call("seirp", 0.0, rates.T, ic, float(K), kw["dt"], np.zeros((6, 1)), nlhs=5)      # K x 7 column-major
    want = api.SEIRP(**kw)
    for g, w in zip(got, want):
        _bits(np.asarray(g).reshape(w.shape), w, "seirp")
    rng = np.random.default_rng(3)
    j0, j1 = rng.random(40), rng.random(40)
    front, iopt = mexhost.call("pareto", j0, j1, nlhs=2)
    m, io = api.get_engine(0).pareto(j0.reshape(1, -1), j1.reshape(1, -1))
    assert np.array_equal(front.ravel(), np.asarray(m).ravel().astype(bool)) and int(iopt[0, 0]) - 1 == int(np.asarray(io)[0])
    rc = cases.rollout_case(noisy=False)
    w = np.outer(np.linspace(0.5, 1.5, 12), np.ones(rc["K"]))
    nc = rng.random(rc["K"])
    J = mexhost.call("npicost", nc.reshape(1, -1), rc["u"], w, nlhs=2)
    Jw = api.NPICost(nc, rc["u"], w)
    assert float(J[0][0, 0]) == Jw[0] and float(J[1][0, 0]) == Jw[1]
