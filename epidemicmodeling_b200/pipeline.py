"""CSV -> prescription CSV without MATLAB: the per-region chain of
Tools/TrainPredictPrescribeNPI.m with the training/regression rounds replaced by the reference's
own trained (a, b) per region (xprize-sample-data/prescription_trained_params_*.mat, committed as
data/regions_nonnegls.npz):

    read_oxcgrt (:62-90)  ->  Engine.preprocess (:121-128, :162-187, :200-201, :240)
    ->  3-state EKF/EKS on the historic NPIs (:360-382)  ->  fused optimal-NPI sweep over the
    epsilon grid (:415-495)  ->  Pareto front + knee (:624-633)  ->  write_prescriptions.

Everything numeric runs through the C ABI on the device; this module only assembles the inputs
the way the reference's script does."""
import numpy as np

from . import _capi as K
from . import synthetic as syn
from . import workloads as wl
from . import xprize_io as xio
from .engine import pack_params


def setup_from_preprocessed(pre, b, N, a, bb, npi_max, weights_row, T_hist, T_fore):
    """One region's `inputs` record (the structure synthetic.sweep_inputs builds) from preprocessed
    data: x = NewCasesSmoothedNormalized followed by T_fore missing days (:458), R_v per day with the
    last historic value held over the horizon, the EKF set-ups of :202-239 and :423-458."""
    I0 = float(pre["I0"][b])
    reg = dict(N=np.array([N]), a=np.array([a]), b=np.array([bb]), npi_max=npi_max,
               cost_weights=np.array([weights_row]), names=np.array(["r"]))
    setup3 = syn.ekf3_setup(reg, 0)
    # ekf3_setup uses the synthetic I0; re-derive the I0-dependent pieces from the data (:201,:229-236)
    s_std, i_std, a_std = 10.0 * I0 / N, 30.0 * I0 / N, 1e-2
    setup3 = dict(setup3)
    setup3["Q_w"] = np.diag(np.array([s_std, i_std, a_std]) ** 2)
    setup3["Ps_init"] = np.diag(np.array([10 * s_std, 10 * i_std, 10 * a_std]) ** 2)
    setup3["s_init"] = np.array([(N - I0) / N, I0 / N, syn.ALPHA0])
    setup3["noise_std"] = (s_std, i_std, a_std)
    setup6 = syn.ekf6_setup(reg, 0, setup3)
    T = T_hist + T_fore
    x = np.concatenate([pre["normalized"][:T_hist, b], np.full(T_fore, np.nan)])
    Rv = pre["R_v"][:T_hist, b] + 1e-30
    R = np.concatenate([Rv, np.full(T_fore, Rv.mean())])                         # :357 horizon = mean(R_v)
    u_hist = np.ascontiguousarray(pre["ip_filled"][:T_hist, :, b].T)            # L x T_hist
    u_fixed = np.concatenate([u_hist, np.repeat(u_hist[:, -1:], T_fore, axis=1)], axis=1)
    weights = np.repeat(np.asarray(weights_row, dtype=np.float64)[:, None], T, axis=1)
    return dict(T=T, T_hist=T_hist, u_hist=u_hist, u_fixed=u_fixed, x=x, R_v=R, setup3=setup3, setup6=setup6,
                weights=weights)


def train_regions(engine, pre, pops, npi_max, T_hist, n_regression_days, max_alt=100):
    """Rounds 1-5 of the driver for every region at once (NONNEGATIVELS branch):
    :247 3-state EKF/EKS with zero inputs  ->  :250-251,264-278 regress the smoothed alpha on
    (NPI_MAXES - InterventionPlans) over the last n_regression_days  ->  :291-302 EKF/EKS with the
    real inputs and the fitted (a, b)  ->  :305-306,326-339 second regression.  Returns dict(a1, b1, a2, b2)."""
    B = pre["I0"].shape[0]
    zeros12 = np.zeros(npi_max.size)

    def smoothed_alpha(a, b, real_inputs):
        inputs = [setup_from_preprocessed(pre, r, float(pops[r]), a[:, r], float(b[r]), npi_max, zeros12, T_hist, 0)
                  for r in range(B)]
        if not real_inputs:
            for rin in inputs:
                rin["u_fixed"] = np.zeros_like(rin["u_fixed"])        # :203 the first round assumes zero inputs
        return wl.run_fixed_input(engine, inputs)[:, 2, :]             # S_SMOOTH(3, :) per region  [T, B]

    n = int(n_regression_days)
    X = np.ascontiguousarray(npi_max[None, :, None] - pre["ip_filled"][T_hist - n:T_hist])   # [n, L, B]
    out = {}
    a, b = np.zeros((npi_max.size, B)), np.zeros(B)
    for rnd, real in ((1, False), (2, True)):
        y = np.ascontiguousarray(smoothed_alpha(a, b, real)[T_hist - n:T_hist])
        a, b, k = engine.nnls_affine(X, y, max_alt=max_alt)
        out[f"a{rnd}"], out[f"b{rnd}"], out[f"alt{rnd}"] = a, b, k
    return out


def prescribe_from_csv(engine, data_file, start_date, end_date, T_fore, eps, regions, out_file=None,
                       forecast_dates=None, n_indexes=1):
    """regions: dict geo id -> dict(N, a, b, weights[12]).  Returns dict(ids, J0, J1, on_front, I_opt,
    u_knee [B, T_fore, 12]) and, with out_file, writes the knee schedules as prescription index 0."""
    ids, dates, cc, _, ip = xio.read_oxcgrt(data_file, start_date, end_date, geo_ids=set(regions))
    B, T_hist = len(ids), cc.shape[0]
    pop = np.array([regions[g]["N"] for g in ids], dtype=np.float64)
    pre = engine.preprocess(cc, pop, ip)
    inputs = [setup_from_preprocessed(pre, b, regions[g]["N"], regions[g]["a"], regions[g]["b"], xio.NPI_MAXES,
                                      regions[g]["weights"], T_hist, T_fore) for b, g in enumerate(ids)]
    S = wl.run_fixed_input(engine, inputs)
    batch = wl.sweep_batch(inputs, S)
    res = wl.run_sweep(engine, batch, eps, want_front=True, want_u_knee=True)
    out = dict(ids=ids, dates=dates, pre=pre, J0=res["J0"], J1=res["J1"], on_front=res["on_front"],
               I_opt=res["I_opt"], u_knee=res["u_knee"])
    if out_file:
        fd = forecast_dates or [f"day+{k + 1}" for k in range(T_fore)]
        xio.write_prescriptions(out_file, ids, fd, [res["u_knee"]])
    return out
