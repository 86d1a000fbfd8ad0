"""Host-side mirror of the reference's MATLAB function signatures.

The reference (alphanumericslab/EpidemicModeling) is interpreted MATLAB with no
FFI layer; its "operator API" for the hot path is the set of function signatures
below.  Each function here keeps the reference's name, argument
order, argument meaning, output order, output shapes and error behaviour, and
runs on the B200 through the C ABI (include/epi_b200.h) as a batch of one
trajectory.  The batched forms live in engine.Engine.  matlab/*.m + matlab/epi_mex.cpp
are the same mirror for a MATLAB/Octave host (uncompilable here: no mex.h).

Arrays use MATLAB shapes: u is L x T, x is 1 x T, S_* are m x T, P_* are
m x m x T, K_GAIN is m x 1 x T, innovations 1 x T, rho T x 1.

Differences forced by the boundary (all documented in DESIGN.md):
  * SIalpha_Controlled draws randn in-line in the reference; here the draws are
    the explicit keyword `noise` (3 x K, call order s, i, alpha); None = zeros.
  * GenericExtendedKalmanFilter accepts arbitrary function handles in the
    reference; only the four known handle sets can run on the device, named by
    the HANDLES_* constants.  Anything else raises NotImplementedError.
"""
import atexit

import numpy as np

from . import _capi as K
from .engine import Engine, pack_params

_engines = {}


def get_engine(device=0):
    """Process-wide engine of GPU `device` (one context per host thread / GPU)."""
    device = int(device)
    if device not in _engines:
        if not _engines:
            atexit.register(_close_engine)
        _engines[device] = Engine(device)
    return _engines[device]


def _close_engine():
    for dev in list(_engines):
        _engines.pop(dev).close()


def _K_of(T, dt):
    return int(np.floor(T / dt + 0.5))  # MATLAB round(), SEIRP.m:13


def _row(v, K_):
    v = np.asarray(v, dtype=np.float64).ravel()
    if v.size == 1:
        return np.full(K_, v[0])
    if v.size < K_ - 1:
        raise IndexError("Index exceeds the number of array elements")  # MATLAB's error
    out = np.zeros(K_)
    n = min(K_, v.size)
    out[:n] = v[:n]  # only samples 1..K-1 are read (SEIRP.m:26)
    return out


def SEIRP(alpha_e, alpha_i, kappa, rho, beta, mu, gamma, s0, e0, i0, r0, p0, T, dt):
    """[s,e,i,r,p] = SEIRP(...)  -- Tools/SEIRP.m:1"""
    K_ = _K_of(T, dt)
    rates = np.stack([_row(v, K_) for v in (alpha_e, alpha_i, kappa, rho, beta, mu, gamma)])
    ic = np.array([[s0], [e0], [i0], [r0], [p0]], dtype=np.float64)
    out = get_engine().seirp(rates, ic, K_, dt, rate_mode=K.RATES_SHARED_SERIES)
    return tuple(out[f, :, 0].reshape(1, K_) for f in range(5))


def SEIRPSaturatedResource(alpha_e, alpha_i, kappa, rho, gamma, s0, e0, i0, r0, p0, T, dt,
                           beta_0, beta_s, mu_0, mu_s, sigma, i_0):
    """[s,e,i,r,p] = SEIRPSaturatedResource(...)  -- Tools/SEIRPSaturatedResource.m:1"""
    K_ = _K_of(T, dt)
    z = np.zeros(K_)
    rates = np.stack([_row(alpha_e, K_), _row(alpha_i, K_), _row(kappa, K_), _row(rho, K_), z, z,
                      _row(gamma, K_)])
    ic = np.array([[s0], [e0], [i0], [r0], [p0]], dtype=np.float64)
    sat = dict(beta_0=beta_0, beta_s=beta_s, mu_0=mu_0, mu_s=mu_s, sigma=sigma, i_0=i_0)
    out = get_engine().seirp(rates, ic, K_, dt, rate_mode=K.RATES_SHARED_SERIES, saturated=sat)
    return tuple(out[f, :, 0].reshape(1, K_) for f in range(5))


def SIalpha_Controlled(u, s0, i0, alpha0, u_max, alpha_min, alpha_max, gamma, a, b, beta,
                       s_noise_std, i_noise_std, alpha_noise_std, K_, dt, noise=None):
    """[s,i,alpha] = SIalpha_Controlled(...)  -- Tools/SIalpha_Controlled.m:1"""
    u = np.asarray(u, dtype=np.float64)
    L = u.shape[0]
    prm = pack_params([dict(dt=dt, beta=beta, gamma=gamma, b=b, alpha_min=alpha_min,
                            alpha_max=alpha_max, a=a, u_max=u_max)], L)
    r = get_engine().rollout_cost(
        prm, np.array([s0, i0, alpha0], dtype=np.float64), np.ascontiguousarray(u[:, :K_].T)[:, :, None],
        K_, L, G=1, B=1, noise_std=np.array([s_noise_std, i_noise_std, alpha_noise_std], dtype=np.float64),
        noise=None if noise is None else np.ascontiguousarray(np.asarray(noise, dtype=np.float64).T)[:, :, None],
        want_traj=True)
    return r["s"].reshape(1, K_), r["i"].reshape(1, K_), r["alpha"].reshape(1, K_)


def SI_Controlled(alpha, beta, s0, i0, K_, dt):
    """[s,i] = SI_Controlled(alpha, beta, s0, i0, K, dt)  -- Tools/SI_Controlled.m:1"""
    al = np.asarray(alpha, dtype=np.float64).ravel()[:K_].reshape(K_, 1)
    s, i = get_engine().si_controlled(al, np.array([beta], dtype=np.float64),
                                      np.array([s0], dtype=np.float64),
                                      np.array([i0], dtype=np.float64), K_, dt)
    return s.reshape(1, K_), i.reshape(1, K_)


def Rt_ExpFitEKF(x, s_init, params, w_bar, v_bar, Ps_init, Q_w, R_v, beta, gamma, inv_monitor_len, order):
    """[S_MINUS, S_PLUS, P_MINUS, P_PLUS, K_GAIN, S_SMOOTH, P_SMOOTH, innovations, rho] =
    Rt_ExpFitEKF(x, s_init, params, w_bar, v_bar, Ps_init, Q_w, R_v, beta, gamma, inv_monitor_len, order)
    -- Tools/Rt_ExpFitEKF.m:1 (x is 1 x T; outputs in the reference's shapes, rho squeezed to T)."""
    if order not in (1, 2):
        raise ValueError("Undefined order")                      # Rt_ExpFitEKF.m:47,75
    x = np.asarray(x, dtype=np.float64)
    if x.ndim == 2 and x.shape[0] != 1:
        raise ValueError("Rt_ExpFitEKF: x must be 1 x T (one observation)")
    x = x.ravel()
    T = x.size
    cm = lambda P: np.ascontiguousarray(np.asarray(P, dtype=np.float64).T).ravel()   # column-major page
    r = get_engine().rt_expfit(x.reshape(T, 1), np.asarray(s_init, dtype=np.float64).reshape(2, 1),
                               np.asarray(params, dtype=np.float64).ravel()[:3], np.asarray(w_bar, dtype=np.float64).ravel()[:2],
                               cm(Ps_init), cm(Q_w), np.asarray(R_v, dtype=np.float64).ravel()[:1], T=T, v_bar=float(v_bar),
                               beta=float(beta), gamma=float(gamma), W=int(inv_monitor_len), order=int(order))
    P = lambda a: np.transpose(a[:, :, 0].reshape(T, 2, 2), (2, 1, 0)).copy()   # [T][col][row] -> row, col, T
    return (r["S_MINUS"][:, :, 0].T.copy(), r["S_PLUS"][:, :, 0].T.copy(), P(r["P_MINUS"]), P(r["P_PLUS"]),
            r["K_GAIN"][:, :, 0].T.reshape(2, 1, T).copy(), r["S_SMOOTH"][:, :, 0].T.copy(), P(r["P_SMOOTH"]),
            r["innovations"][:, 0].reshape(1, T), r["rho"][:, 0].copy())


def NPICost(newcases, inputs, weights):
    """[J0,J1] = NPICost(newcases, inputs, weights)  -- Tools/NPICost.m:1"""
    inp = np.asarray(inputs, dtype=np.float64)
    L, T = inp.shape
    wt = np.broadcast_to(np.asarray(weights, dtype=np.float64), (L, T))
    nc = np.asarray(newcases, dtype=np.float64).ravel()
    J0, J1 = get_engine().npicost(nc.reshape(T, 1), np.ascontiguousarray(inp.T)[:, :, None],
                                  np.ascontiguousarray(wt.T), T, L)
    return float(J0[0]), float(J1[0])


def ParetoFront(J0, J1):
    """Pareto mask and knee index (1-based, like MATLAB's I_opt) of
    Tools/TrainPredictPrescribeNPI.m:624-633."""
    J0 = np.asarray(J0, dtype=np.float64).reshape(1, -1)
    J1 = np.asarray(J1, dtype=np.float64).reshape(1, -1)
    mask, iopt = get_engine().pareto(J0, J1)
    return mask[0].astype(bool), int(iopt[0]) + 1


# --- EKF / EKS ---------------------------------------------------------------------
HANDLES_SIALPHA = "SIAlphaModelEKF"
HANDLES_SIALPHA_BACKWARD = "SIAlphaModelBackwardEKF"
HANDLES_OPTCTRL = "SIAlphaModelEKFOptControlled"
HANDLES_OPTCTRL_BACKWARD = "SIAlphaModelBackwardEKFOptControlled"
_HANDLE_MODEL = {HANDLES_SIALPHA: K.MODEL_SIALPHA, HANDLES_SIALPHA_BACKWARD: K.MODEL_SIALPHA_FLIPPED,
                 HANDLES_OPTCTRL: K.MODEL_OPTCTRL, HANDLES_OPTCTRL_BACKWARD: K.MODEL_OPTCTRL_FLIPPED}


def _classify_QR(Q_w, R_v, T, m):
    """Q/R shape dispatch of Tools/GenericExtendedKalmanFilter.m:64-91."""
    Q = np.asarray(Q_w, dtype=np.float64)
    Q2 = np.atleast_2d(Q)
    if Q.ndim == 3 and Q.shape[0] == Q.shape[1]:
        if Q.shape[2] != T:
            raise ValueError("Process noise covariance noise mismatch")
        q_mode, Qf = K.Q_PERDAY_FULL, np.ascontiguousarray(np.transpose(Q, (2, 1, 0))).ravel()
    elif Q.ndim <= 2 and Q2.shape[0] == Q2.shape[1]:
        Qm = np.eye(m) * Q2[0, 0] if Q2.shape[0] == 1 else Q2  # scalar: B*q*B' with B = I
        if Qm.shape[0] != m:
            raise ValueError("Process noise covariance noise mismatch")
        q_mode, Qf = K.Q_CONST, np.ascontiguousarray(Qm.T).ravel()
    elif Q.ndim <= 2 and min(Q2.shape) == 1 and Q.size == T:
        q_mode, Qf = K.Q_PERDAY_SCALAR, np.ascontiguousarray(Q).ravel()
    else:
        raise ValueError("Process noise covariance noise mismatch")  # :75
    R = np.asarray(R_v, dtype=np.float64)
    R2 = np.atleast_2d(R)
    if R.ndim == 3 and R.shape[0] == R.shape[1] == 1 and R.shape[2] == T:
        r_mode, fixed_R, Rf = K.R_PERDAY, 1, np.ascontiguousarray(R).ravel()
    elif R.ndim <= 2 and R2.shape == (1, 1):
        r_mode, fixed_R, Rf = K.R_CONST, 1, R2.ravel().copy()
    elif R.ndim <= 2 and min(R2.shape) == 1 and R.size == T:
        r_mode, fixed_R, Rf = K.R_PERDAY, 0, np.ascontiguousarray(R).ravel()
    else:
        raise ValueError("Observation noise covariance noise mismatch")  # :90
    return q_mode, Qf, r_mode, fixed_R, Rf


def _ekf_call(model, u, x, params, s_init, Ps_init, s_final, Ps_final, w_bar, v_bar, Q_w, R_v,
              beta, gamma, inv_monitor_len, order):
    u = np.asarray(u, dtype=np.float64)
    L, T = u.shape
    m = 6 if model >= K.MODEL_OPTCTRL else 3
    x = np.asarray(x, dtype=np.float64).ravel()
    if x.size != T:
        raise ValueError("x must have T entries")
    if order not in (1, 2):
        raise ValueError("Undefined order")  # GenericExtendedKalmanFilter.m:111
    legacy = model >= K.MODEL_LEGACY_TOOLS
    if legacy:
        Qm = np.asarray(Q_w, dtype=np.float64)
        Qm = np.eye(m) * Qm.ravel()[0] if Qm.size == 1 else Qm
        q_mode, Qf = K.Q_CONST, np.ascontiguousarray(Qm.T).ravel()
        r_mode, fixed_R, Rf = K.R_CONST, 1, np.asarray(R_v, dtype=np.float64).ravel()[:1].copy()
    else:
        q_mode, Qf, r_mode, fixed_R, Rf = _classify_QR(Q_w, R_v, T, m)
    prm = pack_params([params], L)
    cm = lambda P: np.ascontiguousarray(np.asarray(P, dtype=np.float64).reshape(m, m).T).ravel()
    o = get_engine().ekf_eks(
        model, prm, np.ascontiguousarray(u.T), x, Rf, Qf,
        np.asarray(s_init, dtype=np.float64).ravel(), cm(Ps_init),
        np.asarray(s_final, dtype=np.float64).ravel(), cm(Ps_final),
        B=1, T=T, L=L, G=1, r_mode=r_mode, fixed_R=bool(fixed_R), q_mode=q_mode,
        v_bar=float(np.asarray(v_bar).ravel()[0]), beta=beta, gamma=gamma, W=int(inv_monitor_len),
        order=int(order))
    res = {k: o[k][:, :, 0].T.copy() for k in ("u_opt", "S_MINUS", "S_PLUS", "S_SMOOTH")}
    if not legacy:
        res["u_opt_smooth"] = o["u_opt_smooth"][:, :, 0].T.copy()
    for k in ("P_MINUS", "P_PLUS", "P_SMOOTH"):
        res[k] = np.transpose(o[k][:, :, 0].reshape(T, m, m), (2, 1, 0)).copy()  # [t][col][row] -> row,col,t
    res["K_GAIN"] = o["K_GAIN"][:, :, 0].T.reshape(m, 1, T).copy()
    res["innovations"] = o["innovations"][:, 0].reshape(1, T).copy()
    res["rho"] = o["rho"][:, 0].reshape(T, 1).copy()
    return res


_GEN_ORDER = ("u_opt", "u_opt_smooth", "S_MINUS", "S_PLUS", "S_SMOOTH", "P_MINUS", "P_PLUS",
              "P_SMOOTH", "K_GAIN", "innovations", "rho")


def GenericExtendedKalmanFilter(u, x, handles, params, s_init, Ps_init, s_final, Ps_final, w_bar,
                                v_bar, Q_w, R_v, beta, gamma, inv_monitor_len, order):
    """Tools/GenericExtendedKalmanFilter.m:1 for the four known handle sets."""
    if handles not in _HANDLE_MODEL:
        raise NotImplementedError(
            "only the reference's four handle sets (HANDLES_*) run on the device; arbitrary "
            "function handles must stay on the MATLAB implementation")
    model = _HANDLE_MODEL[handles]
    if handles in (HANDLES_SIALPHA_BACKWARD, HANDLES_OPTCTRL_BACKWARD):
        # the *handles* of the backward wrappers are the sign-reversed dynamics only;
        # the time flip lives in the wrapper function, so undo the model's built-in flip
        raise NotImplementedError("call SIAlphaModelBackwardEKF / ...OptControlled for the backward models")
    r = _ekf_call(model, u, x, params, s_init, Ps_init, s_final, Ps_final, w_bar, v_bar, Q_w, R_v,
                  beta, gamma, inv_monitor_len, order)
    return tuple(r[k] for k in _GEN_ORDER)


def SIAlphaModelEKF(u, x, params, s_init, Ps_init, s_final, Ps_final, w_bar, v_bar, Q_w, R_v,
                    beta, gamma, inv_monitor_len, order):
    """Tools/SIAlphaModelEKF.m:1 (11 outputs)."""
    r = _ekf_call(K.MODEL_SIALPHA, u, x, params, s_init, Ps_init, s_final, Ps_final, w_bar, v_bar,
                  Q_w, R_v, beta, gamma, inv_monitor_len, order)
    return tuple(r[k] for k in _GEN_ORDER)


def SIAlphaModelBackwardEKF(u, x, params, s_init, Ps_init, s_final, Ps_final, w_bar, v_bar, Q_w,
                            R_v, beta, gamma, inv_monitor_len, order):
    """Tools/SIAlphaModelBackwardEKF.m:1 (time-reversed model on flipped data)."""
    r = _ekf_call(K.MODEL_SIALPHA_FLIPPED, u, x, params, s_init, Ps_init, s_final, Ps_final, w_bar,
                  v_bar, Q_w, R_v, beta, gamma, inv_monitor_len, order)
    return tuple(r[k] for k in _GEN_ORDER)


def SIAlphaModelEKFOptControlled(u, x, params, s_init, Ps_init, s_final, Ps_final, w_bar, v_bar,
                                 Q_w, R_v, beta, gamma, inv_monitor_len, order):
    """Tools/SIAlphaModelEKFOptControlled.m:1 (6-state state+costate model)."""
    r = _ekf_call(K.MODEL_OPTCTRL, u, x, params, s_init, Ps_init, s_final, Ps_final, w_bar, v_bar,
                  Q_w, R_v, beta, gamma, inv_monitor_len, order)
    return tuple(r[k] for k in _GEN_ORDER)


def SIAlphaModelBackwardEKFOptControlled(u, x, params, s_init, Ps_init, s_final, Ps_final, w_bar,
                                         v_bar, Q_w, R_v, beta, gamma, inv_monitor_len, order):
    """Tools/SIAlphaModelBackwardEKFOptControlled.m:1."""
    r = _ekf_call(K.MODEL_OPTCTRL_FLIPPED, u, x, params, s_init, Ps_init, s_final, Ps_final, w_bar,
                  v_bar, Q_w, R_v, beta, gamma, inv_monitor_len, order)
    return tuple(r[k] for k in _GEN_ORDER)


def NewCaseEKFEstimatorWithOptimalNPI(u, x, params, s_init, Ps_init, s_final, Ps_final, w_bar,
                                      v_bar, Q_w, R_v, beta, gamma, inv_monitor_len, order,
                                      variant="tools"):
    """Legacy monolithic 6-state EKF/EKS (10 outputs).
    variant="tools":   Tools/NewCaseEKFEstimatorWithOptimalNPI.m:1 output order
        [u_opt, S_MINUS, S_PLUS, S_SMOOTH, P_MINUS, P_PLUS, P_SMOOTH, K_GAIN, innovations, rho]
    variant="codegen": MatlabCodeGenerator/NewCaseEKFEstimatorWithOptimalNPI.m:1 order
        [u_opt, S_MINUS, S_PLUS, P_MINUS, P_PLUS, K_GAIN, S_SMOOTH, P_SMOOTH, innovations, rho]
        with the codegen callbacks (identity ObsHardMargins, NEWCASES only)."""
    if variant not in ("tools", "codegen"):
        raise ValueError("variant must be 'tools' or 'codegen'")
    model = K.MODEL_LEGACY_TOOLS if variant == "tools" else K.MODEL_LEGACY_CODEGEN
    r = _ekf_call(model, u, x, params, s_init, Ps_init, s_final, Ps_final, w_bar, v_bar, Q_w, R_v,
                  beta, gamma, inv_monitor_len, order)
    if variant == "tools":
        order_ = ("u_opt", "S_MINUS", "S_PLUS", "S_SMOOTH", "P_MINUS", "P_PLUS", "P_SMOOTH",
                  "K_GAIN", "innovations", "rho")
    else:
        order_ = ("u_opt", "S_MINUS", "S_PLUS", "P_MINUS", "P_PLUS", "K_GAIN", "S_SMOOTH",
                  "P_SMOOTH", "innovations", "rho")
    return tuple(r[k] for k in order_)
