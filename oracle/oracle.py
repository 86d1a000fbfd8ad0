"""ctypes front-end of the CPU ORACLE (test infrastructure, NOT product code).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.  It loads oracle/libepi_oracle.so (built
by `make -C oracle`) and exposes the reference's MATLAB signatures with
MATLAB-shaped numpy arrays (column-major semantics: u is L x T, S_* m x T,
P_* m x m x T).  See oracle/epi_oracle.h for the arithmetic contract and the
"parity unpinned" note.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libepi_oracle.so")
LMAX = 12

SIALPHA, SIALPHA_FLIPPED, OPTCTRL, OPTCTRL_FLIPPED, LEGACY_TOOLS, LEGACY_CODEGEN = range(6)
OBS = {"NEWCASES": 0, "TOTALCASES": 1}

_dp = C.POINTER(C.c_double)


class OrcParams(C.Structure):
    _fields_ = [("dt", C.c_double), ("beta", C.c_double), ("gamma", C.c_double), ("b", C.c_double),
                ("alpha_min", C.c_double), ("alpha_max", C.c_double), ("s_min", C.c_double),
                ("i_min", C.c_double), ("epsilon", C.c_double), ("sigma", C.c_double),
                ("a", C.c_double * LMAX), ("u_min", C.c_double * LMAX),
                ("u_max", C.c_double * LMAX), ("w", C.c_double * LMAX),
                ("L", C.c_int), ("obs_type", C.c_int)]


class OrcSweepRegion(C.Structure):
    _fields_ = [("prm", OrcParams), ("T", C.c_int), ("T_hist", C.c_int),
                ("u_hist", _dp), ("x", _dp), ("R", _dp),
                ("s_init", C.c_double * 6), ("Ps_init", C.c_double * 36),
                ("s_final", C.c_double * 6), ("Ps_final", C.c_double * 36),
                ("Q", C.c_double * 36), ("beta_ekf", C.c_double), ("gamma_ekf", C.c_double),
                ("W", C.c_int), ("s_h", C.c_double), ("i_h", C.c_double), ("alpha_h", C.c_double),
                ("newcases_hist", _dp), ("weights", _dp), ("noise_std", C.c_double * 3),
                ("noise", _dp)]


def build(force=False):
    if force or not os.path.exists(_SO) or \
            os.path.getmtime(_SO) < os.path.getmtime(os.path.join(_HERE, "epi_oracle.c")):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.orc_pinv_sym.restype = C.c_int
        _lib.orc_ekf_eks.restype = C.c_int
        _lib.orc_num_threads.restype = C.c_int
    return _lib


def _f(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def _p(a):
    return a.ctypes.data_as(_dp) if a is not None else None


def _vecL(v, L, fill=np.nan):
    out = np.full(LMAX, fill, dtype=np.float64)
    if v is not None:
        v = np.asarray(v, dtype=np.float64)
        if v.ndim == 2 and v.shape[1] > 1:
            # `params.w` given as an L x T matrix: phi(kk) linear-indexes column 1
            # (SIAlphaModelEKFOptControlled.m:49-52)
            v = v[:, 0]
        v = v.ravel()
        if v.size == 1:
            out[:L] = v[0]
        else:
            out[:L] = v[:L]
    return out


def make_params(params, L):
    """dict (MATLAB `params` struct fields) -> OrcParams."""
    g = lambda k, d=np.nan: float(np.asarray(params.get(k, d)).ravel()[0])
    p = OrcParams()
    p.dt, p.beta, p.gamma, p.b = g("dt"), g("beta"), g("gamma"), g("b", 0.0)
    p.alpha_min, p.alpha_max = g("alpha_min"), g("alpha_max")
    p.s_min, p.i_min = g("s_min", 0.0), g("i_min", 0.0)
    p.epsilon, p.sigma = g("epsilon"), g("sigma")
    for name in ("a", "u_min", "u_max", "w"):
        arr = _vecL(params.get(name), L)
        getattr(p, name)[:] = list(arr)
    p.L = L
    ot = params.get("obs_type", "NEWCASES")
    if ot not in OBS:
        raise ValueError("unknown observation type")  # SIAlphaModelEKF.m:57
    p.obs_type = OBS[ot]
    return p


def SEIRP(alpha_e, alpha_i, kappa, rho, beta, mu, gamma, s0, e0, i0, r0, p0, T, dt):
    K = int(np.floor(T / dt + 0.5))  # MATLAB round for positive values (SEIRP.m:13)
    rates = [_f(np.broadcast_to(np.asarray(v, dtype=np.float64).ravel(), (K,)) if np.size(v) == 1
                else v).ravel() for v in (alpha_e, alpha_i, kappa, rho, beta, mu, gamma)]
    out = [np.zeros(K) for _ in range(5)]
    lib().orc_seirp(*[_p(r) for r in rates], C.c_int(1), *[C.c_double(v) for v in (s0, e0, i0, r0, p0)],
                    C.c_int(K), C.c_double(dt), *[_p(o) for o in out])
    return tuple(o.reshape(1, K) for o in out)


def SEIRPSaturatedResource(alpha_e, alpha_i, kappa, rho, gamma, s0, e0, i0, r0, p0, T, dt,
                           beta_0, beta_s, mu_0, mu_s, sigma, i_0):
    K = int(np.floor(T / dt + 0.5))
    rates = [_f(np.broadcast_to(np.asarray(v, dtype=np.float64).ravel(), (K,)) if np.size(v) == 1
                else v).ravel() for v in (alpha_e, alpha_i, kappa, rho, gamma)]
    out = [np.zeros(K) for _ in range(5)]
    lib().orc_seirp_saturated(*[_p(r) for r in rates], C.c_int(1),
                              *[C.c_double(v) for v in (s0, e0, i0, r0, p0)], C.c_int(K),
                              C.c_double(dt), *[C.c_double(v) for v in
                                                (beta_0, beta_s, mu_0, mu_s, sigma, i_0)],
                              *[_p(o) for o in out])
    return tuple(o.reshape(1, K) for o in out)


def SIalpha_Controlled(u, s0, i0, alpha0, u_max, alpha_min, alpha_max, gamma, a, b, beta,
                       s_noise_std, i_noise_std, alpha_noise_std, K, dt, noise=None):
    """`noise` (3 x K standard normals, randn call order s,i,alpha) replaces the
    reference's in-line randn draws (SIalpha_Controlled.m:25-27); None = zeros."""
    u = np.asarray(u, dtype=np.float64)
    L = u.shape[0]
    uc = _f(u.T)  # column-major L x K
    nz = _f(np.asarray(noise).T) if noise is not None else None
    s, i, al = np.zeros(K), np.zeros(K), np.zeros(K)
    lib().orc_sialpha_controlled(_p(uc), C.c_int(L), C.c_double(s0), C.c_double(i0),
                                 C.c_double(alpha0), _p(_f(u_max).ravel()), C.c_double(alpha_min),
                                 C.c_double(alpha_max), C.c_double(gamma), _p(_f(a).ravel()),
                                 C.c_double(b), C.c_double(beta), C.c_double(s_noise_std),
                                 C.c_double(i_noise_std), C.c_double(alpha_noise_std),
                                 C.c_int(K), C.c_double(dt), _p(nz), _p(s), _p(i), _p(al))
    return s.reshape(1, K), i.reshape(1, K), al.reshape(1, K)


def SI_Controlled(alpha, beta, s0, i0, K, dt):
    al = _f(alpha).ravel()
    s, i = np.zeros(K), np.zeros(K)
    lib().orc_si_controlled(_p(al), C.c_double(beta), C.c_double(s0), C.c_double(i0), C.c_int(K),
                            C.c_double(dt), _p(s), _p(i))
    return s.reshape(1, K), i.reshape(1, K)


def NPICost(newcases, inputs, weights):
    nc = _f(newcases).ravel()
    inp = np.asarray(inputs, dtype=np.float64)
    L, T = inp.shape
    wt = np.broadcast_to(np.asarray(weights, dtype=np.float64), (L, T))
    J0, J1 = C.c_double(), C.c_double()
    lib().orc_npicost(_p(nc), C.c_int(T), _p(_f(inp.T)), _p(_f(wt.T)), C.c_int(L),
                      C.byref(J0), C.byref(J1))
    return J0.value, J1.value


def philox4x32_10(ctr, key):
    """Philox4x32-10 block: ctr (4 uint32), key (2 uint32) -> 4 uint32."""
    c = (C.c_uint32 * 4)(*[int(v) & 0xFFFFFFFF for v in ctr])
    k = (C.c_uint32 * 2)(*[int(v) & 0xFFFFFFFF for v in key])
    o = (C.c_uint32 * 4)()
    lib().orc_philox4x32_10(c, k, o)
    return [int(v) for v in o]


def random_schedule(seed, region, scenario, n_scenarios, L, K, u_min, u_max):
    """TrainPredictPrescribeNPI.m:499-510 on the Philox stream: uint8 [L, K] (MATLAB shape)."""
    u = np.zeros((K, L), dtype=np.uint8)
    lib().orc_random_schedule(C.c_uint64(int(seed)), C.c_uint32(int(region)), C.c_uint32(int(scenario)),
                              C.c_int(int(n_scenarios)), C.c_int(int(L)), C.c_int(int(K)),
                              _p(_f(u_min)), _p(_f(u_max)), u.ctypes.data_as(C.c_void_p))
    return u.T.copy()


def Rt_ExpFitEKF(x, s_init, params, w_bar, v_bar, Ps_init, Q_w, R_v, beta, gamma, inv_monitor_len, order):
    """Tools/Rt_ExpFitEKF.m:1 -- returns the 9 outputs in the reference's order and shapes."""
    x = _f(x).ravel()
    T = x.size
    o = dict(S_MINUS=np.zeros((T, 2)), S_PLUS=np.zeros((T, 2)), P_MINUS=np.zeros((T, 4)), P_PLUS=np.zeros((T, 4)),
             K_GAIN=np.zeros((T, 2)), S_SMOOTH=np.zeros((T, 2)), P_SMOOTH=np.zeros((T, 4)),
             innovations=np.zeros(T), rho=np.zeros(T))
    lib().orc_rt_expfit_ekf.restype = C.c_int
    rc = lib().orc_rt_expfit_ekf(
        _p(x), C.c_int(T), _p(_f(s_init).ravel()), _p(_f(params).ravel()), _p(_f(w_bar).ravel()),
        C.c_double(float(v_bar)), _p(_f(np.asarray(Ps_init, dtype=np.float64)).ravel()),
        _p(_f(np.asarray(Q_w, dtype=np.float64)).ravel()), C.c_double(float(np.asarray(R_v).ravel()[0])),
        C.c_double(float(beta)), C.c_double(float(gamma)), C.c_int(int(inv_monitor_len)), C.c_int(int(order)),
        *[_p(o[k]) for k in ("S_MINUS", "S_PLUS", "P_MINUS", "P_PLUS", "K_GAIN", "S_SMOOTH", "P_SMOOTH",
                             "innovations", "rho")])
    if rc == -2:
        raise ValueError("Undefined order")
    P = lambda a: np.transpose(a.reshape(T, 2, 2), (1, 2, 0)).copy()     # [T][row][col] -> 2 x 2 x T
    return (o["S_MINUS"].T.copy(), o["S_PLUS"].T.copy(), P(o["P_MINUS"]), P(o["P_PLUS"]),
            o["K_GAIN"].T.reshape(2, 1, T).copy(), o["S_SMOOTH"].T.copy(), P(o["P_SMOOTH"]),
            o["innovations"].reshape(1, T), o["rho"].reshape(T))


def preprocess_region(cc, ip, N, W=7, n_first=7, min_cases=1.0):
    """TrainPredictPrescribeNPI.m:121-128,162-187,200-201,240 for one region.  cc [T], ip [T, L].
    Returns dict(refined, smoothed, zerolag, normalized, confirmed_norm, R_v [T]; ip [T, L]; I0)."""
    cc = _f(cc).ravel()
    T = cc.size
    ip = np.array(ip, dtype=np.float64, order="C").reshape(T, -1).copy()
    o = {k: np.zeros(T) for k in ("refined", "smoothed", "zerolag", "normalized", "confirmed_norm", "R_v")}
    I0 = C.c_double()
    lib().orc_preprocess_region.restype = C.c_int
    rc = lib().orc_preprocess_region(_p(cc), C.c_int(T), C.c_double(float(N)), C.c_int(int(W)), C.c_int(int(n_first)),
                                     C.c_double(float(min_cases)), _p(ip), C.c_int(ip.shape[1]),
                                     *[_p(o[k]) for k in ("refined", "smoothed", "zerolag", "normalized",
                                                          "confirmed_norm", "R_v")], C.byref(I0))
    if rc:
        raise ValueError("Insufficient data" if rc == -1 else "Data length must be larger than 3 times the filter order")
    o["ip"], o["I0"] = ip, I0.value
    return o


def lsqnonneg(X, y):
    X = np.ascontiguousarray(X, dtype=np.float64); y = _f(y).ravel()
    a = np.zeros(X.shape[1])
    lib().orc_lsqnonneg(_p(X), _p(y), C.c_int(X.shape[0]), C.c_int(X.shape[1]), _p(a))
    return a


def nnls_affine(X, y, max_alt=100):
    """TrainPredictPrescribeNPI.m:264-278: (a >= 0, b, accepted alternations)."""
    X = np.ascontiguousarray(X, dtype=np.float64); y = _f(y).ravel()
    a = np.zeros(X.shape[1]); b = C.c_double()
    lib().orc_nnls_affine.restype = C.c_int
    k = lib().orc_nnls_affine(_p(X), _p(y), C.c_int(X.shape[0]), C.c_int(X.shape[1]), C.c_int(int(max_alt)), _p(a), C.byref(b))
    return a, b.value, k


def pareto(J0, J1):
    J0, J1 = _f(J0).ravel(), _f(J1).ravel()
    n = J0.size
    mask = np.zeros(n, dtype=np.uint8)
    iopt = C.c_int()
    lib().orc_pareto(_p(J0), _p(J1), C.c_int(n), mask.ctypes.data_as(C.POINTER(C.c_ubyte)),
                     C.byref(iopt))
    return mask.astype(bool), iopt.value


def pinv_sym(A):
    A = np.asarray(A, dtype=np.float64)
    m = A.shape[0]
    X = np.zeros((m, m))
    rank = C.c_int()
    sweeps = lib().orc_pinv_sym(_p(_f(A.T)), C.c_int(m), _p(X), C.byref(rank))
    return X.T.copy(), rank.value, sweeps


def mrdivide(B, A):
    A = np.asarray(A, dtype=np.float64)
    m = A.shape[0]
    X = np.zeros((m, m))
    lib().orc_mrdivide(_p(_f(np.asarray(B, dtype=np.float64).T)), _p(_f(A.T)), C.c_int(m), _p(X))
    return X.T.copy()


def classify_QR(Q_w, R_v, T, m):
    """Q/R shape dispatch of GenericExtendedKalmanFilter.m:64-91.
    Returns (q_mode, Q_flat, r_mode, fixed_R, R_flat)."""
    Q = np.asarray(Q_w, dtype=np.float64)
    Q2 = np.atleast_2d(Q)
    if Q2.shape[0] == Q2.shape[1]:
        if Q.ndim == 3:  # repmat of an n x n x T array: page k of the result is page k
            q_mode, Qf = 2, _f(np.transpose(Q, (2, 1, 0))).ravel()
        elif Q2.shape[0] == 1:  # scalar -> B*q*B' = q*I
            q_mode, Qf = 0, _f(np.eye(m) * Q2[0, 0]).ravel()
        else:
            q_mode, Qf = 0, _f(Q2.T).ravel()
    elif Q.ndim <= 2 and min(Q2.shape) == 1 and Q.size == T:
        q_mode, Qf = 1, _f(Q).ravel()
    else:
        raise ValueError("Process noise covariance noise mismatch")  # :75
    R = np.asarray(R_v, dtype=np.float64)
    R2 = np.atleast_2d(R)
    if R2.shape[0] == R2.shape[1]:
        if R.ndim == 3:
            r_mode, fixed_R, Rf = 1, 1, _f(R).ravel()
        else:
            r_mode, fixed_R, Rf = 0, 1, _f(R2).ravel()[:1]
    elif R.ndim <= 2 and min(R2.shape) == 1 and R.size == T:
        r_mode, fixed_R, Rf = 1, 0, _f(R).ravel()
    else:
        raise ValueError("Observation noise covariance noise mismatch")  # :90
    return q_mode, Qf, r_mode, fixed_R, Rf


def ekf_eks(model, u, x, params, s_init, Ps_init, s_final, Ps_final, w_bar, v_bar, Q_w, R_v,
            beta, gamma, inv_monitor_len, order):
    """Returns a dict of all 11 outputs with MATLAB shapes."""
    u = np.asarray(u, dtype=np.float64)
    L, T = u.shape
    m = 6 if model >= OPTCTRL else 3
    x = _f(x).ravel()
    assert x.size == T
    prm = make_params(params, L)
    q_mode, Qf, r_mode, fixed_R, Rf = classify_QR(Q_w, R_v, T, m)
    out = dict(u_opt=np.zeros((T, L)), u_opt_smooth=np.zeros((T, L)),
               S_MINUS=np.zeros((T, m)), S_PLUS=np.zeros((T, m)), S_SMOOTH=np.zeros((T, m)),
               P_MINUS=np.zeros((T, m, m)), P_PLUS=np.zeros((T, m, m)), P_SMOOTH=np.zeros((T, m, m)),
               K_GAIN=np.zeros((T, m)), innovations=np.zeros(T), rho=np.zeros(T))
    rc = lib().orc_ekf_eks(
        C.c_int(model), C.byref(prm), C.c_int(T), _p(_f(u.T)), _p(x),
        _p(_f(s_init).ravel()), _p(_f(np.asarray(Ps_init, dtype=np.float64).T)),
        _p(_f(s_final).ravel()), _p(_f(np.asarray(Ps_final, dtype=np.float64).T)),
        C.c_double(float(np.asarray(v_bar).ravel()[0])), C.c_int(q_mode), _p(Qf), C.c_int(r_mode),
        C.c_int(fixed_R), _p(Rf), C.c_double(beta), C.c_double(gamma), C.c_int(inv_monitor_len),
        C.c_int(order), *[_p(out[k]) for k in
                          ("u_opt", "u_opt_smooth", "S_MINUS", "S_PLUS", "S_SMOOTH", "P_MINUS",
                           "P_PLUS", "P_SMOOTH", "K_GAIN", "innovations", "rho")])
    if rc == -2:
        raise ValueError("Undefined order")
    if rc:
        raise ValueError(f"oracle ekf_eks failed rc={rc}")
    # to MATLAB shapes
    res = {k: out[k].T.copy() for k in ("u_opt", "u_opt_smooth", "S_MINUS", "S_PLUS", "S_SMOOTH")}
    for k in ("P_MINUS", "P_PLUS", "P_SMOOTH"):
        res[k] = np.transpose(out[k], (2, 1, 0)).copy()  # [T][col][row] -> row, col, T
    res["K_GAIN"] = out["K_GAIN"].T.reshape(m, 1, T).copy()
    res["innovations"] = out["innovations"].reshape(1, T)
    res["rho"] = out["rho"].reshape(T, 1)
    return res


class SweepRegion:
    """Keeps the numpy buffers of one OrcSweepRegion alive."""

    def __init__(self, params, T, T_hist, u_hist, x, R, s_init, Ps_init, s_final, Ps_final, Q,
                 beta_ekf, gamma_ekf, W, s_h, i_h, alpha_h, newcases_hist, weights,
                 noise_std=(0.0, 0.0, 0.0), noise=None):
        u_hist = np.asarray(u_hist, dtype=np.float64)
        L = u_hist.shape[0]
        self.L, self.T, self.T_hist = L, T, T_hist
        self._bufs = dict(u=_f(u_hist.T), x=_f(x).ravel(), R=_f(R).ravel(),
                          nc=_f(newcases_hist).ravel(),
                          w=_f(np.broadcast_to(np.asarray(weights, dtype=np.float64), (L, T)).T),
                          nz=_f(noise) if noise is not None else None)
        r = OrcSweepRegion()
        r.prm = make_params(params, L)
        r.T, r.T_hist = T, T_hist
        r.u_hist, r.x, r.R = _p(self._bufs["u"]), _p(self._bufs["x"]), _p(self._bufs["R"])
        r.s_init[:] = list(_f(s_init).ravel())
        r.Ps_init[:] = list(_f(np.asarray(Ps_init).T).ravel())
        r.s_final[:] = list(_f(s_final).ravel())
        r.Ps_final[:] = list(_f(np.asarray(Ps_final).T).ravel())
        r.Q[:] = list(_f(np.asarray(Q).T).ravel())
        r.beta_ekf, r.gamma_ekf, r.W = beta_ekf, gamma_ekf, W
        r.s_h, r.i_h, r.alpha_h = s_h, i_h, alpha_h
        r.newcases_hist, r.weights = _p(self._bufs["nc"]), _p(self._bufs["w"])
        r.noise_std[:] = list(noise_std)
        r.noise = _p(self._bufs["nz"])
        self.c = r


def sweep_batch(regions, eps, n_threads=0, want_front=True):
    """OpenMP batch driver; returns J0, J1 [R x E], front mask, I_opt."""
    eps = _f(eps).ravel()
    R, E = len(regions), eps.size
    arr = (OrcSweepRegion * R)(*[r.c for r in regions])
    J0, J1 = np.zeros((R, E)), np.zeros((R, E))
    mask = np.zeros((R, E), dtype=np.uint8)
    iopt = np.zeros(R, dtype=np.int32)
    lib().orc_sweep_batch(arr, C.c_int(R), _p(eps), C.c_int(E), _p(J0), _p(J1),
                          mask.ctypes.data_as(C.POINTER(C.c_ubyte)) if want_front else None,
                          iopt.ctypes.data_as(C.POINTER(C.c_int)) if want_front else None,
                          C.c_int(n_threads))
    return J0, J1, mask.astype(bool), iopt


def sweep_region(region, eps, want_u=False):
    eps = _f(eps).ravel()
    E = eps.size
    L, Tf = region.L, region.T - region.T_hist
    J0, J1 = np.zeros(E), np.zeros(E)
    mask = np.zeros(E, dtype=np.uint8)
    iopt = C.c_int()
    uf = np.zeros((E, Tf, L)) if want_u else None
    lib().orc_sweep_region_run(C.byref(region.c), _p(eps), C.c_int(E), _p(J0), _p(J1), _p(uf),
                               mask.ctypes.data_as(C.POINTER(C.c_ubyte)), C.byref(iopt))
    return J0, J1, mask.astype(bool), iopt.value, (np.transpose(uf, (0, 2, 1)) if want_u else None)


def num_threads():
    return lib().orc_num_threads()
