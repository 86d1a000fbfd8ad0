"""Batched engine: a thin, typed front-end of the C ABI (include/epi_b200.h).

Arrays are either numpy float64 (EPI_MEM_HOST: the blocking MATLAB-style call,
H2D/D2H inside) or torch CUDA float64 tensors (EPI_MEM_DEVICE: kernels are
enqueued on the engine's stream; call `sync()`), always C-contiguous in the
trajectory-minor layout of the header: X[t][field][b].

torch is plumbing only (device memory, streams, torch.distributed); every
computation happens inside libepi_b200.so.  No CPU fallback exists.
"""
import ctypes as C
import sys

import numpy as np

from . import _capi as K

try:  # torch is optional for the host-memory path
    import torch
except Exception:  # pragma: no cover
    torch = None


def _is_torch(x):
    return torch is not None and isinstance(x, torch.Tensor)


def pack_params(param_dicts, L):
    """list of MATLAB-`params`-like dicts -> ctypes array of epi_model_params."""
    arr = (K.ModelParams * len(param_dicts))()
    for p, d in zip(arr, param_dicts):
        g = lambda k, dflt=np.nan: float(np.asarray(d.get(k, dflt), dtype=np.float64).ravel()[0])
        p.dt, p.beta, p.gamma, p.b = g("dt"), g("beta"), g("gamma"), g("b", 0.0)
        p.alpha_min, p.alpha_max = g("alpha_min"), g("alpha_max")
        p.s_min, p.i_min = g("s_min", 0.0), g("i_min", 0.0)
        p.epsilon, p.sigma = g("epsilon"), g("sigma")
        for name in ("a", "u_min", "u_max", "w"):
            v = d.get(name)
            full = np.full(K.LMAX, np.nan)
            if v is not None:
                v = np.asarray(v, dtype=np.float64)
                if v.ndim == 2 and v.shape[1] > 1:
                    # params.w given as an L x T matrix: phi(kk) linear-indexes column 1
                    # (SIAlphaModelEKFOptControlled.m:49-52)
                    v = v[:, 0]
                v = v.ravel()
                full[:L] = v[0] if v.size == 1 else v[:L]
            getattr(p, name)[:] = list(full)
        p.L = L
        ot = d.get("obs_type", "NEWCASES")
        if isinstance(ot, str):
            if ot not in ("NEWCASES", "TOTALCASES"):
                raise ValueError("unknown observation type")  # SIAlphaModelEKF.m:57
            ot = K.OBS_NEWCASES if ot == "NEWCASES" else K.OBS_TOTALCASES
        p.obs_type = int(ot)
    return arr


def params_to_device(arr, device):
    """ctypes epi_model_params array -> torch uint8 CUDA tensor (same bytes)."""
    buf = np.frombuffer(memoryview(arr), dtype=np.uint8).copy()
    return torch.from_numpy(buf).to(device)


class Engine:
    def __init__(self, device=0):
        self._lib = K.load()
        h = C.c_void_p()
        rc = self._lib.epi_create(int(device), C.byref(h))
        if rc != K.OK:
            raise K.EpiError(rc, self._lib.epi_last_error(None).decode())
        self._h = h
        self.device = int(device)
        self._keep = []

    # -- context ------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._lib.epi_destroy(self._h)
            self._h = None

    def __del__(self):
        # never call into CUDA while the interpreter (and the CUDA runtime with it) is
        # being torn down: api.get_engine() registers an atexit close for the shared engine
        if sys is None or sys.is_finalizing():
            return
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != K.OK:
            raise K.EpiError(rc, self._lib.epi_last_error(self._h).decode())

    def sync(self):
        self._ck(self._lib.epi_sync(self._h))

    def set_stream(self, cuda_stream_ptr):
        self._ck(self._lib.epi_set_stream(self._h, C.c_void_p(cuda_stream_ptr or 0)))
        self._bound_stream = cuda_stream_ptr or None

    def use_torch_stream(self):
        """Run on torch's current stream so torch.cuda.Event timing brackets the kernels.
        torch's default stream is the legacy NULL stream; epi_set_stream treats NULL as
        "the context's own stream", so it is passed as the cudaStreamLegacy handle (0x1)."""
        ptr = torch.cuda.current_stream(self.device).cuda_stream or 0x1
        if ptr != getattr(self, "_bound_stream", None):
            self.set_stream(ptr)
            self._bound_stream = ptr

    def set_scratch_limit(self, nbytes):
        self._ck(self._lib.epi_set_scratch_limit(self._h, int(nbytes)))

    def release_cache(self):
        """Return the context's cached device blocks to the driver."""
        self._ck(self._lib.epi_release_cache(self._h))

    @property
    def launch_count(self):
        return int(self._lib.epi_launch_count(self._h))

    def last_kernel_times(self):
        ms = (C.c_float * 16)()
        names = (C.c_char_p * 16)()
        n = self._lib.epi_last_kernel_times(self._h, ms, names, 16)
        return {names[i].decode(): float(ms[i]) for i in range(n)}

    def fp64_probe(self, iters=4096):
        v = C.c_double()
        self._ck(self._lib.epi_fp64_probe(self._h, int(iters), C.byref(v)))
        return v.value

    # -- array plumbing -------------------------------------------------------
    def _mode(self, *arrays):
        dev = [a for a in arrays if _is_torch(a)]
        if dev:
            for a in dev:
                if not a.is_cuda:
                    raise ValueError("torch inputs must be CUDA tensors (EPI_MEM_DEVICE)")
            # Device-memory calls are asynchronous: they are enqueued on torch's CURRENT stream, so they are
            # ordered after whatever produced the input tensors and torch's allocator recycles a tensor only
            # in stream order (a private stream would race with both).
            self.use_torch_stream()
            return K.MEM_DEVICE
        return K.MEM_HOST

    def _in(self, a, mem, dtype=np.float64, n=None):
        """Validated pointer to an input array (None -> NULL)."""
        if a is None:
            return None
        if isinstance(a, C.Array):
            self._keep.append(a)
            return C.addressof(a)
        if mem == K.MEM_DEVICE:
            if not _is_torch(a):
                raise ValueError("EPI_MEM_DEVICE call: every array must be a torch CUDA tensor")
            if not a.is_contiguous():
                raise ValueError("device arrays must be contiguous")
            want = {np.float64: torch.float64, np.uint8: torch.uint8, np.int32: torch.int32}[dtype]
            if a.dtype != want:
                raise ValueError(f"expected dtype {want}, got {a.dtype}")
            if n is not None and a.numel() != n:
                raise ValueError(f"array has {a.numel()} elements, expected {n}")
            self._keep.append(a)
            return a.data_ptr()
        a = np.ascontiguousarray(a, dtype=dtype)
        if n is not None and a.size != n:
            raise ValueError(f"array has {a.size} elements, expected {n}")
        self._keep.append(a)
        return a.ctypes.data

    def _out(self, shape, mem, dtype=np.float64):
        if mem == K.MEM_DEVICE:
            want = {np.float64: torch.float64, np.uint8: torch.uint8, np.int32: torch.int32}[dtype]
            t = torch.empty(shape, dtype=want, device=f"cuda:{self.device}")
            self._keep.append(t)
            return t, t.data_ptr()
        a = np.empty(shape, dtype=dtype)
        self._keep.append(a)
        return a, a.ctypes.data

    def _prm(self, prm, mem):
        if mem == K.MEM_DEVICE:
            if isinstance(prm, C.Array):
                prm = params_to_device(prm, f"cuda:{self.device}")
            self._keep.append(prm)
            return prm.data_ptr()
        self._keep.append(prm)
        return C.addressof(prm)

    def _done(self):
        # The arrays of a call need no keeping alive past its return: host-memory calls are blocking, and
        # device-memory calls run on torch's current stream (_mode), where torch's caching allocator
        # recycles a freed tensor in stream order, i.e. after the kernels that still read it.
        self._keep = []

    # -- SEIRP ----------------------------------------------------------------
    def seirp(self, rates, ic, K_, dt, rate_mode=K.RATES_CONST, saturated=None,
              out_mode=K.SEIRP_OUT_FULL):
        """Tools/SEIRP.m / SEIRPSaturatedResource.m for B trajectories.
        rates: CONST [7,B] | SHARED_SERIES [7,K] | SERIES [K,7,B]; ic [5,B].
        Returns out [5,K,B] (FULL) or [5,B] (FINAL)."""
        mem = self._mode(rates, ic)
        B = int(ic.shape[-1])
        a = K.SeirpArgs()
        a.mem, a.B, a.K, a.dt, a.rate_mode = mem, B, int(K_), float(dt), int(rate_mode)
        n_rates = {K.RATES_CONST: 7 * B, K.RATES_SHARED_SERIES: 7 * K_, K.RATES_SERIES: K_ * 7 * B}[rate_mode]
        a.rates = self._in(rates, mem, n=n_rates)
        a.ic = self._in(ic, mem, n=5 * B)
        if saturated is not None:
            a.saturated = 1
            for k in ("beta_0", "beta_s", "mu_0", "mu_s", "sigma", "i_0"):
                setattr(a, k, float(saturated[k]))
        a.out_mode = int(out_mode)
        out, a.out = self._out((5, K_, B) if out_mode == K.SEIRP_OUT_FULL else (5, B), mem)
        try:
            self._ck(self._lib.epi_seirp_batch(self._h, C.byref(a)))
        finally:
            self._done()
        return out

    # -- rollout + cost ---------------------------------------------------------
    def rollout_cost(self, prm, x0, u, K_, L, G=1, B=None, noise_std=None, noise=None,
                     want_traj=True, want_cost=False, T_total=None, j0_prefix=None,
                     j1_prefix=None, w=None, seed=None, first=0):
        """Tools/SIalpha_Controlled.m (+ Tools/NPICost.m fused).  u [K,L,B] float64 or uint8, or
        u=None with `seed`: the random schedules of TrainPredictPrescribeNPI.m:499-510 are
        generated in the kernel (EPI_U_PHILOX; `first` = global index of trajectory 0)."""
        mem = self._mode(x0, u, noise)
        if u is None:
            if seed is None or B is None:
                raise ValueError("generated schedules need seed= and B=")
        if B is None:
            B = int(u.shape[-1])
        a = K.RolloutArgs()
        a.mem, a.B, a.K, a.L, a.G = mem, B, int(K_), int(L), int(G)
        ng = (B + G - 1) // G
        a.prm = self._prm(prm, mem)
        a.x0 = self._in(x0, mem, n=3 * ng)
        a.noise_std = self._in(noise_std, mem, n=3 * ng)
        if u is None:
            a.u_kind, a.seed, a.first = K.U_PHILOX, int(seed), int(first)
        else:
            is_u8 = (u.dtype == np.uint8) if not _is_torch(u) else (u.dtype == torch.uint8)
            a.u_kind = K.U_U8 if is_u8 else K.U_F64
            a.u = self._in(u, mem, dtype=np.uint8 if is_u8 else np.float64, n=K_ * L * B)
        a.noise = self._in(noise, mem, n=K_ * 3 * B)
        res = {}
        if want_traj:
            res["s"], a.s = self._out((K_, B), mem)
            res["i"], a.i = self._out((K_, B), mem)
            res["alpha"], a.alpha = self._out((K_, B), mem)
        if want_cost:
            a.T_total = int(T_total if T_total is not None else K_)
            a.j0_prefix = self._in(j0_prefix, mem, n=ng)
            a.j1_prefix = self._in(j1_prefix, mem, n=ng)
            a.w = self._in(w, mem, n=ng * K_ * L)
            res["J0"], a.J0 = self._out((B,), mem)
            res["J1"], a.J1 = self._out((B,), mem)
        try:
            self._ck(self._lib.epi_rollout_cost_batch(self._h, C.byref(a)))
        finally:
            self._done()
        return res

    def random_schedules(self, prm, B, K_, L, G, seed, first=0, device=False):
        """The schedules EPI_U_PHILOX integrates, written out: uint8 [K,L,B]."""
        mem = K.MEM_DEVICE if device else K.MEM_HOST
        if device:
            # device-memory calls run on torch's CURRENT stream (see _mode): the output tensor and the
            # temporary params tensor belong to it, so the kernel must be ordered on it too
            self.use_torch_stream()
        a = K.SchedulesArgs()
        a.mem, a.B, a.K, a.L, a.G = mem, int(B), int(K_), int(L), int(G)
        a.prm = self._prm(prm, mem)
        a.seed, a.first = int(seed), int(first)
        u, a.u = self._out((K_, L, B), mem, dtype=np.uint8)
        try:
            self._ck(self._lib.epi_random_schedules(self._h, C.byref(a)))
        finally:
            self._done()
        return u

    def npicost(self, newcases, inputs, weights, T, L, G=1):
        """Tools/NPICost.m: newcases [T,B], inputs [T,L,B], weights per group [T,L]."""
        mem = self._mode(newcases, inputs, weights)
        B = int(newcases.shape[-1])
        a = K.NpiCostArgs()
        a.mem, a.B, a.T, a.L, a.G = mem, B, int(T), int(L), int(G)
        ng = (B + G - 1) // G
        a.newcases = self._in(newcases, mem, n=T * B)
        a.inputs = self._in(inputs, mem, n=T * L * B)
        a.weights = self._in(weights, mem, n=ng * T * L)
        J0, a.J0 = self._out((B,), mem)
        J1, a.J1 = self._out((B,), mem)
        try:
            self._ck(self._lib.epi_npicost_batch(self._h, C.byref(a)))
        finally:
            self._done()
        return J0, J1

    def si_controlled(self, alpha, beta, s0, i0, K_, dt):
        """Tools/SI_Controlled.m: alpha [K,B]; beta, s0, i0 [B]."""
        mem = self._mode(alpha, beta, s0, i0)
        B = int(alpha.shape[-1])
        a = K.SiArgs()
        a.mem, a.B, a.K, a.dt = mem, B, int(K_), float(dt)
        a.alpha = self._in(alpha, mem, n=K_ * B)
        a.beta = self._in(beta, mem, n=B)
        a.s0 = self._in(s0, mem, n=B)
        a.i0 = self._in(i0, mem, n=B)
        s, a.s = self._out((K_, B), mem)
        i, a.i = self._out((K_, B), mem)
        try:
            self._ck(self._lib.epi_si_controlled_batch(self._h, C.byref(a)))
        finally:
            self._done()
        return s, i

    def preprocess(self, cc, population, ip, *, W=7, n_first=7, min_cases=1.0):
        """Data-cleaning block of Tools/TrainPredictPrescribeNPI.m (:121-128, :162-187, :200-201, :240)
        for B regions: cc [T,B] cumulative cases, population [B], ip [T,L,B] NPI levels (NaN allowed).
        Returns dict(ip_filled [T,L,B]; refined, smoothed, zerolag, normalized, confirmed_norm, R_v [T,B]; I0 [B])."""
        mem = self._mode(cc, population, ip)
        T, B = int(cc.shape[0]), int(cc.shape[1])
        L = int(ip.shape[1])
        a = K.PreprocessArgs()
        a.mem, a.B, a.T, a.L, a.W, a.n_first, a.min_cases = mem, B, T, L, int(W), int(n_first), float(min_cases)
        a.cc = self._in(cc, mem, n=T * B)
        a.population = self._in(population, mem, n=B)
        a.ip = self._in(ip, mem, n=T * L * B)
        res = {}
        res["ip_filled"], a.ip_filled = self._out((T, L, B), mem)
        for k in ("refined", "smoothed", "zerolag", "normalized", "confirmed_norm", "R_v"):
            res[k], ptr = self._out((T, B), mem)
            setattr(a, k, ptr)
        res["I0"], a.I0 = self._out((B,), mem)
        try:
            self._ck(self._lib.epi_preprocess_batch(self._h, C.byref(a)))
        finally:
            self._done()
        return res

    def nnls_affine(self, X, y, max_alt=100):
        """lsqnonneg + alternating intercept of TrainPredictPrescribeNPI.m:264-278 for B regions:
        X [n,p,B], y [n,B] -> a [p,B] >= 0, b [B], n_alt [B]."""
        mem = self._mode(X, y)
        n, p, B = int(X.shape[0]), int(X.shape[1]), int(X.shape[2])
        q = K.NnlsArgs()
        q.mem, q.B, q.n, q.p, q.max_alt = mem, B, n, p, int(max_alt)
        q.X = self._in(X, mem, n=n * p * B)
        q.y = self._in(y, mem, n=n * B)
        a, q.a = self._out((p, B), mem)
        b, q.b = self._out((B,), mem)
        k, q.n_alt = self._out((B,), mem, dtype=np.int32)
        try:
            self._ck(self._lib.epi_nnls_affine_batch(self._h, C.byref(q)))
        finally:
            self._done()
        return a, b, k

    RT_OUTPUTS = ("S_MINUS", "S_PLUS", "P_MINUS", "P_PLUS", "K_GAIN", "S_SMOOTH", "P_SMOOTH", "innovations", "rho")

    def rt_expfit(self, x, s_init, params, w_bar, Ps_init, Q, R, *, T, G=1, v_bar=0.0, beta=1.0, gamma=1.0,
                  W=21, order=1, outputs=RT_OUTPUTS):
        """Tools/Rt_ExpFitEKF.m for B trajectories: x [T,B], s_init [2,B]; per group of G
        trajectories params [3] (time_scale, alpha, sigma), w_bar [2], Ps_init / Q [4]
        column-major, R [1].  Returns {name: array} with S_* [T,2,B], P_* [T,4,B], K_GAIN [T,2,B],
        innovations / rho [T,B]."""
        mem = self._mode(x, s_init)
        B = int(x.shape[-1])
        ng = (B + G - 1) // G
        a = K.RtExpFitArgs()
        a.mem, a.B, a.T, a.G = mem, B, int(T), int(G)
        a.x = self._in(x, mem, n=T * B)
        a.s_init = self._in(s_init, mem, n=2 * B)
        a.params = self._in(params, mem, n=3 * ng)
        a.w_bar = self._in(w_bar, mem, n=2 * ng)
        a.Ps_init = self._in(Ps_init, mem, n=4 * ng)
        a.Q = self._in(Q, mem, n=4 * ng)
        a.R = self._in(R, mem, n=ng)
        a.v_bar, a.beta, a.gamma, a.W, a.order = float(v_bar), float(beta), float(gamma), int(W), int(order)
        shapes = dict(S_MINUS=(T, 2, B), S_PLUS=(T, 2, B), P_MINUS=(T, 4, B), P_PLUS=(T, 4, B), K_GAIN=(T, 2, B),
                      S_SMOOTH=(T, 2, B), P_SMOOTH=(T, 4, B), innovations=(T, B), rho=(T, B))
        res = {}
        for k in outputs:
            res[k], ptr = self._out(shapes[k], mem)
            setattr(a, k, ptr)
        try:
            self._ck(self._lib.epi_rt_expfit_batch(self._h, C.byref(a)))
        finally:
            self._done()
        return res

    # -- EKF + smoother -----------------------------------------------------------
    ALL_OUTPUTS = ("u_opt", "u_opt_smooth", "S_MINUS", "S_PLUS", "S_SMOOTH", "P_MINUS", "P_PLUS",
                   "P_SMOOTH", "K_GAIN", "innovations", "rho")

    def ekf_eks(self, model, prm, u, x, R, Q, s_init, Ps_init, s_final, Ps_final, *, B, T, L,
                G=1, epsilon=None, u_per_traj=False, x_per_traj=False, r_mode=K.R_CONST,
                fixed_R=True, r_per_traj=False, q_mode=K.Q_CONST, init_per_traj=False,
                v_bar=0.0, beta=1.0, gamma=1.0, W=21, order=1, outputs=ALL_OUTPUTS,
                want_status=False):
        """GenericExtendedKalmanFilter.m with one of the four known handle sets, or the
        legacy NewCaseEKFEstimatorWithOptimalNPI.m.  Returns a dict of the requested
        outputs in trajectory-minor layout ([T,L,B], [T,m,B], [T,m*m,B], [T,B])."""
        mem = self._mode(u, x, R, Q, s_init, Ps_init, s_final, Ps_final, epsilon)
        m = 6 if model >= K.MODEL_OPTCTRL else 3
        ng = (B + G - 1) // G
        a = K.EkfArgs()
        a.mem, a.model, a.B, a.T, a.L, a.G = mem, int(model), int(B), int(T), int(L), int(G)
        a.prm = self._prm(prm, mem)
        a.epsilon = self._in(epsilon, mem, n=B)
        a.u_per_traj, a.x_per_traj, a.r_per_traj = int(u_per_traj), int(x_per_traj), int(r_per_traj)
        a.u = self._in(u, mem, n=T * L * (B if u_per_traj else ng))
        a.x = self._in(x, mem, n=T * (B if x_per_traj else ng))
        a.r_mode, a.fixed_R, a.q_mode = int(r_mode), int(fixed_R), int(q_mode)
        a.R = self._in(R, mem, n=(1 if r_mode == K.R_CONST else T) * (B if r_per_traj else ng))
        qn = {K.Q_CONST: m * m, K.Q_PERDAY_SCALAR: T, K.Q_PERDAY_FULL: T * m * m}.get(q_mode)
        a.Q = self._in(Q, mem, n=None if qn is None else qn * ng)
        a.init_per_traj = int(init_per_traj)
        nb = B if init_per_traj else ng
        a.s_init = self._in(s_init, mem, n=m * nb)
        a.Ps_init = self._in(Ps_init, mem, n=m * m * nb)
        a.s_final = self._in(s_final, mem, n=m * nb)
        a.Ps_final = self._in(Ps_final, mem, n=m * m * nb)
        a.v_bar, a.beta, a.gamma, a.W, a.order = float(v_bar), float(beta), float(gamma), int(W), int(order)
        shapes = dict(u_opt=(T, L, B), u_opt_smooth=(T, L, B), S_MINUS=(T, m, B), S_PLUS=(T, m, B),
                      S_SMOOTH=(T, m, B), P_MINUS=(T, m * m, B), P_PLUS=(T, m * m, B),
                      P_SMOOTH=(T, m * m, B), K_GAIN=(T, m, B), innovations=(T, B), rho=(T, B))
        res = {}
        legacy = model >= K.MODEL_LEGACY_TOOLS
        for name in outputs:
            if name == "u_opt_smooth" and legacy:
                continue
            res[name], ptr = self._out(shapes[name], mem)
            setattr(a, name, ptr)
        if want_status:
            res["status"], a.status = self._out((B,), mem, dtype=np.int32)
        try:
            self._ck(self._lib.epi_ekf_eks_batch(self._h, C.byref(a)))
        finally:
            self._done()
        return res

    # -- Pareto -------------------------------------------------------------------
    def pareto(self, J0, J1, out=None):
        """TrainPredictPrescribeNPI.m:624-633 for n_sets point sets: J0, J1 [n_sets, n].
        `out` = (on_front uint8 [n_sets, n], I_opt int32 [n_sets]) reuses preallocated device buffers."""
        mem = self._mode(J0, J1)
        n_sets, n = (int(v) for v in J0.shape)
        a = K.ParetoArgs()
        a.mem, a.n_sets, a.n = mem, n_sets, n
        a.J0 = self._in(J0, mem, n=n_sets * n)
        a.J1 = self._in(J1, mem, n=n_sets * n)
        if out is not None and mem == K.MEM_DEVICE:
            mask, iopt = out
            a.on_front = self._in(mask, mem, dtype=np.uint8, n=n_sets * n)
            a.I_opt = self._in(iopt, mem, dtype=np.int32, n=n_sets)
        else:
            mask, a.on_front = self._out((n_sets, n), mem, dtype=np.uint8)
            iopt, a.I_opt = self._out((n_sets,), mem, dtype=np.int32)
        try:
            self._ck(self._lib.epi_pareto_batch(self._h, C.byref(a)))
        finally:
            self._done()
        return mask, iopt

    # -- fused sweep ------------------------------------------------------------------
    def sweep(self, prm, eps, u, x, R, s_init, Ps_init, s_final, Ps_final, Q, x0, newcases_hist,
              weights, *, n_regions, T, T_hist, L, beta_ekf=1.0, gamma_ekf=0.995, W=21,
              noise_std=None, noise=None, want_front=True, want_u_knee=False, want_u_fore=False,
              want_P_first=False, out=None, lean=False, peers=None):
        """The optimal-NPI Pareto sweep (TrainPredictPrescribeNPI.m:421-495,624-633) for all
        regions x all epsilon.  Per-region arrays are [n_regions, ...] row-major with the
        MATLAB column-major page inside ([T,L] days-major, [36] column-major).  `out`
        may carry preallocated J0/J1/on_front/I_opt buffers (device mode) for reuse.
        `peers`: Engines on OTHER GPUs -- the one blocking call then shards the regions over
        [self] + peers (epi_sweep_multi; host arrays only) and returns the same arrays, bit for bit."""
        mem = self._mode(eps, u, x, R, s_init, Ps_init, s_final, Ps_final, Q, x0, newcases_hist,
                         weights, noise)
        nR, nE = int(n_regions), int(eps.shape[0])
        B, Tf = nR * nE, T - T_hist
        a = K.SweepArgs()
        a.mem, a.n_regions, a.n_eps, a.T, a.T_hist, a.L = mem, nR, nE, int(T), int(T_hist), int(L)
        a.prm = self._prm(prm, mem)
        a.eps = self._in(eps, mem, n=nE)
        a.u = self._in(u, mem, n=nR * T * L)
        a.x = self._in(x, mem, n=nR * T)
        a.R = self._in(R, mem, n=nR * T)
        a.s_init = self._in(s_init, mem, n=nR * 6)
        a.Ps_init = self._in(Ps_init, mem, n=nR * 36)
        a.s_final = self._in(s_final, mem, n=nR * 6)
        a.Ps_final = self._in(Ps_final, mem, n=nR * 36)
        a.Q = self._in(Q, mem, n=nR * 36)
        a.beta_ekf, a.gamma_ekf, a.W = float(beta_ekf), float(gamma_ekf), int(W)
        a.lean = int(bool(lean))
        a.x0 = self._in(x0, mem, n=nR * 3)
        a.newcases_hist = self._in(newcases_hist, mem, n=nR * T_hist)
        a.weights = self._in(weights, mem, n=nR * T * L)
        a.noise_std = self._in(noise_std, mem, n=nR * 3)
        a.noise = self._in(noise, mem, n=Tf * 3 * B)
        res = {}
        out = out or {}

        def outbuf(name, shape, dtype=np.float64):
            if name in out:
                o = out[name]
                if mem == K.MEM_HOST:
                    # the result is written through this pointer: never hand the library a converted copy
                    if not (isinstance(o, np.ndarray) and o.flags["C_CONTIGUOUS"] and o.flags["WRITEABLE"]
                            and o.dtype == np.dtype(dtype)):
                        raise ValueError(f"out[{name!r}] must be a writeable C-contiguous {np.dtype(dtype)} array")
                res[name] = o
                return self._in(o, mem, dtype=dtype, n=int(np.prod(shape)))
            res[name], ptr = self._out(shape, mem, dtype=dtype)
            return ptr

        a.J0 = outbuf("J0", (nR, nE))
        a.J1 = outbuf("J1", (nR, nE))
        if want_front:
            a.on_front = outbuf("on_front", (nR, nE), np.uint8)
            a.I_opt = outbuf("I_opt", (nR,), np.int32)
        if want_u_knee:
            a.u_knee = outbuf("u_knee", (nR, Tf, L))
        if want_u_fore:
            a.u_fore = outbuf("u_fore", (Tf, L, B))
        if want_P_first:
            a.P_first = outbuf("P_first", (36, B))
        try:
            if peers:
                if mem != K.MEM_HOST:
                    raise ValueError("sweep(peers=...): host arrays only (device arrays belong to one GPU)")
                hs = (C.c_void_p * (1 + len(peers)))(self._h, *[p._h for p in peers])
                self._ck(self._lib.epi_sweep_multi(hs, 1 + len(peers), C.byref(a)))
            else:
                self._ck(self._lib.epi_sweep(self._h, C.byref(a)))
        finally:
            self._done()
        return res
