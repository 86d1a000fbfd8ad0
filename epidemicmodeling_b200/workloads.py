"""Host-side drivers that batch the reference's sweep loops onto the engine.

These mirror the loop bodies of Tools/TrainPredictPrescribeNPI.m:370-392 (the
3-state "fixed input" EKF/EKS that supplies the historic estimate and the rollout
start) and :415-495,624-633 (the epsilon sweep + Pareto step), turning the
region x epsilon loops into one batch.  Input construction is numpy glue; every
filter / smoother / rollout / cost / Pareto computation runs in libepi_b200.so.
"""
import numpy as np

from . import _capi as K
from .engine import pack_params

try:
    import torch
except Exception:  # pragma: no cover
    torch = None


def _cm(P):
    """m x m matrix -> column-major flat (symmetric matrices are unchanged)."""
    return np.ascontiguousarray(np.asarray(P, dtype=np.float64).T).ravel()


def fixed_input_batch(inputs):
    """Arrays of the 3-state fixed-input EKF/EKS (one trajectory per region)."""
    L = inputs[0]["u_fixed"].shape[0]
    return dict(
        prm=pack_params([r["setup3"]["params"] for r in inputs], L),
        u=np.stack([np.ascontiguousarray(r["u_fixed"].T) for r in inputs]),     # [nR,T,L]
        x=np.stack([r["x"] for r in inputs]),                                    # [nR,T]
        R=np.stack([r["R_v"] for r in inputs]),                                  # [nR,T]
        Q=np.stack([_cm(r["setup3"]["Q_w"]) for r in inputs]),
        s_init=np.stack([r["setup3"]["s_init"] for r in inputs]),
        Ps_init=np.stack([_cm(r["setup3"]["Ps_init"]) for r in inputs]),
        s_final=np.stack([r["setup3"]["s_final"] for r in inputs]),
        Ps_final=np.stack([_cm(r["setup3"]["Ps_final"]) for r in inputs]),
        T=inputs[0]["T"], L=L, n_regions=len(inputs),
        beta=inputs[0]["setup3"]["beta_ekf"], gamma=inputs[0]["setup3"]["gamma_ekf"],
        W=inputs[0]["setup3"]["W"])


def run_fixed_input(engine, inputs):
    """S_SMOOTH [T,3,nR] of the fixed-input round (TrainPredictPrescribeNPI.m:377)."""
    b = fixed_input_batch(inputs)
    out = engine.ekf_eks(K.MODEL_SIALPHA, b["prm"], b["u"], b["x"], b["R"], b["Q"], b["s_init"],
                         b["Ps_init"], b["s_final"], b["Ps_final"], B=b["n_regions"], T=b["T"],
                         L=b["L"], G=1, r_mode=K.R_PERDAY, fixed_R=False, q_mode=K.Q_CONST,
                         beta=b["beta"], gamma=b["gamma"], W=b["W"], order=1, outputs=("S_SMOOTH",))
    return out["S_SMOOTH"]


def sweep_batch(inputs, S_SMOOTH_fixed, x0=None):
    """Per-region arrays of Engine.sweep from the synthetic inputs and the
    fixed-input smoothed states [T,3,nR] (:380-382 historic estimates, :481 start).
    `x0` [nR,3] overrides the rollout start; it is REQUIRED when there is no historic day
    (the reference takes the start from the last historic day, :481)."""
    L = inputs[0]["u_hist"].shape[0]
    T, Th = inputs[0]["T"], inputs[0]["T_hist"]
    S = np.asarray(S_SMOOTH_fixed)
    hist = S[:Th]                                             # [Th,3,nR]
    newcases_hist = ((hist[:, 0] * hist[:, 1]) * hist[:, 2]).T  # s.*i.*alpha  [nR,Th]
    if x0 is None:
        if Th < 1:
            raise ValueError("sweep_batch: no historic day to start the rollout from -- pass x0")
        x0 = hist[Th - 1].T                                   # [nR,3]
    x0 = np.asarray(x0, dtype=np.float64)
    u = np.stack([np.concatenate([np.ascontiguousarray(r["u_hist"].T), np.full((T - Th, L), np.nan)])
                  for r in inputs])                           # :458
    return dict(
        prm=pack_params([r["setup6"]["params"] for r in inputs], L),
        u=u, x=np.stack([r["x"] for r in inputs]), R=np.stack([r["R_v"] for r in inputs]),
        s_init=np.stack([r["setup6"]["s_init"] for r in inputs]),
        Ps_init=np.stack([_cm(r["setup6"]["Ps_init"]) for r in inputs]),
        s_final=np.stack([r["setup6"]["s_final"] for r in inputs]),
        Ps_final=np.stack([_cm(r["setup6"]["Ps_final"]) for r in inputs]),
        Q=np.stack([_cm(r["setup6"]["Q_w"]) for r in inputs]),
        x0=np.ascontiguousarray(x0), newcases_hist=np.ascontiguousarray(newcases_hist),
        weights=np.stack([np.ascontiguousarray(r["weights"].T) for r in inputs]),  # [nR,T,L]
        n_regions=len(inputs), T=T, T_hist=Th, L=L,
        beta_ekf=inputs[0]["setup6"]["beta_ekf"], gamma_ekf=inputs[0]["setup6"]["gamma_ekf"],
        W=inputs[0]["setup6"]["W"])


_SWEEP_ARRAYS = ("u", "x", "R", "s_init", "Ps_init", "s_final", "Ps_final", "Q", "x0",
                 "newcases_hist", "weights")


def sweep_to_device(batch, eps, device):
    """Upload a sweep batch once (bench `value` leg: inputs resident in HBM)."""
    from .engine import params_to_device
    d = dict(batch)
    for k in _SWEEP_ARRAYS:
        d[k] = torch.from_numpy(np.ascontiguousarray(batch[k])).to(device)
    d["prm"] = params_to_device(batch["prm"], device)
    d["eps"] = torch.from_numpy(np.ascontiguousarray(eps, dtype=np.float64)).to(device)
    return d


def run_sweep(engine, batch, eps, **kw):
    """One pass of the fused sweep; `batch` from sweep_batch() (host) or sweep_to_device()."""
    e = batch.get("eps", eps)
    return engine.sweep(batch["prm"], e, batch["u"], batch["x"], batch["R"], batch["s_init"],
                        batch["Ps_init"], batch["s_final"], batch["Ps_final"], batch["Q"],
                        batch["x0"], batch["newcases_hist"], batch["weights"],
                        n_regions=batch["n_regions"], T=batch["T"], T_hist=batch["T_hist"],
                        L=batch["L"], beta_ekf=batch["beta_ekf"], gamma_ekf=batch["gamma_ekf"],
                        W=batch["W"], **kw)


def forecast_quality(engine, inputs, num_forecast_days, max_look_ahead_days, truth=None):
    """The masked-horizon loop of Tools/ForecastQualityAssessment.m:383-394 as ONE batch:
    for start = 1..num_forecast_days the last `start` observations are set to NaN and the
    3-state EKF/EKS is re-run (SIAlphaModelEKF with the region's full NPI history).
    One group per region, `num_forecast_days` trajectories per group, x per trajectory.

    inputs: per-region dicts with u (L x T, key "u_fixed"), x (T), R_v (T), setup3 (synthetic.py).
    truth:  [nR, T] ground-truth new cases (fractions of N); default = the unmasked observations.
    Returns S_PLUS, S_SMOOTH [T,3,nR*nf] and EstError_PLUS / EstError_SMOOTH [nR, nf, max_look_ahead]
    (percent errors; NaN where the reference leaves its preallocated zeros untouched is kept as 0)."""
    b = fixed_input_batch(inputs)
    nR, T, nf = b["n_regions"], b["T"], int(num_forecast_days)
    x = np.repeat(b["x"].T[:, :, None], nf, axis=2)                 # [T, nR, nf]
    for start in range(1, nf + 1):
        x[T - start:, :, start - 1] = np.nan                        # observations_PARTIAL(LL-start+1:LL) = nan
    x = np.ascontiguousarray(x.reshape(T, nR * nf))
    out = engine.ekf_eks(K.MODEL_SIALPHA, b["prm"], b["u"], x, b["R"], b["Q"], b["s_init"], b["Ps_init"],
                         b["s_final"], b["Ps_final"], B=nR * nf, T=T, L=b["L"], G=nf, x_per_traj=True,
                         r_mode=K.R_PERDAY, fixed_R=False, q_mode=K.Q_CONST, beta=b["beta"], gamma=b["gamma"],
                         W=b["W"], order=1, outputs=("S_PLUS", "S_SMOOTH"))
    truth = b["x"] if truth is None else np.asarray(truth, dtype=np.float64)
    err = {}
    for key in ("S_PLUS", "S_SMOOTH"):
        S = np.asarray(out[key]).reshape(T, 3, nR, nf)
        est = (S[:, 0] * S[:, 1]) * S[:, 2]                         # s.*i.*alpha  [T, nR, nf]
        e = 100.0 * np.abs(truth.T[:, :, None] - est) / truth.T[:, :, None]
        E = np.zeros((nR, nf, max_look_ahead_days))
        for start in range(1, nf + 1):                               # :392-394
            last = min(T, T - start + max_look_ahead_days)
            n = last - T + start
            E[:, start - 1, :n] = e[T - start:last, :, start - 1].T
        err[key] = E
    return dict(S_PLUS=out["S_PLUS"], S_SMOOTH=out["S_SMOOTH"], EstError_PLUS=err["S_PLUS"],
                EstError_SMOOTH=err["S_SMOOTH"])


def shard_regions(n_regions, world_size, rank):
    """Contiguous blocks of ceil(n/G) regions per rank."""
    per = (n_regions + world_size - 1) // world_size
    lo = min(n_regions, rank * per)
    return lo, min(n_regions, lo + per)


def sharded_sweep(compute, inputs, eps, rank, world_size, group=None):
    """Strong-scaling form of the sweep: rank r takes the contiguous block of regions
    shard_regions() gives it, runs `compute(local_inputs, eps) -> (J0, J1)` ([n_local, n_eps] torch
    tensors on the rank's device) and all ranks end up with the full (J0, J1) through the path's single
    collective.  `compute` is the engine on a GPU box (see bench.py); the CPU tests inject a stand-in."""
    lo, hi = shard_regions(len(inputs), world_size, rank)
    J0, J1 = compute(inputs[lo:hi], eps)
    if world_size == 1:
        return J0, J1
    return gather_costs(J0, J1, group=group)


def engine_sweep_compute(engine, device=None, lean=False):
    """compute() for sharded_sweep backed by the engine: fixed-input smoother -> sweep batch -> epi_sweep."""
    def compute(local_inputs, eps):
        n_eps = len(eps)
        if not local_inputs:
            z = torch.zeros((0, n_eps), dtype=torch.float64, device=device or "cpu")
            return z, z.clone()
        S = run_fixed_input(engine, local_inputs)
        res = run_sweep(engine, sweep_batch(local_inputs, S), np.asarray(eps, dtype=np.float64), want_front=False,
                        lean=lean)
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device or "cpu")
        return t(res["J0"]), t(res["J1"])
    return compute


def gather_costs(J0, J1, group=None):
    """The path's one collective: all-gather of the per-shard (J0, J1) arrays.
    J0, J1 [n_local_regions, n_eps] torch tensors (CUDA -> NCCL over NVLink; CPU -> gloo).
    Shards may be ragged (last rank), so sizes are exchanged first."""
    import torch.distributed as dist
    ws = dist.get_world_size(group)
    n_eps = J0.shape[1]
    sizes = [torch.zeros(1, dtype=torch.int64, device=J0.device) for _ in range(ws)]
    dist.all_gather(sizes, torch.tensor([J0.shape[0]], dtype=torch.int64, device=J0.device), group=group)
    sizes = [int(s.item()) for s in sizes]
    mx = max(sizes)
    send = torch.zeros((2, mx, n_eps), dtype=J0.dtype, device=J0.device)
    send[0, :J0.shape[0]] = J0
    send[1, :J1.shape[0]] = J1
    recv = torch.empty((ws, 2, mx, n_eps), dtype=J0.dtype, device=J0.device)
    dist.all_gather_into_tensor(recv.view(-1), send.view(-1), group=group)  # flat: gloo and nccl agree
    g0 = torch.cat([recv[r, 0, :sizes[r]] for r in range(ws)])
    g1 = torch.cat([recv[r, 1, :sizes[r]] for r in range(ws)])
    return g0, g1


def gather_fronts(on_front, I_opt, group=None):
    """Config 5 at N GPUs (TrainPredictPrescribeNPI.m:500-521 + :624-633 per region): every rank scores the random
    schedules of ITS regions and extracts their Pareto fronts; this is the path's one collective -- all-gather of the
    per-region front masks (uint8 [n_local_regions, n_schedules]) and knee indices (int32 [n_local_regions]).
    Ragged shards (last rank) are padded to the largest shard; returns the masks and knees of all regions in order."""
    import torch.distributed as dist
    ws = dist.get_world_size(group)
    n_loc, nS = int(on_front.shape[0]), int(on_front.shape[1])
    sizes = [torch.zeros(1, dtype=torch.int64, device=on_front.device) for _ in range(ws)]
    dist.all_gather(sizes, torch.tensor([n_loc], dtype=torch.int64, device=on_front.device), group=group)
    sizes = [int(s.item()) for s in sizes]
    mx = max(sizes)
    send_m = torch.zeros((mx, nS), dtype=torch.uint8, device=on_front.device)
    send_i = torch.zeros((mx,), dtype=torch.int32, device=on_front.device)
    send_m[:n_loc] = on_front
    send_i[:n_loc] = I_opt
    recv_m = torch.empty((ws, mx, nS), dtype=torch.uint8, device=on_front.device)
    recv_i = torch.empty((ws, mx), dtype=torch.int32, device=on_front.device)
    dist.all_gather_into_tensor(recv_m.view(-1), send_m.view(-1), group=group)
    dist.all_gather_into_tensor(recv_i.view(-1), send_i.view(-1), group=group)
    return (torch.cat([recv_m[r, :sizes[r]] for r in range(ws)]), torch.cat([recv_i[r, :sizes[r]] for r in range(ws)]))
