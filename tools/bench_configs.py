"""Secondary measurements for the other BASELINE.json configs (2, 3, 5) on one B200.
bench.py carries the headline (config 4); this script records the remaining kernels'
throughput and roofline fractions into gpurun_out/configs.jsonl (copied to profiles/).

    python tools/bench_configs.py [--scale 1.0]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from epidemicmodeling_b200 import _capi as K  # noqa: E402
from epidemicmodeling_b200 import synthetic as syn, workloads as wl  # noqa: E402
from epidemicmodeling_b200.engine import Engine, pack_params, params_to_device  # noqa: E402

HBM = 6550.1
try:
    HBM = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def _orc():
    from oracle import oracle as orc   # the checker (test infrastructure): spot checks only
    return orc


def _bits(a, b):
    return bool(np.array_equal(np.asarray(a), np.asarray(b)))


def measure(eng, only=0, scale=1.0, spot_check=True, fp64=None):
    """Time configs 2, 3 and 5 on the engine's device; returns a list of records.  With
    `spot_check` a handful of trajectories of every config is compared bit for bit with the
    CPU oracle (outside the timed calls)."""
    class A:
        pass
    a = A()
    a.scale, a.only = scale, only
    eng.use_torch_stream()
    fp64 = fp64 or eng.fp64_probe(4096)
    out = []
    dev = f"cuda:{eng.device}"

    nR = 236
    holder = {}
    t = lambda v: torch.from_numpy(np.ascontiguousarray(v)).to(dev)
    # ---- config 2: SEIRP ensemble, 1M x 365 days
    if a.only in (0, 2):
        B, Kn = int(1_000_000 * a.scale), 365
        rates, ic = syn.seirp_ensemble(B)
        rd, icd = torch.from_numpy(rates).to(dev), torch.from_numpy(ic).to(dev)
        for mode, name in ((K.SEIRP_OUT_FULL, "full"), (K.SEIRP_OUT_FINAL, "final")):
            ms_call = timed(lambda: eng.seirp(rd, icd, Kn, 1.0, out_mode=mode))
            ms = sum(eng.last_kernel_times().values())   # CUDA events around the kernel, on its stream
            steps = B * (Kn - 1)
            rec = {"config": 2, "kernel": f"seirp[{name}]", "B": B, "K": Kn, "ms": ms,
                   "ms_call_incl_output_alloc": ms_call, "steps_per_s": steps / ms * 1e3,
                   "hbm_gbs": (40.0 * B * Kn / ms * 1e3 / 1e9) if mode == K.SEIRP_OUT_FULL else 0.0,
                   "fp64_tflops": 37.0 * steps / ms * 1e3 / 1e12}
            rec["hbm_frac"] = rec["hbm_gbs"] / HBM
            rec["fp64_frac"] = rec["fp64_tflops"] / fp64
            if spot_check and mode == K.SEIRP_OUT_FULL:
                o, full = _orc(), eng.seirp(rd, icd, Kn, 1.0, out_mode=mode)
                eng.sync()
                ok = True
                for b in (0, 31, B // 2, B - 1):
                    want = o.SEIRP(*rates[:, b], *ic[:, b], float(Kn), 1.0)
                    got = full[:, :, b].cpu().numpy()
                    ok = ok and all(_bits(got[f], want[f][0]) for f in range(5))
                rec["oracle_spot_check"] = {"trajectories": 4, "bit_exact": ok}
                del full
            out.append(rec)
        del rd, icd
        sat = dict(beta_0=0.1, beta_s=0.01, mu_0=0.02, mu_s=0.2, sigma=1.0, i_0=0.1)
        rd, icd = torch.from_numpy(rates).to(dev), torch.from_numpy(ic).to(dev)
        timed(lambda: eng.seirp(rd, icd, Kn, 1.0, saturated=sat, out_mode=K.SEIRP_OUT_FINAL))
        ms = sum(eng.last_kernel_times().values())
        out.append({"config": 2, "kernel": "seirp_saturated[final]", "B": B, "K": Kn, "ms": ms,
                    "steps_per_s": B * (Kn - 1) / ms * 1e3})
        del rd, icd
        torch.cuda.empty_cache()
        eng.release_cache()

    # ---- config 3: SI-alpha EKF + smoother, 236 regions x replicates x 400 days
    if a.only in (0, 3):
        nR, nRep, T = 236, int(10_000 * a.scale), 400
        inp = syn.sweep_inputs(n_regions=nR, T_hist=T, T_fore=0)
        b = wl.fixed_input_batch(inp)
        Bt = nR * nRep
        clean = torch.from_numpy(np.stack([r["x"] for r in inp])).to(dev)            # [nR,T]
        g = torch.Generator(device=dev).manual_seed(3)
        x = torch.empty((T, Bt), dtype=torch.float64, device=dev)
        for r in range(nR):
            x[:, r * nRep:(r + 1) * nRep] = torch.clamp_min(
                clean[r][:, None] * (1.0 + 0.05 * torch.randn((T, nRep), dtype=torch.float64, device=dev, generator=g)), 0.0)
        t = lambda v: torch.from_numpy(np.ascontiguousarray(v)).to(dev)
        args = (params_to_device(b["prm"], dev), t(b["u"]), x, t(b["R"]), t(b["Q"]), t(b["s_init"]), t(b["Ps_init"]),
                t(b["s_final"]), t(b["Ps_final"]))
        kw = dict(B=Bt, T=T, L=b["L"], G=nRep, x_per_traj=True, r_mode=K.R_PERDAY, fixed_R=False, beta=1.0,
                  gamma=b["gamma"], W=b["W"], outputs=("S_SMOOTH",))
        holder = {}

        eng.set_scratch_limit(48 << 30)   # keep the waves' scratch in the pool instead of fighting torch's allocator

        def run3():
            holder.clear()
            holder["o"] = eng.ekf_eks(K.MODEL_SIALPHA, *args, **kw)
        ms = timed(run3, reps=3, warm=1)
        eng.set_scratch_limit(0)
        kt = eng.last_kernel_times()
        units = Bt * T
        spot3 = None
        if spot_check:
            o, S, ok = _orc(), holder["o"]["S_SMOOTH"], True
            picks = (0, nRep - 1, nRep, Bt // 2 + 7, Bt - 1)
            for bb in picks:
                r = inp[bb // nRep]
                s3 = r["setup3"]
                want = o.ekf_eks(o.SIALPHA, r["u_fixed"], x[:, bb].cpu().numpy(), s3["params"], s3["s_init"], s3["Ps_init"],
                                 s3["s_final"], s3["Ps_final"], s3["w_bar"], 0.0, s3["Q_w"], r["R_v"], 1.0, s3["gamma_ekf"],
                                 s3["W"], 1)
                ok = ok and _bits(S[:, :, bb].cpu().numpy().T, want["S_SMOOTH"])
            spot3 = {"trajectories": len(picks), "bit_exact": ok}
        out.append({"config": 3, "oracle_spot_check": spot3, "kernel": "ekf_eks<3> lean (S_SMOOTH out)", "B": Bt, "T": T, "ms": ms,
                    "trajectory_days_per_s": units / ms * 1e3, "kernel_ms": kt, "kernel_ms_sum": sum(kt.values()),
                    "hbm_gbs_algorithmic(616B)": 616.0 * units / ms * 1e3 / 1e9,
                    "hbm_frac": 616.0 * units / ms * 1e3 / 1e9 / HBM,
                    "fp64_frac(818flop)": 818.0 * units / ms * 1e3 / 1e12 / fp64})
        holder.clear()
        del x, args
        torch.cuda.empty_cache()
        eng.release_cache()

    # ---- config 5: random-NPI Monte-Carlo scoring, 236 regions x 100k schedules x 120 days (uint8 NPIs)
    if a.only in (0, 5):
        nS, Kn, L = int(100_000 * a.scale), 120, 12
        reg = syn.load_regions(nR)
        Bm = nR * nS
        prm = pack_params([dict(dt=1.0, beta=syn.BETA, gamma=syn.GAMMA, b=reg["b"][r], a=reg["a"][r], u_max=reg["npi_max"],
                                u_min=np.zeros(L), alpha_min=1e-8, alpha_max=100.0) for r in range(nR)], L)
        x0 = t(np.array([[(reg["N"][r] - 10) / reg["N"][r], 10 / reg["N"][r], syn.ALPHA0] for r in range(nR)]))
        w = t(np.stack([np.repeat(reg["cost_weights"][r][None, :], Kn, axis=0) for r in range(nR)]))
        j0p, j1p = torch.zeros(nR, dtype=torch.float64, device=dev), torch.zeros(nR, dtype=torch.float64, device=dev)
        prmd = params_to_device(prm, dev)
        # the schedules of TrainPredictPrescribeNPI.m:500-510 (first half of a region's scenarios constant in time,
        # second half per day), written out as uint8 by the library's counter-based rule (seed 5)
        u8 = eng.random_schedules(prmd, Bm, Kn, L, nS, 5, device=True)

        def run5():
            holder["o"] = eng.rollout_cost(prmd, x0, u8, Kn, L, G=nS, B=Bm, want_traj=False, want_cost=True,
                                           T_total=Kn, j0_prefix=j0p, j1_prefix=j1p, w=w)
        ms_call = timed(run5, reps=3, warm=1)
        ms = sum(eng.last_kernel_times().values())
        units = Bm * Kn
        spot5 = None
        if spot_check:
            o, ok = _orc(), True
            x0h, wh = x0.cpu().numpy(), w.cpu().numpy()
            picks = (0, nS - 1, nS, Bm // 2 + 3, Bm - 1)
            for bb in picks:
                r = bb // nS
                ub = u8[:, :, bb].cpu().numpy().T.astype(float)
                s_, i_, al_ = o.SIalpha_Controlled(ub, *x0h[r], reg["npi_max"], 1e-8, 100.0, syn.GAMMA, reg["a"][r], reg["b"][r],
                                                   syn.BETA, 0.0, 0.0, 0.0, Kn, 1.0)
                j0, j1 = o.NPICost((s_ * i_) * al_, ub, wh[r].T)
                ok = ok and float(holder["o"]["J0"][bb]) == j0 and float(holder["o"]["J1"][bb]) == j1
            spot5 = {"trajectories": len(picks), "bit_exact": bool(ok)}
        out.append({"config": 5, "kernel": "rollout_cost[u8]", "oracle_spot_check": spot5, "B": Bm, "K": Kn, "ms": ms,
                    "trajectory_days_per_s": units / ms * 1e3, "hbm_gbs_algorithmic(12B)": 12.0 * units / ms * 1e3 / 1e9,
                    "hbm_frac": 12.0 * units / ms * 1e3 / 1e9 / HBM, "fp64_frac(91flop)": 91.0 * units / ms * 1e3 / 1e12 / fp64})

        # the same scoring with the schedules drawn in the kernel (Philox, seed 5: SURVEY 8d config 5,
        # no HBM stream at all) -- and the supplied-u8 path fed with exactly those schedules must agree
        del u8
        torch.cuda.empty_cache()

        def run5g():
            holder["g"] = eng.rollout_cost(prmd, x0, None, Kn, L, G=nS, B=Bm, want_traj=False, want_cost=True,
                                           T_total=Kn, j0_prefix=j0p, j1_prefix=j1p, w=w, seed=5)
        timed(run5g, reps=3, warm=1)
        ms = sum(eng.last_kernel_times().values())
        eng.sync()
        same = bool(torch.equal(holder["o"]["J0"], holder["g"]["J0"]) and torch.equal(holder["o"]["J1"], holder["g"]["J1"]))
        out.append({"config": 5, "kernel": "rollout_cost[philox, generated in-kernel]", "B": Bm, "K": Kn, "ms": ms,
                    "trajectory_days_per_s": units / ms * 1e3, "fp64_frac(91flop)": 91.0 * units / ms * 1e3 / 1e12 / fp64,
                    "bit_identical_to_supplied_u8": same})
        holder["o"] = holder["g"]

        J0, J1 = holder["o"]["J0"].view(nR, nS), holder["o"]["J1"].view(nR, nS)
        ms = timed(lambda: holder.__setitem__("p", eng.pareto(J0, J1)), reps=3, warm=1)
        mask, iopt = holder["p"]
        out.append({"config": 5, "kernel": "pareto[sorted]", "n_sets": nR, "n": nS, "ms": ms,
                    "points_per_s": nR * nS / ms * 1e3, "front_sizes_mean": float(mask.sum(dim=1).double().mean().item())})

    for r in out:
        r["hbm_peak_gbs"], r["fp64_peak_tflops_measured"] = HBM, fp64
    holder.clear()
    torch.cuda.empty_cache()
    eng.release_cache()
    return out



def measure_config5_sharded(eng, rank, world, dev, steps=5, warm=2, scale=1.0):
    """BASELINE config 5 at N GPUs (every rank calls this): the 236 regions x 100k random NPI schedules x 120 days are
    sharded by region (TrainPredictPrescribeNPI.m:421,500 are the loops), every rank draws its schedules in the kernel
    (Philox counters are global: `first` = the shard's first trajectory, so the job is the same whatever N), scores
    them (SIalpha_Controlled + NPICost), extracts the Pareto front and knee of each of ITS regions, and one all-gather
    (two NCCL calls: front masks and knee indices) assembles the per-region fronts on every rank.  Device-timed, max
    over ranks; rank 0 spot-checks its first and last trajectory against the oracle."""
    import torch.distributed as dist
    from epidemicmodeling_b200 import workloads as wl
    nR, nS, Kn, L = 236, int(100_000 * scale), 120, 12
    lo, hi = wl.shard_regions(nR, world, rank)
    per, n_loc = (nR + world - 1) // world, hi - lo
    reg = syn.load_regions(nR)
    t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
    prm = pack_params([dict(dt=1.0, beta=syn.BETA, gamma=syn.GAMMA, b=reg["b"][r], a=reg["a"][r], u_max=reg["npi_max"],
                            u_min=np.zeros(L), alpha_min=1e-8, alpha_max=100.0) for r in range(lo, hi)], L)
    prmd = params_to_device(prm, dev)
    x0 = t(np.array([[(reg["N"][r] - 10) / reg["N"][r], 10 / reg["N"][r], syn.ALPHA0] for r in range(lo, hi)]))
    w = t(np.stack([np.repeat(reg["cost_weights"][r][None, :], Kn, axis=0) for r in range(lo, hi)]))
    j0p = torch.zeros(n_loc, dtype=torch.float64, device=dev)
    j1p = torch.zeros(n_loc, dtype=torch.float64, device=dev)
    Bm = n_loc * nS
    send_mask = torch.zeros((per, nS), dtype=torch.uint8, device=dev)
    send_iopt = torch.zeros((per,), dtype=torch.int32, device=dev)
    full_mask = torch.empty((world * per, nS), dtype=torch.uint8, device=dev)
    full_iopt = torch.empty((world * per,), dtype=torch.int32, device=dev)
    holder = {}

    def step():
        o = eng.rollout_cost(prmd, x0, None, Kn, L, G=nS, B=Bm, want_traj=False, want_cost=True, T_total=Kn,
                             j0_prefix=j0p, j1_prefix=j1p, w=w, seed=5, first=lo * nS)
        eng.pareto(o["J0"].view(n_loc, nS), o["J1"].view(n_loc, nS), out=(send_mask[:n_loc], send_iopt[:n_loc]))
        dist.all_gather_into_tensor(full_mask, send_mask)
        dist.all_gather_into_tensor(full_iopt, send_iopt)
        holder["o"] = o

    for _ in range(warm):
        step()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    dist.barrier(); torch.cuda.synchronize()
    tm = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=dev)
    dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    ms = float(tm.item())
    kt = {k: round(v, 4) for k, v in eng.last_kernel_times().items()}   # (the Pareto call)
    eng.rollout_cost(prmd, x0, None, Kn, L, G=nS, B=Bm, want_traj=False, want_cost=True, T_total=Kn,
                     j0_prefix=j0p, j1_prefix=j1p, w=w, seed=5, first=lo * nS)
    torch.cuda.synchronize()
    kt.update({k: round(v, 4) for k, v in eng.last_kernel_times().items()})
    spot = None
    if rank == 0:
        o, ok = _orc(), True
        x0h, wh = x0.cpu().numpy(), w.cpu().numpy()
        for bb in (0, nS // 2, Bm - 1):
            r = bb // nS
            ub = o.random_schedule(5, lo + r, bb % nS, nS, L, Kn, np.zeros(L), reg["npi_max"]).astype(float)
            s_, i_, al_ = o.SIalpha_Controlled(ub, *x0h[r], reg["npi_max"], 1e-8, 100.0, syn.GAMMA, reg["a"][lo + r],
                                               reg["b"][lo + r], syn.BETA, 0.0, 0.0, 0.0, Kn, 1.0)
            j0, j1 = o.NPICost((s_ * i_) * al_, ub, wh[r].T)
            ok = ok and float(holder["o"]["J0"][bb]) == j0 and float(holder["o"]["J1"][bb]) == j1
        spot = {"trajectories": 3, "bit_exact": bool(ok)}
    front_sizes = float(full_mask[:nR].sum(dim=1).double().mean().item())
    units = nR * nS * Kn
    res = {"config": 5, "n_gpus": world, "scaling": "strong", "regions_this_gpu": n_loc, "schedules_per_region": nS, "days": Kn,
           "ms_per_step": ms, "trajectory_days_per_s": units / ms * 1e3, "kernel_ms_rank0": kt,
           "front_sizes_mean": front_sizes, "oracle_spot_check": spot,
           "collective": "all_gather_into_tensor of the per-region front masks (uint8 [regions, schedules]) and knee indices",
           "note": "schedules drawn in the kernel (Philox, seed 5, global counters); costs stay on the rank that made them"}
    holder.clear()
    del send_mask, full_mask
    torch.cuda.empty_cache()
    eng.release_cache()
    return res

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--only", type=int, default=0, help="run a single config (2, 3 or 5)")
    ap.add_argument("--no-check", action="store_true", help="skip the oracle spot checks")
    a = ap.parse_args()
    eng = Engine(0)
    out = measure(eng, a.only, a.scale, spot_check=not a.no_check)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "configs.jsonl"), "w") as f:
        for r in out:
            f.write(json.dumps(r) + "\n")
            print(json.dumps(r))
    eng.close()


if __name__ == "__main__":
    main()
