"""Survivor fraction of the J0-bucket pruning on REAL config-5 costs (30 regions x 100k Philox schedules)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from epidemicmodeling_b200 import synthetic as syn
from epidemicmodeling_b200.engine import Engine, pack_params, params_to_device
nR, nS, Kn, L = 30, 100_000, 120, 12
dev = "cuda:0"
eng = Engine(0); eng.use_torch_stream()
reg = syn.load_regions(236)
t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
prm = pack_params([dict(dt=1.0, beta=syn.BETA, gamma=syn.GAMMA, b=reg["b"][r], a=reg["a"][r], u_max=reg["npi_max"], u_min=np.zeros(L),
                        alpha_min=1e-8, alpha_max=100.0) for r in range(nR)], L)
prmd = params_to_device(prm, dev)
x0 = t(np.array([[(reg["N"][r] - 10) / reg["N"][r], 10 / reg["N"][r], syn.ALPHA0] for r in range(nR)]))
w = t(np.stack([np.repeat(reg["cost_weights"][r][None, :], Kn, axis=0) for r in range(nR)]))
z = torch.zeros(nR, dtype=torch.float64, device=dev)
o = eng.rollout_cost(prmd, x0, None, Kn, L, G=nS, B=nR * nS, want_traj=False, want_cost=True, T_total=Kn, j0_prefix=z, j1_prefix=z, w=w, seed=5)
J0, J1 = o["J0"].view(nR, nS), o["J1"].view(nR, nS)
for _ in range(3):
    m, io = eng.pareto(J0, J1)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    m, io = eng.pareto(J0, J1)
e1.record(); torch.cuda.synchronize()
print("pareto ms", e0.elapsed_time(e1) / 5, "front mean", float(m.sum(1).double().mean()))
for NB in (1024, 4096, 16384):
    lo, hi = J0.min(1, keepdim=True).values, J0.max(1, keepdim=True).values
    b = ((J0 - lo) * (NB / (hi - lo))).clamp(max=NB - 1).long()
    bmin = torch.full((nR, NB), float("inf"), dtype=torch.float64, device=dev)
    bmin.scatter_reduce_(1, b, J1, "amin")
    pm = torch.cummin(torch.cat([torch.full((nR, 1), float("inf"), dtype=torch.float64, device=dev), bmin[:, :-1]], 1), 1).values
    surv = ~(pm.gather(1, b) < J1)
    print("buckets", NB, "survivors per set mean", float(surv.sum(1).double().mean()), "max", int(surv.sum(1).max()))
for NB in (1024, 4096):
    o64 = J0.view(torch.int64)                      # positive doubles: the bit pattern is order preserving
    lo, hi = o64.min(1, keepdim=True).values, o64.max(1, keepdim=True).values
    sh = 20
    b = (((o64 - lo) >> sh) * NB // (((hi - lo) >> sh) + 1)).clamp(max=NB - 1)
    bmin = torch.full((nR, NB), float("inf"), dtype=torch.float64, device=dev)
    bmin.scatter_reduce_(1, b, J1, "amin")
    pm = torch.cummin(torch.cat([torch.full((nR, 1), float("inf"), dtype=torch.float64, device=dev), bmin[:, :-1]], 1), 1).values
    surv2 = ~(pm.gather(1, b) < J1)
    print("bit-space buckets", NB, "survivors per set mean", float(surv2.sum(1).double().mean()), "max", int(surv2.sum(1).max()),
          "both tests", float((surv2 & surv).sum(1).double().mean()))
# held vs per-day halves
print("J0 range", float(J0.min()), float(J0.max()), "J1 range", float(J1.min()), float(J1.max()))
print("distinct J0 per set (mean)", float(np.mean([len(torch.unique(J0[r])) for r in range(3)])))
eng.close()
