"""Generate tests/golden/*.npz from the CPU oracle (run once, outputs committed).

The reference ships no numeric golden vectors and neither MATLAB nor Octave is
available ("parity unpinned"), so the goldens freeze the ORACLE's
outputs on the reference's own parameter sets.  They pin (i) the oracle against
drift (tests/test_oracle.py) and (ii) the CUDA path on the GPU box, where
/root/reference does not exist (tests/test_gpu_parity.py).

Usage:  python tools/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cases  # noqa: E402
from oracle import oracle as orc  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
EKF_KEYS = ("u_opt", "u_opt_smooth", "S_MINUS", "S_PLUS", "S_SMOOTH", "P_MINUS", "P_PLUS", "P_SMOOTH",
            "K_GAIN", "innovations", "rho")


def ekf_args(c):
    return (c["u"], c["x"], c["params"], c["s_init"], c["Ps_init"], c["s_final"], c["Ps_final"],
            c["w_bar"], c["v_bar"], c["Q_w"], c["R_v"], c["beta"], c["gamma"], c["inv_monitor_len"],
            c["order"])


def main():
    os.makedirs(OUT, exist_ok=True)
    # SEIRP: final sample of every scenario + the full scenario-A / Y trajectories
    g = {}
    for name, kw in cases.seirp_scenarios(short=True).items():
        out = orc.SEIRP(**kw)
        g[f"{name}_last"] = np.array([o[0, -1] for o in out])
        if name in ("A", "Y"):
            g[f"{name}_full"] = np.concatenate(out)
    sat = orc.SEIRPSaturatedResource(**cases.seirp_saturated_case())
    g["SAT_last"] = np.array([o[0, -1] for o in sat])
    g["SAT_every100"] = np.concatenate(sat)[:, ::100]
    np.savez_compressed(os.path.join(OUT, "seirp.npz"), **g)

    # rollout + cost + Pareto
    rc = cases.rollout_case()
    s, i, al = orc.SIalpha_Controlled(**rc)
    w = np.outer(np.linspace(0.5, 1.5, 12), np.ones(rc["K"]))
    J0, J1 = orc.NPICost(s * i * al, rc["u"], w)
    rng = np.random.default_rng(5)
    pj0, pj1 = rng.random(64), rng.random(64)
    pj0[10] = pj0[3]; pj1[10] = pj1[3]  # an exact tie survives the strict filter
    mask, iopt = orc.pareto(pj0, pj1)
    np.savez_compressed(os.path.join(OUT, "rollout.npz"), s=s, i=i, alpha=al, J=np.array([J0, J1]),
                        pj0=pj0, pj1=pj1, mask=mask, iopt=iopt)

    # EKF/EKS: every model variant
    variants = {
        "ekf3_perday": (orc.SIALPHA, cases.ekf3_case(0, variant="perday")),
        "ekf3_adaptive": (orc.SIALPHA, cases.ekf3_case(1, variant="adaptive")),
        "ekf3_totalcases": (orc.SIALPHA, cases.ekf3_case(2, variant="totalcases")),
        "ekf3_endpoint": (orc.SIALPHA, cases.ekf3_case(3, variant="endpoint")),
        "ekf3_flipped": (orc.SIALPHA_FLIPPED, cases.ekf3_case(4, variant="backward")),
        "ekf6_optctrl": (orc.OPTCTRL, cases.ekf6_case(0)),
        "ekf6_flipped": (orc.OPTCTRL_FLIPPED, cases.ekf6_case(1, backward=True)),
        "legacy_tools": (orc.LEGACY_TOOLS, cases.legacy_case(0)),
        "legacy_codegen": (orc.LEGACY_CODEGEN, cases.legacy_case(1)),
    }
    for name, (model, c) in variants.items():
        o = orc.ekf_eks(model, *ekf_args(c))
        np.savez_compressed(os.path.join(OUT, f"{name}.npz"), **{k: o[k] for k in EKF_KEYS})

    # fused sweep: 3 regions x 12 epsilon
    inp, eps = cases.sweep_case()
    J0s, J1s, masks, iopts, ufs = [], [], [], [], []
    for r in inp:
        s3 = r["setup3"]
        o3 = orc.ekf_eks(orc.SIALPHA, r["u_fixed"], r["x"], s3["params"], s3["s_init"], s3["Ps_init"],
                         s3["s_final"], s3["Ps_final"], s3["w_bar"], s3["v_bar"], s3["Q_w"], r["R_v"],
                         s3["beta_ekf"], s3["gamma_ekf"], s3["W"], 1)
        Th = r["T_hist"]
        S = o3["S_SMOOTH"]
        s6 = r["setup6"]
        reg = orc.SweepRegion(s6["params"], r["T"], Th, r["u_hist"], r["x"], r["R_v"], s6["s_init"],
                              s6["Ps_init"], s6["s_final"], s6["Ps_final"], s6["Q_w"], s6["beta_ekf"],
                              s6["gamma_ekf"], s6["W"], S[0, Th - 1], S[1, Th - 1], S[2, Th - 1],
                              (S[0, :Th] * S[1, :Th]) * S[2, :Th], r["weights"])
        j0, j1, m, io, uf = orc.sweep_region(reg, eps, want_u=True)
        J0s.append(j0); J1s.append(j1); masks.append(m); iopts.append(io); ufs.append(uf)
    np.savez_compressed(os.path.join(OUT, "sweep.npz"), eps=eps, J0=np.array(J0s), J1=np.array(J1s),
                        mask=np.array(masks), iopt=np.array(iopts), u_fore=np.array(ufs))
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
