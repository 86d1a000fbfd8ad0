"""Jacobi-ordering study on the sweep workload's P_MINUS matrices (CPU, test infrastructure:
uses oracle/).  Counts rotations / sweeps per matrix and the warp-level cost (32 consecutive
epsilons of one region and day share a warp in eks_gain_kernel) for the row-cyclic pair order
and for the round-robin order whose 3 pairs per set are disjoint.

    python tools/jacobi_study.py [n_regions] [T_hist]
"""
import sys, os
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from epidemicmodeling_b200 import synthetic as syn  # noqa: E402
from oracle import oracle as orc  # noqa: E402

REL = 2.0 ** -62

ROW_CYCLIC = [[(p, q)] for p in range(5) for q in range(p + 1, 6)]
ROUND_ROBIN = [[(0, 1), (2, 3), (4, 5)], [(0, 2), (1, 4), (3, 5)], [(0, 4), (1, 3), (2, 5)],
               [(0, 3), (1, 5), (2, 4)], [(0, 5), (1, 2), (3, 4)]]


def jacobi_stats(A, sets, warp=32, local=None):
    """A [N,6,6] symmetric.  Vectorised threshold Jacobi.  Returns per-matrix rotations, sweeps,
    and warp-level executed set count (a set executes for a warp if any lane rotates any pair)."""
    a = A.copy()
    N = a.shape[0]
    rot = np.zeros(N, int)
    sweeps = np.zeros(N, int)
    nw = (N + warp - 1) // warp
    warp_sets = np.zeros(nw, int)
    warp_sweeps = np.zeros(nw, int)
    alive = np.ones(N, bool)
    iu = np.triu_indices(6, 1)
    for sweep in range(30):
        dmax = np.abs(np.einsum("nii->ni", a)).max(1)
        thr = dmax * REL
        off = np.abs(a[:, iu[0], iu[1]]).max(1)
        if local is None:
            alive &= off > thr
        else:
            d = np.abs(np.einsum("nii->ni", a))
            with np.errstate(all="ignore"):
                crit = (a[:, iu[0], iu[1]] ** 2) > local * d[:, iu[0]] * d[:, iu[1]]
            alive &= crit.any(1)
        if not alive.any():
            break
        sweeps[alive] += 1
        wa = np.zeros(nw * warp, bool); wa[:N] = alive
        warp_sweeps += wa.reshape(nw, warp).any(1)
        for st in sets:
            act_set = np.zeros(N, bool)
            for (p, q) in st:
                apq = a[:, p, q]
                if local is None:
                    act = alive & (np.abs(apq) > thr)
                else:
                    with np.errstate(all="ignore"):
                        act = alive & (apq * apq > local * np.abs(a[:, p, p] * a[:, q, q]))
                act_set |= act
                if not act.any():
                    continue
                idx = np.nonzero(act)[0]
                app, aqq, apq = a[idx, p, p], a[idx, q, q], a[idx, p, q]
                with np.errstate(all="ignore"):
                    theta = 0.5 * (aqq - app) / apq
                    t = 1.0 / (np.abs(theta) + np.sqrt(theta * theta + 1.0))
                    t = np.where(theta < 0, -t, t)
                    c = 1.0 / np.sqrt(t * t + 1.0)
                    s = t * c
                a[idx, p, p] = app - t * apq
                a[idx, q, q] = aqq + t * apq
                a[idx, p, q] = 0.0; a[idx, q, p] = 0.0
                for r in range(6):
                    if r in (p, q):
                        continue
                    g, h = a[idx, r, p], a[idx, r, q]
                    gp, hp = c * g - s * h, s * g + c * h
                    a[idx, r, p] = gp; a[idx, p, r] = gp
                    a[idx, r, q] = hp; a[idx, q, r] = hp
                rot[idx] += 1
            wa = np.zeros(nw * warp, bool); wa[:N] = act_set
            warp_sets += wa.reshape(nw, warp).any(1)
    return rot, sweeps, warp_sets, warp_sweeps, a


def main():
    nR = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    Th = int(sys.argv[2]) if len(sys.argv) > 2 else 441
    inputs = syn.sweep_inputs(n_regions=nR, T_hist=Th, T_fore=120)
    eps = syn.epsilon_grid_xprize02(250)
    for r, inp in enumerate(inputs):
        s6 = inp["setup6"]
        u = np.concatenate([inp["u_hist"], np.full((syn.L_NPI, 120), np.nan)], axis=1)
        for e0 in (0, 96, 128, 192):
            mats = []
            for e in eps[e0:e0 + 32]:
                prm = dict(s6["params"]); prm["epsilon"] = float(e)
                out = orc.ekf_eks(orc.OPTCTRL, u, inp["x"], prm, s6["s_init"], s6["Ps_init"], s6["s_final"],
                                  s6["Ps_final"], s6["w_bar"], 0.0, s6["Q_w"], inp["R_v"], 1.0, 0.995, 21, 1)
                mats.append(np.transpose(out["P_MINUS"], (2, 0, 1))[1:])  # [T-1,6,6]
            # warp = 32 eps of one day: order matrices day-major, eps-minor
            A = np.stack(mats, 1).reshape(-1, 6, 6)
            for name, sets, local in (("round-robin", ROUND_ROBIN, None), ("rr local 2^-106", ROUND_ROBIN, 2.0 ** -106),
                                      ("rr local 2^-100", ROUND_ROBIN, 2.0 ** -100)):
                rot, sw, wsets, wsw, ad = jacobi_stats(A, sets, local=local)
                lam = np.sort(np.abs(np.einsum("nii->ni", ad)), 1)
                if local is None:
                    lam_ref = lam
                with np.errstate(all="ignore"):
                    rel = np.nanmax(np.abs(lam - lam_ref) / np.maximum(lam_ref, 1e-300), axis=1)
                print(f"region {r} eps[{e0}:{e0+32}] {name:16s} rot/matrix {rot.mean():6.2f} (max {rot.max()}) "
                      f"sweeps {sw.mean():5.2f} (max {sw.max()})  warp: sets {wsets.mean():6.2f} sweeps {wsw.mean():5.2f} "
                      f"| eig rel diff vs global: median {np.median(rel):.1e} max {rel.max():.1e}")


if __name__ == "__main__":
    main()
