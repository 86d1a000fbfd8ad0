"""Write tests/golden/ref_inputs.mat: the inputs of the parity cases (tests/cases.py) as MATLAB structs,
for oracle/ref_fixtures.m -- the script that runs the UNMODIFIED reference .m files on them under
MATLAB/Octave and saves tests/golden/ref_outputs.mat.  tests/test_oracle.py compares the oracle with that
file when it exists (neither interpreter is available in the build image, so it does not exist yet:
"parity unpinned").

    python tools/make_ref_inputs.py
    octave --eval "cd oracle; ref_fixtures('/path/to/EpidemicModeling')"     # or matlab -batch
    python -m pytest tests/test_oracle.py -k reference_fixtures
"""
import os
import sys

import numpy as np
import scipy.io

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "ref_inputs.mat")


def col(v):
    return np.asarray(v, dtype=np.float64).reshape(-1, 1)


def params_struct(p):
    """The reference's `params` struct (column vectors for w, a, u_min, u_max as the drivers build them)."""
    s = {}
    for k, v in p.items():
        if k == "obs_type":
            s[k] = str(v)
        elif np.ndim(v) >= 1:
            s[k] = col(v)
        else:
            s[k] = float(v)
    return s


def ekf_struct(c):
    x = np.asarray(c["x"], dtype=np.float64).reshape(1, -1)
    R = np.asarray(c["R_v"], dtype=np.float64)
    return dict(u=np.asarray(c["u"], dtype=np.float64), x=x, params=params_struct(c["params"]),
                s_init=col(c["s_init"]), Ps_init=np.asarray(c["Ps_init"], dtype=np.float64),
                s_final=col(c["s_final"]), Ps_final=np.asarray(c["Ps_final"], dtype=np.float64),
                w_bar=col(c["w_bar"]), v_bar=float(c["v_bar"]), Q_w=np.asarray(c["Q_w"], dtype=np.float64),
                R_v=R.reshape(1, -1) if R.ndim == 1 else R, beta=float(c["beta"]), gamma=float(c["gamma"]),
                inv_monitor_len=float(c["inv_monitor_len"]), order=float(c["order"]))


def ekf_cases():
    """name -> (reference function, inputs); the same cases tools/make_golden.py freezes."""
    out = {}
    for v in ("perday", "adaptive", "totalcases", "endpoint"):
        out[f"ekf3_{v}"] = ("SIAlphaModelEKF", cases.ekf3_case(variant=v))
    out["ekf3_flipped"] = ("SIAlphaModelBackwardEKF", cases.ekf3_case(variant="backward"))
    out["ekf6_optctrl"] = ("SIAlphaModelEKFOptControlled", cases.ekf6_case())
    out["ekf6_flipped"] = ("SIAlphaModelBackwardEKFOptControlled", cases.ekf6_case(backward=True))
    out["legacy_tools"] = ("NewCaseEKFEstimatorWithOptimalNPI", cases.legacy_case())
    return out


def main():
    m = {}
    seirp = {}
    for name, kw in cases.seirp_scenarios(short=True).items():
        seirp[name] = {k: (np.asarray(v, dtype=np.float64).reshape(1, -1) if np.ndim(v) else float(v))
                       for k, v in kw.items()}
    m["seirp"] = seirp
    sat = cases.seirp_saturated_case()
    m["seirp_sat"] = {k: (np.asarray(v, dtype=np.float64).reshape(1, -1) if np.ndim(v) else float(v))
                      for k, v in sat.items()}
    ekf = {}
    for name, (fn, c) in ekf_cases().items():
        s = ekf_struct(c)
        s["fn"] = fn
        ekf[name] = s
    m["ekf"] = ekf
    rc = cases.rollout_case()
    m["rollout"] = dict(u=rc["u"], s0=rc["s0"], i0=rc["i0"], alpha0=rc["alpha0"], u_max=col(rc["u_max"]),
                        alpha_min=rc["alpha_min"], alpha_max=rc["alpha_max"], gamma=rc["gamma"], a=col(rc["a"]),
                        b=float(rc["b"]), beta=float(rc["beta"]), s_noise_std=rc["s_noise_std"],
                        i_noise_std=rc["i_noise_std"], alpha_noise_std=rc["alpha_noise_std"], K=float(rc["K"]),
                        dt=rc["dt"],
                        # randn is called in the order s, i, alpha inside the day loop (SIalpha_Controlled.m:25-27)
                        randn_stream=np.asarray(rc["noise"], dtype=np.float64).T.reshape(1, -1),
                        weights=np.outer(np.linspace(0.5, 1.5, 12), np.ones(rc["K"])))
    rt = cases.rt_expfit_case()
    m["rt_expfit"] = dict(x=rt["x"], s_init=col(rt["s_init"]), params=np.asarray(rt["params"]).reshape(1, -1),
                          w_bar=col(rt["w_bar"]), v_bar=float(rt["v_bar"]), Ps_init=rt["Ps_init"], Q_w=rt["Q_w"],
                          R_v=float(rt["R_v"]), beta=rt["beta"], gamma=rt["gamma"],
                          inv_monitor_len=float(rt["inv_monitor_len"]), order=float(rt["order"]))
    scipy.io.savemat(OUT, m, format="5", do_compression=True, oned_as="column")
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
