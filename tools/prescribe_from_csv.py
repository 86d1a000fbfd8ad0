"""CSV -> prescription CSV on the GPU, no MATLAB (SURVEY 8f-2 demo).

The Oxford data file the reference downloads is not bundled, so `--make-synthetic` first writes one
in the same wire format (CountryName, RegionName, Date, ConfirmedCases, ConfirmedDeaths, 12 NPI
columns) from the synthetic region histories, with reporting gaps (N/A cells) and a negative
correction in it; the pipeline then reads it back like any OxCGRT_latest.csv.

    python tools/prescribe_from_csv.py --regions 8 --out gpurun_out/prescriptions.csv
"""
import argparse
import datetime
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from epidemicmodeling_b200 import synthetic as syn, xprize_io as xio  # noqa: E402


def synthetic_oxcgrt(path, n_regions, T_hist, start="2020-03-15", seed=9):
    """Write an OxCGRT-format CSV; returns {geo id: dict(N, a, b, weights)}."""
    import pandas as pd
    reg = syn.load_regions(n_regions)
    inp = syn.sweep_inputs(n_regions=n_regions, T_hist=T_hist, T_fore=1)
    rng = np.random.default_rng(seed)
    d0 = datetime.date.fromisoformat(start)
    rows, regions = [], {}
    for r in range(n_regions):
        name = f"{str(reg['names'][r]).split('|')[0].strip()} {r:03d}"   # unique country names
        N = float(reg["N"][r])
        new = np.round(np.nan_to_num(inp[r]["x"][:T_hist]) * N)
        cc = np.cumsum(new)
        cc[rng.integers(20, T_hist - 5)] -= 3.0                     # a downward correction (negative diff)
        cc_out = cc.astype(object)
        for t in rng.integers(5, T_hist - 1, 3):
            cc_out[t] = ""                                          # reporting gaps
        if r % 3 == 0:
            cc_out[T_hist - 1] = ""                                 # missing last day (:168-171)
        ip = inp[r]["u_hist"].T.astype(object)                      # [T, L]
        for _ in range(6):
            ip[rng.integers(1, T_hist), rng.integers(0, 12)] = ""   # N/A NPI cells (:121-128)
        ip[0, r % 12] = ""
        for t in range(T_hist):
            date = (d0 + datetime.timedelta(days=t)).strftime("%Y%m%d")
            rows.append([name, "", date, cc_out[t], ""] + list(ip[t]))
        regions[xio.geo_id(name, "")] = dict(N=N, a=reg["a"][r], b=float(reg["b"][r]), weights=reg["cost_weights"][r])
    pd.DataFrame(rows, columns=["CountryName", "RegionName", "Date", "ConfirmedCases", "ConfirmedDeaths"]
                 + xio.NPI_COLUMNS).to_csv(path, index=False)
    end = (d0 + datetime.timedelta(days=T_hist - 1)).isoformat()
    return regions, start, end


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--regions", type=int, default=8)
    ap.add_argument("--t-hist", type=int, default=120)
    ap.add_argument("--t-fore", type=int, default=30)
    ap.add_argument("--eps", type=int, default=50)
    ap.add_argument("--data", default=os.path.join(ROOT, "gpurun_out", "synthetic_oxcgrt.csv"))
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "prescriptions.csv"))
    a = ap.parse_args()
    os.makedirs(os.path.dirname(a.data), exist_ok=True)
    regions, start, end = synthetic_oxcgrt(a.data, a.regions, a.t_hist)
    from epidemicmodeling_b200.engine import Engine
    from epidemicmodeling_b200 import pipeline
    eng = Engine(0)
    d1 = datetime.date.fromisoformat(end)
    fdates = [(d1 + datetime.timedelta(days=k + 1)).isoformat() for k in range(a.t_fore)]
    res = pipeline.prescribe_from_csv(eng, a.data, start, end, a.t_fore, syn.epsilon_grid_xprize02(a.eps), regions,
                                      out_file=a.out, forecast_dates=fdates)
    print(f"{len(res['ids'])} regions x {a.eps} eps: front sizes {res['on_front'].sum(axis=1).tolist()}, "
          f"knee indexes {res['I_opt'].tolist()}")
    print("kernels (ms):", {k: round(v, 3) for k, v in eng.last_kernel_times().items()})
    print("wrote", a.out)
    eng.close()


if __name__ == "__main__":
    main()
