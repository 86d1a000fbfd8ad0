"""Extract the per-region fixture the synthetic workloads are shaped on.

Run ONCE in the build container (needs /root/reference, which does not exist on
the GPU box); the output is committed.  Source data (reference, read-only):
  xprize-sample-data/prescription_trained_params_nonnegls.mat  (235 regions:
      Country, Region, N_population, b, a[12], b2, a2[12]   -- SURVEY.md section 0.7)
  xprize-sample-data/uniform_random_costs.csv                  (236 x 12 NPI cost weights)
Only numbers and names are extracted; no reference source code is copied.
"""
import csv
import sys

import numpy as np
import scipy.io as sio

REF = "/root/reference/xprize-sample-data"
OUT = "epidemicmodeling_b200/data/regions_nonnegls.npz"


def main():
    cell = sio.loadmat(f"{REF}/prescription_trained_params_nonnegls.mat")["TrainedModelParams"]
    rows = cell[1:]
    names, N, b1, a1, b2, a2 = [], [], [], [], [], []
    for r in rows:
        c = str(r[0][0]) if r[0].size else ""
        g = str(r[1][0]) if r[1].size else ""
        names.append(f"{c}|{g}")
        N.append(float(r[2].ravel()[0]))
        b1.append(float(r[3].ravel()[0]))
        a1.append(np.asarray(r[4], dtype=np.float64).ravel())
        b2.append(float(r[5].ravel()[0]))
        a2.append(np.asarray(r[6], dtype=np.float64).ravel())
    costs = {}
    with open(f"{REF}/uniform_random_costs.csv") as f:
        rd = csv.reader(f)
        next(rd)
        for row in rd:
            costs[f"{row[0]}|{row[1]}"] = np.array([float(v) for v in row[2:14]])
    w = np.stack([costs.get(n, np.ones(12)) for n in names])
    np.savez_compressed(
        OUT, names=np.array(names), N=np.array(N), b1=np.array(b1), a1=np.stack(a1),
        b2=np.array(b2), a2=np.stack(a2), cost_weights=w,
        npi_max=np.array([3, 3, 2, 4, 2, 3, 2, 4, 2, 3, 2, 4], dtype=np.float64))
    print("wrote", OUT, len(names), "regions; cost rows matched:",
          sum(n in costs for n in names))


if __name__ == "__main__":
    sys.exit(main())
