"""Diagnostic: where does the blocking host-memory (e2e) sweep call spend its wall time?
Run with EPI_TRACE_HOST=1 to get the library's own host-side timeline per call on stderr."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from epidemicmodeling_b200 import synthetic as syn, workloads as wl
from epidemicmodeling_b200.engine import Engine

eng = Engine(0)
NREG = int(os.environ.get("DIAG_REGIONS", "236"))
inp = syn.sweep_inputs(n_regions=NREG, T_hist=441, T_fore=120)
eps = syn.epsilon_grid_xprize02(250)
S = wl.run_fixed_input(eng, inp)
batch = wl.sweep_batch(inp, S)
hb = dict(batch)
for k in wl._SWEEP_ARRAYS:
    hb[k] = torch.from_numpy(np.ascontiguousarray(batch[k])).pin_memory().numpy()
peps = torch.from_numpy(np.ascontiguousarray(eps)).pin_memory().numpy()
nR, nE = NREG, 250
hout = {"J0": torch.empty((nR, nE), dtype=torch.float64).pin_memory().numpy(),
        "J1": torch.empty((nR, nE), dtype=torch.float64).pin_memory().numpy(),
        "on_front": torch.empty((nR, nE), dtype=torch.uint8).pin_memory().numpy(),
        "I_opt": torch.empty((nR,), dtype=torch.int32).pin_memory().numpy()}
# raw pinned H2D of the same volume, for scale
src = torch.empty(28_700_000 // 8, dtype=torch.float64).pin_memory()
dst = torch.empty_like(src, device="cuda")
for it in range(5):
    torch.cuda.synchronize(); t0 = time.perf_counter(); dst.copy_(src, non_blocking=True); torch.cuda.synchronize()
    print(f"raw pinned H2D 28.7 MB: {1e3*(time.perf_counter()-t0):.3f} ms", flush=True)
print("affinity", len(os.sched_getaffinity(0)), "cpus; loadavg", os.getloadavg(), flush=True)
for it in range(int(os.environ.get("DIAG_ITERS", "30"))):
    t0 = time.perf_counter()
    r = wl.run_sweep(eng, hb, peps, out=hout)
    t1 = time.perf_counter()
    kt = eng.last_kernel_times()
    print(f"iter {it}: wall {1e3*(t1-t0):.2f} ms; kernels {sum(kt.values()):.2f} ms", flush=True)
print("loadavg", os.getloadavg(), flush=True)
eng.close()
