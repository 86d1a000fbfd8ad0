"""Diagnostic: where does the host-memory (e2e) sweep call spend its time?"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from epidemicmodeling_b200 import synthetic as syn, workloads as wl
from epidemicmodeling_b200.engine import Engine

eng = Engine(0)
mode = sys.argv[1] if len(sys.argv) > 1 else "own"
if mode == "legacy":
    eng.use_torch_stream()
inp = syn.sweep_inputs(n_regions=236, T_hist=441, T_fore=120)
eps = syn.epsilon_grid_xprize02(250)
S = wl.run_fixed_input(eng, inp)
batch = wl.sweep_batch(inp, S)
hb = dict(batch)
for k in wl._SWEEP_ARRAYS:
    hb[k] = torch.from_numpy(np.ascontiguousarray(batch[k])).pin_memory().numpy()
for it in range(40):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = wl.run_sweep(eng, hb, eps)
    t1 = time.perf_counter()
    kt = eng.last_kernel_times()
    print(f"[{mode}] iter {it}: wall {1e3*(t1-t0):.2f} ms; kernels {sum(kt.values()):.2f} ms", flush=True)
eng.close()
