"""Per-kernel times of the sorted Pareto path: python tools/pareto_probe.py n_sets n  (run under ncu for the launch list)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from epidemicmodeling_b200.engine import Engine
n_sets, n = int(sys.argv[1]), int(sys.argv[2])
eng = Engine(0); eng.use_torch_stream()
g = torch.Generator(device="cuda").manual_seed(1)
J0 = torch.rand((n_sets, n), dtype=torch.float64, device="cuda", generator=g)
J1 = 1.0 / (J0 + 0.05) + 0.1 * torch.rand((n_sets, n), dtype=torch.float64, device="cuda", generator=g)
for _ in range(3):
    m, io = eng.pareto(J0, J1)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    m, io = eng.pareto(J0, J1)
e1.record(); torch.cuda.synchronize()
print("pareto", n_sets, "x", n, "ms per call", e0.elapsed_time(e1) / 5, "front mean", float(m.sum(dim=1).double().mean()))
eng.close()
