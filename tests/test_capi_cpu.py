"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and
exports every symbol include/epi_b200.h declares; the ctypes structs match the C
layout; without a GPU the product path FAILS LOUDLY (no CPU fallback); the
host-side sharding/gather logic works at world_size 2 over gloo."""
import ctypes as C
import os
import re
import subprocess
import sys
import textwrap

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from epidemicmodeling_b200 import _build, _capi
    _build.build()
    return _capi.load()


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "epi_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(epi_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol(lib):
    from epidemicmodeling_b200 import _capi
    declared = _declared_symbols()
    assert len(declared) >= 16
    for name in declared:
        assert hasattr(lib, name), f"libepi_b200.so lacks {name}"
    assert sorted(_capi.SYMBOLS) == declared  # the binding covers the header exactly


def test_ctypes_structs_match_c_layout(lib, tmp_path):
    """Compile a tiny C program against the header and compare sizeof/offsetof."""
    from epidemicmodeling_b200 import _capi as K
    probes = {
        "epi_model_params": (K.ModelParams, ["dt", "a", "w", "L", "obs_type"]),
        "epi_seirp_args": (K.SeirpArgs, ["rates", "saturated", "i_0", "out"]),
        "epi_rollout_args": (K.RolloutArgs, ["prm", "u_kind", "u", "T_total", "J1"]),
        "epi_npicost_args": (K.NpiCostArgs, ["newcases", "J1"]),
        "epi_si_args": (K.SiArgs, ["dt", "alpha", "i"]),
        "epi_ekf_args": (K.EkfArgs, ["prm", "u", "R", "Q", "v_bar", "W", "u_opt", "rho", "status"]),
        "epi_pareto_args": (K.ParetoArgs, ["J0", "I_opt"]),
        "epi_sweep_args": (K.SweepArgs, ["prm", "beta_ekf", "W", "x0", "noise", "J0", "P_first"]),
    }
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "epi_b200.h"', "int main(){"]
    for cname, (_, fields) in probes.items():
        lines.append(f'printf("{cname} %zu\\n", sizeof({cname}));')
        for f in fields:
            lines.append(f'printf("{cname}.{f} %zu\\n", offsetof({cname}, {f}));')
    lines.append("return 0;}")
    src = tmp_path / "probe.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "probe"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = dict(l.split() for l in subprocess.check_output([str(exe)], text=True).splitlines())
    for cname, (st, fields) in probes.items():
        assert int(got[cname]) == C.sizeof(st), cname
        for f in fields:
            assert int(got[f"{cname}.{f}"]) == getattr(st, f).offset, (cname, f)


def test_no_cpu_fallback_without_device(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from epidemicmodeling_b200 import _capi as K
    h = C.c_void_p()
    rc = lib.epi_create(0, C.byref(h))
    assert rc == K.ERR_NO_DEVICE and not h.value
    assert b"no CPU fallback" in lib.epi_last_error(None)
    from epidemicmodeling_b200 import api
    with pytest.raises(K.EpiError):
        api.SEIRP(0.65, 0.005, 0.05, 0.08, 0.1, 0.02, 0.0, 1 - 1e-6, 1e-6, 0, 0, 0, 50, 0.1)


def test_product_package_never_imports_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may touch oracle/."""
    pkg = os.path.join(ROOT, "epidemicmodeling_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in txt and "from oracle" not in txt, f
                assert "epi_oracle" not in txt, f
    for f in os.listdir(os.path.join(ROOT, "matlab")):
        assert "oracle" not in open(os.path.join(ROOT, "matlab", f), errors="ignore").read(), f


def test_pack_params_and_qr_dispatch():
    from epidemicmodeling_b200 import api
    from epidemicmodeling_b200 import _capi as K
    from epidemicmodeling_b200.engine import pack_params
    p = pack_params([dict(dt=1.0, beta=0.2, gamma=0.1, b=0.05, alpha_min=1e-8, alpha_max=100.0,
                          a=np.arange(12.0), u_max=np.ones(12), w=np.arange(24.0).reshape(12, 2),
                          obs_type="TOTALCASES")], 12)
    assert p[0].L == 12 and p[0].obs_type == K.OBS_TOTALCASES
    assert list(p[0].w) == list(np.arange(0, 24, 2.0))      # a 12xT `w` keeps column 1 only
    with pytest.raises(ValueError, match="unknown observation type"):
        pack_params([dict(obs_type="DEATHS")], 12)
    T, m = 40, 3
    q, Qf, r, fixed, Rf = api._classify_QR(np.eye(3) * 2.0, np.ones((1, T)), T, m)
    assert (q, r, fixed) == (K.Q_CONST, K.R_PERDAY, 0) and Qf.size == 9 and Rf.size == T
    q, Qf, r, fixed, Rf = api._classify_QR(np.ones(T), np.array([[0.5]]), T, m)
    assert (q, r, fixed) == (K.Q_PERDAY_SCALAR, K.R_CONST, 1)
    q, Qf, r, fixed, Rf = api._classify_QR(np.ones((3, 3, T)), np.ones((1, 1, T)), T, m)
    assert (q, r, fixed) == (K.Q_PERDAY_FULL, K.R_PERDAY, 1)  # square pages => fixed_R stays true (:79-81)
    with pytest.raises(ValueError, match="Process noise"):
        api._classify_QR(np.ones(7), np.array([[0.5]]), T, m)
    with pytest.raises(ValueError, match="Observation noise"):
        api._classify_QR(np.eye(3), np.ones(7), T, m)


def test_shard_regions_covers_everything():
    from epidemicmodeling_b200.workloads import shard_regions
    for n in (1, 7, 236):
        for ws in (1, 2, 4, 8):
            spans = [shard_regions(n, ws, r) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
                assert a1 == b0 and a0 <= a1


_GLOO_SWEEP_SCRIPT = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
    import numpy as np, torch, torch.distributed as dist
    import cases
    from oracle import oracle as orc            # stand-in compute for the CPU test (no GPU here)
    from epidemicmodeling_b200.workloads import sharded_sweep
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
    rank = dist.get_rank()
    inp, eps = cases.sweep_case(n_regions=3, n_eps=4, T_hist=25, T_fore=10)

    def compute(local, eps):
        J0s, J1s = [], []
        for r in local:
            s3, s6, Th = r["setup3"], r["setup6"], r["T_hist"]
            S = orc.ekf_eks(orc.SIALPHA, r["u_fixed"], r["x"], s3["params"], s3["s_init"], s3["Ps_init"], s3["s_final"],
                            s3["Ps_final"], s3["w_bar"], 0.0, s3["Q_w"], r["R_v"], 1.0, s3["gamma_ekf"], s3["W"], 1)["S_SMOOTH"]
            reg = orc.SweepRegion(s6["params"], r["T"], Th, r["u_hist"], r["x"], r["R_v"], s6["s_init"], s6["Ps_init"],
                                  s6["s_final"], s6["Ps_final"], s6["Q_w"], 1.0, 0.995, 21, S[0, Th - 1], S[1, Th - 1],
                                  S[2, Th - 1], (S[0, :Th] * S[1, :Th]) * S[2, :Th], r["weights"])
            j0, j1, _, _, _ = orc.sweep_region(reg, eps)
            J0s.append(j0); J1s.append(j1)
        return torch.from_numpy(np.array(J0s).reshape(len(local), len(eps))), torch.from_numpy(np.array(J1s).reshape(len(local), len(eps)))

    g0, g1 = sharded_sweep(compute, inp, eps, rank, 2)
    f0, f1 = compute(inp, eps)                   # the unsharded answer
    assert torch.equal(g0, f0) and torch.equal(g1, f1), rank
    dist.barrier(); dist.destroy_process_group()
    print("ok", rank)
""")

_GLOO_SCRIPT = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, {root!r})
    import torch, torch.distributed as dist
    from epidemicmodeling_b200.workloads import shard_regions, gather_costs
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
    rank, n, ne = dist.get_rank(), 7, 5            # ragged: 4 + 3 regions
    full0 = torch.arange(n * ne, dtype=torch.float64).reshape(n, ne)
    full1 = -full0
    lo, hi = shard_regions(n, 2, rank)
    g0, g1 = gather_costs(full0[lo:hi].clone(), full1[lo:hi].clone())
    assert torch.equal(g0, full0) and torch.equal(g1, full1), (rank, g0)
    dist.barrier()
    dist.destroy_process_group()
    print("ok", rank)
""")


def test_gather_costs_world_size_2_gloo(tmp_path):
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "gloo_gather.py"
    script.write_text(_GLOO_SCRIPT.format(root=ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=180)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o



_GLOO_C5_SCRIPT = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, {root!r})
    import numpy as np, torch, torch.distributed as dist
    from oracle import oracle as orc            # stand-in compute for the CPU test (no GPU here)
    from epidemicmodeling_b200 import synthetic as syn
    from epidemicmodeling_b200.workloads import shard_regions, gather_fronts
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
    rank = dist.get_rank()
    nR, nS, Kn, L, seed = 3, 14, 9, 12, 5         # ragged: 2 + 1 regions
    reg = syn.load_regions(nR)

    def score(regions):
        # the schedules are a function of the GLOBAL (region, scenario) pair: a shard draws exactly what the whole job draws
        masks, knees = [], []
        for r in regions:
            J0, J1 = np.zeros(nS), np.zeros(nS)
            for sc in range(nS):
                u = orc.random_schedule(seed, r, sc, nS, L, Kn, np.zeros(L), reg["npi_max"]).astype(float)
                s_, i_, al_ = orc.SIalpha_Controlled(u, (reg["N"][r] - 10) / reg["N"][r], 10 / reg["N"][r], syn.ALPHA0,
                                                     reg["npi_max"], 1e-8, 100.0, syn.GAMMA, reg["a"][r], reg["b"][r], syn.BETA,
                                                     0.0, 0.0, 0.0, Kn, 1.0)
                J0[sc], J1[sc] = orc.NPICost((s_ * i_) * al_, u, np.repeat(reg["cost_weights"][r][:, None], Kn, axis=1))
            m, k = orc.pareto(J0, J1)
            masks.append(m.astype(np.uint8)); knees.append(k)
        return torch.from_numpy(np.array(masks, dtype=np.uint8).reshape(len(regions), nS)), torch.tensor(knees, dtype=torch.int32)

    lo, hi = shard_regions(nR, 2, rank)
    m_loc, k_loc = score(range(lo, hi))
    m_all, k_all = gather_fronts(m_loc, k_loc)
    m_ref, k_ref = score(range(nR))               # the unsharded answer
    assert torch.equal(m_all, m_ref) and torch.equal(k_all, k_ref), rank
    dist.barrier(); dist.destroy_process_group()
    print("ok", rank)
""")


def test_config5_front_gather_world_size_2_gloo(tmp_path):
    """BASELINE config 5 at N > 1 (bench.py `config5_sharded`): regions sharded over the ranks, schedules drawn from
    global counters, per-region Pareto on the owning rank, ONE all-gather of front masks and knees -- equal to the
    unsharded job (oracle as the stand-in compute on the CPU)."""
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "gloo_c5.py"
    script.write_text(_GLOO_C5_SCRIPT.format(root=ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o


def test_mex_gateway_type_checks_against_stub_header():
    """matlab/epi_mex.cpp cannot be built here (no MATLAB/Octave); at least prove it is well-formed C++
    against the documented MEX API subset (tests/stubs/mex.h) and the real C-ABI header."""
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-Werror", "-I", os.path.join(ROOT, "tests", "stubs"),
                        "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "matlab", "epi_mex.cpp")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    # every reference signature named in SURVEY 8b has a shim of the same name
    for name in ("SEIRP", "SEIRPSaturatedResource", "SIAlphaModelEKF", "SIAlphaModelBackwardEKF",
                 "SIAlphaModelEKFOptControlled", "SIAlphaModelBackwardEKFOptControlled",
                 "GenericExtendedKalmanFilter", "NewCaseEKFEstimatorWithOptimalNPI", "SIalpha_Controlled",
                 "SI_Controlled", "NPICost"):
        txt = open(os.path.join(ROOT, "matlab", name + ".m")).read()
        assert txt.lstrip().startswith("function") and f"= {name}(" in txt.splitlines()[0], name


def test_sharded_sweep_world_size_2_gloo(tmp_path):
    """The N > 1 path end to end on CPU: shard regions, compute (oracle stand-in), all-gather (gloo)."""
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "gloo_sweep.py"
    script.write_text(_GLOO_SWEEP_SCRIPT.format(root=ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o


def test_missing_library_fails_loudly(tmp_path):
    """No silent fallback: without the built .so the product import path raises."""
    code = ("import os, sys; sys.path.insert(0, %r); os.environ['EPI_B200_LIB'] = %r; "
            "from epidemicmodeling_b200 import _capi\n"
            "try:\n    _capi.load()\nexcept ImportError as e:\n    print('RAISED', e)\n" % (ROOT, str(tmp_path / "nope.so")))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert "RAISED" in out.stdout and "no cpu fallback" in out.stdout.lower(), out.stdout + out.stderr
