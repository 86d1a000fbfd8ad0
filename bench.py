#!/usr/bin/env python
"""bench.py -- the optimal-NPI Pareto sweep (BASELINE.json config 4, the configuration
the headline metric "EKF/EKS+optimal-NPI trajectory-days/sec" is quoted on).

One "step" = one pass of the hot path over one batch: for every (region, epsilon) the
6-state EKF + fixed-interval smoother over T = T_hist + T_fore days, the SIalpha rollout
of the smoothed schedule, NPICost, then the per-region Pareto front + knee, and (N > 1)
the path's single collective: an all-gather of the per-shard (J0, J1).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

  value : trajectory-days/s, whole job, inputs resident in HBM (device-memory C-ABI mode)
  e2e   : the same through the blocking host-memory C-ABI call (pinned host buffers,
          H2D + D2H inside the timed region) -- the call a MATLAB/Octave host makes
  roofline / cpu_baseline : see DESIGN.md "Measurement"
Multi-GPU (torchrun, one process per GPU): STRONG scaling -- the ONE 236-region sweep is sharded
by region (contiguous blocks of ceil(236/N) regions, workloads.shard_regions), each rank keeps its
shard's inputs resident, and a step ends with the path's single collective, an all-gather of the
per-shard (J0, J1) over NCCL, followed by the Pareto step on the gathered costs.  `value` =
236 x 250 x 561 trajectory-days / step time (max over ranks).  The replicated (weak) form of round 1
is reported under "replica_weak" for reference.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# canonical work per trajectory-day (DESIGN.md "Measurement")
FLOPS_EKF_EKS_M6 = 4758.0
BYTES_FUSED_M6 = 864.0            # packed tape written once + read once
KERNEL_ALGO = {                   # per trajectory-day: (algorithmic bytes, canonical flops)
    "ekf_forward": (432.0, 2314.0),   # tape write (S-, S+, packed P-, P+)
    "eks_gain": (384.0, 1310.0),      # tape read: packed P-, P+ and S+ (S- is the backward pass's); P+A' + pinv + product + Jacobian
    "ekf_forward_piped": (432.0, 2314.0),  # small shards: the forward pass with the gains of its finished chunks beside it
    "eks_gain_tail": (0.0, 0.0),           # ... and what is left of the gains when it ends (csrc/capi.cu, piped schedule)
    "eks_backward": (48.0, 204.0),    # S- read; S_SMOOTH recursion + schedule; P_SMOOTH is not needed for (J0, J1)
    "rollout_cost": (0.0, 91.0),
    "pareto": (0.0, 0.0),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--regions", type=int, default=236)
    ap.add_argument("--eps", type=int, default=250)
    ap.add_argument("--t-hist", type=int, default=441)
    ap.add_argument("--t-fore", type=int, default=120)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-lean", action="store_true", help="skip the extra lean-mode leg (profiling runs)")
    ap.add_argument("--no-secondary", action="store_true", help="skip the configs 2/3/5 block")
    ap.add_argument("--no-replica", action="store_true", help="N > 1: skip the replicated (weak-scaling) extra leg")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU-baseline sample time")
    return ap.parse_args()


def workload_name(a):
    return (f"optimal-NPI Pareto sweep (testPrescribeXPRIZE02 shape): {a.regions} regions x {a.eps} eps x "
            f"12 NPIs x ({a.t_hist}+{a.t_fore}) days")


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.proc:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


class NvmlSampler:
    """The same samples through NVML (nvidia_ml_py) from a thread: a query takes well under a
    millisecond, so even a 100 ms timed region gets dozens of samples (the nvidia-smi process above
    needs longer than that just to start).  Falls back to ClockSampler when NVML is unavailable."""
    REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))

    def __init__(self, index):
        self.index, self.h, self.fallback = index, None, None
        self.sm, self.mx, self.mask, self._stop, self.thread = [], None, 0, threading.Event(), None
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            self.nv = pynvml
            try:  # the CUDA ordinal is not the NVML index under CUDA_VISIBLE_DEVICES: go through the UUID
                uuid = "GPU-" + str(torch.cuda.get_device_properties(index).uuid)
                self.h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode() if hasattr(uuid, "encode") else uuid)
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.h, self.fallback = None, ClockSampler(index)

    def _one(self):
        nv = self.nv
        self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
        get = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        self.mask |= int(get(self.h))

    def _run(self):
        while not self._stop.is_set():
            try:
                self._one()
            except Exception:
                pass
            self._stop.wait(0.004)

    def start(self):
        if self.fallback:
            return self.fallback.start()
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def stop(self):
        if self.fallback:
            return self.fallback.stop()
        self._stop.set()
        self.thread.join(timeout=2)
        out = {"sm_mhz": None, "sm_max_mhz": self.mx, "reasons": [], "samples": len(self.sm), "source": "nvml"}
        if self.sm:
            out["sm_mhz"] = float(np.median(self.sm))
            out["reasons"] = sorted(name for bit, name in self.REASONS if self.mask & bit)
        return out


def build_inputs(a, rank=0):
    """The ONE synthetic workload (same seeds on every rank); `rank` only offsets the seeds of the
    replicated extra leg."""
    from epidemicmodeling_b200 import synthetic as syn
    inp = syn.sweep_inputs(n_regions=a.regions, T_hist=a.t_hist, T_fore=a.t_fore, seed_u=2 + 100003 * rank,
                           seed_x=3 + 100003 * rank)
    eps = syn.epsilon_grid_xprize02(a.eps)
    return inp, eps


def host_threads():
    """Threads the CPU arm may use: the process's affinity mask, NOT OpenMP's default --
    torchrun exports OMP_NUM_THREADS=1 to its children, which round 1's reference arm obeyed."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


# ----------------------------------------------------------------------------- CPU reference arm
def oracle_regions(inp, S_fixed):
    """OrcSweepRegion objects from the inputs and a [T,3,nR] fixed-input smoother result."""
    from oracle import oracle as orc
    regs = []
    for r, rin in enumerate(inp):
        s6, Th = rin["setup6"], rin["T_hist"]
        Sr = S_fixed[:, :, r].T
        regs.append(orc.SweepRegion(s6["params"], rin["T"], Th, rin["u_hist"], rin["x"], rin["R_v"], s6["s_init"],
                                    s6["Ps_init"], s6["s_final"], s6["Ps_final"], s6["Q_w"], s6["beta_ekf"],
                                    s6["gamma_ekf"], s6["W"], Sr[0, Th - 1], Sr[1, Th - 1], Sr[2, Th - 1],
                                    (Sr[0, :Th] * Sr[1, :Th]) * Sr[2, :Th], rin["weights"]))
    return regs


def oracle_fixed_input(inp):
    from oracle import oracle as orc
    out = []
    for rin in inp:
        s3 = rin["setup3"]
        o3 = orc.ekf_eks(orc.SIALPHA, rin["u_fixed"], rin["x"], s3["params"], s3["s_init"], s3["Ps_init"],
                         s3["s_final"], s3["Ps_final"], s3["w_bar"], 0.0, s3["Q_w"], rin["R_v"], s3["beta_ekf"],
                         s3["gamma_ekf"], s3["W"], 1)
        out.append(o3["S_SMOOTH"].T)
    return np.stack(out, axis=2)  # [T,3,nR]


def cpu_sample(a, inp, eps, seconds):
    """Size a bounded sample of the workload (the first n regions x all epsilon) for the CPU oracle
    on all host threads (OpenMP over trajectories).  Returns (n, cores, seconds for one region)."""
    from oracle import oracle as orc
    cores = host_threads()
    probe = inp[:1]
    regs = oracle_regions(probe, oracle_fixed_input(probe))
    t0 = time.perf_counter()
    orc.sweep_batch(regs, eps, n_threads=cores)
    t1 = time.perf_counter() - t0
    # one region alone cannot use more threads than it has trajectories; scale the estimate accordingly
    n = int(max(1, min(len(inp), round(seconds / max(t1, 1e-6)))))
    return n, cores, t1


def run_reference(a):
    """--impl reference: the reference's CPU implementation of the path.  Octave/MATLAB do not
    exist in this image, so this is the C oracle (a line-by-line port of the .m files) with all
    host threads -- kind "port".  Each step = a bounded sample (n regions x all epsilon).
    Under torchrun rank 0 alone runs it, with every host thread the process may use."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import oracle as orc
    inp, eps = build_inputs(a)
    T = a.t_hist + a.t_fore
    budget = max(2.0, min(a.cpu_seconds, 120.0 / max(1, a.steps + a.warmup)))
    n, cores, _ = cpu_sample(a, inp, eps, budget)
    sample = inp[:n]
    regs = oracle_regions(sample, oracle_fixed_input(sample))
    for _ in range(a.warmup):
        orc.sweep_batch(regs, eps, n_threads=cores)
    t0 = time.perf_counter()
    for _ in range(a.steps):
        orc.sweep_batch(regs, eps, n_threads=cores)
    dt = (time.perf_counter() - t0) / max(1, a.steps)
    units = n * a.eps * T
    val = units / dt
    sample_s = f"{n} of {a.regions} regions x {a.eps} eps x {T} days per step (OpenMP over trajectories)"
    line = {"impl": "reference", "metric": "trajectory_days_per_sec", "value": val, "unit": "trajectory-days/s",
            "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": dt * 1e3,
            "higher_is_better": True, "scaling": "strong" if a.gpus > 1 else "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(a), "note": "CPU oracle port of the reference .m files; "
                       "Octave/MATLAB absent from the image; a CPU job: the same host threads whatever --gpus says"},
            "cpu_baseline": {"value": val, "unit": "trajectory-days/s", "cores": cores, "kind": "port",
                             "sample": sample_s, "omp_num_threads_env": os.environ.get("OMP_NUM_THREADS")},
            "e2e": {"value": val, "unit": "trajectory-days/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------- our arm
def timed_steps(step, n, fence, drain=None):
    """CUDA-event time of n calls of step() on torch's current stream, fenced on both sides."""
    import torch
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fence()
    e0.record()
    for _ in range(n):
        step()
    if drain:
        drain()
    e1.record()
    fence()
    return e0.elapsed_time(e1)


def main():
    a = parse()
    if a.impl == "reference":
        return run_reference(a)

    import torch
    import torch.distributed as dist
    from epidemicmodeling_b200 import workloads as wl
    from epidemicmodeling_b200.engine import Engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the engine has no CPU fallback")
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))

    def fence():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    T = a.t_hist + a.t_fore
    nR, nE = a.regions, a.eps
    inp_all, eps = build_inputs(a)
    # strong scaling: this rank's contiguous block of regions (TrainPredictPrescribeNPI.m:93 is the loop)
    lo, hi = wl.shard_regions(nR, world, rank)
    per = (nR + world - 1) // world
    inp = inp_all[lo:hi]
    n_loc = hi - lo
    eng = Engine(local)
    eng.use_torch_stream()

    # driver glue before the sweep (TrainPredictPrescribeNPI.m:377-392): fixed-input 3-state EKF/EKS
    S_fixed = wl.run_fixed_input(eng, inp)
    batch = wl.sweep_batch(inp, S_fixed)
    dbatch = wl.sweep_to_device(batch, eps, dev)
    f64, u8, i32 = torch.float64, torch.uint8, torch.int32

    # Output buffers.  N = 1: epi_sweep writes J0/J1/front/knee itself.  N > 1: the shard's costs go into the first
    # n_loc rows of a [per, nE] send buffer (double-buffered); the all-gather lands them in [N*per, nE], whose first
    # nR rows are the full (J0, J1) because shards are contiguous blocks; the Pareto step runs on those.
    nbuf = 2 if world > 1 else 1
    loc = [{"J0": torch.zeros((per, nE), dtype=f64, device=dev), "J1": torch.zeros((per, nE), dtype=f64, device=dev)}
           for _ in range(nbuf)]
    full = [{"J0": torch.empty((world * per, nE), dtype=f64, device=dev),
             "J1": torch.empty((world * per, nE), dtype=f64, device=dev),
             "on_front": torch.empty((nR, nE), dtype=u8, device=dev),
             "I_opt": torch.empty((nR,), dtype=i32, device=dev)} for _ in range(nbuf)]
    out1 = {"J0": full[0]["J0"][:nR], "J1": full[0]["J1"][:nR], "on_front": full[0]["on_front"], "I_opt": full[0]["I_opt"]}
    comm = torch.cuda.Stream(device=dev) if world > 1 else None
    eng_tail = Engine(local) if world > 1 else None      # its own context: the gather/Pareto tail runs on the side stream
    comm_ev = [None] * nbuf
    step_no = [0]

    def sweep_local(i, b=dbatch, e=None):
        if world == 1:
            return wl.run_sweep(eng, b, e, out=out1)
        return wl.run_sweep(eng, b, e, out={"J0": loc[i]["J0"][:n_loc], "J1": loc[i]["J1"][:n_loc]}, want_front=False)

    def tail(i):
        """the path's single collective + the Pareto step on the gathered costs (current stream)"""
        dist.all_gather_into_tensor(full[i]["J0"].view(-1), loc[i]["J0"].view(-1))
        dist.all_gather_into_tensor(full[i]["J1"].view(-1), loc[i]["J1"].view(-1))
        eng_tail.pareto(full[i]["J0"][:nR], full[i]["J1"][:nR], out=(full[i]["on_front"], full[i]["I_opt"]))

    def tail_async(i):
        ready = torch.cuda.Event()
        ready.record()
        with torch.cuda.stream(comm):
            comm.wait_event(ready)
            tail(i)
            done = torch.cuda.Event()
            done.record(comm)
        comm_ev[i] = done

    def drain():
        for e in comm_ev:
            if e is not None:
                torch.cuda.current_stream().wait_event(e)

    def step():
        """one pass of the hot path; N > 1: the tail of step i overlaps the sweep of step i+1"""
        i = step_no[0] % nbuf
        step_no[0] += 1
        if comm_ev[i] is not None:      # buffer set i is free once its previous tail has finished
            torch.cuda.current_stream().wait_event(comm_ev[i])
        sweep_local(i)
        if world > 1:
            tail_async(i)

    def step_serial():
        """the same with the tail on the launching stream: the latency of ONE sweep"""
        sweep_local(0)
        if world > 1:
            tail(0)

    W = max(3, a.warmup)
    for _ in range(W):
        step()
    drain()
    fence()
    launches0 = eng.launch_count + (eng_tail.launch_count if eng_tail else 0)
    sampler = NvmlSampler(local)
    sampler.start()
    ms_total = timed_steps(step, a.steps, fence, drain)
    clocks = sampler.stop()
    launches = eng.launch_count + (eng_tail.launch_count if eng_tail else 0) - launches0
    ms_step = max_over_ranks(ms_total) / a.steps
    units_total = nR * nE * T
    units_rank = n_loc * nE * T
    value = units_total / (ms_step * 1e-3)

    # per-kernel durations (CUDA events recorded by the library on the launching stream): extra untimed steps
    ktimes, per_kernel_runs = {}, []
    drain()
    for _ in range(3):
        sweep_local(0)
        torch.cuda.synchronize()
        per_kernel_runs.append(eng.last_kernel_times())
    for k in per_kernel_runs[0]:
        ktimes[k] = float(np.mean([r[k] for r in per_kernel_runs]))

    # N > 1: the serial step and the floor of its tail (all-gather + Pareto alone)
    strong = None
    if world > 1:
        for _ in range(3):
            step_serial()
        ms_serial = max_over_ranks(timed_steps(step_serial, a.steps, fence)) / a.steps
        with torch.cuda.stream(torch.cuda.current_stream()):
            for _ in range(3):
                tail(0)
            ms_tail = max_over_ranks(timed_steps(lambda: tail(0), a.steps, fence)) / a.steps
        kt_all = [None] * world
        dist.all_gather_object(kt_all, {k: round(v, 4) for k, v in ktimes.items()})
        strong = {"ms_per_step_pipelined": ms_step, "ms_per_step_serial": ms_serial,
                  "collective_and_pareto_floor_ms": ms_tail,
                  "regions_per_rank": [wl.shard_regions(nR, world, r)[1] - wl.shard_regions(nR, world, r)[0] for r in range(world)],
                  "tiles_per_rank": (per * nE + 31) // 32,
                  "kernel_ms_per_rank": kt_all,
                  "schedule": ("piped: eks_gain runs on a second stream beside ekf_forward, chunk by chunk behind a stream wait "
                               "on the forward pass's progress counter; eks_gain_tail is what is left when the forward pass ends")
                              if any("ekf_forward_piped" in (k_ or {}) for k_ in kt_all) else "one stream",
                  "limiter": "the time-sequential passes: ekf_forward and eks_backward walk the T days of a trajectory with one "
                             "warp per 32 trajectories (about 2 us and 1 us a day), so a shard of %d tiles on 148 SMs is latency- "
                             "not throughput-bound; the smoother tail (all-gather + Pareto) is the fixed floor" % ((per * nE + 31) // 32)}

    # extra (not the headline): the lean sweep mode -- identical outputs, smoother only on the days
    # whose schedule is optimised (include/epi_b200.h: epi_sweep_args.lean)
    ms_lean, lean_same = None, None
    if not a.no_lean and world == 1:
        out_lean = {k: torch.empty_like(v) for k, v in out1.items()}
        run_lean = lambda: wl.run_sweep(eng, dbatch, None, out=out_lean, lean=True)
        for _ in range(3):
            run_lean()
        ms_lean = timed_steps(run_lean, a.steps, fence) / a.steps
        lean_same = all(bool(torch.equal(out_lean[k], out1[k])) for k in out1)
    elif not a.no_lean:
        # N > 1: the same pipelined step (shard sweep + gather + Pareto tail) with the lean smoother
        drain()
        ref_J0 = full[(step_no[0] - 1) % nbuf]["J0"][:nR].clone()
        ref_front = full[(step_no[0] - 1) % nbuf]["on_front"].clone()

        def step_lean():
            i = step_no[0] % nbuf
            step_no[0] += 1
            if comm_ev[i] is not None:
                torch.cuda.current_stream().wait_event(comm_ev[i])
            wl.run_sweep(eng, dbatch, None, out={"J0": loc[i]["J0"][:n_loc], "J1": loc[i]["J1"][:n_loc]}, want_front=False,
                         lean=True)
            tail_async(i)
        for _ in range(3):
            step_lean()
        drain()
        ms_lean = max_over_ranks(timed_steps(step_lean, a.steps, fence, drain)) / a.steps
        last = (step_no[0] - 1) % nbuf
        lean_same = bool(torch.equal(full[last]["J0"][:nR], ref_J0) and torch.equal(full[last]["on_front"], ref_front))

    # ---- e2e: the blocking host-memory C-ABI call (H2D + kernels + D2H inside the call), pinned and pageable buffers
    def e2e_leg(pin):
        mk = (lambda t: t.pin_memory()) if pin else (lambda t: t)
        hb, h2d = dict(batch), 0
        for k in wl._SWEEP_ARRAYS:
            tt = mk(torch.from_numpy(np.array(batch[k], dtype=np.float64, order="C", copy=True)))
            hb[k] = tt.numpy()
            h2d += tt.numel() * 8
        peps = mk(torch.from_numpy(np.array(eps, dtype=np.float64, copy=True))).numpy()
        h2d += peps.size * 8 + len(bytes(memoryview(batch["prm"])))
        def hostbufs():
            o = {"J0": mk(torch.empty((n_loc, nE), dtype=f64)).numpy(), "J1": mk(torch.empty((n_loc, nE), dtype=f64)).numpy()}
            if world == 1:
                o["on_front"] = mk(torch.empty((nR, nE), dtype=u8)).numpy()
                o["I_opt"] = mk(torch.empty((nR,), dtype=i32)).numpy()
            return o
        houts = [hostbufs() for _ in range(nbuf)]
        d2h = sum(v.nbytes for v in houts[0].values())
        if world > 1:   # the gathered front/knee come back to the host as well
            hfront = [{"on_front": mk(torch.empty((nR, nE), dtype=u8)), "I_opt": mk(torch.empty((nR,), dtype=i32))}
                      for _ in range(nbuf)]
            d2h += nR * nE + nR * 4
            h2d += 2 * n_loc * nE * 8   # the shard's costs go back up for the NCCL gather
        drain()
        for i in range(nbuf):
            comm_ev[i] = None
        hstep_no = [0]

        def step_host():
            i = hstep_no[0] % nbuf
            hstep_no[0] += 1
            if comm_ev[i] is not None:   # the D2H of this call may overwrite houts[i] only after its last tail
                torch.cuda.current_stream().wait_event(comm_ev[i])
                comm_ev[i].synchronize()
            wl.run_sweep(eng, hb, peps, out=houts[i], want_front=(world == 1))   # blocking: H2D, kernels, D2H, sync
            if world > 1:
                loc[i]["J0"][:n_loc].copy_(torch.from_numpy(houts[i]["J0"]), non_blocking=True)
                loc[i]["J1"][:n_loc].copy_(torch.from_numpy(houts[i]["J1"]), non_blocking=True)
                tail_async(i)
                with torch.cuda.stream(comm):
                    hfront[i]["on_front"].copy_(full[i]["on_front"], non_blocking=True)
                    hfront[i]["I_opt"].copy_(full[i]["I_opt"], non_blocking=True)
                    comm_ev[i] = torch.cuda.Event()
                    comm_ev[i].record(comm)

        for _ in range(3):
            step_host()
        walls = []

        def one():
            tw = time.perf_counter()
            step_host()
            walls.append((time.perf_counter() - tw) * 1e3)
        ms_e2e = max_over_ranks(timed_steps(one, a.steps, fence, drain)) / a.steps
        if world == 1:  # parity of the two legs: same bits from the host-memory and device-memory modes
            assert np.array_equal(houts[0]["J0"], out1["J0"].cpu().numpy()), "host/device legs disagree"
        return {"value": units_total / (ms_e2e * 1e-3), "unit": "trajectory-days/s",
                "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "ms_per_step": ms_e2e,
                "ms_per_step_median": float(np.median(walls)), "ms_per_step_min": float(min(walls)),
                "ms_per_step_max": float(max(walls)), "host_buffers": "pinned" if pin else "pageable",
                "note": "value uses the total over all K steps (one attempt); bytes are this rank's"}

    e2e = e2e_pageable = None
    if not a.no_e2e:
        e2e = e2e_leg(True)
        e2e_pageable = e2e_leg(False)   # what a MATLAB/Octave host hands over (mxGetPr memory is pageable)
        e2e["pageable"] = {k: e2e_pageable[k] for k in ("value", "ms_per_step", "ms_per_step_median", "ms_per_step_min",
                                                         "ms_per_step_max")}

    # N > 1 extra: round 1's replicated (weak-scaling) form -- every rank a full 236-region replica
    replica = None
    if world > 1 and not a.no_replica:
        drain()
        inp_r, _ = build_inputs(a, rank)
        S_r = wl.run_fixed_input(eng, inp_r)
        db_r = wl.sweep_to_device(wl.sweep_batch(inp_r, S_r), eps, dev)
        out_r = {"J0": torch.empty((nR, nE), dtype=f64, device=dev), "J1": torch.empty((nR, nE), dtype=f64, device=dev),
                 "on_front": torch.empty((nR, nE), dtype=u8, device=dev), "I_opt": torch.empty((nR,), dtype=i32, device=dev)}
        run_r = lambda: wl.run_sweep(eng, db_r, None, out=out_r)
        for _ in range(3):
            run_r()
        k_r = max(3, a.steps // 2)
        ms_r = max_over_ranks(timed_steps(run_r, k_r, fence)) / k_r
        replica = {"value": units_total * world / (ms_r * 1e-3), "unit": "trajectory-days/s", "ms_per_step": ms_r,
                   "scaling": "weak", "note": "every rank sweeps its own 236-region replica; no collective in the step"}
        del db_r, out_r
        torch.cuda.empty_cache()
        eng.release_cache()

    # N > 1 extra: the whole 236-region sweep through ONE blocking host-memory call on all N GPUs
    # (epi_sweep_multi: what a single-process MATLAB/Octave host gets).  Rank 0 drives every GPU; the other
    # ranks wait on a CPU (gloo) barrier so that their devices are idle.
    host_multi = None
    if world > 1 and not a.no_e2e:
        drain()
        torch.cuda.synchronize()
        cpu_group = dist.new_group(backend="gloo")
        dist.barrier(group=cpu_group)
        if rank == 0:
            try:
                peers = [Engine(d) for d in range(world) if d != local]
                S_all = wl.run_fixed_input(eng, inp_all)
                full_batch = wl.sweep_batch(inp_all, S_all)
                hb = dict(full_batch)
                for k in wl._SWEEP_ARRAYS:
                    hb[k] = torch.from_numpy(np.array(full_batch[k], dtype=np.float64, order="C", copy=True)).pin_memory().numpy()
                heps = np.array(eps, dtype=np.float64, copy=True)
                hout = {"J0": np.empty((nR, nE)), "J1": np.empty((nR, nE)), "on_front": np.empty((nR, nE), dtype=np.uint8),
                        "I_opt": np.empty((nR,), dtype=np.int32)}
                call = lambda: wl.run_sweep(eng, hb, heps, out=hout, peers=peers)
                for _ in range(3):
                    call()
                walls = []
                for _ in range(a.steps):
                    tw0 = time.perf_counter()
                    call()
                    walls.append((time.perf_counter() - tw0) * 1e3)
                ms_hm = float(np.mean(walls))
                host_multi = {"value": units_total / (ms_hm * 1e-3), "unit": "trajectory-days/s", "ms_per_call": ms_hm,
                              "ms_per_call_min": float(min(walls)), "ms_per_call_max": float(max(walls)), "gpus": world,
                              "timing": "host wall clock around the blocking call (H2D, kernels, D2H on every GPU inside)",
                              "note": "one process, one call: include/epi_b200.h epi_sweep_multi; pinned host buffers"}
                for p_ in peers:
                    p_.close()
            except Exception as exc:
                host_multi = {"error": repr(exc)}
        dist.barrier(group=cpu_group)

    # ---- roofline of the dominant kernel
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        hbm_peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    else:
        hbm_peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    fp64_tflops = eng.fp64_probe(4096)
    dom = max((k for k in ktimes if k in KERNEL_ALGO), key=lambda k: ktimes[k])
    kern = {}
    for k, ms in ktimes.items():
        ab, fl = KERNEL_ALGO.get(k, (0.0, 0.0))
        kern[k] = {"ms": ms, "share": ms / max(1e-9, sum(ktimes.values())),
                   "hbm_gbs": ab * units_rank / (ms * 1e-3) / 1e9 if ms > 0 else 0.0,
                   "fp64_tflops": fl * units_rank / (ms * 1e-3) / 1e12 if ms > 0 else 0.0}
    achieved = kern[dom]["hbm_gbs"]
    traffic, traffic_src = None, None  # DRAM bytes per launch of the dominant kernel, from the committed ncu --set full capture
    try:
        if (a.regions, a.eps, a.t_hist, a.t_fore, world) == (236, 250, 441, 120, 1):
            for name in ("r02_traffic.json", "r01_traffic.json"):
                tp = os.path.join(ROOT, "profiles", name)
                if os.path.exists(tp) and dom in json.load(open(tp)):
                    traffic, traffic_src = json.load(open(tp))[dom]["traffic"], f"profiles/{name} (ncu --set full, per launch)"
                    break
    except Exception:
        traffic = None
    roofline = {"kernel": dom, "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                "frac": achieved / hbm_peak, "traffic": traffic, "traffic_source": traffic_src,
                "algorithmic_bytes_per_launch": KERNEL_ALGO[dom][0] * units_rank, "peak_source": peak_src,
                "algorithmic_bytes_per_trajectory_day": KERNEL_ALGO[dom][0],
                "fp64": {"achieved_tflops": kern[dom]["fp64_tflops"], "peak_tflops_measured": fp64_tflops,
                         "frac": kern[dom]["fp64_tflops"] / fp64_tflops if fp64_tflops else None,
                         "canonical_flops_per_trajectory_day": KERNEL_ALGO[dom][1]},
                "step": {"hbm_frac": BYTES_FUSED_M6 * units_total / world / (ms_step * 1e-3) / 1e9 / hbm_peak,
                         "fp64_frac": FLOPS_EKF_EKS_M6 * units_total / world / (ms_step * 1e-3) / 1e12 / fp64_tflops
                         if fp64_tflops else None,
                         "note": "per GPU: canonical work of the whole job / N / step time"},
                "kernels": kern}

    # ---- CPU baseline beside it (rank 0, N = 1 only)
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        from oracle import oracle as orc
        n, cores, _ = cpu_sample(a, inp, eps, a.cpu_seconds)
        sample = inp[:n]
        regs = oracle_regions(sample, S_fixed[:, :, :n])
        t0 = time.perf_counter()
        j0, j1, _, _ = orc.sweep_batch(regs, eps, n_threads=cores)
        dt = time.perf_counter() - t0
        # B3: the interpreter regime the reference actually runs in (MATLAB/Octave are absent):
        # the NumPy twin of the same .m files, one core, a handful of trajectories
        from oracle import numpy_twin as tw
        rin = inp[0]
        s6 = rin["setup6"]
        u6 = np.concatenate([rin["u_hist"], np.full((12, a.t_fore), np.nan)], axis=1)
        tw0 = time.perf_counter()
        n_tw = 0
        for e in eps[:: max(1, nE // 3)][:3]:
            tw.ekf_eks("optctrl", False, u6, rin["x"], dict(s6["params"], epsilon=float(e)), s6["s_init"], s6["Ps_init"],
                       s6["s_final"], s6["Ps_final"], 0.0, s6["Q_w"], rin["R_v"], s6["beta_ekf"], s6["gamma_ekf"], s6["W"])
            n_tw += 1
        tw_dt = time.perf_counter() - tw0
        cpu = {"value": n * nE * T / dt, "unit": "trajectory-days/s", "cores": cores, "kind": "port",
               "interpreted_proxy": {"value": n_tw * T / tw_dt, "unit": "trajectory-days/s", "cores": 1,
                                     "what": "NumPy twin of the same .m files (interpreter regime; EKF/EKS only)",
                                     "sample": f"{n_tw} trajectories x {T} days, {tw_dt:.2f} s"},
               "sample": f"{n} of {nR} regions x {nE} eps x {T} days, {dt:.1f} s (C oracle, OpenMP)",
               "parity_with_gpu_bit_exact": bool(np.array_equal(j0, out1["J0"].cpu().numpy()[:n]) and
                                                 np.array_equal(j1, out1["J1"].cpu().numpy()[:n]))}

    # ---- the other BASELINE configs (2, 3, 5), outside the headline timing: ms, rate, roofline fraction and an
    # oracle spot check each (tools/bench_configs.py)
    secondary = None
    if rank == 0 and world == 1 and not a.no_secondary:
        try:
            del dbatch
            torch.cuda.empty_cache()
            eng.release_cache()
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import bench_configs
            secondary = bench_configs.measure(eng, spot_check=True, fp64=fp64_tflops)
        except Exception as exc:  # the headline line must still be printed
            secondary = {"error": repr(exc)}

    # N > 1: BASELINE config 5 ("random-NPI Monte-Carlo scoring ... at 1/2/4/8 GPUs with Pareto all-gather") sharded by
    # region, outside the headline timing (tools/bench_configs.py measure_config5_sharded); N = 1 is in `secondary`
    config5 = None
    if world > 1 and not a.no_secondary:
        try:
            del dbatch
            torch.cuda.empty_cache()
            eng.release_cache()
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import bench_configs
            config5 = bench_configs.measure_config5_sharded(eng, rank, world, dev)
        except Exception as exc:
            config5 = {"error": repr(exc)}

    if rank == 0:
        par = ("1 GPU: the whole sweep, Pareto inside epi_sweep" if world == 1 else
               f"strong: {nR} regions sharded in contiguous blocks of {per} over {world} GPUs; per step ONE all-gather of the "
               "per-shard (J0,J1) (two NCCL calls, J0 and J1) then the Pareto step on the gathered costs, on a side stream "
               "overlapping the next step's sweep and drained inside the timed region")
        line = {"metric": "trajectory_days_per_sec", "value": value, "unit": "trajectory-days/s", "n_gpus": world,
                "steps": a.steps, "warmup": W, "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload_name(a), "trajectories_total": nR * nE, "trajectories_this_gpu": n_loc * nE,
                           "days": T, "parallelism": par,
                           "l2": "per-step tape traffic (>30 GB at N = 1, >3 GB per GPU at N = 8) far exceeds the 126 MB L2; no explicit flush",
                           "mode_value": "EPI_MEM_DEVICE (inputs resident in HBM)",
                           "mode_e2e": "EPI_MEM_HOST (host buffers, blocking call; pinned = headline, pageable beside it)"},
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
                "cpu_baseline": cpu, "strong_scaling": strong, "host_multi_call": host_multi, "replica_weak": replica, "secondary": secondary, "config5_sharded": config5,
                "lean_mode": None if ms_lean is None else {
                              "value": units_total / (ms_lean * 1e-3), "unit": "trajectory-days/s",
                              "ms_per_step": ms_lean, "outputs_bit_identical_to_full": lean_same,
                              "note": "not the headline: smoother gains/backward only on the days to optimise "
                                      "(epi_sweep_args.lean); same J0/J1/front/knee bits"}}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if eng_tail:
        eng_tail.close()
    eng.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
