"""Summarise ncu captures brought back in gpurun_out/ into profiles/ (tracked).

    python tools/ncu_summary.py r01            # round tag

Reads  gpurun_out/launches.csv   (ncu --metrics gpu__time_duration.sum launch list)
       gpurun_out/prof_*.ncu-rep  (ncu --set full, one kernel each)
Writes profiles/<tag>_launches.csv (copy), profiles/<tag>_launch_shares.md and
       profiles/<tag>_<kernel>.md (key raw metrics + top stall reasons per source line).
"""
import csv
import glob
import io
import os
import re
import shutil
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GO = os.path.join(ROOT, "gpurun_out")
PR = os.path.join(ROOT, "profiles")

KEYS = [
    "gcc__cache_requests_type_instruction.sum", "gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed",
    "sm__icc_request_hit_rate.pct", "sm__icc_requests.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "launch__waves_per_multiprocessor", "launch__grid_size", "launch__block_size",
    "sm__inst_executed_pipe_fp64.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
    "lts__t_sector_hit_rate.pct", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct",
    "smsp__warp_issue_stalled_wait_per_warp_active.pct", "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct",
    "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_no_instruction_per_warp_active.pct",
    "smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct", "smsp__warp_issue_stalled_dispatch_stall_per_warp_active.pct",
    "smsp__warp_issue_stalled_not_selected_per_warp_active.pct", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
    "smsp__warp_issue_stalled_drain_per_warp_active.pct", "smsp__warp_issue_stalled_imc_miss_per_warp_active.pct",
    "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct", "local_load", "local_store",
]


def ncu(*args):
    return subprocess.run(["ncu", *args], capture_output=True, text=True).stdout


def launch_shares(tag):
    src = os.path.join(GO, "launches.csv")
    if not os.path.exists(src):
        return
    os.makedirs(PR, exist_ok=True)
    shutil.copy(src, os.path.join(PR, f"{tag}_launches.csv"))
    txt = open(src).read()
    start = txt.find('"ID"')
    rows = list(csv.DictReader(io.StringIO(txt[start:])))
    tot = defaultdict(float)
    cnt = defaultdict(int)
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        tot[name] += v
        cnt[name] += 1
    allms = sum(tot.values())
    with open(os.path.join(PR, f"{tag}_launch_shares.md"), "w") as f:
        f.write(f"# ncu launch list ({tag}): gpu__time_duration per kernel (cold-cache, serialised: compare SHARES)\n\n")
        f.write("| kernel | launches | total ms | mean ms | share |\n|---|---|---|---|---|\n")
        for k in sorted(tot, key=lambda k: -tot[k]):
            f.write(f"| `{k}` | {cnt[k]} | {tot[k]:.3f} | {tot[k] / cnt[k]:.3f} | {100 * tot[k] / allms:.1f}% |\n")
    print(open(os.path.join(PR, f"{tag}_launch_shares.md")).read())


def full(tag):
    for rep in sorted(glob.glob(os.path.join(GO, "prof_*.ncu-rep"))):
        kname = os.path.basename(rep)[5:-8]
        raw = ncu("-i", rep, "--page", "raw", "--csv")
        rows = list(csv.reader(io.StringIO(raw)))
        if len(rows) < 3:
            print("no data in", rep)
            continue
        hdr, units, vals = rows[0], rows[1], rows[2]
        out = [f"# ncu --set full: {kname} ({tag})\n", "| metric | value | unit |", "|---|---|---|"]
        for h, u, v in zip(hdr, units, vals):
            if any(h == k or (k in ("local_load", "local_store") and k in h) for k in KEYS) or "warp_issue_stalled" in h and h.endswith("per_warp_active.pct"):
                out.append(f"| {h} | {v} | {u} |")
        # source page (CUDA lines correlated with SASS): instructions and stall samples per line
        srcp = ncu("-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass")
        cur, agg = None, []
        for r in csv.reader(io.StringIO(srcp)):
            if len(r) == 2 and r[0] == "File Path":
                cur = os.path.basename(r[1])
                continue
            if len(r) > 8 and r[0].isdigit() and r[2] == "-":
                try:
                    inst, samp = int(r[7]), int(r[6])
                except ValueError:
                    continue
                if inst > 0 or samp > 0:
                    agg.append((samp, inst, cur, int(r[0]), r[1].strip()[:110]))
        if agg:
            tot_i = sum(a_[1] for a_ in agg) or 1
            tot_s = sum(a_[0] for a_ in agg) or 1
            agg.sort(reverse=True)
            out.append("\n## hottest source lines (share of warp-stall samples | share of executed instructions)\n")
            for samp, inst, f, ln, txt in agg[:30]:
                out.append(f"- {100 * samp / tot_s:4.1f}% | {100 * inst / tot_i:4.1f}%  `{f}:{ln}`  `{txt}`")
        path = os.path.join(PR, f"{tag}_{kname}.md")
        open(path, "w").write("\n".join(out) + "\n")
        print("wrote", path)


def traffic(tag):
    """profiles/<tag>_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum per launch of every captured kernel
    (bench.py reads roofline.traffic from it)."""
    import json
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    out = {}
    for rep in sorted(glob.glob(os.path.join(GO, "prof_*.ncu-rep"))):
        kname = os.path.basename(rep)[5:-8]
        rows = list(csv.reader(io.StringIO(ncu("-i", rep, "--page", "raw", "--csv"))))
        if len(rows) < 3:
            continue
        col = {h: i for i, h in enumerate(rows[0])}
        try:
            rd = float(rows[2][col["dram__bytes_read.sum"]]) * scale[rows[1][col["dram__bytes_read.sum"]]]
            wr = float(rows[2][col["dram__bytes_write.sum"]]) * scale[rows[1][col["dram__bytes_write.sum"]]]
        except (KeyError, ValueError):
            continue
        out[kname] = {"dram_bytes_read": rd, "dram_bytes_write": wr, "traffic": rd + wr}
    out["_note"] = ("dram__bytes_read.sum + dram__bytes_write.sum per launch from ncu --set full (profiles/%s_*.md); forward/gain/backward "
                    "on the bench workload 236x250x561; seirp_staged at B=500000, rollout_staged at B=5.9M" % tag)
    json.dump(out, open(os.path.join(PR, f"{tag}_traffic.json"), "w"), indent=1)
    print("wrote", os.path.join(PR, f"{tag}_traffic.json"))


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    launch_shares(tag)
    full(tag)
    traffic(tag)
