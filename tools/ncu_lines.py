"""Attribute an `ncu --page source --csv` dump (SASS view) to SOURCE LINES, offline.

ncu's CSV carries no line column; `nvdisasm -g` of the same cubin does.  The two list the kernel's
instructions in the same order (16 bytes apart), so instruction i of the CSV is instruction i of the
disassembly.

    cuobjdump -xelf all epidemicmodeling_b200/build/eks_gain.o      # -> eks_gain.sm_100a.cubin
    nvdisasm -g -c eks_gain.sm_100a.cubin > gain.sass
    ncu -i rep.ncu-rep --page source --csv > gain.csv
    python tools/ncu_lines.py gain.csv gain.sass <mangled-kernel-substring> [top_n]

Prints executed warp instructions / stall samples per (file:line) with the FP64 share, and per file.
"""
import collections
import csv
import re
import sys


def sass_lines(path, kernel):
    """[(opcode, file, line)] for the instructions of the first .text section whose name contains `kernel`."""
    out, on, cur = [], False, ("?", 0)
    for ln in open(path):
        if ln.startswith(".text."):
            if on:
                break
            on = kernel in ln
            continue
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
        if m:
            out.append((m.group(3), cur[0], cur[1]))
    return out


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    sass = sass_lines(sys.argv[2], sys.argv[3])
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    col = {n: i for i, n in enumerate(rows[hi])}
    body = [r for r in rows[hi + 1:] if len(r) >= len(rows[hi])]
    if len(body) != len(sass):
        print(f"warning: {len(body)} CSV instructions vs {len(sass)} disassembled", file=sys.stderr)
    per = collections.defaultdict(lambda: [0, 0, 0, 0])   # inst, fp64 inst, samples, static
    tot = [0, 0]
    for r, (op, f, l) in zip(body, sass):
        ne, ns = int(float(r[col["Instructions Executed"]])), int(float(r[col["# Samples"]]))
        p = per[(f, l)]
        p[0] += ne
        p[1] += ne if op[0] == "D" and op not in ("DEPBAR",) else 0
        p[2] += ns
        p[3] += 1
        tot[0] += ne
        tot[1] += ns
    print(f"total warp instructions {tot[0]}, samples {tot[1]}, static instructions {len(sass)}")
    print(f"{'file:line':34s} {'inst%':>6s} {'fp64%of line':>12s} {'samples%':>8s} {'static':>6s}")
    for (f, l), p in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{f + ':' + str(l):34s} {100 * p[0] / tot[0]:6.2f} {100 * p[1] / max(1, p[0]):12.1f} "
              f"{100 * p[2] / max(1, tot[1]):8.2f} {p[3]:6d}")
    byf = collections.defaultdict(lambda: [0, 0])
    for (f, l), p in per.items():
        byf[f][0] += p[0]
        byf[f][1] += p[2]
    print("per file:")
    for f, p in sorted(byf.items(), key=lambda kv: -kv[1][0]):
        print(f"  {f:28s} {100 * p[0] / tot[0]:6.2f}%  samples {100 * p[1] / max(1, tot[1]):6.2f}%")


if __name__ == "__main__":
    main()
