"""Aggregate an `ncu --page source --csv` dump: executed instructions and stall samples per opcode
class and per stall reason (SASS view; source mapping needs --import-source + -lineinfo).
    ncu -i rep --page source --csv > x.csv ; python tools/ncu_src.py x.csv
"""
import csv, sys, collections, re
rows = list(csv.reader(open(sys.argv[1])))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
col = {n: i for i, n in enumerate(hdr)}
ops = collections.Counter(); samp = collections.Counter(); stalls = collections.Counter()
tot_inst = 0; tot_samp = 0
stall_cols = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
for r in rows[hdr_i + 1:]:
    if len(r) < len(hdr): continue
    src = r[col["Source"]]
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", src)
    op = m.group(2) if m else "?"
    try:
        ne = int(float(r[col["Instructions Executed"]])); ns = int(float(r[col["# Samples"]]))
    except ValueError:
        continue
    ops[op] += ne; samp[op] += ns; tot_inst += ne; tot_samp += ns
    for s in stall_cols:
        try: stalls[s] += int(float(r[col[s]]))
        except ValueError: pass
print("total warp instructions", tot_inst, "samples", tot_samp)
print("by opcode (executed share | sample share):")
for op, n in ops.most_common(28):
    print(f"  {op:10s} {100*n/tot_inst:5.1f}%  {100*samp[op]/max(1,tot_samp):5.1f}%")
print("stall reasons (share of samples):")
ts = sum(stalls.values())
for s, n in stalls.most_common(12):
    print(f"  {s:24s} {100*n/max(1,ts):5.1f}%")
