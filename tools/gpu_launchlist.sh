B="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-lean"
$B > gpurun_out/plain.log 2> gpurun_out/plain.err || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum,launch__grid_size,launch__registers_per_thread,launch__occupancy_limit_registers --clock-control none -c 60 --csv --log-file gpurun_out/launches_short.csv $B > gpurun_out/ncu_launches.log 2>&1
python - <<PY
import csv,io
txt=open("gpurun_out/launches_short.csv").read(); txt=txt[txt.find('"ID"'):]
for r in csv.DictReader(io.StringIO(txt)):
    if int(r["ID"])>35: print(r["ID"], r["Kernel Name"][:60], r["Metric Name"], r["Metric Value"])
PY
