timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout=200 -k "lane_group" 2>&1 | tail -1
for R in 30 59; do
python bench.py --regions $R --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-lean --no-secondary > gpurun_out/rowsd_${R}.log 2> gpurun_out/rowsd_${R}.err
python - <<PY
import json
d=json.loads(open("gpurun_out/rowsd_${R}.log").read().strip().splitlines()[-1])
print("regions $R default ms/step", round(d["ms_per_step"],3), {k:round(x["ms"],3) for k,x in d["roofline"]["kernels"].items()})
PY
done
