timeout 900 python -m pytest tests -m gpu -q -x --timeout=600 2>&1 | tail -2
for R in 30 59 118 236; do for PF in 0 3; do
EPI_BWD_PREFETCH=$PF python bench.py --regions $R --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-lean --no-secondary > gpurun_out/sm_${R}.log 2> gpurun_out/sm_${R}.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/sm_${R}.log").read().strip().splitlines()[-1])
    print("regions $R bwd_prefetch=$PF ms/step", round(d["ms_per_step"],3), {k:round(x["ms"],3) for k,x in d["roofline"]["kernels"].items()})
except Exception as e:
    print("regions $R FAILED", e)
PY
done; done
