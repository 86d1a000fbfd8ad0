# full GPU suite, then the bench at the strong-scaling shard sizes and the 1-GPU size
timeout 900 python -m pytest tests -m gpu -q -x --timeout=300 > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
for R in ${FULL_REGIONS:-30 59 236}; do
timeout 120 python bench.py --regions $R --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-lean --no-secondary > gpurun_out/full_${R}.log 2> gpurun_out/full_${R}.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/full_${R}.log").read().strip().splitlines()[-1])
    print("regions $R ms/step", round(d["ms_per_step"],3), {k:round(x["ms"],3) for k,x in d["roofline"]["kernels"].items()})
except Exception as e:
    print("regions $R FAILED", e)
PY
done
