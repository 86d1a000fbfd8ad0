# lane-group (EPI_ROWS=1) vs one-thread (EPI_ROWS=0) kernels at shard sizes of the strong-scaling sweep
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout=200 -k "lane_group" 2>&1 | tail -2
for R in ${ROWS_REGIONS:-30 59 118}; do for M in 0 1; do
EPI_ROWS=$M python bench.py --regions $R --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-lean --no-secondary > gpurun_out/rows_${R}_$M.log 2> gpurun_out/rows_${R}_$M.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/rows_${R}_$M.log").read().strip().splitlines()[-1])
    print("regions $R rows=$M ms/step", round(d["ms_per_step"],3), {k:round(x["ms"],3) for k,x in d["roofline"]["kernels"].items()})
except Exception as e:
    print("regions $R rows=$M FAILED", e)
PY
done; done
