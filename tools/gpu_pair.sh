timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout=200 -k "two_warp or sweep_matches or guard" 2>&1 | tail -3
for R in ${PAIR_REGIONS:-30 59 118}; do for M in 0 1; do
EPI_PAIR=$M python bench.py --regions $R --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-lean --no-secondary > gpurun_out/pair_${R}_$M.log 2> gpurun_out/pair_${R}_$M.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/pair_${R}_$M.log").read().strip().splitlines()[-1])
    print("regions $R pair=$M ms/step", round(d["ms_per_step"],3), {k:round(x["ms"],3) for k,x in d["roofline"]["kernels"].items()})
except Exception as e:
    print("regions $R pair=$M FAILED", e)
PY
done; done
