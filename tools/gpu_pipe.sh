# piped schedule (forward || gains) on the strong-scaling shard sizes: parity, then ms/step with and without
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout=300 -k "piped or sweep_matches or two_warp" 2>&1 | tail -3
for R in ${PIPE_REGIONS:-30 59}; do for M in ${PIPE_MODES:-0 1}; do
EPI_PIPE=$M timeout 300 python bench.py --regions $R --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-lean --no-secondary > gpurun_out/pipe_${R}_$M.log 2> gpurun_out/pipe_${R}_$M.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/pipe_${R}_$M.log").read().strip().splitlines()[-1])
    print("regions $R pipe=$M chunks=${EPI_PIPE_CHUNKS:-8} ms/step", round(d["ms_per_step"],3), {k:round(x["ms"],3) for k,x in d["roofline"]["kernels"].items()})
except Exception as e:
    print("regions $R pipe=$M FAILED", e)
PY
done; done
