timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout=100 -k "staged_backward or piped or sweep_matches" 2>&1 | tail -3
run() { # name, regions, env...
  name=$1; R=$2; shift; shift
  env "$@" timeout 60 python bench.py --regions $R --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-lean --no-secondary > gpurun_out/bwd_$name.log 2> gpurun_out/bwd_$name.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bwd_$name.log").read().strip().splitlines()[-1])
    print("$name regions $R ms/step", round(d["ms_per_step"],3), {k:round(x["ms"],3) for k,x in d["roofline"]["kernels"].items()})
except Exception as e:
    print("$name FAILED", e)
PY
}
run auto_30 30
run s4_30 30 EPI_BWD_STAGES=4
run s6_59 59 EPI_BWD_STAGES=4
run auto_59 59
