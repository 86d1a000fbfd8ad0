# usage: bash tools/gpu_variants.sh v0 v1 ...   (libraries built by tools/build_variants.py)
for v in "$@"; do
  export EPI_B200_LIB=$PWD/epidemicmodeling_b200/variants/$v/libepi_b200.so
  python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-lean --no-secondary > gpurun_out/var_$v.log 2> gpurun_out/var_$v.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/var_$v.log").read().strip().splitlines()[-1])
    print("$v", "ms/step", round(d["ms_per_step"],3), {k:round(x["ms"],3) for k,x in d["roofline"]["kernels"].items()})
except Exception as e:
    print("$v", "FAILED", e)
PY
  if [ -n "$PARITY" ]; then timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "sweep or ekf6 or optctrl" 2>&1 | tail -1; fi
done
