for cfg in "1 1184" "3 1184" "7 1184" "12 1184" "24 1184" "7 592" "7 888" "14 2368"; do
  set -- $cfg
  EPI_FWD_SEGMENTS=$1 EPI_FWD_GRID=$2 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-lean > gpurun_out/fx.log 2> gpurun_out/fx.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/fx.log").read().strip().splitlines()[-1])
print("S=$1 grid=$2", {k:round(x["ms"],3) for k,x in d["roofline"]["kernels"].items() if k=="ekf_forward"})
PY
done
